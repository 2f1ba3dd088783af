// compat forwarding header: `#include "json.hpp"` + `using json = nlohmann::json;` of the reference's sources.
#pragma once
#include "hnh/json.h"
namespace nlohmann {
using json = hnh::Json;
}
