// hnh/json.h -- a minimal ordered JSON value (objects, arrays, numbers, strings, bools) used
// for the benchmark records and algorithm info that the reference emits with nlohmann::json
// (distributed_sparse.h:131-179,245-261; benchmark_dist.cpp:144-162).  Not a parser.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

namespace hnh {

class Json {
public:
    enum Kind { Null, Bool, Int, Double, String, Array, Object };
    Json() : kind_(Null) {}
    Json(bool b) : kind_(Bool), b_(b) {}
    Json(int v) : kind_(Int), i_(v) {}
    Json(long v) : kind_(Int), i_(v) {}
    Json(long long v) : kind_(Int), i_(v) {}
    Json(unsigned v) : kind_(Int), i_(v) {}
    Json(unsigned long v) : kind_(Int), i_((long long)v) {}
    Json(unsigned long long v) : kind_(Int), i_((long long)v) {}
    Json(double v) : kind_(Double), d_(v) {}
    Json(const char *s) : kind_(String), s_(s) {}
    Json(const std::string &s) : kind_(String), s_(s) {}
    static Json array() { Json j; j.kind_ = Array; return j; }
    static Json object() { Json j; j.kind_ = Object; return j; }

    Json &operator[](const std::string &key) {
        if (kind_ == Null) kind_ = Object;
        for (auto &kv : members_)
            if (kv.first == key) return kv.second;
        members_.emplace_back(key, Json());
        return members_.back().second;
    }
    void push_back(const Json &v) {
        if (kind_ == Null) kind_ = Array;
        items_.push_back(v);
    }
    bool is_null() const { return kind_ == Null; }

    std::string dump(int indent = -1) const {
        std::string out;
        write(out, indent, 0);
        return out;
    }

private:
    static void escape(std::string &out, const std::string &s) {
        out += '"';
        for (char ch : s) {
            switch (ch) {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\t': out += "\\t"; break;
                default:
                    if ((unsigned char)ch < 0x20) {
                        char buf[8];
                        snprintf(buf, sizeof buf, "\\u%04x", ch);
                        out += buf;
                    } else {
                        out += ch;
                    }
            }
        }
        out += '"';
    }
    static void newline(std::string &out, int indent, int depth) {
        if (indent < 0) return;
        out += '\n';
        out.append((size_t)indent * depth, ' ');
    }
    void write(std::string &out, int indent, int depth) const {
        char buf[64];
        switch (kind_) {
            case Null: out += "null"; break;
            case Bool: out += b_ ? "true" : "false"; break;
            case Int: snprintf(buf, sizeof buf, "%lld", i_); out += buf; break;
            case Double:
                if (std::isfinite(d_)) { snprintf(buf, sizeof buf, "%.17g", d_); out += buf; }
                else out += "null";
                break;
            case String: escape(out, s_); break;
            case Array:
                out += '[';
                for (size_t i = 0; i < items_.size(); i++) {
                    if (i) out += ',';
                    newline(out, indent, depth + 1);
                    items_[i].write(out, indent, depth + 1);
                }
                if (!items_.empty()) newline(out, indent, depth);
                out += ']';
                break;
            case Object:
                out += '{';
                for (size_t i = 0; i < members_.size(); i++) {
                    if (i) out += ',';
                    newline(out, indent, depth + 1);
                    escape(out, members_[i].first);
                    out += indent < 0 ? ":" : ": ";
                    members_[i].second.write(out, indent, depth + 1);
                }
                if (!members_.empty()) newline(out, indent, depth);
                out += '}';
                break;
        }
    }
    Kind kind_;
    bool b_ = false;
    long long i_ = 0;
    double d_ = 0.0;
    std::string s_;
    std::vector<Json> items_;
    std::vector<std::pair<std::string, Json>> members_;
};

}  // namespace hnh

using json = hnh::Json;
