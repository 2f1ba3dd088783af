#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref).  Run in the
container that has /root/reference; the files are the fallback when libhnh_ref.so is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests import mp_util as U  # noqa: E402
from tests.test_multirank_cpu import CASES_2, CASES_4  # noqa: E402
from tests.test_multirank_gpu import CASES  # noqa: E402

assert ref.build(), "oracle/_ref could not be built (no /root/reference?)"
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
todo = [(2, c) for c in CASES_2 if c["logM"] <= 9] + [(4, c) for c in CASES_4] + [(p, c) for p, cs in CASES.items() for c in cs]
seen = set()
for p, c in todo:
    key = (c["name"], p)
    if c["name"].startswith("nogolden"):
        continue
    if key in seen:
        continue
    seen.add(key)
    ranks, _ = U.reference_for(c, p)
    path = os.path.join(ROOT, "tests", "golden", f"{c['name']}_p{p}.npz")
    np.savez_compressed(path, **U.flatten_ref(ranks))
    print(path, os.path.getsize(path))
