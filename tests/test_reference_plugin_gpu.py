"""INTEGRATION.md route B, executed: the REFERENCE's own host code (its algorithm classes, redistribution, value
plumbing, shift loops -- compiled unmodified into oracle/_ref) with ONE change, the one a maintainer would make:
`StandardKernel` replaced by `CudaKernel` (include/hnh/reference_plugin/cuda_kernel.h), which sends the two local
kernels to libhnh_b200.so through the reference's KernelImplementation interface.  Every public operation of every
algorithm must then give the same per-rank results as the unmodified reference."""
import os

import numpy as np
import pytest

from oracle import hnh_oracle as orc
from oracle import ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref.available() and os.path.exists(ref.SO_CUDA)), reason="oracle/_ref builds are absent")]

OPS = ["sddmmA", "spmmA", "spmmB", "fusedA", "sddmmB", "fusedB"]


@pytest.mark.parametrize("alg,p,c", [("15d_fusion1", 1, 1), ("15d_fusion2", 1, 1), ("15d_sparse", 1, 1), ("25d_dense_replicate", 1, 1),
                                     ("25d_sparse_replicate", 1, 1), ("15d_fusion1", 4, 2), ("15d_fusion2", 4, 1), ("15d_fusion2", 8, 4),
                                     ("15d_sparse", 4, 1), ("15d_sparse", 8, 2), ("25d_dense_replicate", 4, 1),
                                     ("25d_dense_replicate", 8, 2), ("25d_sparse_replicate", 4, 1), ("25d_sparse_replicate", 8, 2)])
def test_reference_host_code_with_cuda_kernels_plugged_in(alg, p, c):
    logM, npr, R = 7, 5, 8
    N = 1 << logM
    rows, cols, _ = orc.er_tuples(logM, npr, 0xC0FFEE + 6)
    rng = np.random.default_rng(11)
    A, B = rng.uniform(-1, 1, (N, R)), rng.uniform(-1, 1, (N, R))
    sv = rng.uniform(0.5, 1.5, len(rows))
    want = ref.run(alg, p, c, R, N, N, rows, cols, sv, A, B, OPS)
    got = ref.run(alg, p, c, R, N, N, rows, cols, sv, A, B, OPS, plugin=True)
    for r, (g, w) in enumerate(zip(got, want)):
        for key in ("S_rows", "S_cols", "ST_rows", "ST_cols"):  # recovered THROUGH the plugged-in SDDMM: exact integers
            assert np.array_equal(g[key], w[key]), (r, key)
        for t, op in enumerate(OPS):
            for f in ("A", "B", "values"):
                a, b = g["ops"][t][f], w["ops"][t][f]
                assert a.shape == b.shape, (r, op, f)
                if a.size:
                    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
                    assert err < 1e-11, (r, op, f, err)
