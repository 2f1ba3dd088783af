"""bench.py's parity checker (sample_parity / pattern_local) on CPU: a stand-in algorithm object with an R-split,
multi-block local layout computes the FusedMM with scipy on the global pattern operands; the checker must accept it
and must catch a corrupted result.  (The real algorithm objects need a GPU; this pins the checker's own layout logic.)"""
import numpy as np
import pytest
import scipy.sparse as sp

import bench
from distributed_sddmm_b200 import lib
from oracle import hnh_oracle as orc
from oracle import ref


class _Dense:
    def __init__(self, shape):
        self.shape = shape
        self.host = np.zeros(shape)

    def from_host(self, a):
        assert a.shape == self.shape
        self.host = np.array(a, dtype=np.float64)

    def to_host(self):
        return self.host.copy()


class _FakeAlg:
    """Holds row blocks `blocks` of A (each `rows` high) restricted to columns [left, left + nc) of R, stacked."""

    def __init__(self, logM, npr, R, blocks, rows, left, nc, corrupt=False):
        self.N, self.R, self.left, self.nc, self.rows, self.blocks, self.corrupt = 1 << logM, R, left, nc, rows, blocks, corrupt
        r, c, v = orc.er_tuples(logM, npr, bench.SEED)
        self.S = sp.csr_matrix((v, (r.astype(np.int64), c.astype(np.int64))), shape=(self.N, self.N))
        self.shifted = 0

    def submatrices(self, which):
        return np.array([[b * self.rows, self.left, self.rows, self.nc] for b in self.blocks], dtype=np.int32)

    def initial_shift(self, A, B, mode):
        self.shifted += 1

    def de_shift(self, A, B, mode):
        self.shifted -= 1

    def fusedSpMM(self, A, B, Sv, res, mode):
        # the operands this object was given must be its slices of the global pattern matrices
        GA, GB = ref.pattern(self.N, self.R, 1), ref.pattern(self.N, self.R, 2)
        want_in = np.vstack([GA[b * self.rows:(b + 1) * self.rows, self.left:self.left + self.nc] for b in self.blocks])
        assert np.array_equal(A.host, want_in)
        vals = self.S.multiply(GA @ GB.T).tocsr()
        out = (vals @ GB)[:, self.left:self.left + self.nc]
        A.host = np.vstack([out[b * self.rows:(b + 1) * self.rows] for b in self.blocks])
        if self.corrupt:
            A.host[self.rows // 2 + 3, 0] += 0.5 * np.abs(A.host).max()


@pytest.mark.parametrize("left,nc", [(0, 16), (4, 4), (8, 8)])
def test_sample_parity_accepts_a_correct_r_split_layout_and_catches_a_wrong_one(hnh, left, nc):
    logM, npr, R, rows = 10, 8, 16, 128
    for corrupt in (False, True):
        alg = _FakeAlg(logM, npr, R, blocks=[1, 5, 6], rows=rows, left=left, nc=nc, corrupt=corrupt)
        A, B = _Dense((3 * rows, nc)), _Dense((3 * rows, nc))
        _, rec = bench.sample_parity(logM, npr, R, alg, A, B, None, None, 0, 1, shifts=True, rows_per_block=rows)
        assert alg.shifted == 0 and rec["rows"] == 3 * rows and rec["nnz"] > 0
        if corrupt:
            assert rec["max_rel_err"] > bench.PARITY_RTOL
        else:
            assert rec["max_rel_err"] < 1e-12, rec


def test_pattern_local_zeroes_padding_rows():
    m = bench.pattern_local(np.array([[96, 2, 64, 3]]), (64, 3), 1, nrows_global=100)
    assert np.array_equal(m[:4], ref.pattern(4, 3, 1, row0=96, col0=2)) and not m[4:].any()
