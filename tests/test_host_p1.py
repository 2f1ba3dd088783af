"""Host classes at p = 1 on a CPU-only box (Self transport): tuple loading, redistribution,
block splitting and CSR construction of every algorithm, against the oracle's COO->CSR.
(No kernel runs here: compute needs the GPU and is covered by the -m gpu tests.)"""
import numpy as np
import pytest

from distributed_sddmm_b200 import driver as D
from oracle import hnh_oracle as orc

SEED = 0xC0FFEE + 1


@pytest.fixture(scope="module")
def world():
    r, n = D.world_init("self")
    assert (r, n) == (0, 1)
    yield
    D.world_finalize()


def test_load_er_matches_oracle(world):
    S = D.SpmatLocal.load_er(10, 8, SEED)
    ref = orc.er_tuples(10, 8, SEED)
    info = S.info()
    assert info["M"] == info["N"] == 1024 and info["dist_nnz"] == len(ref[0]) == info["local_tuples"]
    r, c, v = S.tuples()
    assert np.array_equal(r, ref[0]) and np.array_equal(c, ref[1]) and np.array_equal(v, ref[2])


@pytest.mark.parametrize("name,transposed", [("15d_fusion1", True), ("15d_fusion2", False), ("15d_sparse", False),
                                             ("25d_dense_replicate", True), ("25d_sparse_replicate", False)])
# (14, 40): 650k tuples -- the multi-threaded paths of the tuple bucketing and of the CSR counting sort (host_sort.h)
@pytest.mark.parametrize("logM,npr", [(9, 6), (14, 40)])
def test_p1_layout_and_blocks(world, name, transposed, logM, npr):
    R = 16
    N = 1 << logM
    S = D.SpmatLocal.load_er(logM, npr, SEED)
    rows, cols, vals = orc.er_tuples(logM, npr, SEED)
    alg = D.Algorithm(name, S, R, 1)
    d = alg.dims
    assert (d.M, d.N, d.R, d.p, d.c) == (N, N, R, 1, 1)
    assert (d.localArows, d.localAcols, d.localBrows, d.localBcols) == (N, R, N, R)
    assert alg.submatrices("A").tolist() == [[0, 0, N, R]]
    assert d.s_values == d.st_values == len(rows)
    # S block == CSR of S (or of S stored transposed); ST block == CSR of S^T (or transposed again)
    refS = orc.coo_to_csr(N, N, rows, cols, vals, transpose=transposed)
    order = np.lexsort((cols, rows))  # ST tuples (c, r) sorted column-major == sorted by (r, c) of S
    refST = orc.coo_to_csr(N, N, cols[order], rows[order], vals[order], transpose=transposed)
    for which, ref in (("S", refS), ("ST", refST)):
        blocks = alg.blocks(which)
        assert len(blocks) == 1 and blocks[0] is not None
        b = blocks[0]
        assert b["transpose"] == transposed and b["rows"] == N
        assert np.array_equal(b["rowStart"], ref.rowStart)
        assert np.array_equal(b["col_idx"], ref.col_idx)
        assert np.array_equal(b["row_idx"], ref.row_idx)
        assert np.array_equal(b["values"], ref.values)
    info = alg.info()
    assert info["nnz"] == len(rows) and info["p"] == 1 and info["nnz_procs"] == [len(rows)]


def test_bad_arguments_raise(world):
    S = D.SpmatLocal.load_er(6, 4, SEED)
    with pytest.raises(RuntimeError, match="unknown algorithm"):
        D.Algorithm("nope", S, 8, 1)
    with pytest.raises(RuntimeError, match="c divide"):
        D.Algorithm("15d_fusion1", S, 8, 2)
    with pytest.raises(RuntimeError, match="perfect square"):
        D.Algorithm("25d_dense_replicate", S, 8, 3)


def test_load_matrix_market_file(world, tmp_path):
    """loadTuples(readFromFile = true): general and symmetric coordinate files, 1-based indices."""
    gen = tmp_path / "g.mtx"
    gen.write_text("%%MatrixMarket matrix coordinate real general\n% comment\n4 5 3\n1 1 2.5\n4 5 -1\n2 3 7\n")
    S = D.SpmatLocal.load_file(str(gen))
    assert S.info() == {"M": 4, "N": 5, "dist_nnz": 3, "local_tuples": 3}
    r, c, v = S.tuples()
    assert sorted(zip(r.tolist(), c.tolist(), v.tolist())) == [(0, 0, 2.5), (1, 2, 7.0), (3, 4, -1.0)]
    sym = tmp_path / "s.mtx"
    sym.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n3 3 2\n2 1\n3 3\n")
    S = D.SpmatLocal.load_file(str(sym))
    r, c, v = S.tuples()
    assert sorted(zip(r.tolist(), c.tolist(), v.tolist())) == [(0, 1, 1.0), (1, 0, 1.0), (2, 2, 1.0)]
    alg = D.Algorithm("15d_fusion2", S, 4, 1)
    assert alg.dims.M == 3 and alg.dims.s_values == 3
    with pytest.raises(RuntimeError, match="cannot open"):
        D.SpmatLocal.load_file(str(tmp_path / "missing.mtx"))
