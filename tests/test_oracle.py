"""CPU tests of the oracle itself (no GPU): the C restatement against scipy on global
matrices, against the dummyInitialize closed form, and the CSR
construction order (SpmatLocal.hpp:78-188)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import hnh_oracle as orc


@pytest.mark.parametrize("logM,npr,R", [(8, 4, 8), (10, 8, 16), (9, 6, 5), (10, 32, 128)])
def test_oracle_matches_scipy(logM, npr, R):
    rows, cols, vals = orc.er_tuples(logM, npr, seed=0xC0FFEE + 1)
    N = 1 << logM
    rng = np.random.default_rng(logM * 100 + R)
    A = rng.uniform(-1, 1, (N, R))
    B = rng.uniform(-1, 1, (N, R))
    sv = rng.uniform(0.5, 1.5, len(rows))
    S, sddmm, spmmA, spmmB, fused = orc.global_reference(rows, cols, sv, A, B)
    csr = orc.coo_to_csr(N, N, rows, cols, sv)
    # CSR indices bit-exact against scipy's canonical (row asc, col asc) order
    assert np.array_equal(csr.rowStart, S.indptr.astype(np.int64))
    assert np.array_equal(csr.col_idx, S.indices.astype(np.int64))
    assert np.array_equal(csr.values, S.data)
    assert np.array_equal(csr.row_idx, np.repeat(np.arange(N), np.diff(S.indptr)))
    # SDDMM (values zeroed first, then Hadamard with SValues as 15D_dense_shift.hpp:366)
    v = np.zeros(csr.nnz)
    orc.sddmm_coo(csr.row_idx, csr.col_idx, v, A, B)
    np.testing.assert_allclose(sv * v, sddmm, rtol=1e-12, atol=1e-13)
    # SpMM A = S B (beta = 1: accumulates into a non-zero Y)
    Y0 = rng.uniform(-1, 1, (N, R))
    Y = Y0.copy()
    orc.spmm_csr(csr.rowStart, csr.col_idx, csr.values, B, Y)
    np.testing.assert_allclose(Y - Y0, spmmA, rtol=1e-11, atol=1e-12)
    # SpMM B = S^T A through the transposed block
    csrT = orc.coo_to_csr(N, N, rows, cols, sv, transpose=True)
    ST = sp.csr_matrix(S.T)
    ST.sort_indices()
    assert np.array_equal(csrT.rowStart, ST.indptr.astype(np.int64))
    assert np.array_equal(csrT.col_idx, ST.indices.astype(np.int64))
    Z = np.zeros((N, R))
    orc.spmm_csr(csrT.rowStart, csrT.col_idx, csrT.values, A, Z)
    np.testing.assert_allclose(Z, spmmB, rtol=1e-11, atol=1e-12)


def test_oracle_fused_block_is_sddmm_then_spmm():
    logM, npr, R = 9, 8, 32
    rows, cols, vals = orc.er_tuples(logM, npr, seed=7)
    N = 1 << logM
    rng = np.random.default_rng(3)
    A = rng.uniform(-1, 1, (N, R))
    B = rng.uniform(-1, 1, (N, R))
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    _, sddmm, _, _, fused = orc.global_reference(rows, cols, vals, A, B)
    v = np.zeros(csr.nnz)
    out = np.zeros((N, R))
    orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, v, A, B, out)
    np.testing.assert_allclose(v, sddmm, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out, fused, rtol=1e-11, atol=1e-12)


def test_dummy_initialize_closed_form():
    """Known-answer test: with S == 1 and X[g_row,g_col] = g_row*R + g_col the SDDMM value at
    (i,j) has a closed form that is exact in fp64 (N=2^14, R=16 -> < 2^41)."""
    logM, npr, R = 14, 8, 16
    rows, cols, vals = orc.er_tuples(logM, npr, seed=0xC0FFEE)
    N = 1 << logM
    A = orc.dummy_matrix(0, N, R)
    B = orc.dummy_matrix(0, N, R)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    v = np.zeros(csr.nnz)
    orc.sddmm_coo(csr.row_idx, csr.col_idx, v, A, B)
    expect = orc.dummy_sddmm_closed_form(csr.row_idx, csr.col_idx, R)
    assert np.array_equal(v, expect)
    # SpMM rows: A[i,k] = sum_{j in row i} (jR + k)
    Y = np.zeros((N, R))
    orc.spmm_csr(csr.rowStart, csr.col_idx, csr.values, B, Y)
    deg = np.diff(csr.rowStart).astype(np.float64)
    colsum = np.add.reduceat(np.append(csr.col_idx, 0).astype(np.float64), csr.rowStart[:-1])[:N]
    colsum[deg == 0] = 0.0
    expectY = colsum[:, None] * R + deg[:, None] * np.arange(R)[None, :]
    assert np.array_equal(Y, expectY)


def test_coo_to_csr_keeps_duplicates_and_is_stable():
    # duplicates are not merged (mkl_sparse_convert_csr does not), input order kept
    r = np.array([1, 0, 1, 1, 0], dtype=np.uint64)
    c = np.array([0, 2, 0, 3, 2], dtype=np.uint64)
    v = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    csr = orc.coo_to_csr(2, 4, r, c, v)
    assert csr.rowStart.tolist() == [0, 2, 5]
    assert csr.col_idx.tolist() == [2, 2, 0, 0, 3]
    assert csr.values.tolist() == [2.0, 5.0, 1.0, 3.0, 4.0]
    t = orc.coo_to_csr(2, 4, r, c, v, transpose=True)
    assert (t.rows, t.cols) == (4, 2)
    assert t.rowStart.tolist() == [0, 2, 2, 4, 5]
    assert t.col_idx.tolist() == [1, 1, 0, 0, 1]
    with pytest.raises(ValueError):
        orc.coo_to_csr(2, 4, np.array([2], dtype=np.uint64), np.array([0], dtype=np.uint64), np.array([1.0]))


def test_empty_block():
    csr = orc.coo_to_csr(4, 4, np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0))
    assert csr.nnz == 0 and csr.rowStart.tolist() == [0] * 5
    Y = np.ones((4, 3))
    orc.spmm_csr(csr.rowStart, csr.col_idx, csr.values, np.ones((4, 3)), Y)
    assert np.array_equal(Y, np.ones((4, 3)))


def test_er_generator_is_partition_independent_and_unique():
    logM, npr = 10, 16
    full = orc.er_tuples(logM, npr, seed=0xC0FFEE)
    parts = [orc.er_tuples(logM, npr, 0xC0FFEE, lo, lo + 256) for lo in range(0, 1024, 256)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), full[0])
    assert np.array_equal(np.concatenate([p[1] for p in parts]), full[1])
    key = full[0].astype(np.int64) * (1 << logM) + full[1].astype(np.int64)
    assert len(np.unique(key)) == len(key) and np.all(np.diff(key) > 0)
    assert 0.97 * 1024 * npr < len(key) <= 1024 * npr
    assert np.all(full[2] == 1.0)
