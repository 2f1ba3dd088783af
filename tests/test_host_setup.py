"""Host-side setup helpers of the product (no GPU needed) against the oracle: the ER tuple
generator and the COO->CSR conversion must be bit-exact (integer / index work)."""
import numpy as np
import pytest

from oracle import hnh_oracle as orc


def gen(hnh, logM, npr, seed, lo, hi):
    cap = (hi - lo) * npr
    r = np.zeros(cap, np.uint64)
    c = np.zeros(cap, np.uint64)
    v = np.zeros(cap, np.float64)
    n = hnh.hnh_er_generate_host(logM, npr, seed, lo, hi, r.ctypes.data, c.ctypes.data, v.ctypes.data, cap)
    assert n >= 0
    return r[:n], c[:n], v[:n]


@pytest.mark.parametrize("logM,npr", [(6, 3), (10, 8), (12, 32), (14, 8)])
def test_er_generator_matches_oracle_bit_exact(hnh, logM, npr):
    seed = 0xC0FFEE + logM
    N = 1 << logM
    ref = orc.er_tuples(logM, npr, seed)
    got = gen(hnh, logM, npr, seed, 0, N)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    # rank-sliced generation concatenates to the same thing (layout-independent input)
    parts = [gen(hnh, logM, npr, seed, lo, lo + N // 4) for lo in range(0, N, N // 4)]
    assert np.array_equal(np.concatenate([p[1] for p in parts]), ref[1])


def test_er_generator_capacity_and_argument_errors(hnh):
    r = np.zeros(4, np.uint64)
    assert hnh.hnh_er_generate_host(8, 8, 1, 0, 256, r.ctypes.data, r.ctypes.data, r.ctypes.data, 4) == -1
    assert hnh.hnh_er_generate_host(8, 8, 1, 10, 5, None, None, None, 0) == -1
    assert hnh.hnh_er_generate_host(8, 8, 1, 5, 5, None, None, None, 0) == 0


@pytest.mark.parametrize("transpose", [False, True])
def test_coo_to_csr_host_matches_oracle(hnh, transpose):
    rng = np.random.default_rng(11)
    M, N, nnz = 37, 91, 2000
    r = rng.integers(0, M, nnz).astype(np.uint64)
    c = rng.integers(0, N, nnz).astype(np.uint64)  # duplicates likely: must be kept
    order = np.lexsort((r, c))
    r, c = r[order], c[order]
    v = rng.uniform(-1, 1, nnz)
    ref = orc.coo_to_csr(M, N, r, c, v, transpose=transpose)
    out_rows = N if transpose else M
    rs = np.zeros(out_rows + 1, np.int64)
    ci = np.zeros(nnz, np.int64)
    ri = np.zeros(nnz, np.int64)
    vv = np.zeros(nnz)
    rc = hnh.hnh_coo_to_csr_host(M, N, nnz, r.ctypes.data, c.ctypes.data, v.ctypes.data, int(transpose),
                                 rs.ctypes.data, ci.ctypes.data, ri.ctypes.data, vv.ctypes.data)
    assert rc == 0
    assert np.array_equal(rs, ref.rowStart) and np.array_equal(ci, ref.col_idx)
    assert np.array_equal(ri, ref.row_idx) and np.array_equal(vv, ref.values)
    bad = np.array([M], np.uint64)
    assert hnh.hnh_coo_to_csr_host(M, N, 1, bad.ctypes.data, bad.ctypes.data, v.ctypes.data, 0,
                                   rs.ctypes.data, ci.ctypes.data, None, vv.ctypes.data) == -1


@pytest.mark.parametrize("transpose", [False, True])
@pytest.mark.parametrize("shape", ["wide", "one_row", "sparse_rows"])
def test_coo_to_csr_host_parallel_path_matches_oracle(hnh, transpose, shape):
    """Enough entries for the multi-threaded histogram path of the stable counting sort (csrc/host_sort.h): every
    thread's chunk must land behind the previous thread's entries of the same row, duplicates in input order."""
    rng = np.random.default_rng(23)
    nnz = 600_000
    M, N = {"wide": (5000, 3000), "one_row": (1, 4096), "sparse_rows": (1 << 20, 1 << 10)}[shape]
    if transpose and shape == "one_row":
        M, N = N, M
    r = rng.integers(0, M, nnz).astype(np.uint64)
    c = rng.integers(0, N, nnz).astype(np.uint64)
    order = np.lexsort((r, c))
    r, c = np.ascontiguousarray(r[order]), np.ascontiguousarray(c[order])
    v = np.arange(nnz, dtype=np.float64)  # distinct values: the order of duplicates is observable
    ref = orc.coo_to_csr(M, N, r, c, v, transpose=transpose)
    out_rows = N if transpose else M
    rs, ci, ri, vv = np.zeros(out_rows + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
    assert hnh.hnh_coo_to_csr_host(M, N, nnz, r.ctypes.data, c.ctypes.data, v.ctypes.data, int(transpose), rs.ctypes.data,
                                   ci.ctypes.data, ri.ctypes.data, vv.ctypes.data) == 0
    assert np.array_equal(rs, ref.rowStart) and np.array_equal(ci, ref.col_idx)
    assert np.array_equal(ri, ref.row_idx) and np.array_equal(vv, ref.values)
    # one bad coordinate in the last thread's chunk is still found
    r2 = r.copy()
    r2[-3] = M + 7
    assert hnh.hnh_coo_to_csr_host(M, N, nnz, r2.ctypes.data, c.ctypes.data, v.ctypes.data, int(transpose), rs.ctypes.data,
                                   ci.ctypes.data, ri.ctypes.data, vv.ctypes.data) == -1


def test_coo_to_csr_host_random_shapes_property(hnh):
    """Property test over small random blocks (including empty ones, single rows / columns, heavy duplication):
    the host COO -> CSR equals the oracle's, rowStart is a monotone partition of [0, nnz], and transposing twice
    returns the original entries."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 40), st.integers(0, 300), st.booleans(), st.integers(0, 2 ** 31))
    def check(M, N, nnz, transpose, seed):
        rng = np.random.default_rng(seed)
        r = rng.integers(0, M, nnz).astype(np.uint64)
        c = rng.integers(0, N, nnz).astype(np.uint64)
        order = np.lexsort((r, c))
        r, c = np.ascontiguousarray(r[order]), np.ascontiguousarray(c[order])
        v = rng.uniform(-1, 1, nnz)
        out_rows = N if transpose else M
        rs = np.full(out_rows + 1, -1, np.int64)
        ci, ri, vv = np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1), np.int64), np.zeros(max(nnz, 1))
        assert hnh.hnh_coo_to_csr_host(M, N, nnz, r.ctypes.data, c.ctypes.data, v.ctypes.data, int(transpose), rs.ctypes.data,
                                       ci.ctypes.data, ri.ctypes.data, vv.ctypes.data) == 0
        assert rs[0] == 0 and rs[-1] == nnz and np.all(np.diff(rs) >= 0)
        if nnz:
            ref = orc.coo_to_csr(M, N, r, c, v, transpose=transpose)
            assert np.array_equal(rs, ref.rowStart) and np.array_equal(ci[:nnz], ref.col_idx)
            assert np.array_equal(ri[:nnz], ref.row_idx) and np.array_equal(vv[:nnz], ref.values)
            # the multiset of (row, col, value) is preserved
            a = sorted(zip((c if transpose else r).tolist(), (r if transpose else c).tolist(), v.tolist()))
            b = sorted(zip(ri[:nnz].tolist(), ci[:nnz].tolist(), vv[:nnz].tolist()))
            assert a == b

    check()
