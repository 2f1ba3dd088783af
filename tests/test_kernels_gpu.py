"""Parity of the sm_100a kernels (through the C ABI) with the CPU oracle.

Tolerances: CSR indices are inputs here (bit-exact by construction); fp64 values must agree
within 1e-5 relative (BASELINE.json north_star) -- we assert the much tighter 1e-12 the
arithmetic actually achieves, and bit-exactness for SpMM where the summation order matches
the oracle's.
"""
import os

import numpy as np
import pytest
import torch

from distributed_sddmm_b200 import lib
from oracle import hnh_oracle as orc
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # asserted; the contract is 1e-5


def rel_err(got, ref):
    scale = max(np.abs(ref).max(), 1e-300)
    return np.abs(got - ref).max() / scale


def make_problem(logM, npr, R, seed=1, rect=None):
    rows, cols, vals = orc.er_tuples(logM, npr, seed=0xC0FFEE + seed)
    N = 1 << logM
    rng = np.random.default_rng(seed * 7919 + R)
    A = rng.uniform(-1, 1, (N, R))
    B = rng.uniform(-1, 1, (N, R))
    return N, rows, cols, vals, A, B, rng


R_TABLE = [4, 8, 16, 32, 64, 128, 256]
R_ODD = [1, 3, 5, 12, 100, 130]


@pytest.mark.parametrize("R", R_TABLE + R_ODD)
@pytest.mark.parametrize("transpose", [False, True])
def test_sddmm_matches_oracle(R, transpose):
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R)
    csr = orc.coo_to_csr(N, N, rows, cols, vals, transpose=transpose)
    # role swap for transposed blocks (sparse_kernels.cpp:29-38)
    X, Y = (B, A) if transpose else (A, B)
    v0 = rng.uniform(-1, 1, csr.nnz)  # += semantics: start from non-zero values
    ref = orc.sddmm_coo(csr.row_idx, csr.col_idx, v0.copy(), X, Y)
    got = gu.run_sddmm(csr, X, Y, v0)
    assert rel_err(got, ref) < RTOL
    got_coo = gu.run_sddmm(csr, X, Y, v0, coo=True)
    assert rel_err(got_coo, ref) < RTOL
    got_gen = gu.run_sddmm(csr, X, Y, v0, flags=1)
    assert rel_err(got_gen, ref) < RTOL


@pytest.mark.parametrize("R", R_TABLE + R_ODD)
def test_spmm_matches_oracle(R):
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R, seed=2)
    sv = rng.uniform(0.5, 1.5, len(rows))
    csr = orc.coo_to_csr(N, N, rows, cols, sv)
    Y0 = rng.uniform(-1, 1, (N, R))  # beta = 1: accumulate into non-zero output
    ref = orc.spmm_csr(csr.rowStart, csr.col_idx, csr.values, B, Y0.copy())
    got = gu.run_spmm(csr, csr.values, B, Y0)
    assert rel_err(got, ref) < RTOL
    got_gen = gu.run_spmm(csr, csr.values, B, Y0, flags=1)
    assert rel_err(got_gen, ref) < RTOL
    # transposed block: B += S^T A (sparse_kernels.cpp:108-121)
    csrT = orc.coo_to_csr(N, N, rows, cols, sv, transpose=True)
    refT = orc.spmm_csr(csrT.rowStart, csrT.col_idx, csrT.values, A, Y0.copy())
    gotT = gu.run_spmm(csrT, csrT.values, A, Y0)
    assert rel_err(gotT, refT) < RTOL


@pytest.mark.parametrize("R", R_TABLE + [5, 100])
def test_fused_matches_oracle(R):
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R, seed=3)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    v0 = rng.uniform(-1, 1, csr.nnz)
    O0 = rng.uniform(-1, 1, (N, R))
    vref, oref = orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, v0.copy(), A, B, O0.copy())
    v, o = gu.run_fused(csr, v0, A, B, O0)
    assert rel_err(v, vref) < RTOL and rel_err(o, oref) < RTOL
    v, o = gu.run_fused(csr, v0, A, B, O0, flags=1)
    assert rel_err(v, vref) < RTOL and rel_err(o, oref) < RTOL


def test_fused_equals_sddmm_then_spmm_on_device():
    R = 128
    N, rows, cols, vals, A, B, rng = make_problem(11, 16, R, seed=4)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    v = gu.run_sddmm(csr, A, B)
    out = gu.run_spmm(csr, v, B, np.zeros((N, R)))
    vf, of = gu.run_fused(csr, np.zeros(csr.nnz), A, B, np.zeros((N, R)))
    assert rel_err(vf, v) < RTOL and rel_err(of, out) < RTOL


def test_dummy_initialize_known_answer_on_device():
    """Closed-form KAT: exact in fp64, so the GPU must be bit-exact."""
    logM, npr, R = 14, 8, 16
    rows, cols, vals = orc.er_tuples(logM, npr, seed=0xC0FFEE)
    N = 1 << logM
    A = orc.dummy_matrix(0, N, R)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    got = gu.run_sddmm(csr, A, A)
    assert np.array_equal(got, orc.dummy_sddmm_closed_form(csr.row_idx, csr.col_idx, R))


@pytest.mark.parametrize("R", [4, 32, 128])
def test_ragged_rows_and_empty_rows(R):
    """A power-law-ish block: one very long row, many empty rows, duplicates kept."""
    rng = np.random.default_rng(5)
    M, N = 300, 700
    r = np.concatenate([np.full(5000, 7), rng.integers(100, 120, 400), np.array([299, 299, 299])])
    c = np.concatenate([rng.integers(0, N, 5000), rng.integers(0, N, 400), np.array([5, 5, 699])])
    order = np.lexsort((r, c))  # column-major order like SpmatLocal.hpp:458
    r, c = r[order].astype(np.uint64), c[order].astype(np.uint64)
    v = rng.uniform(0.5, 1.5, len(r))
    A = rng.uniform(-1, 1, (M, R))
    B = rng.uniform(-1, 1, (N, R))
    for tr in (False, True):
        csr = orc.coo_to_csr(M, N, r, c, v, transpose=tr)
        X, Y = (B, A) if tr else (A, B)
        ref = orc.sddmm_coo(csr.row_idx, csr.col_idx, np.zeros(csr.nnz), X, Y)
        assert rel_err(gu.run_sddmm(csr, X, Y), ref) < RTOL
        O0 = rng.uniform(-1, 1, X.shape)
        oref = orc.spmm_csr(csr.rowStart, csr.col_idx, csr.values, Y, O0.copy())
        assert rel_err(gu.run_spmm(csr, csr.values, Y, O0), oref) < RTOL
        vref, fref = orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, np.zeros(csr.nnz), X, Y, O0.copy())
        vf, of = gu.run_fused(csr, np.zeros(csr.nnz), X, Y, O0)
        assert rel_err(vf, vref) < RTOL and rel_err(of, fref) < RTOL


def test_empty_block_is_noop(hnh):
    csr = orc.coo_to_csr(8, 8, np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0))
    Y0 = np.ones((8, 16))
    assert np.array_equal(gu.run_spmm(csr, np.zeros(0), np.ones((8, 16)), Y0), Y0)
    assert gu.run_sddmm(csr, Y0, Y0).shape == (0,)


def test_empty_block_with_overwrite_flags_zeroes_the_output(hnh):
    """BETA0 == 'zero the output first' (hnh_b200.h): an empty block must leave Y / Out = 0, not stale."""
    csr = orc.coo_to_csr(8, 8, np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0))
    ones = np.ones((8, 16))
    assert np.array_equal(gu.run_spmm(csr, np.zeros(0), ones, ones.copy(), flags=4), np.zeros((8, 16)))
    _, out = gu.run_fused(csr, np.zeros(0), ones, ones, ones.copy(), flags=32)  # BETA0_OUT
    assert np.array_equal(out, np.zeros((8, 16)))
    _, out = gu.run_fused(csr, np.zeros(0), ones, ones, ones.copy(), flags=16)  # BETA0_VALUES only: Out kept
    assert np.array_equal(out, ones)


def test_unaligned_operands_fall_back_to_scalar_kernel(hnh):
    """Row pointers that are only 8-byte aligned (odd element offset) must still be correct."""
    R = 32
    N, rows, cols, vals, A, B, rng = make_problem(9, 8, R, seed=6)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    ref = orc.sddmm_coo(csr.row_idx, csr.col_idx, np.zeros(csr.nnz), A, B)
    padA = torch.zeros(N * R + 1, dtype=torch.float64, device="cuda")
    padB = torch.zeros(N * R + 1, dtype=torch.float64, device="cuda")
    padA[1:] = gu.dev(A).flatten()
    padB[1:] = gu.dev(B).flatten()
    rs, ci = gu.dev(csr.rowStart), gu.dev(csr.col_idx)
    v = torch.zeros(csr.nnz, dtype=torch.float64, device="cuda")
    rc = hnh.hnh_sddmm_f64(rs.data_ptr(), ci.data_ptr(), v.data_ptr(), N, csr.nnz,
                           padA.data_ptr() + 8, padB.data_ptr() + 8, R, 0, gu.stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(v.cpu().numpy(), ref) < RTOL


def test_row_range_call(hnh):
    """Kernels accept a row sub-range (rowStart + row0, Y + row0*r): used for chunked pipelining."""
    R = 64
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R, seed=8)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    ref = orc.spmm_csr(csr.rowStart, csr.col_idx, csr.values, B, np.zeros((N, R)))
    rs, ci, dv, dB = gu.dev(csr.rowStart), gu.dev(csr.col_idx), gu.dev(csr.values), gu.dev(B)
    out = torch.zeros((N, R), dtype=torch.float64, device="cuda")
    half = N // 2
    for row0, nrows in ((0, half), (half, N - half)):
        nnz = int(csr.rowStart[row0 + nrows] - csr.rowStart[row0])
        rc = hnh.hnh_spmm_f64(rs.data_ptr() + 8 * row0, ci.data_ptr(), dv.data_ptr(), nrows, nnz,
                              dB.data_ptr(), out.data_ptr() + 8 * R * row0, R, 0, gu.stream())
        assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), ref) < RTOL


def test_plumbing_and_row_algebra(hnh):
    rng = np.random.default_rng(9)
    n, rows, R = 10007, 513, 24
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    da, db = gu.dev(a), gu.dev(b)
    dd = torch.empty(n, dtype=torch.float64, device="cuda")
    st = gu.stream()
    assert hnh.hnh_hadamard_f64(dd.data_ptr(), da.data_ptr(), db.data_ptr(), n, st) == 0
    assert np.array_equal(dd.cpu().numpy(), a * b)
    assert hnh.hnh_fill_f64(dd.data_ptr(), n, 0.001, st) == 0
    assert np.all(dd.cpu().numpy() == 0.001)
    assert hnh.hnh_fill_f64(dd.data_ptr(), n, 0.0, st) == 0
    assert np.all(dd.cpu().numpy() == 0.0)
    A, B = rng.uniform(-1, 1, (rows, R)), rng.uniform(-1, 1, (rows, R))
    dA, dB = gu.dev(A), gu.dev(B)
    out = torch.empty(rows, dtype=torch.float64, device="cuda")
    assert hnh.hnh_batch_dot_f64(out.data_ptr(), dA.data_ptr(), dB.data_ptr(), rows, R, st) == 0
    np.testing.assert_allclose(out.cpu().numpy(), (A * B).sum(1), rtol=1e-13, atol=1e-15)
    s = rng.uniform(-1, 1, rows)
    ds = gu.dev(s)
    D = torch.empty_like(dA)
    assert hnh.hnh_row_axpy_f64(D.data_ptr(), dA.data_ptr(), -1.0, ds.data_ptr(), dB.data_ptr(), rows, R, st) == 0
    np.testing.assert_allclose(D.cpu().numpy(), A - s[:, None] * B, rtol=1e-14, atol=1e-16)
    q = torch.empty(rows, dtype=torch.float64, device="cuda")
    assert hnh.hnh_vec_quotient_f64(q.data_ptr(), ds.data_ptr(), 1e-8, out.data_ptr(), 1e-8, rows, st) == 0
    np.testing.assert_allclose(q.cpu().numpy(), (s + 1e-8) / ((A * B).sum(1) + 1e-8), rtol=1e-12)
    nrm = torch.empty(1, dtype=torch.float64, device="cuda")
    assert hnh.hnh_squared_norm_f64(nrm.data_ptr(), da.data_ptr(), n, st) == 0
    np.testing.assert_allclose(nrm.item(), (a * a).sum(), rtol=1e-12)
    # expand_row_idx
    csr = orc.coo_to_csr(64, 64, *orc.er_tuples(6, 4, 1))
    ri = torch.empty(csr.nnz, dtype=torch.int64, device="cuda")
    assert hnh.hnh_expand_row_idx(gu.dev(csr.rowStart).data_ptr(), 64, csr.nnz, ri.data_ptr(), st) == 0
    assert np.array_equal(ri.cpu().numpy(), csr.row_idx)


def test_host_buffer_block_api(hnh):
    """hnh_block_run_host: the e2e entry point (host buffers, copies inside the call)."""
    import ctypes as C
    R = 128
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R, seed=10)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    blk = C.c_void_p()
    assert hnh.hnh_block_create_host(csr.rowStart.ctypes.data, csr.col_idx.ctypes.data, N, N, csr.nnz, R, C.byref(blk)) == 0
    v = np.zeros(csr.nnz)
    out = np.zeros((N, R))
    vref, oref = orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, v.copy(), A, B, out.copy())
    assert hnh.hnh_block_run_host(blk, 2, A.ctypes.data, B.ctypes.data, v.ctypes.data, out.ctypes.data, R, 0, None) == 0
    assert rel_err(v, vref) < RTOL and rel_err(out, oref) < RTOL
    v2 = np.zeros(csr.nnz)
    assert hnh.hnh_block_run_host(blk, 0, A.ctypes.data, B.ctypes.data, v2.ctypes.data, None, R, 0, None) == 0
    assert rel_err(v2, vref) < RTOL
    out2 = np.zeros((N, R))
    assert hnh.hnh_block_run_host(blk, 1, A.ctypes.data, B.ctypes.data, v2.ctypes.data, out2.ctypes.data, R, 0, None) == 0
    assert rel_err(out2, oref) < RTOL
    hnh.hnh_block_destroy(blk)


@pytest.mark.parametrize("R", [4, 16, 128, 256, 12])
def test_beta0_overwrite_variants(hnh, R):
    """HNH_FLAG_BETA0 == zero the output first, without the extra pass; fused may run in place."""
    BETA0 = 4
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R, seed=12)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    junk_v = rng.uniform(-1, 1, csr.nnz)
    junk_o = rng.uniform(-1, 1, (N, R))
    vref, oref = orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, np.zeros(csr.nnz), A, B, np.zeros((N, R)))
    assert rel_err(gu.run_sddmm(csr, A, B, junk_v, flags=BETA0), vref) < RTOL
    assert rel_err(gu.run_spmm(csr, vref, B, junk_o, flags=BETA0), oref) < RTOL
    v, o = gu.run_fused(csr, junk_v, A, B, junk_o, flags=BETA0)
    assert rel_err(v, vref) < RTOL and rel_err(o, oref) < RTOL
    if R != 12:  # in place: Out aliases X
        rs, ci, dA, dB = gu.dev(csr.rowStart), gu.dev(csr.col_idx), gu.dev(A), gu.dev(B)
        dv = gu.dev(junk_v)
        rc = hnh.hnh_fused_f64(rs.data_ptr(), ci.data_ptr(), dv.data_ptr(), N, csr.nnz, dA.data_ptr(),
                               dB.data_ptr(), dA.data_ptr(), R, BETA0, gu.stream())
        assert rc == 0
        torch.cuda.synchronize()
        assert rel_err(dv.cpu().numpy(), vref) < RTOL and rel_err(dA.cpu().numpy(), oref) < RTOL
    else:
        dA = gu.dev(A)
        assert hnh.hnh_fused_f64(dA.data_ptr(), dA.data_ptr(), dA.data_ptr(), N, csr.nnz, dA.data_ptr(),
                                 dA.data_ptr(), dA.data_ptr(), R, BETA0, gu.stream()) == -1


@pytest.mark.parametrize("R", [4, 32, 128, 256, 12])
def test_scaled_epilogues_match_oracle(hnh, R):
    """hnh_sddmm_scaled_f64 / hnh_fused_scaled_f64: the Hadamard product with the caller's S values folded into the
    kernel (reference: a separate `SValues.cwiseProduct(getCSRValues())` pass, 15D_dense_shift.hpp:364-368).  Must equal
    kernel-then-multiply exactly (one rounding of the product either way); R = 12 takes the generic fallback."""
    N, rows, cols, vals, A, B, rng = make_problem(10, 8, R, seed=13)
    csr = orc.coo_to_csr(N, N, rows, cols, vals)
    scale = rng.uniform(0.5, 1.5, csr.nnz)
    rs, ci = gu.dev(csr.rowStart), gu.dev(csr.col_idx)
    dA, dB, dS = gu.dev(A), gu.dev(B), gu.dev(scale)
    st = gu.stream()
    dots = gu.run_sddmm(csr, A, B, flags=4)
    BETA0, SCALE_VALUES, BETA0_VALUES, BETA0_OUT = 4, 256, 16, 32

    def sddmm(flags, with_out, v0=None):
        v = gu.dev(np.zeros(csr.nnz) if v0 is None else v0)
        out = torch.full((csr.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        rc = hnh.hnh_sddmm_scaled_f64(rs.data_ptr(), ci.data_ptr(), v.data_ptr(), N, csr.nnz, dA.data_ptr(), dB.data_ptr(), R,
                                      flags, dS.data_ptr(), out.data_ptr() if with_out else None, st)
        assert rc == 0, hnh.hnh_last_error_string()
        torch.cuda.synchronize()
        return v.cpu().numpy(), out.cpu().numpy()

    v, out = sddmm(BETA0, True)              # values = dot, scaled_out = scale * dot
    assert np.array_equal(v, dots) and np.array_equal(out, scale * dots)
    v, out = sddmm(BETA0 | SCALE_VALUES, True)   # both receive the product
    assert np.array_equal(v, scale * dots) and np.array_equal(out, scale * dots)
    v, _ = sddmm(BETA0, False)               # no second output: the product replaces the values
    assert np.array_equal(v, scale * dots)
    v0 = rng.uniform(-1, 1, csr.nnz)         # accumulate form: old value added before scaling
    acc = gu.run_sddmm(csr, A, B, v0)
    v, out = sddmm(0, True, v0)
    assert np.array_equal(v, acc) and np.array_equal(out, scale * acc)

    if R == 12:  # no scaled fused kernel outside the dispatch table: the call must say so
        vz = gu.dev(np.zeros(csr.nnz)); o = gu.dev(np.zeros((N, R)))
        assert hnh.hnh_fused_scaled_f64(rs.data_ptr(), ci.data_ptr(), vz.data_ptr(), N, csr.nnz, dA.data_ptr(), dB.data_ptr(),
                                        o.data_ptr(), R, BETA0_VALUES, dS.data_ptr(), None, st) == -1
        return
    # fused: values = scale * dot, Out (+)= sum values * Y, scaled_out = values -- against SDDMM, multiply, SpMM
    want_vals = scale * dots
    O0 = rng.uniform(-1, 1, (N, R))
    want_out = gu.run_spmm(csr, want_vals, B, O0)
    for bo in (0, BETA0_OUT):
        vz, o = gu.dev(np.zeros(csr.nnz)), gu.dev(O0)
        so = torch.full((csr.nnz,), float("nan"), dtype=torch.float64, device="cuda")
        rc = hnh.hnh_fused_scaled_f64(rs.data_ptr(), ci.data_ptr(), vz.data_ptr(), N, csr.nnz, dA.data_ptr(), dB.data_ptr(),
                                      o.data_ptr(), R, BETA0_VALUES | bo, dS.data_ptr(), so.data_ptr(), st)
        assert rc == 0, hnh.hnh_last_error_string()
        torch.cuda.synchronize()
        assert np.array_equal(vz.cpu().numpy(), want_vals) and np.array_equal(so.cpu().numpy(), want_vals)
        ref_out = want_out if bo == 0 else gu.run_spmm(csr, want_vals, B, np.zeros((N, R)))
        assert rel_err(o.cpu().numpy(), ref_out) < RTOL
    # needs the first-visit form
    vz, o = gu.dev(np.zeros(csr.nnz)), gu.dev(O0)
    assert hnh.hnh_fused_scaled_f64(rs.data_ptr(), ci.data_ptr(), vz.data_ptr(), N, csr.nnz, dA.data_ptr(), dB.data_ptr(),
                                    o.data_ptr(), R, 0, dS.data_ptr(), None, st) == -1


@pytest.mark.parametrize("R", [128, 256])
def test_tma_staged_variants_match_oracle(hnh, R):
    """HNH_FLAG_TMA_STAGE: the X tile arrives by cp.async.bulk + mbarrier; results must be identical in
    value to the direct-load kernels (same arithmetic order).  Row count not a multiple of the tile."""
    TMA, BETA0 = 64, 4
    rng = np.random.default_rng(21)
    M, N = 1003, 777
    r = rng.integers(0, M, 9000).astype(np.uint64)
    c = rng.integers(0, N, 9000).astype(np.uint64)
    order = np.lexsort((r, c))
    r, c = r[order], c[order]
    csr = orc.coo_to_csr(M, N, r, c, np.ones(len(r)))
    A, B = rng.uniform(-1, 1, (M, R)), rng.uniform(-1, 1, (N, R))
    v0, O0 = rng.uniform(-1, 1, csr.nnz), rng.uniform(-1, 1, (M, R))
    ref = orc.sddmm_coo(csr.row_idx, csr.col_idx, v0.copy(), A, B)
    direct = gu.run_sddmm(csr, A, B, v0, flags=2)  # HNH_FLAG_FORCE_DIRECT
    tma = gu.run_sddmm(csr, A, B, v0, flags=TMA)
    assert rel_err(tma, ref) < RTOL and np.array_equal(tma, direct)
    assert rel_err(gu.run_sddmm(csr, A, B, v0, flags=TMA | BETA0), orc.sddmm_coo(csr.row_idx, csr.col_idx, np.zeros(csr.nnz), A, B)) < RTOL
    vref, oref = orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, v0.copy(), A, B, O0.copy())
    v, o = gu.run_fused(csr, v0, A, B, O0, flags=TMA)
    vd, od = gu.run_fused(csr, v0, A, B, O0, flags=2)
    assert rel_err(v, vref) < RTOL and rel_err(o, oref) < RTOL and np.array_equal(v, vd) and np.array_equal(o, od)
    vref0, oref0 = orc.fused_block(csr.rowStart, csr.row_idx, csr.col_idx, np.zeros(csr.nnz), A, B, np.zeros((M, R)))
    v, o = gu.run_fused(csr, v0, A, B, O0, flags=TMA | BETA0)
    assert rel_err(v, vref0) < RTOL and rel_err(o, oref0) < RTOL


@pytest.mark.parametrize("R", [128, 256])
def test_tma_per_warp_variant_matches_direct(hnh, R):
    """HNH_FLAG_TMA_WARP (experimental, never selected automatically): per-warp bulk-copy slots."""
    WARP, DIRECT, BETA0 = 128, 2, 4
    rng = np.random.default_rng(22)
    M, N = 2051, 901
    r = rng.integers(0, M, 30000).astype(np.uint64)
    c = rng.integers(0, N, 30000).astype(np.uint64)
    order = np.lexsort((r, c))
    csr = orc.coo_to_csr(M, N, r[order], c[order], np.ones(len(r)))
    A, B = rng.uniform(-1, 1, (M, R)), rng.uniform(-1, 1, (N, R))
    v0, O0 = rng.uniform(-1, 1, csr.nnz), rng.uniform(-1, 1, (M, R))
    for extra in (0, BETA0):
        assert np.array_equal(gu.run_sddmm(csr, A, B, v0, flags=WARP | extra), gu.run_sddmm(csr, A, B, v0, flags=DIRECT | extra))
        v, o = gu.run_fused(csr, v0, A, B, O0, flags=WARP | extra)
        vd, od = gu.run_fused(csr, v0, A, B, O0, flags=DIRECT | extra)
        assert np.array_equal(v, vd) and np.array_equal(o, od)


@pytest.mark.parametrize("logM,npr,lo,hi", [(10, 8, 0, 1024), (12, 32, 100, 3000), (6, 64, 0, 64), (14, 1, 5, 6)])
def test_device_er_generator_is_bit_identical_with_host(logM, npr, lo, hi):
    """hnh_er_generate_device against hnh_er_generate_host (rows, columns, values, count); npr = 64 on 64 columns
    forces many duplicate draws."""
    import torch
    L = lib()
    cap = (hi - lo) * npr
    hr, hc, hv = np.empty(cap, np.uint64), np.empty(cap, np.uint64), np.empty(cap, np.float64)
    n = L.hnh_er_generate_host(logM, npr, 0xC0FFEE + 2, lo, hi, hr.ctypes.data, hc.ctypes.data, hv.ctypes.data, cap)
    assert n > 0
    dev = torch.device("cuda:0")
    dr = torch.zeros(cap, dtype=torch.int64, device=dev)
    dc = torch.zeros(cap, dtype=torch.int64, device=dev)
    dv = torch.zeros(cap, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    m = L.hnh_er_generate_device(logM, npr, 0xC0FFEE + 2, lo, hi, dr.data_ptr(), dc.data_ptr(), dv.data_ptr(), cap, st)
    assert m == n
    assert np.array_equal(dr.cpu().numpy()[:n].view(np.uint64), hr[:n])
    assert np.array_equal(dc.cpu().numpy()[:n].view(np.uint64), hc[:n])
    assert np.array_equal(dv.cpu().numpy()[:n], hv[:n])
    assert L.hnh_er_generate_device(logM, npr, 1, lo, hi, dr.data_ptr(), dc.data_ptr(), dv.data_ptr(), 1, st) == -1 or n <= 1


@pytest.mark.parametrize("transpose", [0, 1])
@pytest.mark.parametrize("rows,cols,logM,npr", [(1024, 1024, 10, 8), (300, 4096, 12, 16), (1, 64, 6, 64), (512, 512, 9, 1)])
def test_device_coo_to_csr_is_bit_identical_with_host(rows, cols, logM, npr, transpose):
    """hnh_coo_to_csr_device against hnh_coo_to_csr_host on tuples in (col, row) order -- the order redistributed
    tuples arrive in (SpmatLocal.hpp:458) -- with and without transposition; duplicates keep their input order."""
    import torch
    L = lib()
    from oracle import hnh_oracle as orc
    r, c, _ = orc.er_tuples(logM, npr, 0xC0FFEE + 4, 0, rows)
    order = np.lexsort((r, c))
    r, c = np.ascontiguousarray(r[order]), np.ascontiguousarray(c[order])
    # duplicates with distinct values: stability is observable
    r, c = np.concatenate([r, r[:7]]), np.concatenate([c, c[:7]])
    v = np.arange(len(r), dtype=np.float64) + 0.5
    nnz = len(r)
    out_rows = cols if transpose else rows
    h_rs, h_ci, h_ri, h_v = np.empty(out_rows + 1, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.int64), np.empty(nnz)
    assert L.hnh_coo_to_csr_host(rows, cols, nnz, r.ctypes.data, c.ctypes.data, v.ctypes.data, transpose, h_rs.ctypes.data,
                                 h_ci.ctypes.data, h_ri.ctypes.data, h_v.ctypes.data) == 0
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)  # noqa: E731
    dr, dc, dv = T(r), T(c), T(v)
    d_rs = torch.full((out_rows + 1,), -1, dtype=torch.int64, device=dev)
    d_ci, d_ri = torch.full((nnz,), -1, dtype=torch.int64, device=dev), torch.full((nnz,), -1, dtype=torch.int64, device=dev)
    d_v = torch.zeros(nnz, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert L.hnh_coo_to_csr_device(rows, cols, nnz, dr.data_ptr(), dc.data_ptr(), dv.data_ptr(), transpose, d_rs.data_ptr(),
                                   d_ci.data_ptr(), d_ri.data_ptr(), d_v.data_ptr(), st) == 0
    assert np.array_equal(d_rs.cpu().numpy(), h_rs)
    assert np.array_equal(d_ci.cpu().numpy(), h_ci)
    assert np.array_equal(d_ri.cpu().numpy(), h_ri)
    assert np.array_equal(d_v.cpu().numpy(), h_v)
    # a coordinate outside the block is refused
    bad = T(np.array([rows], np.uint64)) if not transpose else T(np.array([cols + 5], np.uint64))
    one = T(np.array([0], np.uint64))
    args = (bad, one) if not transpose else (one, bad)
    assert L.hnh_coo_to_csr_device(rows, cols, 1, args[0].data_ptr(), args[1].data_ptr(), dv.data_ptr(), transpose,
                                   d_rs.data_ptr(), d_ci.data_ptr(), d_ri.data_ptr(), d_v.data_ptr(), st) == -1
