"""Multi-process worker for the distributed parity tests (launched by tests/mp_util.py through
torch.distributed.run).  Every rank builds the requested algorithms through the product's driver
ABI, and writes what it sees -- layout, local CSR blocks and, on a GPU, the outputs of the
operation script -- to <outdir>/<case>_rank<r>.npz in the same shape oracle/ref.py returns for the
reference, so that the test process can compare rank by rank.

Transports: "gloo" (External transport over torch.distributed gloo; CPU boxes and several ranks
sharing ONE GPU), "nccl" (one GPU per rank)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sval(r, c):
    """Deterministic nonzero value as a function of the global coordinate."""
    return 0.5 + ((r.astype(np.int64) * 31 + c.astype(np.int64) * 17) % 97) / 97.0


def global_inputs(N, R, seed, M=None):
    """A (M x R) and B (N x R); M defaults to N.  The draws for B do not depend on M."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(-1, 1, (N, R))
    B = rng.uniform(-1, 1, (N, R))
    if M is not None and M != N:
        A = np.random.default_rng(seed + 1).uniform(-1, 1, (M, R))
    return A, B


def gat_inputs(N, layers, seed):
    """Deterministic GAT test inputs: (layers as tuples, X0 (N x in), weights[layer][head] (in x per_head))."""
    layers = [tuple(int(x) for x in l) for l in layers]
    rng = np.random.default_rng(seed + 77)
    X0 = rng.uniform(-1, 1, (N, layers[0][0]))
    weights = [[rng.uniform(-1, 1, (fin, fph)) for _ in range(heads)] for fin, fph, heads in layers]
    return layers, X0, weights


def als_inputs(N, R, seed):
    """Deterministic ALS test inputs (global): ground-truth factors and starting embeddings."""
    rng = np.random.default_rng(seed + 99)
    Agt, Bgt = rng.uniform(-1, 1, (N, R)) / R, rng.uniform(-1, 1, (N, R)) / R
    A0, B0 = 1.4 * rng.uniform(-1, 1, (N, R)) / R, rng.uniform(-1, 1, (N, R)) / R / 1.3
    return Agt, Bgt, A0, B0


def gather_local(G, subs, shape):
    out = np.zeros(shape)
    flat = out.reshape(-1)
    at = 0
    for top, left, nr, nc in subs:
        blk = np.zeros((nr, nc))
        hi = min(top + nr, G.shape[0])
        if hi > top:
            blk[:hi - top] = G[top:hi, left:left + nc]
        flat[at:at + nr * nc] = blk.reshape(-1)
        at += nr * nc
    return out


def main():
    cases = json.loads(sys.argv[1])
    outdir = sys.argv[2]
    transport = sys.argv[3]
    import torch
    have_gpu = torch.cuda.is_available()
    if have_gpu:
        ngpu = torch.cuda.device_count()
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % ngpu if transport == "nccl" else 0)
    from distributed_sddmm_b200 import driver as D
    rank, world = D.world_init(transport)
    for case in cases:
        name, alg_name, c, R, logM, npr, seed, script = (case[k] for k in ("name", "alg", "c", "R", "logM", "npr", "seed", "script"))
        N = case.get("n") or (1 << logM)  # "n": a size that does not divide evenly among the ranks
        M = case.get("m") or N            # "m": a rectangular M x N matrix (rows of the generator below M)
        if case.get("load", "tuples") == "er":
            # SpmatLocal::loadTuples(false, logM, npr, ""): every rank generates its row slice, values 1.0
            S = D.SpmatLocal.load_er(logM, npr, seed)
        else:
            # caller-provided tuples with coordinate-dependent values, dealt to the ranks in contiguous slices
            from oracle import hnh_oracle as orc
            deal = case.get("deal", "slices")
            if deal == "slices":      # contiguous row slices
                per = M // world
                lo, hi = per * rank, (M if rank == world - 1 else per * (rank + 1))
                tr, tc, _ = orc.er_tuples(logM, npr, seed, lo, hi)
            else:
                tr, tc, _ = orc.er_tuples(logM, npr, seed, 0, M)
                if deal == "last":    # everything starts on the last rank
                    mine = np.full(len(tr), rank == world - 1)
                else:                 # "scatter": rows dealt round-robin-ish, handed over in reverse order
                    mine = (tr.astype(np.int64) * 7 + 3) % world == rank
                tr, tc = tr[mine][::-1].copy(), tc[mine][::-1].copy()
            keep = tc < N
            tr, tc = tr[keep], tc[keep]
            S = D.SpmatLocal.from_tuples(M, N, tr, tc, sval(tr, tc))
        alg = D.Algorithm(alg_name, S, R, c)
        d = alg.dims
        out = dict(i=d.grid_i, j=d.grid_j, k=d.grid_k, localArows=d.localArows, localAcols=d.localAcols,
                   localBrows=d.localBrows, localBcols=d.localBcols, aSubmatrices=alg.submatrices("A"),
                   bSubmatrices=alg.submatrices("B"), s_values=d.s_values, st_values=d.st_values, p=d.p)
        for key in ("S", "ST"):
            blocks = alg.blocks(key)
            out[key + "_nblocks"] = len(blocks)
            for b, blk in enumerate(blocks):
                out[f"{key}_b{b}_null"] = blk is None
                if blk is not None:
                    for f in ("rows", "cols", "transpose", "rowStart", "col_idx", "row_idx", "values"):
                        out[f"{key}_b{b}_{f}"] = blk[f]
        if have_gpu and script:
            A, B = alg.like_A_matrix(), alg.like_B_matrix()
            GA, GB = global_inputs(N, R, seed, M)  # same sizes as the reference side
            shapeA, shapeB = (d.localArows, d.localAcols), (d.localBrows, d.localBcols)
            subsA, subsB = alg.submatrices("A"), alg.submatrices("B")

            def first_col(subs, shape, index):
                m = np.zeros(shape)
                flat = m.reshape(-1)
                at = 0
                for top, left, nr, nc in subs:
                    blk = np.zeros((nr, nc))
                    if left == 0:
                        blk[:, 0] = np.arange(top, top + nr) if index else 1.0
                    flat[at:at + nr * nc] = blk.reshape(-1)
                    at += nr * nc
                return m

            coords = {}
            for which, mode, like in (("S", "sddmmA", alg.like_S_values), ("ST", "sddmmB", alg.like_ST_values)):
                ones, res = like(1.0), like(0.0)
                got = []
                for p_ in range(2):
                    A.from_host(first_col(subsA, shapeA, p_ == 0))
                    B.from_host(first_col(subsB, shapeB, p_ == 1))
                    alg.initial_shift(A, B, mode)
                    (alg.sddmmA if which == "S" else alg.sddmmB)(A, B, ones, res)
                    alg.de_shift(A, B, mode)
                    got.append(np.rint(res.to_host()).astype(np.int64))
                coords[which] = got
                out[which + "_rows"], out[which + "_cols"] = got
            Sv, STv = alg.like_S_values(0.0), alg.like_ST_values(0.0)
            if case.get("load", "tuples") == "er":  # generator path: every nonzero is 1.0
                Sv.fill(1.0)
                STv.fill(1.0)
            else:
                Sv.from_host(sval(*coords["S"]))
                STv.from_host(sval(*coords["ST"]))
            res_s, res_st = alg.like_S_values(0.0), alg.like_ST_values(0.0)
            for t, op in enumerate(script):
                A.from_host(gather_local(GA, subsA, shapeA))
                B.from_host(gather_local(GB, subsB, shapeB))
                vals = np.zeros(0)
                if op == "sddmmA":
                    alg.initial_shift(A, B, "sddmmA"); alg.sddmmA(A, B, Sv, res_s); alg.de_shift(A, B, "sddmmA"); vals = res_s.to_host()
                elif op == "sddmmB":
                    alg.initial_shift(A, B, "sddmmB"); alg.sddmmB(A, B, STv, res_st); alg.de_shift(A, B, "sddmmB"); vals = res_st.to_host()
                elif op == "spmmA":
                    alg.initial_shift(A, B, "spmmA"); alg.spmmA(A, B, Sv); alg.de_shift(A, B, "spmmA")
                elif op == "spmmB":
                    alg.initial_shift(A, B, "spmmB"); alg.spmmB(A, B, STv); alg.de_shift(A, B, "spmmB")
                elif op == "fusedA":
                    res_s.fill(0.0)
                    alg.initial_shift(A, B, "sddmmA"); alg.fusedSpMM(A, B, Sv, res_s, "A"); alg.de_shift(A, B, "sddmmA"); vals = res_s.to_host()
                elif op == "fusedB":
                    res_st.fill(0.0)
                    alg.initial_shift(A, B, "sddmmB"); alg.fusedSpMM(A, B, STv, res_st, "B"); alg.de_shift(A, B, "sddmmB"); vals = res_st.to_host()
                else:
                    raise ValueError(op)
                out[f"op{t}_A"], out[f"op{t}_B"], out[f"op{t}_values"] = A.to_host(), B.to_host(), vals
            out["perf"] = json.dumps(alg.perf())
            if case.get("als"):
                out["als"] = np.array(alg.als_residuals(1))
        if have_gpu and "hostpipe" in case:
            # fusedSpMM_host (upload / kernels / download pipelined, riding operand all-gathered) against the device path
            A, B = alg.like_A_matrix(), alg.like_B_matrix()
            GA, GB = global_inputs(N, R, seed)
            shapeA, shapeB = (d.localArows, d.localAcols), (d.localBrows, d.localBcols)
            hA = gather_local(GA, alg.submatrices("A"), shapeA)
            hB = gather_local(GB, alg.submatrices("B"), shapeB)
            for mode, like in (("A", alg.like_S_values), ("B", alg.like_ST_values)):
                Sv, res = like(1.0), like(0.0)
                A.from_host(hA)
                B.from_host(hB)
                alg.fusedSpMM(A, B, Sv, res, mode)
                out[f"hostpipe_{mode}_want"] = (A if mode == "A" else B).to_host()
                got = np.full(shapeA if mode == "A" else shapeB, np.nan)
                alg.fusedSpMM_host(A, B, Sv, res, hA, hB, got, mode, chunk_rows=case["hostpipe"])
                out[f"hostpipe_{mode}_got"] = got
                out[f"hostpipe_{mode}_staged"] = (A if mode == "A" else B).to_host()
        if have_gpu and case.get("als_parity"):
            # Distributed_ALS on given inputs (hnhd_als_run), to be compared with the reference's own ALS run
            shapeA, shapeB = (d.localArows, d.localAcols), (d.localBrows, d.localBcols)
            subsA, subsB = alg.submatrices("A"), alg.submatrices("B")
            Agt, Bgt, A0, B0 = als_inputs(N, R, seed)
            res, locA, locB = D.als_run(alg, gather_local(Agt, subsA, shapeA), gather_local(Bgt, subsB, shapeB),
                                        gather_local(A0, subsA, shapeA), gather_local(B0, subsB, shapeB), 1, 10)
            out["als_res"], out["als_A"], out["als_B"] = np.array(res), locA, locB
        if have_gpu and case.get("gat"):
            # GAT forward pass on this algorithm object (dense-shift layouts: one row block per rank, full width)
            g = case["gat"]
            layers, X0, weights = gat_inputs(N, g["layers"], seed)
            net = D.GAT(alg, layers, g["alpha"])
            for i, (fin, fph, heads) in enumerate(layers):
                for h in range(heads):
                    net.set_weight(i, h, weights[i][h])
            top, _, nr, _ = alg.submatrices("B")[0]
            net.set_input(gather_local(X0, [(top, 0, nr, layers[0][0])], net.buffer_shape(0)))
            net.forward()
            out["gat_out"] = net.buffer(len(layers))
            del net
        out["info"] = json.dumps(alg.info())
        out["setup_times"] = json.dumps(D.setup_times(reset=True))
        np.savez(os.path.join(outdir, f"{name}_rank{rank}.npz"), **out)
        del alg, S
    D.world_finalize()  # also destroys the torch.distributed group (clean gloo teardown)


if __name__ == "__main__":
    main()
