"""One rank of the multi-process GPU tests of jets_dist_op_* (tests/test_gpu_dist.py starts WORLD_SIZE of these).

Every rank builds the WHOLE operator on its own GPU as an ordinary JopBlock and applies it with jets_apply
(the single-GPU result), then builds its block rows, runs the distributed apply through the C ABI and
compares its shards of the result VECTORS with the single-GPU ones: bit for bit wherever the row sums have
the same order (forward always; adjoint with halo 1), to the north_star tolerance otherwise.

Ranks may share one GPU (rank % device_count): the exchange goes through CUDA IPC either way and the host
bootstrap (gloo) replaces NCCL, which refuses two ranks on one device.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    case = json.loads(os.environ["JETS_DIST_CASE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    torch.cuda.set_device(dev)
    import jets_b200 as B
    B.init(dev)
    D = B.dist
    D.init_host_bootstrap(B, dist, torch, rank, world)
    use_nccl = bool(case.get("nccl")) and ndev >= world
    if use_nccl:
        # the torch group is gloo-only here, so ship the ncclUniqueId through it by hand
        ident = C.create_string_buffer(128)
        if rank == 0:
            B.check(B.lib.jets_dist_unique_id(ident))
        t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        B.check(B.lib.jets_dist_init(rank, world, C.create_string_buffer(bytes(t.numpy().tobytes()), 128)))

    T = np.dtype(case["dtype"])
    nblk, halo, blk = case["nblk"], case["halo"], case["block_len"]
    lens = [blk + case.get("ragged", 0) * (((i + 1) // 2) % 3) for i in range(nblk)]   # neighbours in pairs of equal length
    sp = [B.JetSpace(T, n) for n in lens]
    g = np.random.default_rng(case.get("seed", 0))
    # operator state per global block, identical on every rank
    state = {}
    for r in range(nblk):
        for c in range(nblk):
            if abs(r - c) <= halo and lens[r] == lens[c]:
                state[(r, c)] = (g.random(lens[r]).astype(T), float(g.random()))
    host_w = {k: B.to_device(v[0]) for k, v in state.items()}

    def make_block(r, c):
        if (r, c) not in state:
            return B.JopZeroBlock(sp[c], sp[r])
        w, a = host_w[(r, c)], state[(r, c)][1]
        kind = (r + 2 * c) % 4 if r != c else 0
        if kind == 0:
            return B.JopDiagonal(w)
        if kind == 1:
            return B.JopStencil(T, lens[r], "fdiff")
        if kind == 2:
            return a * B.JopStencil(T, lens[r], "lap")
        return B.JopDiagonal(w) @ B.JopStencil(T, lens[r], "fdiff")

    # single-GPU result on this rank's device
    A = B.blockop([[make_block(r, c) for c in range(nblk)] for r in range(nblk)])
    part = D.RowPartition(nblk, world, rank, halo=halo)
    n0, n1 = part.r0, part.r1
    # halo-extended local operator: columns outside the global operator are zero blocks over a neighbour-sized space
    def ext_space(j):
        c = part.global_col(j)
        return sp[c] if c is not None else sp[n0 if j < halo else n1 - 1]
    rows = []
    for i in range(part.nloc):
        row = []
        for j in range(part.next_cols):
            c = part.global_col(j)
            r = n0 + i
            row.append(make_block(r, c) if c is not None and abs(r - c) <= halo else B.JopZeroBlock(ext_space(j), sp[r]))
        rows.append(row)
    A_loc = B.blockop(rows)
    op = D.DistOp(B, A_loc, halo)
    off = np.concatenate([[0], np.cumsum(lens)])
    sl = slice(off[n0], off[n1])
    own_dom = B.JetBSpace(sp[n0:n1])
    fails = []
    niter = case.get("iters", 5)
    for it in range(niter):
        x = g.random(off[-1]).astype(T) - 0.5        # same stream on every rank -> same vectors
        y = g.random(off[-1]).astype(T) - 0.5
        xd = B.to_device(x, B.domain(A))
        d_ref = (A * xd).to_host()
        m_ref = (A.T * B.to_device(y, B.range_(A))).to_host()
        x_own = B.to_device(x[sl], own_dom)
        if it % 2 == 1:          # every other iteration on a REGISTERED vector: the pull path (halo read in place)
            op.register(x_own)
        d_own = op.forward(B.zeros(own_dom), x_own).to_host()
        if it % 2 == 1 and it >= 3:    # ... and again on the same registered vector after overwriting it
            x2 = g.random(off[-1]).astype(T) - 0.5
            x_own.from_host(x2[sl])
            d2 = op.forward(B.zeros(own_dom), x_own).to_host()
            if not np.array_equal(d2, (A * B.to_device(x2, B.domain(A))).to_host()[sl]):
                fails.append(f"iter {it}: forward on a re-used registered vector differs")
        y_own = B.to_device(y[sl], own_dom)
        m_own = op.adjoint(B.zeros(own_dom), y_own).to_host()
        if not np.array_equal(d_own, d_ref[sl]):
            fails.append(f"iter {it}: forward differs (max {np.abs(d_own - d_ref[sl]).max():.3e})")
        if halo == 1:
            if not np.array_equal(m_own, m_ref[sl]):
                fails.append(f"iter {it}: adjoint differs bitwise (max {np.abs(m_own - m_ref[sl]).max():.3e})")
        else:
            tol = 1e-12 if T == np.float64 else 1e-5
            err = np.linalg.norm(m_own.astype(np.float64) - m_ref[sl]) / max(np.linalg.norm(m_ref[sl]), 1e-300)
            if not err <= tol:
                fails.append(f"iter {it}: adjoint off by {err:.3e} > {tol}")
    # host-buffer pipeline: A'(A x) on pinned shards, twice (cross-step hazards), against the single-GPU result
    x = g.random(off[-1]).astype(T) - 0.5
    n_ref = (A.T * (A * B.to_device(x, B.domain(A)))).to_host()
    h_in = torch.from_numpy(x[sl].copy()).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    for rep in range(2):
        h_out.zero_()
        op.normal_host(h_out.data_ptr(), h_in.data_ptr(), case.get("chunks", 3))
        op.join()
        B.sync()
        got = h_out.numpy()
        if halo == 1:
            ok = np.array_equal(got, n_ref[sl])
        else:
            tol = 1e-12 if T == np.float64 else 1e-5
            ok = np.linalg.norm(got.astype(np.float64) - n_ref[sl]) <= tol * np.linalg.norm(n_ref[sl])
        if not ok:
            fails.append(f"host pipeline rep {rep}: differs (max {np.abs(got - n_ref[sl]).max():.3e})")
    # a whole solver on the partitioned operator: LSQR with one jets_dist_apply per product and rank-ordered norms
    # (the iterate v is a REGISTERED vector that the solver overwrites every iteration: the pull path's exit handshake)
    if case.get("lsqr", 1):
        rhs = g.random(off[-1]).astype(T) + 0.5
        its = 12
        x_ref, h_ref = B.solvers.lsqr(A, B.to_device(rhs, B.range_(A)), its)
        x_d, h_d = B.solvers.lsqr_dist(op, B.to_device(rhs[sl], own_dom), its)
        tol = 1e-9 if T == np.float64 else 2e-3
        xr = x_ref.to_host()[sl].astype(np.float64)
        err = np.linalg.norm(x_d.to_host().astype(np.float64) - xr) / max(np.linalg.norm(xr), 1e-300)
        if not err <= tol:
            fails.append(f"distributed LSQR iterate off by {err:.3e} > {tol}")
        if not all(abs(a - ar) <= tol * abs(ar) and abs(b - br) <= tol * abs(br) for (a, b), (ar, br) in zip(h_d, h_ref)):
            fails.append("distributed LSQR alpha/beta history differs")
    if op.gate_timeouts != 0:
        fails.append(f"{op.gate_timeouts} work units timed out waiting for a neighbour")
    launches = op.info(4)
    if launches != 1:
        fails.append(f"a distributed apply took {launches} launches, expected 1")
    if use_nccl and case.get("dense"):
        fails += dense_case(B, D, T, g, rank, world)
        fails += dense_case(B, D, np.dtype(np.complex64 if T == np.float32 else np.complex128), g, rank, world)
    dist.barrier()
    op.close()
    B.check(B.lib.jets_dist_shutdown())
    dist.destroy_process_group()
    print(json.dumps({"rank": rank, "device": dev, "fails": fails}), flush=True)
    sys.exit(1 if fails else 0)


def dense_case(B, D, T, g, rank, world):
    """Row-partitioned JopBlock of dense blocks: all-gather forward, reduce-scatter adjoint (NCCL)."""
    fails = []
    nb, k = 2 * world, 96
    cplx = np.issubdtype(T, np.complexfloating)

    def draw(*shape):
        return (g.random(shape) + (1j * g.random(shape) if cplx else 0) - 0.5).astype(T)
    mats = {(r, c): draw(k, k) for r in range(nb) for c in range(nb)}
    A = B.blockop([[B.JopDense(mats[(r, c)]) for c in range(nb)] for r in range(nb)])
    nloc = nb // world
    r0 = rank * nloc
    A_loc = B.blockop([[B.JopDense(mats[(r0 + i, c)]) for c in range(nb)] for i in range(nloc)])
    op = D.DistOp(B, A_loc, dense=True)
    x = draw(nb * k)
    y = draw(nb * k)
    d_ref = (A * B.to_device(x, B.domain(A))).to_host()
    m_ref = (A.T * B.to_device(y, B.range_(A))).to_host()
    sl = slice(r0 * k, (r0 + nloc) * k)
    d = op.forward(B.zeros(B.range_(A_loc)), B.to_device(x[sl], op.own_space)).to_host()
    m = op.adjoint(B.zeros(op.own_space), B.to_device(y[sl], B.range_(A_loc))).to_host()
    tol = 1e-12 if T in (np.float64, np.complex128) else 1e-5
    for name, got, ref in (("forward", d, d_ref[sl]), ("adjoint", m, m_ref[sl])):
        err = np.linalg.norm(got.astype(np.complex128) - ref) / np.linalg.norm(ref)
        if not err <= tol:
            fails.append(f"dense {name} ({T}): off by {err:.3e} > {tol}")
    op.close()
    return fails


if __name__ == "__main__":
    main()
