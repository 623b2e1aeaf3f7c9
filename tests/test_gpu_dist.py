"""GPU tests of the distributed operators behind the C ABI (jets_dist_op_*, csrc/dist_op.cu): WORLD ranks as
separate processes, each comparing its shard of the distributed forward / adjoint VECTORS -- and of the
host-buffer pipeline -- with the single-GPU apply of the whole operator (tests/workers/dist_worker.py).
With fewer GPUs than ranks the ranks share devices (CUDA IPC works between processes on one device; the
kernels' flag waits are served by time-slicing), so the peer-memory path is exercised on a 1-GPU box too;
the NCCL-backed dense-structure case needs one GPU per rank and is skipped otherwise."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "workers", "dist_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_ranks(world, case, timeout=240, extra_env=None):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   JETS_DIST_CASE=json.dumps(case), JETS_B200_GATE_TIMEOUT_MS="20000")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=timeout)
            outs.append((p.returncode, o, e))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for rc, o, e in outs:
        assert rc == 0, f"rank failed (rc {rc})\nstdout:\n{o}\nstderr:\n{e[-3000:]}"
    return [json.loads(o.strip().splitlines()[-1]) for _, o, _ in outs]


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
@pytest.mark.parametrize("world,halo,dtype,ragged", [(2, 1, "float32", 0), (2, 2, "float64", 0), (3, 1, "float64", 1), (4, 1, "float32", 0)])
def test_banded_dist_op_vectors_match_single_gpu(world, halo, dtype, ragged):
    case = {"dtype": dtype, "nblk": 4 * world if halo == 1 else 3 * world, "halo": halo, "block_len": 40000, "ragged": 4 * ragged,
            "iters": 5, "chunks": 3, "seed": 11 * world + halo}
    res = run_ranks(world, case)
    assert all(r["fails"] == [] for r in res), res


@pytest.mark.gpu
def test_banded_dist_op_with_tail_sub_bundles():
    """The same comparison with the fine-grained tail forced on at test size (JETS_B200_TAIL_MIN_UNITS=1) and 16 local
    block rows: gated bundles, their signals' unit counts and the flush markers must survive the re-enumeration."""
    case = {"dtype": "float32", "nblk": 32, "halo": 1, "block_len": 20000, "ragged": 0, "iters": 4, "chunks": 2, "seed": 77}
    res = run_ranks(2, case, extra_env={"JETS_B200_TAIL_MIN_UNITS": "1"})
    assert all(r["fails"] == [] for r in res), res


@pytest.mark.gpu
def test_dense_dist_op_allgather_reduce_scatter():
    if _ngpu() < 2:
        pytest.skip("the NCCL dense-structure exchange needs one GPU per rank")
    case = {"dtype": "float32", "nblk": 8, "halo": 1, "block_len": 8192, "iters": 1, "chunks": 2, "nccl": 1, "dense": 1, "seed": 5}
    res = run_ranks(2, case)
    assert all(r["fails"] == [] for r in res), res


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,n", [("float32", 40_000), ("float64", 1_000_004)])
def test_loopback_block_circulant_single_process(dtype, n, monkeypatch):
    """JETS_B200_DIST_LOOPBACK=1: the rank is its own previous and next neighbour, i.e. the operator is block-
    CIRCULANT.  The whole gated path (push units, flag waits, flush markers, signals, double-buffered arena) runs
    in ONE process on one GPU and must match the explicit circulant JopBlock: bit for bit on every block row /
    column without a wrapped term; the first and last ones sum their wrapped term in halo order (first / last)
    instead of column order, so they are held to the north_star tolerance."""
    import numpy as np
    import jets_b200 as B
    B.init(0)
    T = np.dtype(dtype)
    nb = 6
    g = np.random.default_rng(3)
    W = {(r, c): B.to_device(g.random(n).astype(T)) for r in range(nb) for c in (r, (r + 1) % nb, (r - 1) % nb)}
    sp = B.JetSpace(T, n)

    def blk(r, c):
        if c == r:
            return B.JopDiagonal(W[(r, c)])
        if c == (r + 1) % nb:
            return B.JopStencil(T, n, "fdiff") @ B.JopDiagonal(W[(r, c)])
        if c == (r - 1) % nb:
            return 0.5 * B.JopStencil(T, n, "lap")
        return B.JopZeroBlock(sp, sp)
    A = B.blockop([[blk(r, c) for c in range(nb)] for r in range(nb)])
    # rank-local rows over [block nb-1 | blocks 0..nb-1 | block 0]
    cols = [nb - 1] + list(range(nb)) + [0]
    rows = []
    for r in range(nb):
        row = []
        for j, c in enumerate(cols):
            wrap_lo = j == 0 and r == 0
            wrap_hi = j == nb + 1 and r == nb - 1
            inside = 1 <= j <= nb and (c == r or (abs(c - r) == 1))
            row.append(blk(r, c) if (wrap_lo or wrap_hi or inside) else B.JopZeroBlock(sp, sp))
        rows.append(row)
    A_loc = B.blockop(rows)
    monkeypatch.setenv("JETS_B200_DIST_LOOPBACK", "1")
    op = B.dist.DistOp(B, A_loc, halo=1)
    assert op.info(2) == 3
    own = B.JetBSpace([sp] * nb)
    tol = 1e-12 if T == np.float64 else 1e-5
    for it in range(6):     # more than two epochs: both halves of the double-buffered arena, flags re-armed
        x = B.rand(own, seed=100 + it)
        if it >= 3:         # registered input: the pull path (halo terms read the "neighbour's" vector in place)
            op.register(x)
        y = B.rand(own, seed=200 + it)
        d = op.forward(B.zeros(own), x).to_host()
        m = op.adjoint(B.zeros(own), y).to_host()
        d_ref = (A * x).to_host()
        m_ref = (A.T * y).to_host()
        inner = slice(n, (nb - 1) * n)      # rows / columns without a wrapped term: same order, same bits
        assert np.linalg.norm(d.astype(np.float64) - d_ref) <= tol * np.linalg.norm(d_ref)
        assert np.linalg.norm(m.astype(np.float64) - m_ref) <= tol * np.linalg.norm(m_ref)
        assert np.array_equal(d[inner], d_ref[inner])
        assert np.array_equal(m[inner], m_ref[inner])
    assert op.gate_timeouts == 0 and op.info(4) == 1
    op.close()
