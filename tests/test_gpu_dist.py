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


def run_ranks(world, case, timeout=240):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   JETS_DIST_CASE=json.dumps(case), JETS_B200_GATE_TIMEOUT_MS="20000")
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
def test_dense_dist_op_allgather_reduce_scatter():
    if _ngpu() < 2:
        pytest.skip("the NCCL dense-structure exchange needs one GPU per rank")
    case = {"dtype": "float32", "nblk": 8, "halo": 1, "block_len": 8192, "iters": 1, "chunks": 2, "nccl": 1, "dense": 1, "seed": 5}
    res = run_ranks(2, case)
    assert all(r["fails"] == [] for r in res), res
