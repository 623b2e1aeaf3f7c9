"""world_size 2 / 4 CPU tests of the block-row partition and of the distributed banded protocol over
torch.distributed gloo.  The partition bookkeeping (jets_b200.dist.RowPartition / build_local_operator) is the
product's own host code; the protocol itself runs inside libjets_b200 on GPUs, so here its CPU model
(tests/dist_model.py: same data movement, same order of additions, numpy-oracle arithmetic) is driven over gloo
and compared with the single-process oracle: bit for bit for halo 1 -- the property the device path claims and
tests/test_gpu_dist.py verifies on GPUs -- and to rounding for halo 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class GlooComm:
    """exchange(to_prev, to_next) -> (from_prev, from_next): lists of numpy blocks via gloo send/recv."""

    def __init__(self, part, shapes_like):
        self.part = part
        self.like = shapes_like      # (h blocks expected from prev, h blocks expected from next)

    def exchange(self, to_prev, to_next):
        p = self.part
        reqs, got_prev, got_next = [], None, None
        if p.has_prev:
            reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(a)), p.rank - 1) for a in to_prev]
            got_prev = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype) for a in self.like[0]]
            reqs += [dist.irecv(t, p.rank - 1) for t in got_prev]
        if p.has_next:
            reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(a)), p.rank + 1) for a in to_next]
            got_next = [torch.empty(a.shape, dtype=torch.from_numpy(a).dtype) for a in self.like[1]]
            reqs += [dist.irecv(t, p.rank + 1) for t in got_next]
        for r in reqs:
            r.wait()
        return ([t.numpy() for t in got_prev] if got_prev is not None else None,
                [t.numpy() for t in got_next] if got_next is not None else None)


def _global_problem(nblk, n, T, halo, seed=0):
    g = np.random.default_rng(seed)
    W = {(r, c): g.random(n).astype(T) for r in range(nblk) for c in range(nblk) if abs(r - c) <= halo}
    m = g.random(nblk * n).astype(T)
    d = g.random(nblk * n).astype(T)
    return W, m, d


def _make_block(J, T, n, W):
    def mk(r, c):
        if r == c:
            return J.JopDiagonal(W[(r, c)])
        if abs(r - c) == 2:
            return J.JopDiagonal(W[(r, c)])
        return J.JopStencil(T, n, "fdiff") if c == r + 1 else J.JopStencil(T, n, "lap")
    return mk


def _worker(rank, world, port, nblk, n, halo, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import jets_oracle as J
    import jets_b200.dist as D
    import dist_model as M
    T = np.float64
    W, m, d = _global_problem(nblk, n, T, halo)
    part = D.RowPartition(nblk, world, rank, halo=halo)
    sp = J.JetSpace(T, n)
    A = D.build_local_operator(J, part, _make_block(J, T, n, W), lambda: J.JopZeroBlock(sp, sp))
    assert J.nblocks(A) == (part.nloc, part.nloc + 2 * halo)
    like = ([np.empty(n, dtype=T)] * halo, [np.empty(n, dtype=T)] * halo)
    comm = GlooComm(part, like)
    x_own = [m[(part.r0 + k) * n:(part.r0 + k + 1) * n].copy() for k in range(part.nloc)]
    d_own = [d[(part.r0 + k) * n:(part.r0 + k + 1) * n].copy() for k in range(part.nloc)]
    f = M.forward(J, part, comm, A, x_own)
    t = M.adjoint(J, part, comm, A, d_own)
    np.save(os.path.join(out_dir, f"f{rank}.npy"), np.concatenate(f))
    np.save(os.path.join(out_dir, f"t{rank}.npy"), np.concatenate(t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, 1), (4, 1), (2, 2)])
def test_block_row_partition_matches_single_process(world, halo, tmp_path):
    from oracle import jets_oracle as J
    nblk, n, T = 8, 257, np.float64
    mp.spawn(_worker, args=(world, _free_port(), nblk, n, halo, str(tmp_path)), nprocs=world, join=True)
    W, m, d = _global_problem(nblk, n, T, halo)
    sp = J.JetSpace(T, n)
    mk = _make_block(J, T, n, W)
    A = J.blockop([[mk(r, c) if abs(r - c) <= halo else J.JopZeroBlock(sp, sp) for c in range(nblk)] for r in range(nblk)])
    f_ref = J.to_array(A * J.reshape(m.copy(), J.domain(A)))
    t_ref = J.to_array(A.T * J.reshape(d.copy(), J.range_(A)))
    f = np.concatenate([np.load(tmp_path / f"f{r}.npy") for r in range(world)])
    t = np.concatenate([np.load(tmp_path / f"t{r}.npy") for r in range(world)])
    assert np.array_equal(f, f_ref)
    if halo == 1:
        # one remote term per side, entering first (previous rank) / last (next rank): the single-process order
        assert np.array_equal(t, t_ref)
    else:
        # two remote terms from the next rank arrive as ONE partial sum: same value up to one rounding
        assert np.allclose(t, t_ref, rtol=1e-15, atol=1e-15)


class GlooCollectives:
    """allgather / reduce_scatter of numpy vectors over gloo (rank-ordered concatenation; sum over the ranks)."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def allgather(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(a.real.dtype) if np.iscomplexobj(a) else np.ascontiguousarray(a))
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t)
        full = torch.cat(outs).numpy()
        return full.view(a.dtype) if np.iscomplexobj(a) else full

    def reduce_scatter(self, flat):
        t = torch.from_numpy(np.ascontiguousarray(flat)).clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)          # gloo has no reduce_scatter: reduce, then keep the shard
        n = t.numel() // self.world
        return t[self.rank * n:(self.rank + 1) * n].numpy().copy()


def _dense_problem(T, nb, k, seed=3):
    g = np.random.default_rng(seed)
    cplx = np.issubdtype(np.dtype(T), np.complexfloating)

    def draw(*shape):
        return (g.random(shape) + (1j * g.random(shape) if cplx else 0) - 0.5).astype(T)
    mats = {(r, c): draw(k, k) for r in range(nb) for c in range(nb)}
    return mats, draw(nb * k), draw(nb * k)


def _dense_worker(rank, world, port, tname, nb, k, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import jets_oracle as J
    import dist_model as M
    T = np.dtype(tname)
    mats, x, y = _dense_problem(T, nb, k)
    nloc = nb // world
    r0 = rank * nloc
    A_loc = J.blockop([[J.JopDense(mats[(r0 + i, c)]) for c in range(nb)] for i in range(nloc)])
    comm = GlooCollectives(rank, world)
    sl = slice(r0 * k, (r0 + nloc) * k)
    f = M.dense_forward(J, comm, A_loc, x[sl].copy())
    t = M.dense_adjoint(J, comm, A_loc, y[sl].copy(), world, rank)
    np.save(os.path.join(out_dir, f"f{rank}.npy"), f)
    np.save(os.path.join(out_dir, f"t{rank}.npy"), t)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tname", ["float64", "complex128"])
def test_dense_structure_allgather_reduce_scatter_model(tname, tmp_path):
    """Row-partitioned JopBlock of dense blocks (north_star's all-gather forward / reduce-scatter adjoint) at world size 2:
    the forward shards are bit-identical to the single-process rows (same matrix products on the same gathered vector),
    the adjoint shards agree to rounding (the ranks' partial sums are added in another order than :1049's row order)."""
    from oracle import jets_oracle as J
    world, nb, k = 2, 4, 33
    T = np.dtype(tname)
    mp.spawn(_dense_worker, args=(world, _free_port(), tname, nb, k, str(tmp_path)), nprocs=world, join=True)
    mats, x, y = _dense_problem(T, nb, k)
    A = J.blockop([[J.JopDense(mats[(r, c)]) for c in range(nb)] for r in range(nb)])
    f_ref = J.to_array(A * J.reshape(x.copy(), J.domain(A)))
    t_ref = J.to_array(A.T * J.reshape(y.copy(), J.range_(A)))
    f = np.concatenate([np.load(tmp_path / f"f{r}.npy") for r in range(world)])
    t = np.concatenate([np.load(tmp_path / f"t{r}.npy") for r in range(world)])
    assert np.array_equal(f, f_ref)
    assert np.linalg.norm(t - t_ref) <= 1e-13 * np.linalg.norm(t_ref)


def test_partition_bookkeeping():
    import jets_b200.dist as D
    p = D.RowPartition(256, 8, 3)
    assert (p.r0, p.r1, p.nloc, p.next_cols) == (96, 128, 32, 34)
    mp_ = p.local_block_map()
    assert mp_[0][0] == (96, 95) and mp_[0][1] == (96, 96) and mp_[0][2] == (96, 97) and mp_[0][3] is None
    assert mp_[31][33] == (127, 128) and mp_[31][31] == (127, 126)
    p0 = D.RowPartition(256, 8, 0)
    assert p0.local_block_map()[0][0] is None and not p0.has_prev and p0.has_next
    with pytest.raises(ValueError):
        D.RowPartition(10, 4, 0)
