"""world_size=2 (and 4) CPU tests of the block-row partition + halo protocol (jets_b200.dist) over
torch.distributed gloo.  The partition logic is backend-agnostic: here it drives the numpy oracle
with a gloo communicator; on GPUs the same functions drive libjets_b200 with its NCCL communicator.
The distributed result must equal the single-process oracle bit for bit (the halo adds happen in
rank order and every block sum has at most one remote contribution per side)."""
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
    """halo_exchange / halo_reduce on oracle BlockArrays (numpy) via gloo send/recv."""

    def __init__(self, part):
        self.part = part

    def _xfer(self, sends, recvs):
        reqs = []
        for dst, arr in sends:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(arr)), dst))
        bufs = []
        for src, shape, dtype in recvs:
            t = torch.empty(shape, dtype=dtype)
            bufs.append(t)
            reqs.append(dist.irecv(t, src))
        for r in reqs:
            r.wait()
        return [b.numpy() for b in bufs]

    def halo_exchange(self, x_ext, h, n):
        p = self.part
        blk = x_ext.arrays
        sends, recvs = [], []
        tdt = torch.from_numpy(blk[0]).dtype
        if p.has_next:
            sends += [(p.rank + 1, blk[n + k]) for k in range(h)]          # my last h own blocks
            recvs += [(p.rank + 1, blk[h + n + k].shape, tdt) for k in range(h)]
        if p.has_prev:
            sends += [(p.rank - 1, blk[h + k]) for k in range(h)]          # my first h own blocks
            recvs += [(p.rank - 1, blk[k].shape, tdt) for k in range(h)]
        got = self._xfer(sends, recvs)
        i = 0
        if p.has_next:
            for k in range(h):
                blk[h + n + k][...] = got[i]
                i += 1
        if p.has_prev:
            for k in range(h):
                blk[k][...] = got[i]
                i += 1

    # split form used by dist.OverlappedBanded (no streams on the CPU: the hooks are no-ops)
    def halo_reduce_begin(self, m_ext, h, n):
        p = self.part
        blk = m_ext.arrays
        tdt = torch.from_numpy(blk[0]).dtype
        sends, recvs = [], []
        if p.has_prev:
            sends += [(p.rank - 1, blk[k]) for k in range(h)]
            recvs += [(p.rank - 1, blk[h + k].shape, tdt) for k in range(h)]
        if p.has_next:
            sends += [(p.rank + 1, blk[h + n + k]) for k in range(h)]
            recvs += [(p.rank + 1, blk[n + k].shape, tdt) for k in range(h)]
        self._staged = self._xfer(sends, recvs)

    def halo_reduce_end(self, m_ext, h, n):
        p = self.part
        blk = m_ext.arrays
        got, i = self._staged, 0
        if p.has_prev:
            for k in range(h):
                blk[h + k][...] = blk[h + k] + got[i]
                i += 1
        if p.has_next:
            for k in range(h):
                blk[n + k][...] = blk[n + k] + got[i]
                i += 1

    def halo_exchange_begin(self, x_ext, h, n):
        self.halo_exchange(x_ext, h, n)

    def halo_exchange_end(self):
        pass

    def halo_reduce(self, m_ext, h, n):
        p = self.part
        blk = m_ext.arrays
        tdt = torch.from_numpy(blk[0]).dtype
        sends, recvs = [], []
        if p.has_prev:
            sends += [(p.rank - 1, blk[k]) for k in range(h)]               # partial for prev's last h
            recvs += [(p.rank - 1, blk[h + k].shape, tdt) for k in range(h)]
        if p.has_next:
            sends += [(p.rank + 1, blk[h + n + k]) for k in range(h)]       # partial for next's first h
            recvs += [(p.rank + 1, blk[n + k].shape, tdt) for k in range(h)]
        got = self._xfer(sends, recvs)
        i = 0
        if p.has_prev:                                                        # previous rank first
            for k in range(h):
                blk[h + k][...] = blk[h + k] + got[i]
                i += 1
        if p.has_next:
            for k in range(h):
                blk[n + k][...] = blk[n + k] + got[i]
                i += 1


def _global_problem(nblk, n, T, seed=0):
    g = np.random.default_rng(seed)
    W = [g.random(n).astype(T) for _ in range(nblk)]
    m = g.random(nblk * n).astype(T)
    d = g.random(nblk * n).astype(T)
    return W, m, d


def _make_block(J, T, n, W):
    def mk(r, c):
        if r == c:
            return J.JopDiagonal(W[r])
        return J.JopStencil(T, n, "fdiff") if c == r + 1 else J.JopStencil(T, n, "lap")
    return mk


def _worker(rank, world, port, nblk, n, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import jets_oracle as J
    import jets_b200.dist as D
    T = np.float64
    W, m, d = _global_problem(nblk, n, T)
    part = D.RowPartition(nblk, world, rank)
    sp = J.JetSpace(T, n)
    A = D.build_local_operator(J, part, _make_block(J, T, n, W), lambda: J.JopZeroBlock(sp, sp))
    assert J.nblocks(A) == (part.nloc, part.nloc + 2)
    comm = GlooComm(part)
    x_ext = J.zeros(J.domain(A))
    for k in range(part.nloc):
        x_ext.arrays[1 + k][...] = m[(part.r0 + k) * n:(part.r0 + k + 1) * n]
    d_loc = D.forward(J, part, comm, A, x_ext, J.zeros(J.range_(A)))
    dd = J.zeros(J.range_(A))
    for k in range(part.nloc):
        dd.arrays[k][...] = d[(part.r0 + k) * n:(part.r0 + k + 1) * n]
    m_ext = D.adjoint(J, part, comm, A, J.zeros(J.domain(A)), dd)
    # the overlapped decomposition (interior/boundary rows, halo partials first) must give the same bits
    def view(x, first, cnt):
        idx, o = [], 0
        for a in x.arrays[first:first + cnt]:
            idx.append((o + 1, o + a.size))
            o += a.size
        return J.BlockArray(x.arrays[first:first + cnt], idx)
    x2, d2, m2 = J.zeros(J.domain(A)), J.zeros(J.range_(A)), J.zeros(J.domain(A))
    for k in range(part.nloc):
        x2.arrays[1 + k][...] = m[(part.r0 + k) * n:(part.r0 + k + 1) * n]
    ov = D.OverlappedBanded(J, part, comm, _make_block(J, T, n, W), lambda: J.JopZeroBlock(sp, sp), x2, d2, m2, view)
    ov.forward()
    assert np.array_equal(J.to_array(d2), J.to_array(d_loc))
    for k in range(part.nloc):
        d2.arrays[k][...] = dd.arrays[k]
    ov.adjoint()
    for k in range(part.nloc):
        assert np.array_equal(m2.arrays[1 + k], m_ext.arrays[1 + k])
    np.save(os.path.join(out_dir, f"f{rank}.npy"), J.to_array(d_loc))
    np.save(os.path.join(out_dir, f"t{rank}.npy"), np.concatenate([m_ext.arrays[1 + k] for k in range(part.nloc)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_block_row_partition_matches_single_process(world, tmp_path):
    from oracle import jets_oracle as J
    nblk, n, T = 8, 257, np.float64
    mp.spawn(_worker, args=(world, _free_port(), nblk, n, str(tmp_path)), nprocs=world, join=True)
    W, m, d = _global_problem(nblk, n, T)
    sp = J.JetSpace(T, n)
    mk = _make_block(J, T, n, W)
    A = J.blockop([[mk(r, c) if abs(r - c) <= 1 else J.JopZeroBlock(sp, sp) for c in range(nblk)] for r in range(nblk)])
    f_ref = J.to_array(A * J.reshape(m.copy(), J.domain(A)))
    t_ref = J.to_array(A.T * J.reshape(d.copy(), J.range_(A)))
    f = np.concatenate([np.load(tmp_path / f"f{r}.npy") for r in range(world)])
    t = np.concatenate([np.load(tmp_path / f"t{r}.npy") for r in range(world)])
    assert np.array_equal(f, f_ref)
    # interior blocks: identical; partition-boundary blocks: the remote partial is added last instead
    # of in column order -> same values up to one rounding of a 3-term sum
    assert np.allclose(t, t_ref, rtol=1e-15, atol=1e-15)


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
