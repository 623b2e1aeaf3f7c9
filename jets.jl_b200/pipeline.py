"""Host-buffer apply of a block-banded JopBlock, pipelined over block-row chunks.

``d = A*m`` followed by ``m' = A'*d`` with ``m`` and ``m'`` in (pinned) HOST memory is bound by the
host link (16 GB each way at config 5 against ~16 ms of HBM work), so the only thing that matters
is keeping both directions of the link busy at once: the vector is cut into chunks of consecutive
block rows; chunk k is uploaded on one stream while the forward apply of chunk k-1 and the adjoint
apply of chunk k-2 run on a second stream and the finished chunk of ``m'`` goes back to the host on a
third.  The chunk operators are ordinary JopBlocks over VIEWS of the same device vectors (block
rows a..b of the operator over the halo-extended window of its domain, src/Jets.jl:1015-1030; for
the adjoint the transposed window, :1039-1055), so every output element is computed by exactly the
same fused kernel arithmetic as in the monolithic apply and the result is bit-identical.

The reference has no counterpart (Jets.jl applies operators to arrays already in memory); this is
the caller-side loop a host-resident solver would write around ``mul!``.
"""
from __future__ import annotations

import ctypes as C


class ChunkedBandedApply:
    """Pipelined ``m_host -> d = A m -> m' = A' d -> m'_host`` for the local rows of a block-banded
    operator described by a ``dist.RowPartition`` (world size 1 or the rank-local part).

    ``x_ext`` / ``m_ext`` are the halo-extended domain vectors (nloc + 2*halo blocks) and ``d`` the
    range vector (nloc blocks) the monolithic path uses; ``make_block(r, c)`` / ``zero_block()``
    build the operator blocks exactly as for ``dist.build_local_operator``.
    """

    def __init__(self, B, torch, part, make_block, zero_block, x_ext, d, m_ext, nchunks=16):
        self.B, self.torch, self.part = B, torch, part
        h, n = part.halo, part.nloc
        nchunks = max(1, min(nchunks, n // max(1, h)))
        bounds = [n * k // nchunks for k in range(nchunks + 1)]
        self.chunks = [(bounds[k], bounds[k + 1]) for k in range(nchunks) if bounds[k + 1] > bounds[k]]
        bmap = part.local_block_map()
        Z = zero_block()
        cache = {}

        def blk(rc):
            if rc is None:
                return Z
            if rc not in cache:
                cache[rc] = make_block(*rc)
            return cache[rc]

        def view(x, first, count):
            hd = C.c_void_p()
            B.check(B.lib.jets_buf_view(x._h, first, count, C.byref(hd)))
            return B.DeviceArray(hd, B.JetBSpace(x.space.spaces[first:first + count]), owner=x)

        self.fwd, self.adj, self.up, self.down = [], [], [], []
        for a, b in self.chunks:
            # forward: rows [a,b) read extended columns [a, b+2h)
            Af = B.blockop([[blk(bmap[i][j]) for j in range(a, b + 2 * h)] for i in range(a, b)])
            self.fwd.append((Af, view(d, a, b - a), view(x_ext, a, b - a + 2 * h)))
            # adjoint: own columns [a,b) (= extended [a+h, b+h)) collect rows [a-h, b+h)
            ra, rb = max(0, a - h), min(n, b + h)
            At = B.adjoint(B.blockop([[blk(bmap[i][j]) for j in range(a + h, b + h)] for i in range(ra, rb)]))
            self.adj.append((At, view(m_ext, a + h, b - a), view(d, ra, rb - ra)))
            self.up.append(view(x_ext, a + h, b - a))
            self.down.append(view(m_ext, a + h, b - a))
        self.offsets = []
        off = 0
        for v in self.up:
            self.offsets.append(off)
            off += len(v)
        self.nelem = off
        self.s_up, self.s_comp, self.s_down = (torch.cuda.Stream() for _ in range(3))
        self._last = None   # event: previous step's compute finished (x_ext may be overwritten)
        self._last_down = None

    def _on(self, stream):
        self.B.check(self.B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))

    def start_event(self, after_stream):
        """A timing event at the head of the pipeline (ordered after everything on ``after_stream``)."""
        ev = self.torch.cuda.Event(enable_timing=True)
        self.s_up.wait_stream(after_stream)
        self.s_comp.wait_stream(after_stream)
        self.s_down.wait_stream(after_stream)
        ev.record(self.s_up)
        return ev

    def step(self, h_in, h_out, restore_stream):
        """One pipelined step.  ``h_in``/``h_out``: pinned host tensors of ``nelem`` elements.
        Returns the event that marks the completion of the last download."""
        B, torch = self.B, self.torch
        esz = h_in.element_size()
        K = len(self.chunks)
        ev_up = [torch.cuda.Event() for _ in range(K)]
        ev_fwd = [torch.cuda.Event() for _ in range(K)]
        ev_adj = [torch.cuda.Event() for _ in range(K)]
        try:
            if self._last is not None:
                self.s_up.wait_event(self._last)
            self._on(self.s_up)
            for k in range(K):
                B.check(B.lib.jets_buf_upload_async(self.up[k]._h, -1, C.c_void_p(h_in.data_ptr() + self.offsets[k] * esz),
                                                    len(self.up[k])))
                ev_up[k].record(self.s_up)
            self._on(self.s_comp)
            if self._last_down is not None:
                self.s_comp.wait_event(self._last_down)      # m_ext of the previous step fully downloaded
            for k in range(K + 1):
                if k < K:                                      # forward chunk k needs uploads <= k+1
                    self.s_comp.wait_event(ev_up[min(k + 1, K - 1)])
                    Af, dv, xv = self.fwd[k]
                    B.mul_(dv, Af, xv)
                    ev_fwd[k].record(self.s_comp)
                j = k - 1                                      # adjoint chunk j needs forwards <= j+1 (just issued)
                if 0 <= j < K and (k < K or j == K - 1):
                    At, mv, dv = self.adj[j]
                    B.mul_(mv, At, dv)
                    ev_adj[j].record(self.s_comp)
            done = torch.cuda.Event()
            done.record(self.s_comp)
            self._last = done
            self._on(self.s_down)
            for k in range(K):
                self.s_down.wait_event(ev_adj[k])
                B.check(B.lib.jets_buf_download_async(self.down[k]._h, -1, C.c_void_p(h_out.data_ptr() + self.offsets[k] * esz),
                                                      len(self.down[k])))
            fin = torch.cuda.Event(enable_timing=True)
            fin.record(self.s_down)
            self._last_down = fin
        finally:
            self._on(restore_stream)
        return fin
