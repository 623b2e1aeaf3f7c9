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


def chunk_bounds(nloc, halo, nchunks):
    """Cuts the `nloc` local block rows into at most `nchunks` chunks of consecutive rows, none shorter
    than the halo width (so that a chunk's forward needs its two neighbouring chunks only)."""
    nchunks = max(1, min(nchunks, nloc // max(1, halo)))
    b = [nloc * k // nchunks for k in range(nchunks + 1)]
    return [(b[k], b[k + 1]) for k in range(nchunks) if b[k + 1] > b[k]]


def compute_schedule(K, has_prev, has_next):
    """Order of the compute-stream work of one pipelined step of a rank-local block-banded operator
    cut into K chunks, as a list of

        ("fwd", k)         forward apply of chunk k             (needs uploads of chunks <= k+1)
        ("adj", j)         adjoint apply of own-column chunk j  (needs forwards j-1, j, j+1)
        ("exchange",)      halo gather of x from the neighbours (needs ALL uploads, on every rank)
        ("partials",)      partial sums for the neighbours' columns (needs the boundary forwards)
        ("reduce_begin",), ("reduce_end",)   ship / add those partials

    and the set of chunks whose result is final only after "reduce_end" (their downloads go last).
    Chunks that touch nothing of a neighbouring rank stream through as soon as their uploads land;
    the first chunk (its forward reads the previous rank's LAST block, which arrives at the very end of
    that rank's upload) and the last one are deferred together with the adjoint chunks that depend on
    them.  With one rank this degenerates to the plain k / k-1 software pipeline."""
    late_f = set()
    if has_prev:
        late_f.add(0)
    if has_next:
        late_f.add(K - 1)
    early_f = [k for k in range(K) if k not in late_f]

    def deps(j):
        return [i for i in (j - 1, j, j + 1) if 0 <= i < K]
    late_a = {j for j in range(K) if any(i in late_f for i in deps(j))}
    seq, issued = [], set()
    for k in early_f:
        seq.append(("fwd", k))
        issued.add(k)
        for j in (k - 1, k):          # adjoint chunks completed by this forward
            if 0 <= j < K and j not in late_a and ("adj", j) not in seq and all(i in issued for i in deps(j)):
                seq.append(("adj", j))
    if late_f:
        seq.append(("exchange",))
        for k in sorted(late_f):
            seq.append(("fwd", k))
        seq.append(("partials",))
        seq.append(("reduce_begin",))
        for j in sorted(late_a):
            seq.append(("adj", j))
        seq.append(("reduce_end",))
    final_after_reduce = set()
    if has_prev:
        final_after_reduce.add(0)
    if has_next:
        final_after_reduce.add(K - 1)
    return seq, late_a | final_after_reduce


class ChunkedBandedApply:
    """Pipelined ``m_host -> d = A m -> m' = A' d -> m'_host`` for the local rows of a block-banded
    operator described by a ``dist.RowPartition`` (world size 1 or the rank-local part).

    ``x_ext`` / ``m_ext`` are the halo-extended domain vectors (nloc + 2*halo blocks) and ``d`` the
    range vector (nloc blocks) the monolithic path uses; ``make_block(r, c)`` / ``zero_block()``
    build the operator blocks exactly as for ``dist.build_local_operator``.
    """

    def __init__(self, B, torch, part, make_block, zero_block, x_ext, d, m_ext, nchunks=16, comm=None):
        self.B, self.torch, self.part, self.comm = B, torch, part, comm
        h, n = part.halo, part.nloc
        self.chunks = chunk_bounds(n, h, nchunks)
        if part.world > 1 and comm is None:
            raise ValueError("a multi-rank pipeline needs the communicator (halo exchange / reduce)")
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

        self.x_ext, self.d, self.m_ext = x_ext, d, m_ext
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
        # partial sums this rank contributes to its neighbours' columns (they live in the halo blocks of
        # m_ext until halo_reduce ships them): the boundary rows only, as in dist.OverlappedBanded
        self.partials = []
        if part.has_prev:
            self.partials.append((B.adjoint(B.blockop([[blk(bmap[i][j]) for j in range(0, h)] for i in range(0, h)])),
                                  view(m_ext, 0, h), view(d, 0, h)))
        if part.has_next:
            self.partials.append((B.adjoint(B.blockop([[blk(bmap[i][j]) for j in range(n + h, n + 2 * h)] for i in range(n - h, n)])),
                                  view(m_ext, n + h, h), view(d, n - h, h)))
        self.schedule, self.late_down = compute_schedule(len(self.chunks), part.has_prev, part.has_next)
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
            part, comm = self.part, self.comm
            for item in self.schedule:
                if item[0] == "fwd":                           # forward chunk k needs uploads <= k+1
                    k = item[1]
                    self.s_comp.wait_event(ev_up[min(k + 1, K - 1)])
                    Af, dv, xv = self.fwd[k]
                    B.mul_(dv, Af, xv)
                elif item[0] == "adj":                         # adjoint chunk j: forwards j-1..j+1 were issued before it
                    At, mv, dv = self.adj[item[1]]
                    B.mul_(mv, At, dv)
                    if item[1] not in self.late_down:
                        ev_adj[item[1]].record(self.s_comp)
                elif item[0] == "exchange":                    # every rank has ALL of its x on the device
                    self.s_comp.wait_event(ev_up[K - 1])
                    comm.halo_exchange(self.x_ext, part.halo, part.nloc)
                elif item[0] == "partials":
                    for At, mv, dv in self.partials:
                        B.mul_(mv, At, dv)
                elif item[0] == "reduce_begin":
                    comm.halo_reduce_begin(self.m_ext, part.halo, part.nloc)
                elif item[0] == "reduce_end":
                    comm.halo_reduce_end(self.m_ext, part.halo, part.nloc)
                    for j in self.late_down:
                        ev_adj[j].record(self.s_comp)
            done = torch.cuda.Event()
            done.record(self.s_comp)
            self._last = done
            self._on(self.s_down)
            for k in [j for j in range(K) if j not in self.late_down] + sorted(self.late_down):
                self.s_down.wait_event(ev_adj[k])
                B.check(B.lib.jets_buf_download_async(self.down[k]._h, -1, C.c_void_p(h_out.data_ptr() + self.offsets[k] * esz),
                                                      len(self.down[k])))
            fin = torch.cuda.Event(enable_timing=True)
            fin.record(self.s_down)
            self._last_down = fin
        finally:
            self._on(restore_stream)
        return fin
