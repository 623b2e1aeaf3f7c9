"""Block-row partition of a JopBlock across GPUs (one process per GPU) -- SURVEY §8e.

The reference has no distributed path (Jets.jl is single-process; DistributedJets.jl is a separate
package), so this is the build's own design: rank g owns a contiguous range of block rows, the
matching range blocks of the result and the matching domain blocks.  For a block-banded operator
(bandwidth b: op[r,c] == JopZeroBlock for |r-c| > b) the forward apply needs only b halo blocks
from each neighbour instead of an all-gather of the whole domain, and the adjoint sends its partial
contributions to the b blocks it does not own back to their owners.  All of that -- the halo traffic over
peer memory, the flags, the host-buffer pipeline -- happens INSIDE libjets_b200 behind one call per apply
(``jets_dist_apply`` / ``jets_dist_apply_normal_host``, csrc/dist_op.cu); a dense block structure uses
NCCL all-gather / reduce-scatter inside the same call.

What is left here is host-side bookkeeping: pure functions of (nblk, world, rank) that describe the
rank-local operator (also driven over torch.distributed gloo by the CPU tests), the ``DistOp`` handle
wrapper, and the communicator bootstrap.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass


@dataclass(frozen=True)
class RowPartition:
    """Contiguous block-row ownership: rank g owns rows [r0, r1) of `nblk`."""
    nblk: int
    world: int
    rank: int
    halo: int = 1  # bandwidth of the block-banded operator

    def __post_init__(self):
        if self.nblk % self.world:
            raise ValueError(f"{self.nblk} block rows do not split evenly over {self.world} ranks")
        if self.nblk // self.world < self.halo:
            raise ValueError("fewer local block rows than the halo width")

    @property
    def nloc(self):
        return self.nblk // self.world

    @property
    def r0(self):
        return self.rank * self.nloc

    @property
    def r1(self):
        return self.r0 + self.nloc

    @property
    def next_cols(self):
        """Number of block columns of the halo-extended local domain [lo halo | own | hi halo]."""
        return self.nloc + 2 * self.halo

    def global_col(self, j):
        """Global block column of extended local column j, or None outside the operator."""
        c = self.r0 - self.halo + j
        return c if 0 <= c < self.nblk else None

    def local_block_map(self):
        """[[(r, c) or None]]: the global block behind every entry of the nloc x next_cols local
        operator (None = outside the band or outside the operator -> JopZeroBlock)."""
        rows = []
        for i in range(self.nloc):
            r = self.r0 + i
            row = []
            for j in range(self.next_cols):
                c = self.global_col(j)
                row.append((r, c) if c is not None and abs(r - c) <= self.halo else None)
            rows.append(row)
        return rows

    @property
    def has_prev(self):
        return self.rank > 0

    @property
    def has_next(self):
        return self.rank + 1 < self.world


def build_local_operator(K, part: RowPartition, make_block, zero_block):
    """Local rows of the global operator as a K.blockop over the halo-extended domain.
    ``make_block(r, c)`` builds the global block (r, c); ``zero_block()`` a JopZeroBlock."""
    Z = zero_block()
    rows = [[Z if rc is None else make_block(*rc) for rc in row] for row in part.local_block_map()]
    return K.blockop(rows)


class DistOp:
    """``jets_dist_op``: the rank-local rows of a block-row partitioned JopBlock behind ONE library call per
    apply (include/jets_b200.h).  ``A_loc`` is the nloc x (nloc + 2*halo) JopBlock over the halo-extended
    domain that ``build_local_operator`` makes (banded), or the nloc x ncol JopBlock over the whole domain
    (``dense=True``).  The vectors are the rank's own shards; halo blocks, flags and staging live inside the
    library.  Everything here is argument marshalling."""

    def __init__(self, B, A_loc, halo=1, dense=False):
        self.B, self.A_loc, self.halo, self.dense = B, A_loc, halo, dense
        h = C.c_void_p()
        if dense:
            B.check(B.lib.jets_dist_op_create_dense(A_loc._h.h, C.byref(h)))
        else:
            B.check(B.lib.jets_dist_op_create(A_loc._h.h, halo, C.byref(h)))
        self._h = h
        dom, rng = B.domain(A_loc), B.range_(A_loc)
        if dense:
            world = max(1, B.lib.jets_dist_size())
            self.own_space = B.JetSpace(dom.T, len(dom) // world)
        else:
            n = len(rng.spaces)
            self.own_space = B.JetBSpace(dom.spaces[halo:halo + n])
        self.range_space = rng

    def close(self):
        if self._h:
            self.B.check(self.B.lib.jets_dist_op_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            if self._h and self.B.lib is not None:
                self.B.lib.jets_dist_op_destroy(self._h)
        except Exception:
            pass

    def forward(self, d, x, nonlinear=False):
        """d = (A x)[own rows]; x = own domain shard."""
        self.B.check(self.B.lib.jets_dist_apply(self._h, 0 if nonlinear else 1, d._h, x._h))
        return d

    def adjoint(self, m, d):
        """m = (A' d)[own columns]; d = own range shard."""
        self.B.check(self.B.lib.jets_dist_apply(self._h, 2, m._h, d._h))
        return m

    def register(self, x):
        """Collective: lets the neighbours map this domain shard, so that forward applies on it read their halo
        blocks in place over NVLink (no halo copy).  Returns x."""
        self.B.check(self.B.lib.jets_dist_op_register(self._h, x._h))
        return x

    def normal_host(self, h_out_ptr, h_in_ptr, nchunks=0):
        """host_out = A'(A host_in) on this rank's shards, chunk-pipelined (asynchronous; ``join`` to wait)."""
        self.B.check(self.B.lib.jets_dist_apply_normal_host(self._h, C.c_void_p(h_out_ptr), C.c_void_p(h_in_ptr), nchunks))

    def join(self):
        self.B.check(self.B.lib.jets_dist_op_join(self._h))

    def info(self, what):
        return self.B.lib.jets_dist_op_info(self._h, what)

    @property
    def gate_timeouts(self):
        return self.info(6)


_BOOTSTRAP_KEEPALIVE = []


def init_host_bootstrap(B, dist_mod, torch_mod, rank, world, group=None):
    """jets_dist_init_host with torch.distributed as the out-of-band all-gather (a gloo group, or any group
    that accepts CPU tensors).  Only set-up records and host scalars travel this way."""
    AG = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)

    def _allgather(user, mine, all_, nbytes):
        try:
            t = torch_mod.frombuffer(bytearray(C.string_at(mine, nbytes)), dtype=torch_mod.uint8)
            outs = [torch_mod.empty(nbytes, dtype=torch_mod.uint8) for _ in range(world)]
            dist_mod.all_gather(outs, t, group=group)
            C.memmove(all_, b"".join(bytes(o.numpy().tobytes()) for o in outs), world * nbytes)
            return 0
        except Exception:  # never let an exception unwind through the C frame
            import traceback
            traceback.print_exc()
            return 1
    cb = AG(_allgather)
    _BOOTSTRAP_KEEPALIVE.append(cb)
    B.check(B.lib.jets_dist_init_host(rank, world, cb, None))


def init_nccl_from_torch(B, dist_mod, torch_mod, rank, world):
    """Create the library's NCCL communicator: rank 0 makes the ncclUniqueId, torch.distributed
    (already initialised by the launcher) broadcasts its 128 bytes."""
    ident = C.create_string_buffer(128)
    if rank == 0:
        B.check(B.lib.jets_dist_unique_id(ident))
    t = torch_mod.frombuffer(bytearray(ident.raw), dtype=torch_mod.uint8).cuda()
    dist_mod.broadcast(t, 0)
    ident = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
    B.check(B.lib.jets_dist_init(rank, world, ident))
