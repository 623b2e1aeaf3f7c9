"""CPU model of the distributed block-banded apply of csrc/dist_op.cu -- TEST INFRASTRUCTURE.

The device path moves halo blocks through peer memory and fixes the ORDER in which a block row (forward) or a
block column (adjoint) is summed; this model performs the same movement over any communicator with
``exchange(to_prev, to_next) -> (from_prev, from_next)`` (torch.distributed gloo in tests/test_dist_gloo.py) and
the same additions with the numpy oracle's leaf operators, so that the claim "bit-identical to the
single-process apply for halo 1" is checked on the CPU, world size 2 and 4, without a GPU.

    forward   d_r = sum over extended columns j ascending of A_loc[r, j] * x_ext[j]         (src/Jets.jl:1015-1030)
    adjoint   partial sums for the neighbours' columns are sent; the owner's column sum starts with the previous
              rank's partial, runs over its own rows ascending and ends with the next rank's partial (:1039-1055).
"""
import numpy as np


def forward(J, part, comm, A_loc, x_own):
    """x_own: the rank's nloc domain blocks (numpy).  Returns its nloc range blocks."""
    h, n = part.halo, part.nloc
    lo, hi = comm.exchange(x_own[:h] if part.has_prev else None, x_own[n - h:] if part.has_next else None)
    dom = J.domain(A_loc)
    ext = [np.zeros(len(J.space(dom, j + 1)), dtype=x_own[0].dtype) for j in range(n + 2 * h)]
    for k in range(h):
        if lo is not None:
            ext[k] = lo[k]
        if hi is not None:
            ext[n + h + k] = hi[k]
    for b in range(n):
        ext[h + b] = x_own[b]
    out = []
    for r in range(n):
        acc = np.zeros(len(J.space(J.range_(A_loc), r + 1)), dtype=x_own[0].dtype)
        for j in range(n + 2 * h):
            blk = J.getblock(A_loc, r + 1, j + 1)
            if not J.iszero(blk):
                acc = acc + J.to_array(blk * ext[j])
        out.append(acc)
    return out


def _column(J, A_loc, j, d_own, first=None, last=None):
    n = len(d_own)
    acc = np.zeros(len(J.space(J.domain(A_loc), j + 1)), dtype=d_own[0].dtype)
    if first is not None:
        acc = acc + first
    for i in range(n):
        blk = J.getblock(A_loc, i + 1, j + 1)
        if not J.iszero(blk):
            acc = acc + J.to_array(J.adjoint(blk) * d_own[i])
    if last is not None:
        acc = acc + last
    return acc


def adjoint(J, part, comm, A_loc, d_own):
    """d_own: the rank's nloc range blocks.  Returns its nloc domain blocks of A' d."""
    h, n = part.halo, part.nloc
    to_prev = [_column(J, A_loc, j, d_own) for j in range(h)] if part.has_prev else None
    to_next = [_column(J, A_loc, n + h + k, d_own) for k in range(h)] if part.has_next else None
    from_prev, from_next = comm.exchange(to_prev, to_next)
    out = []
    for b in range(n):
        first = from_prev[b] if (from_prev is not None and b < h) else None
        last = from_next[b - (n - h)] if (from_next is not None and b >= n - h) else None
        out.append(_column(J, A_loc, h + b, d_own, first, last))
    return out


# ---- dense block structure (csrc/dist_op.cu: dense_apply) ------------------------------------------------------
def dense_forward(J, comm, A_loc, x_shard):
    """Every block row needs the whole domain (:1015-1030 without zero blocks): all-gather the equal shards, apply the
    rank's rows.  comm.allgather(a) -> the shards of all ranks concatenated in rank order."""
    full = comm.allgather(x_shard)
    return J.to_array(A_loc * J.reshape(full, J.domain(A_loc)))


def dense_adjoint(J, comm, A_loc, d_rows, nranks, rank):
    """The rank's rows contribute to every block column (:1039-1055): full-length partial, summed over the ranks,
    every rank keeps its shard (reduce-scatter).  Complex shards travel as twice as many reals, like the device path."""
    part = J.to_array(J.adjoint(A_loc) * J.reshape(d_rows, J.range_(A_loc)))
    cplx = np.iscomplexobj(part)
    flat = np.ascontiguousarray(part).view(part.real.dtype) if cplx else part
    mine = comm.reduce_scatter(flat)
    return mine.view(part.dtype) if cplx else mine
