"""CPU replay of the host-buffer pipeline's issue order (csrc/dist_op.cu: pipe_schedule, exported as
jets_dist_pipeline_schedule -- a pure host function, no GPU needed).  Every rank executes ITS schedule on numpy
blocks with NaN-poisoned buffers, in the order the compute stream would, with the in-kernel flag waits modelled
as "this item cannot run yet": a forward issued before the uploads it waits for, an adjoint before its forwards,
a chunk that reads a halo / staging buffer the schedule does not gate, or a cyclic wait between ranks leaves NaNs
(or a deadlock) behind.  The concatenated results must equal A'(A m) of the global block-banded operator."""
import numpy as np
import pytest

import jets_b200 as B   # loads libjets_b200.so (the schedule is host code)

pipeline = B.pipeline
distmod = B.dist


def run(nblk, world, halo, nchunks, bs=3, seed=0):
    g = np.random.default_rng(seed)
    N = nblk * bs
    M = np.zeros((N, N))
    for r in range(nblk):
        for c in range(nblk):
            if abs(r - c) <= halo:
                M[r * bs:(r + 1) * bs, c * bs:(c + 1) * bs] = g.random((bs, bs))
    x = g.random(N)
    want = M.T @ (M @ x)
    blkM = lambda r, c: M[r * bs:(r + 1) * bs, c * bs:(c + 1) * bs]
    ranks = []
    for rk in range(world):
        part = distmod.RowPartition(nblk, world, rk, halo=halo)
        n, h = part.nloc, part.halo
        nz = np.array([[rc is not None for rc in row] for row in part.local_block_map()])
        items, chunks, up_need = pipeline.schedule(nz, h, nchunks, part.has_prev, part.has_next)
        nan = lambda *s: np.full(s, np.nan)
        ranks.append(dict(part=part, n=n, h=h, nz=nz, items=items, chunks=chunks, up_need=up_need, pos=0, uploaded=-1,
                          x=nan(n, bs), lo=nan(h, bs), hi=nan(h, bs), d=nan(n, bs), slo=nan(h, bs), shi=nan(h, bs),
                          m=nan(n, bs), out=nan(n, bs), flags=set()))

    def chunk_of(R, blk):
        return next(k for k, (a, b) in enumerate(R["chunks"]) if a <= blk < b)

    def upload_to(R, k):           # uploads land in chunk order; never further than the schedule waits for
        while R["uploaded"] < k:
            R["uploaded"] += 1
            a, b = R["chunks"][R["uploaded"]]
            R["x"][a:b] = x.reshape(nblk, bs)[R["part"].r0 + a:R["part"].r0 + b]

    def ext(R, j):                 # extended column j of the rank-local domain
        n, h = R["n"], R["h"]
        return R["lo"][j] if j < h else R["hi"][j - n - h] if j >= n + h else R["x"][j - h]

    def try_item(rk):
        R = ranks[rk]
        what, k = R["items"][R["pos"]]
        p, n, h = R["part"], R["n"], R["h"]
        if what == "push_prev":
            upload_to(R, chunk_of(R, h - 1))
            ranks[rk - 1]["hi"][:] = R["x"][:h]
            ranks[rk - 1]["flags"].add("hi")
        elif what == "push_next":
            upload_to(R, chunk_of(R, n - 1))
            ranks[rk + 1]["lo"][:] = R["x"][n - h:]
            ranks[rk + 1]["flags"].add("lo")
        elif what == "fwd":
            a, b = R["chunks"][k]
            reads_lo = R["nz"][a:b, :h].any()
            reads_hi = R["nz"][a:b, n + h:].any()
            if (reads_lo and "lo" not in R["flags"]) or (reads_hi and "hi" not in R["flags"]):
                return False           # the kernel would wait for the neighbour's flag
            upload_to(R, R["up_need"][k])
            for i in range(a, b):
                acc = np.zeros(bs)
                for j in range(n + 2 * h):
                    if R["nz"][i, j]:
                        acc = acc + blkM(p.r0 + i, p.global_col(j)) @ ext(R, j)
                R["d"][i] = acc
        elif what in ("partial_prev", "partial_next"):
            cols = range(0, h) if what == "partial_prev" else range(n + h, n + 2 * h)
            dst = ranks[rk - 1]["shi"] if what == "partial_prev" else ranks[rk + 1]["slo"]
            for kk, j in enumerate(cols):
                acc = np.zeros(bs)
                for i in range(n):
                    if R["nz"][i, j]:
                        acc = acc + blkM(p.r0 + i, p.global_col(j)).T @ R["d"][i]
                dst[kk] = acc
            (ranks[rk - 1] if what == "partial_prev" else ranks[rk + 1])["flags"].add("shi" if what == "partial_prev" else "slo")
        elif what == "adj":
            a, b = R["chunks"][k]
            need_lo = p.has_prev and a < h
            need_hi = p.has_next and b > n - h
            if (need_lo and "slo" not in R["flags"]) or (need_hi and "shi" not in R["flags"]):
                return False
            for bcol in range(a, b):
                acc = np.zeros(bs)
                if p.has_prev and bcol < h:
                    acc = acc + R["slo"][bcol]
                for i in range(n):
                    if R["nz"][i, bcol + h]:
                        acc = acc + blkM(p.r0 + i, p.global_col(bcol + h)).T @ R["d"][i]
                if p.has_next and bcol >= n - h:
                    acc = acc + R["shi"][bcol - (n - h)]
                R["m"][bcol] = acc
            R["out"][a:b] = R["m"][a:b]          # the download follows the chunk's adjoint
        R["pos"] += 1
        return True

    while any(R["pos"] < len(R["items"]) for R in ranks):
        progress = False
        for rk, R in enumerate(ranks):
            while R["pos"] < len(R["items"]) and try_item(rk):
                progress = True
        assert progress, "deadlock: every rank waits for a flag nobody will raise"
    got = np.concatenate([R["out"].reshape(-1) for R in ranks])
    for R in ranks:
        assert sorted(k for w, k in R["items"] if w == "fwd") == list(range(len(R["chunks"])))
        assert sorted(k for w, k in R["items"] if w == "adj") == list(range(len(R["chunks"])))
    assert not np.isnan(got).any(), "a stage ran before its inputs were ready"
    assert np.allclose(got, want, rtol=1e-13, atol=0)


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("nchunks", [1, 2, 3, 4, 5, 8, 32])
@pytest.mark.parametrize("halo", [1, 2])
def test_schedule_respects_every_dependency(world, nchunks, halo):
    run(nblk=world * 8, world=world, halo=halo, nchunks=nchunks, seed=world * 100 + nchunks)


def _tridiag(n, h=1):
    return np.array([[abs(i + h - j) <= h for j in range(n + 2 * h)] for i in range(n)])


def test_single_rank_schedule_is_the_plain_software_pipeline():
    nz = _tridiag(4)
    nz[0, 0] = nz[3, 5] = False     # no neighbours: the halo columns are zero blocks
    seq, chunks, up_need = pipeline.schedule(nz, 1, 4, False, False)
    assert chunks == [(0, 1), (1, 2), (2, 3), (3, 4)] and up_need == [1, 2, 3, 3]
    assert seq == [("fwd", 0), ("fwd", 1), ("adj", 0), ("fwd", 2), ("adj", 1), ("fwd", 3), ("adj", 2), ("adj", 3)]


def test_middle_rank_defers_only_the_boundary_chunks():
    seq, chunks, _ = pipeline.schedule(_tridiag(8), 1, 8, True, True)
    assert seq[0] == ("push_prev", 0)
    early = seq[1:seq.index(("push_next", 0))]
    assert [i for i in early if i[0] == "fwd"] == [("fwd", k) for k in range(1, 7)]
    assert [i for i in early if i[0] == "adj"] == [("adj", j) for j in range(2, 6)]
    tail = seq[seq.index(("push_next", 0)):]
    assert tail == [("push_next", 0), ("fwd", 0), ("fwd", 7), ("partial_prev", 0), ("partial_next", 0), ("adj", 0), ("adj", 1),
                    ("adj", 6), ("adj", 7)]
