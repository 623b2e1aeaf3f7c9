"""CPU check of the chunked host-buffer pipeline's schedule (jets.jl_b200/pipeline.py: compute_schedule):
every rank executes its schedule on numpy blocks with NaN-poisoned buffers -- a forward that runs before
the uploads it needs, an adjoint before its forwards, a download before the halo reduce, or a read of a
halo block before the exchange leaves NaNs in the result -- and the concatenated own blocks must equal
A'(A m) of the global block-banded operator.  The same schedule drives the CUDA streams on the device."""
import numpy as np
import pytest

import jets_b200 as B   # loads libjets_b200.so (no GPU needed for the host-side scheduling logic)

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
        chunks = pipeline.chunk_bounds(n, h, nchunks)
        sched, late_down = pipeline.compute_schedule(len(chunks), part.has_prev, part.has_next)
        ranks.append(dict(part=part, n=n, h=h, chunks=chunks, sched=list(sched), late=late_down, pos=0,
                          x_ext=np.full((n + 2 * h, bs), np.nan), d=np.full((n, bs), np.nan),
                          m_ext=np.full((n + 2 * h, bs), np.nan), uploaded=-1, out=np.full((n, bs), np.nan),
                          downloaded=set(), adj_done=set(), reduced=False))

    def upload_to(R, k):           # uploads land in chunk order; never more than the schedule waits for
        while R["uploaded"] < k:
            R["uploaded"] += 1
            a, b = R["chunks"][R["uploaded"]]
            r0 = R["part"].r0
            R["x_ext"][R["h"] + a:R["h"] + b] = x.reshape(nblk, bs)[r0 + a:r0 + b]

    def fwd(R, k):
        a, b = R["chunks"][k]
        p, h = R["part"], R["h"]
        for i in range(a, b):
            acc = np.zeros(bs)
            for j in range(i, i + 2 * h + 1):
                c = p.global_col(j)
                if c is not None and abs(p.r0 + i - c) <= h:
                    acc = acc + blkM(p.r0 + i, c) @ R["x_ext"][j]
            R["d"][i] = acc

    def adj_cols(R, cols, rows, dst0):
        p, h = R["part"], R["h"]
        for jj, j in enumerate(cols):
            c = p.global_col(j)
            acc = np.zeros(bs)
            for i in rows:
                if c is not None and abs(p.r0 + i - c) <= h:
                    acc = acc + blkM(p.r0 + i, c).T @ R["d"][i]
            R["m_ext"][dst0 + jj] = acc

    def step_until(R, stop):
        while R["pos"] < len(R["sched"]):
            it = R["sched"][R["pos"]]
            if it[0] in stop:
                return it[0]
            R["pos"] += 1
            K, n, h = len(R["chunks"]), R["n"], R["h"]
            if it[0] == "fwd":
                upload_to(R, min(it[1] + 1, K - 1))
                fwd(R, it[1])
            elif it[0] == "adj":
                a, b = R["chunks"][it[1]]
                adj_cols(R, range(a + h, b + h), range(max(0, a - h), min(n, b + h)), a + h)
                R["adj_done"].add(it[1])
                if it[1] not in R["late"]:
                    R["out"][a:b] = R["m_ext"][a + h:b + h]      # early download
            elif it[0] == "partials":
                if R["part"].has_prev:
                    adj_cols(R, range(0, h), range(0, h), 0)
                if R["part"].has_next:
                    adj_cols(R, range(n + h, n + 2 * h), range(n - h, n), n + h)
        return None

    for R in ranks:
        assert step_until(R, {"exchange"}) in ("exchange", None)
    if world > 1:
        for R in ranks:                                  # the fence: every rank has uploaded everything
            upload_to(R, len(R["chunks"]) - 1)
        for rk, R in enumerate(ranks):
            n, h = R["n"], R["h"]
            if R["part"].has_prev:
                R["x_ext"][0:h] = ranks[rk - 1]["x_ext"][n:n + h]
            if R["part"].has_next:
                R["x_ext"][n + h:] = ranks[rk + 1]["x_ext"][h:2 * h]
            R["pos"] += 1
        for R in ranks:
            assert step_until(R, {"reduce_begin"}) == "reduce_begin"
        staged = []
        for rk, R in enumerate(ranks):                   # reduce_begin: pull the neighbours' partials
            n, h = R["n"], R["h"]
            lo = ranks[rk - 1]["m_ext"][n + h:].copy() if R["part"].has_prev else None
            hi = ranks[rk + 1]["m_ext"][0:h].copy() if R["part"].has_next else None
            staged.append((lo, hi))
            R["pos"] += 1
        for R in ranks:
            assert step_until(R, {"reduce_end"}) == "reduce_end"
        for rk, R in enumerate(ranks):                   # reduce_end: previous rank first
            n, h = R["n"], R["h"]
            lo, hi = staged[rk]
            if lo is not None:
                R["m_ext"][h:2 * h] += lo
            if hi is not None:
                R["m_ext"][n:n + h] += hi
            R["pos"] += 1
            for j in R["late"]:
                a, b = R["chunks"][j]
                assert j in R["adj_done"]
                R["out"][a:b] = R["m_ext"][a + h:b + h]
    got = np.concatenate([R["out"].reshape(-1) for R in ranks])
    for R in ranks:
        assert R["pos"] == len(R["sched"])
        assert sorted(k for k in R["adj_done"]) == list(range(len(R["chunks"])))
    assert not np.isnan(got).any(), "a stage ran before its inputs were ready"
    assert np.allclose(got, want, rtol=1e-13, atol=0)


@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("nchunks", [1, 2, 3, 4, 5, 8, 32])
@pytest.mark.parametrize("halo", [1, 2])
def test_schedule_respects_every_dependency(world, nchunks, halo):
    run(nblk=world * 8, world=world, halo=halo, nchunks=nchunks, seed=world * 100 + nchunks)


def test_single_rank_schedule_is_the_plain_software_pipeline():
    seq, late = pipeline.compute_schedule(4, False, False)
    assert seq == [("fwd", 0), ("fwd", 1), ("adj", 0), ("fwd", 2), ("adj", 1), ("fwd", 3), ("adj", 2), ("adj", 3)]
    assert late == set()


def test_middle_rank_defers_only_the_boundary_chunks():
    seq, late = pipeline.compute_schedule(8, True, True)
    early = seq[:seq.index(("exchange",))]
    assert [i for i in early if i[0] == "fwd"] == [("fwd", k) for k in range(1, 7)]
    assert [i for i in early if i[0] == "adj"] == [("adj", j) for j in range(2, 6)]
    assert late == {0, 1, 6, 7}
    tail = seq[seq.index(("exchange",)):]
    assert tail == [("exchange",), ("fwd", 0), ("fwd", 7), ("partials",), ("reduce_begin",), ("adj", 0), ("adj", 1),
                    ("adj", 6), ("adj", 7), ("reduce_end",)]
