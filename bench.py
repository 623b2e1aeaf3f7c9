#!/usr/bin/env python
"""bench.py -- JopBlock forward+adjoint mul! throughput (BASELINE.json metric).

Headline workload (config 5 of BASELINE.json): a 256x256 block-tridiagonal JopBlock over a 16 GB
Float32 domain vector (256 blocks x 62.5 MB): diagonal JopLn blocks on the block diagonal (16 GB of
state), stateless stencil blocks on the +-1 block off-diagonals, JopZeroBlock elsewhere (the
reference skips those, src/Jets.jl:1022,1047).  One "step" = d = A*m followed by m' = A'*d.
Algorithmic bytes per apply = domain + range + state = 48 GB, 96 GB per step.

N GPUs: STRONG scaling -- the same operator is partitioned by block row across ranks (one process
per GPU); the forward gathers one halo block from each neighbour, the adjoint sends its partial
halo contributions back and adds them in rank order (NCCL send/recv over NVLink inside
libjets_b200.so).  value = 96 GB / max-over-ranks step time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S] [--no-extra]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NBLK = 256
BLK = 15_625_000          # 62.5 MB of Float32
SEED_W, SEED_M = 5001, 5002


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["bf16_tflops"])
    except Exception:
        return 1590.0


def traffic_from_profiles(key):
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------- CPU arm -----
def cpu_tridiag(blk, steps, warmup, mode):
    """Times the C restatement of the Jets CPU path (oracle/jets_oracle.c) on the same block
    structure with block length `blk`.  Returns (GB/s, ms/step, threads)."""
    import numpy as np
    from oracle import c_oracle as CO
    CO.use_all_cores()      # launchers (torchrun) export OMP_NUM_THREADS=1: the baseline uses every host core regardless
    rng = np.random.default_rng(1)
    base = rng.random(1 << 20, dtype=np.float32)

    def big(n):
        return np.resize(base, n)
    W = big(NBLK * blk)
    leaves = [[("diag", W[r * blk:(r + 1) * blk]) if r == c else ("fdiff", None) if c == r + 1 else
               ("lap", None) if c == r - 1 else ("zero", None) for c in range(NBLK)] for r in range(NBLK)]
    A = CO.BlockOp(leaves, [blk] * NBLK, [blk] * NBLK, np.float32)
    m = big(NBLK * blk)
    d = np.empty(NBLK * blk, dtype=np.float32)
    m2 = np.empty(NBLK * blk, dtype=np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        A.apply(m, False, mode, out=d)
        A.apply(d, True, mode, out=m2)
        t = time.perf_counter() - t0
        if i >= warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    bytes_step = 2 * 3 * NBLK * blk * 4
    return bytes_step / (ms * 1e-3) / 1e9, ms, (CO.num_threads() if mode != 0 else 1)


CPU_SAMPLE_NOTE = ("C restatement of src/Jets.jl:1010-1057 (oracle/jets_oracle.c; Julia is not installed, so this is a port, "
                   "not Jets).  value = the reference's own passes and temporaries (leaf into dtmp, then `_d .+= dtmp`) with "
                   "every elementwise pass spread over all host threads -- Jets itself runs them on ONE thread")


CPU_SAMPLE_BLK = BLK // 8


def cpu_arm(steps, warmup):
    """The three CPU legs on bounded samples of the config-5 structure: the reference's algorithm threaded
    (headline), the same on one thread (what Jets does today), and a fused+threaded rewrite (best CPU)."""
    gbs2, ms2, thr = cpu_tridiag(CPU_SAMPLE_BLK, steps, warmup, 2)
    gbs1, ms1, _ = cpu_tridiag(BLK // 8, max(1, min(steps, 2)), 1, 1)
    gbs0, ms0, _ = cpu_tridiag(BLK // 32, 1, 0, 0)
    cb = {"value": round(gbs2, 3), "unit": "GB/s", "cores": thr, "kind": "port",
          "sample": f"same 256x256 block-tridiagonal structure, block length {BLK // 8} (1/8 of the GPU workload: "
                    f"{NBLK * (BLK // 8) * 4 / 1e9:.2f} GB vectors). " + CPU_SAMPLE_NOTE,
          "single_thread_as_in_jets": {"value": round(gbs0, 3), "unit": "GB/s", "cores": 1,
                                       "sample": f"block length {BLK // 32}, the reference's own passes/temporaries"},
          "fused_openmp_rewrite": {"value": round(gbs1, 3), "unit": "GB/s", "cores": thr,
                                   "sample": f"block length {BLK // 8}; one fused pass per output element (not what the "
                                             "reference does; the strongest CPU version of the path)"}}
    return cb, ms2


def cpu_secondary():
    """The C restatement timed on configs 1 and 2 as well (bounded: config 1 at full size is only 192 MB per
    apply; config 2 on a 1/10 sample), in the three modes of cpu_arm: the reference's passes on ONE thread (what
    Jets does), the same passes threaded, and a fused threaded rewrite.  GB/s of algorithmic bytes."""
    import numpy as np
    from oracle import c_oracle as CO
    out = {"cores": CO.use_all_cores()}
    g = np.random.default_rng(1)

    def best_ms(fn, reps):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return min(ts)
    # config 1: 4x4 diagonal blocks, 1e6 elements, Float64 -- full size
    n = 1_000_000
    W = [[g.random(n) for _ in range(4)] for _ in range(4)]
    A = CO.BlockOp([[("diag", W[r][c]) for c in range(4)] for r in range(4)], [n] * 4, [n] * 4, np.float64)
    m, d = g.random(4 * n), g.random(4 * n)
    o1, o2 = np.empty(4 * n), np.empty(4 * n)
    c1 = {}
    for name, mode, reps in (("single_thread_as_in_jets", 0, 3), ("reference_passes_threaded", 2, 8), ("fused_openmp_rewrite", 1, 8)):
        ms = best_ms(lambda: (A.apply(m, False, mode, o1), A.apply(d, True, mode, o2)), reps)
        c1[name] = {"ms_per_fwd_adj": round(ms, 3), "gbs": round(384e6 / ms / 1e6, 2)}
    out["config1_blockdiag_4x4_1e6_f64"] = c1
    # config 2: diagonal ∘ fdiff ∘ jacobian(x^2), Float32 -- 1e7 elements (1/10 of the GPU workload)
    n = 10_000_000
    w, mo, x = (g.random(n).astype(np.float32) for _ in range(3))
    o = np.empty(n, dtype=np.float32)
    c2 = {"sample": "1e7 elements (1/10 of the GPU workload)"}
    for name, mode, reps in (("single_thread_as_in_jets", 0, 3), ("reference_passes_threaded", 2, 8), ("fused_openmp_rewrite", 1, 8)):
        ms = best_ms(lambda: (CO.chain_apply(w, mo, x, False, mode, o), CO.chain_apply(w, mo, x, True, mode, o)), reps)
        c2[name] = {"ms_per_fwd_adj": round(ms, 3), "gbs": round(2 * 16.0 * n / ms / 1e6, 2)}
    out["config2_chain_f32"] = c2
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ms = cpu_arm(max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": "JopBlock fwd+adj mul! GB/s", "value": cb["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the SAME workload as the GPU arm (structure, eltype, fwd+adj step); each timed step runs a bounded
        # sample of it -- block length / 8 -- whose true size is stated here (GB/s does not depend on it)
        "config": dict(workload_config(args.gpus, 1.0), sample_block_len=CPU_SAMPLE_BLK,
                       sample_algorithmic_bytes_per_step=int(2 * 3 * NBLK * CPU_SAMPLE_BLK * 4),
                       sample_note="ms_per_step is for the 1/8-size sample; value (GB/s) is size-independent"),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line, default=float), flush=True)


def workload_config(n, scale):
    return {"workload": "config5: JopBlock 256x256 block-tridiagonal (diagonal JopLn on the block diagonal, stencil "
                        "blocks on +-1, JopZeroBlock elsewhere), Float32, 16 GB domain vector, fwd+adj mul!",
            "nblocks": NBLK, "block_len": int(BLK * scale), "partition": f"block-row x{n}",
            "algorithmic_bytes_per_step": int(2 * 3 * NBLK * int(BLK * scale) * 4),
            "l2": "inputs (48 GB working set) far exceed the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- GPU arm -----
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.path = f"/tmp/jets_clocks_{os.getpid()}.csv"
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v == "Active":
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def build_c5(B, rank, world, blk):
    """Rank-local rows [r0, r1) of the 256x256 operator as an R_loc x (R_loc+2) JopBlock over the
    halo-extended local domain [lo halo | own blocks | hi halo], handed to the library as ONE distributed
    operator (jets_dist_op_create): every apply below is one jets_dist_apply call = one kernel launch."""
    import numpy as np
    import ctypes as C
    T = np.float32
    D = B.dist
    part = D.RowPartition(NBLK, world, rank, halo=1)
    rl, r0 = part.nloc, part.r0
    sp = B.JetSpace(T, blk)
    own = B.JetBSpace([sp] * rl)
    W = B.zeros(own)
    # counter-based RNG keyed by the GLOBAL element index: every partition draws the same operator
    B.check(B.lib.jets_buf_rand(W._h, SEED_W, C.c_uint64(r0 * blk), 0))
    Sup = B.JopStencil(T, blk, "fdiff")
    Slo = B.JopStencil(T, blk, "lap")
    Z = B.JopZeroBlock(sp, sp)

    def make_block(r, c):
        if r == c:
            return B.JopDiagonal(B.getblock(W, r - r0 + 1))
        return Sup if c == r + 1 else Slo
    A = D.build_local_operator(B, part, make_block, lambda: Z)
    op = D.DistOp(B, A, halo=1)
    x = B.zeros(own)
    B.check(B.lib.jets_buf_rand(x._h, SEED_M, C.c_uint64(r0 * blk), 0))
    if os.environ.get("JETS_BENCH_NO_REGISTER", "0") != "1":   # (tuning: the push path instead)
        op.register(x)  # collective: forward applies on x read the neighbours' halo blocks in place (NVLink, no copy)
    return dict(A=A, op=op, x=x, d=B.zeros(own), m=B.zeros(own), W=W, rl=rl, part=part, own=own)


def sum_over_ranks(B, world, v):
    """Host scalar summed over the ranks in rank order (bit-stable)."""
    import ctypes as C
    if world == 1:
        return float(v)
    r = C.c_double(float(v))
    B.check(B.lib.jets_dist_sum_scalar(C.byref(r)))
    return r.value


def host_link_probe(torch, dist, world, nloc, h_in, h_out, barrier):
    """What the host link delivers for these two pinned buffers with ALL ranks copying at the same time
    (the GPUs of one node share PCIe switches / root ports): each direction alone and both at once, max over
    ranks.  The e2e step moves h2d + d2h bytes, so its floor is the both-directions time."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    dev_a = torch.empty(nloc, dtype=torch.float32, device="cuda")
    dev_b = torch.empty(nloc, dtype=torch.float32, device="cuda")

    def timed(up, down):
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        s1.wait_event(a0)
        s2.wait_event(a0)
        if up:
            with torch.cuda.stream(s1):
                dev_a.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(dev_b, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        a1.record()
        torch.cuda.synchronize()
        t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    timed(True, True)
    t_up, t_dn, t_both = timed(True, False), timed(False, True), timed(True, True)
    del dev_a, dev_b
    gb = nloc * 4 / 1e9
    return {"ranks_copying_at_once": world, "per_rank_gb_each_way": round(gb, 3),
            "h2d_alone_gbs_per_rank": round(gb / t_up * 1e3, 1), "d2h_alone_gbs_per_rank": round(gb / t_dn * 1e3, 1),
            "h2d_alone_gbs_all_ranks": round(world * gb / t_up * 1e3, 1), "d2h_alone_gbs_all_ranks": round(world * gb / t_dn * 1e3, 1),
            "both_directions_ms": round(t_both, 2),
            "what": "cudaMemcpyAsync of the same pinned buffers on every rank simultaneously (max over ranks): each "
                    "direction alone, then both at once"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libjets_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    import jets_b200 as B
    import ctypes as C
    B.init(local)
    stream = torch.cuda.current_stream()
    B.check(B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        B.dist.init_nccl_from_torch(B, dist, torch, rank, world)
    blk = int(BLK * args.scale)
    blk -= blk % 4
    S = build_c5(B, rank, world, blk)
    op, rl, part = S["op"], S["rl"], S["part"]
    lib = B.lib

    def fwd():
        op.forward(S["d"], S["x"])          # ONE library call = one kernel launch (halo exchange inside)

    def adj():
        op.adjoint(S["m"], S["d"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        fwd()
        adj()
    barrier()
    # parity at full size through size-independent properties: <A m, y> == <m, A' y> over all ranks
    def ddot(x, y):  # f64 dot over all ranks, partials summed in rank order (bit-stable)
        r = C.c_double()
        B.check(lib.jets_dot(x._h, y._h, C.byref(r)))
        return sum_over_ranks(B, world, r.value)

    y = B.zeros(S["own"])
    B.check(lib.jets_buf_rand(y._h, 77, C.c_uint64(part.r0 * blk), 0))
    tmp = B.zeros(S["own"])
    fwd()
    lhs = ddot(S["d"], y)
    op.adjoint(tmp, y)
    rhs = ddot(S["x"], tmp)
    dpt = abs(lhs - rhs) / abs(lhs + rhs)
    adj()
    # partition-independent checksums of the two results (identical for every --gpus N: the distributed
    # row sums have the single-GPU order, the dots are summed in rank order)
    checksum = {"dot(A*m, y)": lhs, "|A'*A*m|^2": ddot(S["m"], S["m"])}
    del y, tmp
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    l0 = B.launch_count()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        fwd()
        ev[k][1].record(stream)
        adj()
        ev[k][2].record(stream)
    t_end.record(stream)
    barrier()
    launches = B.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    total_ms = t_start.elapsed_time(t_end)
    # per-launch duration of the dominant kernel: CUDA events around each one-launch apply, on the launch stream
    k_fwd = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    k_adj = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    tm = torch.tensor([total_ms, k_fwd, k_adj], dtype=torch.float64, device="cuda")
    ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
    total_ms, k_fwd, k_adj = tm.tolist()
    ms_step = total_ms / args.steps
    bytes_apply = 3 * NBLK * blk * 4
    value = 2 * bytes_apply / (ms_step * 1e-3) / 1e9
    gate_timeouts = op.gate_timeouts
    engine = {"launches_per_apply": op.info(4), "neighbours": op.info(2), "gate_timeouts": gate_timeouts}

    if os.environ.get("JETS_B200_TRACE") == "1":   # tuning: per-CTA timelines of the last 64 launches of every rank
        tb = np.zeros(64 * 160 * 8, dtype=np.uint64)
        nt = lib.jets_debug_trace(tb.ctypes.data_as(C.c_void_p), tb.size)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", f"trace_n{world}_rank{rank}.npy"), np.concatenate([[np.uint64(nt)], tb]))
    if args.no_e2e:     # tuning runs only: a line without e2e is not a bench result
        if rank == 0:
            print(json.dumps({"tuning_only": True, "n_gpus": world, "value": round(value, 2), "launch_ms": [round(k_fwd, 4), round(k_adj, 4)],
                              "frac": round((bytes_apply / world) / ((k_fwd + k_adj) / 2 * 1e-3) / 1e9 / peaks()[0], 4),
                              "dpt": dpt, "checksum": checksum, "engine": engine}), flush=True)
        S["op"].close()
        if world > 1:
            B.check(lib.jets_dist_shutdown())
            dist.destroy_process_group()
        return
    # ---- end to end through the C ABI with HOST buffers (pinned): jets_dist_apply_normal_host moves this rank's
    # shard H2D, applies A then A', and moves the result D2H, chunk-pipelined inside the library
    nloc = rl * blk
    gen = torch.Generator().manual_seed(1234 + rank)
    h_in = torch.empty(nloc, dtype=torch.float32, pin_memory=True)
    h_out = torch.empty(nloc, dtype=torch.float32, pin_memory=True)
    h_in.uniform_(0, 1, generator=gen)
    e2e_steps = max(1, min(args.steps, 4))
    nchunks = int(os.environ.get("JETS_BENCH_E2E_CHUNKS", "32"))
    # (1) the pipelined call must equal upload -> jets_dist_apply -> jets_dist_apply -> download, bit for bit
    h_out.zero_()
    op.normal_host(h_out.data_ptr(), h_in.data_ptr(), nchunks)
    op.join()
    barrier()
    B.check(lib.jets_buf_upload(S["x"]._h, -1, C.c_void_p(h_in.data_ptr()), nloc))
    fwd()
    adj()
    B.check(lib.jets_buf_upload(S["d"]._h, -1, C.c_void_p(h_out.data_ptr()), nloc))    # d <- pipelined result
    B.lincomb_(S["d"], [(1.0, S["d"]), (-1.0, S["m"])])
    mn, mx = B.extrema(S["d"])
    same = torch.tensor([1 if (float(mn) == 0.0 and float(mx) == 0.0) else 0], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    e2e_same = bool(same.item())
    e2e_check = sum_over_ranks(B, world, float(h_out[:1000].double().sum().item()))
    npipe = op.info(3)
    # the device-resident vectors are not needed any more: the pipeline owns its work vectors
    S["x"] = S["d"] = S["m"] = None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        op.normal_host(h_out.data_ptr(), h_in.data_ptr(), nchunks)
    op.join()
    e1.record(stream)
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = te.item() / e2e_steps
    e2e_val = 2 * bytes_apply / (e2e_ms * 1e-3) / 1e9
    e2e_what = ("pinned host m -> H2D -> d=A*m -> m'=A'*d -> D2H m' in ONE C-ABI call per step (jets_dist_apply_normal_host): "
                f"{npipe} block-row chunks pipelined on 3 streams inside libjets_b200 (csrc/dist_op.cu); consecutive steps overlap")
    try:
        link = host_link_probe(torch, dist, world, nloc, h_in, h_out, barrier)
        link["e2e_over_link_floor"] = round(link["both_directions_ms"] / e2e_ms, 3)
    except Exception as ex:  # never let the probe break the bench line
        link = {"error": str(ex)}
    del h_in, h_out

    peak, peak_src = peaks()
    k_ms = (k_fwd + k_adj) / 2
    achieved = (bytes_apply / world) / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "JopBlock fwd+adj mul! GB/s", "value": round(value, 2), "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, args.scale),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic_from_profiles("c5_fused_bundle_bytes_per_launch"),
                     "kernel": "jets_fused_bundle_kernel<float,16,2>", "peak_source": peak_src,
                     "launch_ms": {"forward": round(k_fwd, 4), "adjoint": round(k_adj, 4)},
                     "algorithmic_bytes_per_launch": bytes_apply // world},
        "e2e": {"value": round(e2e_val, 2), "unit": "GB/s", "h2d_bytes_per_step": NBLK * blk * 4,
                "d2h_bytes_per_step": NBLK * blk * 4, "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
                "what": e2e_what, "result_probe_sum_first_1000_per_rank": e2e_check, "host_link": link,
                "pipelined_equals_sequential": e2e_same},
        "gpu_launches": int(ln.item()),
        "clocks": clocks,
        "parity": {"dot_product_test_rel": dpt, "tolerance": 1e-5, "checksum": checksum},
        "engine": engine,
    }
    if gate_timeouts:
        line["invalid"] = f"{gate_timeouts} work units gave up waiting for a neighbouring rank"
    if not e2e_same:
        line["e2e"]["invalid"] = "the pipelined host-buffer result differs from upload -> apply -> apply -> download"
    S["op"].close()
    del S, op
    import gc
    gc.collect()
    # Calibration on THIS box: the boxes of the pool differ by several per cent for one binary (clocks, HBM stacks), and
    # `peak` was measured by the driver on another one -- so the line also carries what a plain device-to-device copy
    # of 4 GB (8 GB of traffic per copy, far beyond L2) achieves here, right now, timed the same way.
    try:
        n_cal = 1 << 30
        ca, cb = torch.empty(n_cal, dtype=torch.float32, device="cuda"), torch.empty(n_cal, dtype=torch.float32, device="cuda")
        ca.uniform_()
        ms_cal = time_steps(torch, stream, lambda: cb.copy_(ca), 10, 3)
        same_box = 2 * n_cal * 4 / (ms_cal * 1e-3) / 1e9
        line["roofline"]["same_box_copy_gbps"] = round(same_box, 1)
        line["roofline"]["frac_of_same_box_copy"] = round(achieved / same_box, 4)
        del ca, cb
    except Exception as ex:  # a calibration must never cost the bench line
        line["roofline"]["same_box_copy_gbps"] = None
        line["roofline"]["same_box_copy_error"] = str(ex)
    if world > 1:
        try:
            line["dense_structure"] = dense_structure_workload(B, torch, dist, stream, rank, world, peak, barrier)
        except Exception as e:  # never lose the headline line to a secondary workload
            line["dense_structure"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"], _ = cpu_arm(2, 1)
        try:
            line["cpu_baseline"]["other_configs"] = cpu_secondary()
        except Exception as ex:  # a diagnostic leg must never cost the bench line
            line["cpu_baseline"]["other_configs"] = {"error": str(ex)}
    if rank == 0 and world == 1 and not args.no_extra:
        try:
            line["other_workloads"] = extra_workloads(B, torch, stream, peak)
            ow = line["other_workloads"]
            line["roofline"]["others"] = {   # compact copy of the secondary configs' fractions of the HBM peak
                "c1": ow["config1_blockdiag_4x4_1e6_f64"].get("frac_of_hbm_peak"),
                "c1_synchronised_each_step": ow["config1_blockdiag_4x4_1e6_f64"].get("frac_of_hbm_peak_synchronised_each_step"),
                "c2": ow["config2_chain_1e8_f32"].get("frac_of_hbm_peak"),
                "c3a": ow["config3a_dense_64x64_2048_f32_gemv"].get("frac_of_hbm_peak"),
                "c3b": ow["config3b_dense_64x64_2048_f32_64rhs_tcgen05"].get("frac_of_hbm_peak"),
                "c4": ow["config4_lsqr_200it_blockdiag_plus_sum_f64"].get("frac_of_per_primitive_roofline"),
                "c4_fused_of_its_own_roofline": ow["config4_lsqr_200it_fused_updates"].get("frac_of_fused_roofline")}
        except Exception as e:  # never lose the headline line to a secondary workload
            line["other_workloads"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line, default=float), flush=True)
    if world > 1:
        B.check(lib.jets_dist_shutdown())
        dist.destroy_process_group()


def dense_structure_workload(B, torch, dist, stream, rank, world, peak, barrier):
    """north_star's dense-structure multi-GPU path at reduced size: a 32x32 JopBlock of dense 2048x2048 Float32
    blocks (16 GiB of matrices, config 3 quartered) partitioned by block row; the forward all-gathers the domain
    shards (NCCL over NVLink) before the local GEMV rows, the adjoint reduce-scatters the per-rank partial sums
    (src/Jets.jl:1015-1055 with no zero blocks).  The exchange itself is timed alone as well."""
    import numpy as np
    import ctypes as C
    T = np.float32
    nb, k = 32, 2048
    nloc = nb // world
    Asp = B.JetSpace(T, k, k)
    mats = B.zeros(B.JetBSpace([Asp] * (nloc * nb)))
    # element index of block (r, c) in the global column-major block table, so every partition draws the same matrices
    for i in range(nloc):
        for c in range(nb):
            blkv = B.getblock(mats, 1 + i * nb + c)
            B.check(B.lib.jets_buf_rand(blkv._h, 3001, C.c_uint64(((rank * nloc + i) * nb + c) * k * k), 0))
    A_loc = B.blockop([[B.JopDense(B.getblock(mats, 1 + i * nb + c)) for c in range(nb)] for i in range(nloc)])
    op = B.dist.DistOp(B, A_loc, dense=True)
    shard = B.JetSpace(T, nb * k // world)
    x = B.zeros(shard)
    B.check(B.lib.jets_buf_rand(x._h, 3002, C.c_uint64(rank * (nb * k // world)), 0))
    d, m = B.zeros(B.range_(A_loc)), B.zeros(shard)
    y = B.zeros(B.range_(A_loc))
    B.check(B.lib.jets_buf_rand(y._h, 3003, C.c_uint64(rank * nloc * k), 0))

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    ms_f = timed(lambda: op.forward(d, x))
    ms_t = timed(lambda: op.adjoint(m, d))
    full = B.zeros(B.domain(A_loc))
    ms_ag = timed(lambda: B.check(B.lib.jets_dist_allgather(x._h, full._h)), 20)
    ms_rs = timed(lambda: B.check(B.lib.jets_dist_reduce_scatter(full._h, m._h)), 20)
    # dot-product identity over all ranks: <A x, y> == <x, A' y>
    op.forward(d, x)
    op.adjoint(m, y)
    lhs = sum_over_ranks(B, world, float(B.dot(d, y)))
    rhs = sum_over_ranks(B, world, float(B.dot(x, m)))
    by = nloc * nb * k * k * 4          # matrix bytes per rank per apply
    link_bytes = (world - 1) * (nb * k // world) * 4   # received (forward) / sent (adjoint) per rank
    op.close()
    return {"workload": f"{nb}x{nb} dense {k}x{k} Float32 blocks, block rows over {world} GPUs (config 3 at 1/4 size), single-vector GEMV",
            "forward_ms": round(ms_f, 4), "adjoint_ms": round(ms_t, 4),
            "per_gpu_gbs": round(by / ((ms_f + ms_t) / 2) / 1e6, 1), "frac_of_hbm_peak_per_gpu": round(by / ((ms_f + ms_t) / 2) / 1e6 / peak, 4),
            "aggregate_gbs": round(world * by / ((ms_f + ms_t) / 2) / 1e6, 1),
            "allgather_alone_ms": round(ms_ag, 4), "reduce_scatter_alone_ms": round(ms_rs, 4),
            "nvlink_bytes_per_rank_per_apply": link_bytes, "nvlink_floor_us_at_770GBs": round(link_bytes / 770e9 * 1e6, 3),
            "exchange_is": "latency bound: the shards are KB-sized, the NVLink floor is far below NCCL's launch latency",
            "dot_product_test_rel": abs(lhs - rhs) / max(abs(lhs), abs(rhs)), "tolerance": 1e-5}


def time_steps(torch, stream, fn, steps, warmup, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / steps


def extra_workloads(B, torch, stream, peak):
    """Secondary BASELINE.json configs on one GPU (kernel-only, inputs resident, L2 flushed between
    iterations when the working set is not >> L2)."""
    import numpy as np
    out = {}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    # config 1: 4x4 diagonal blocks, 1e6 elements each, Float64.  192 MB per apply is close to the L2
    # size, so NSETS independent (operator, vectors) sets are cycled: the working set (1.15 GB) is far
    # larger than L2 and no flush kernel leaves dirty lines behind for the timed kernel to write back.
    n, NSETS = 1_000_000, 6
    sp = B.JetSpace(np.float64, n)
    sets = []
    for i in range(NSETS):
        W = B.rand(B.JetBSpace([sp] * 16), seed=1001 + 10 * i)
        A = B.blockop([[B.JopDiagonal(B.getblock(W, 1 + r + 4 * c)) for c in range(4)] for r in range(4)])
        sets.append((A, B.adjoint(A), B.rand(B.domain(A), seed=1002 + 10 * i), B.zeros(B.range_(A)), B.zeros(B.domain(A)), W))
    cnt = [0]

    def step1():
        A_, At_, m_, d_, m2_, _ = sets[cnt[0] % NSETS]
        cnt[0] += 1
        B.mul_(d_, A_, m_)
        B.mul_(m2_, At_, d_)
    ms = time_steps(torch, stream, step1, 4 * NSETS, 2 * NSETS)
    # the same steps issued back to back (one pair of events around 8 rounds over the sets, as a solver loop issues
    # them): launches overlap their ramps through programmatic dependent launch, which a sync per step forbids
    torch.cuda.synchronize()
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record(stream)
    for _ in range(8 * NSETS):
        step1()
    eb1.record(stream)
    torch.cuda.synchronize()
    ms_b2b = eb0.elapsed_time(eb1) / (8 * NSETS)
    A, At, m, d, m2, W = sets[0]

    def step1f():
        B.mul_(d, A, m)
        B.mul_(m2, At, d)
    ms_flush = time_steps(torch, stream, step1f, 20, 3, flush)
    lhs, rhs = B.dot_product_test(A, m, B.rand(B.range_(A), seed=1003))
    # Primary figure: the K steps between ONE pair of events (the bench contract's timing, and how a solver loop issues
    # the applies); the per-step-synchronised figure is kept beside it (it adds one launch latency per step and forbids
    # the overlap of one launch's tail with the next one's prologue).
    ms_sync = ms
    ms = ms_b2b
    gbs = 2 * 192e6 / (ms * 1e-3) / 1e9
    # calibration: what a plain device copy of the SAME traffic (96 MB read + 96 MB write per launch, two
    # launches per step, the same rotation over 6 buffer pairs) takes -- at 40 us per launch the ramp and
    # tail of ANY kernel are a visible part of the time, which the long-copy peak does not include
    cp_src = [torch.empty(96_000_000 // 8, dtype=torch.float64, device="cuda").normal_() for _ in range(NSETS)]
    cp_dst = [torch.empty(96_000_000 // 8, dtype=torch.float64, device="cuda") for _ in range(NSETS)]
    ccnt = [0]

    def step1c():
        i = ccnt[0] % NSETS
        ccnt[0] += 1
        cp_dst[i].copy_(cp_src[i])
        cp_src[i].copy_(cp_dst[(i + 1) % NSETS])
    ms_copy = time_steps(torch, stream, step1c, 4 * NSETS, 2 * NSETS)
    del cp_src, cp_dst
    out["config1_blockdiag_4x4_1e6_f64"] = {"ms_per_step": round(ms, 4), "value": round(gbs, 1), "unit": "GB/s",
                                             "frac_of_hbm_peak": round(gbs / peak, 4), "algorithmic_bytes_per_step": 384_000_000,
                                             "dot_product_test_rel": abs(lhs - rhs) / abs(lhs + rhs), "engine": B.plan_info(A),
                                             "l2": f"{NSETS} independent operator/vector sets cycled ({NSETS * 192} MB working set >> 126 MB L2)",
                                             "timing": "48 steps (8 rounds over the 6 sets) between one pair of CUDA events",
                                             "ms_per_step_synchronised_each_step": round(ms_sync, 4),
                                             "frac_of_hbm_peak_synchronised_each_step": round(2 * 192e6 / (ms_sync * 1e-3) / 1e9 / peak, 4),
                                             "ms_per_step_single_set_after_256MB_write_flush": round(ms_flush, 4),
                                             "size_matched_device_copy_ms_per_step": round(ms_copy, 4),
                                             "frac_of_size_matched_copy": round(ms_copy / ms_sync, 4)}   # both synchronised per step
    del A, At, W, m, d, m2, sets
    # config 2: diagonal ∘ fdiff ∘ jacobian(pointwise square), 1e8 elements, Float32
    n = 100_000_000
    T = np.float32
    sp = B.JetSpace(T, n)
    w, mo = B.rand(sp, seed=2001), B.rand(sp, seed=2002)
    G = B.JopDiagonal(w) @ B.JopStencil(T, n, "fdiff") @ B.JopPointwise(T, n, "square")
    Jc = B.jacobian(G, mo)
    dm, dd = B.rand(sp, seed=2003), B.zeros(sp)
    dm2 = B.zeros(sp)
    Jt = B.adjoint(Jc)

    def step2():
        B.mul_(dd, Jc, dm)
        B.mul_(dm2, Jt, dd)
    ms = time_steps(torch, stream, step2, 20, 3)
    lhs, rhs = B.dot_product_test(Jc, dm, B.rand(sp, seed=2004))
    gbs = 2 * 1.6e9 / (ms * 1e-3) / 1e9
    out["config2_chain_1e8_f32"] = {"ms_per_step": round(ms, 4), "value": round(gbs, 1), "unit": "GB/s",
                                    "frac_of_hbm_peak": round(gbs / peak, 4), "algorithmic_bytes_per_step": 3_200_000_000,
                                    "dot_product_test_rel": abs(lhs - rhs) / abs(lhs + rhs), "engine": B.plan_info(Jc),
                                    "l2": "1.6 GB working set >> L2"}
    del G, Jc, Jt, w, mo, dm, dd, dm2
    # config 4: LSQR-style loop, 200 iterations, A = B - 0.5*S (8x8 block-diagonal of 2^20-element Float64
    # diagonals minus block-diagonal stencils); device-resident scalars, one CUDA graph per iteration.
    # Algorithmic bytes per iteration counted per primitive: 24 N w = 1.611 GB.
    nb, n4 = 8, 1 << 20
    T8 = np.float64
    sp = B.JetSpace(T8, n4)
    W4 = B.rand(B.JetBSpace([sp] * nb), seed=4001)
    Bd = B.blockop([[B.JopDiagonal(B.getblock(W4, i + 1)) if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
    Sd = B.blockop([[B.JopStencil(T8, n4, "lap") if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
    A4 = Bd - 0.5 * Sd
    rhs4 = B.rand(B.range_(A4), seed=4002)
    iters = 200
    import ctypes as C
    torch.cuda.synchronize()
    s4 = torch.cuda.Stream()               # stream capture is not possible on the legacy default stream
    B.check(B.lib.jets_stream_set(C.c_void_p(s4.cuda_stream)))
    try:
        G = B.solvers.LsqrGraph(A4, rhs4)      # start-up + iteration 1 + graph capture (untimed)
        G.run(5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s4)
        G.run(iters)
        e1.record(s4)
        torch.cuda.synchronize()
        ms4 = e0.elapsed_time(e1)
        x4, (a4, b4) = G.result()
        nx4 = float(B.norm(x4))
    finally:
        torch.cuda.synchronize()
        B.check(B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))
    by_iter = 24 * nb * n4 * 8
    out["config4_lsqr_200it_blockdiag_plus_sum_f64"] = {
        "total_ms": round(ms4, 3), "us_per_iteration": round(1e3 * ms4 / iters, 2), "value": round(by_iter * iters / ms4 / 1e6, 1),
        "unit": "GB/s", "frac_of_per_primitive_roofline": round(by_iter * iters / ms4 / 1e6 / peak, 4),
        "algorithmic_bytes_per_iteration": by_iter, "iterations": iters,
        "what": "A*v, A'*u, 2 norms, 4 axpby-class updates per iteration; scalars on device; one CUDA graph per iteration",
        "final_alpha_beta": [a4, b4], "norm_x": nx4, "engine": B.plan_info(A4)}
    # the same loop with the updates folded into the applies (jets_apply_axpby, LsqrGraphFused): 16 instead of
    # 24 vector passes per iteration, so the fraction of the PER-PRIMITIVE roofline may exceed 1
    torch.cuda.synchronize()
    B.check(B.lib.jets_stream_set(C.c_void_p(s4.cuda_stream)))
    try:
        Gf = B.solvers.LsqrGraphFused(A4, rhs4)
        Gf.run(5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s4)
        Gf.run(iters)
        e1.record(s4)
        torch.cuda.synchronize()
        msf = e0.elapsed_time(e1)
        xf, (af, bf) = Gf.result()
        nxf = float(B.norm(xf))
    finally:
        torch.cuda.synchronize()
        B.check(B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))
    out["config4_lsqr_200it_fused_updates"] = {
        "total_ms": round(msf, 3), "us_per_iteration": round(1e3 * msf / iters, 2), "value": round(by_iter * iters / msf / 1e6, 1),
        "unit": "GB/s (per-primitive algorithmic bytes / time)", "frac_of_per_primitive_roofline": round(by_iter * iters / msf / 1e6 / peak, 4),
        "bytes_actually_required_per_iteration": 15 * nb * n4 * 8,
        "bytes_note": "2 fused applies (4 N w each) + 2 norms (1 each, folded into the applies' store epilogue) + the x / w updates in "
                      "one pass (5 instead of 6): 15 N w of DRAM traffic; frac_of_fused_roofline keeps round 1's 16 N w numerator",
        "frac_of_fused_roofline": round(16 * nb * n4 * 8 * iters / msf / 1e6 / peak, 4),
        "what": "u, v unnormalised; u = A v/alpha - (alpha/beta) u and v = A'u/beta - (beta/alpha) v each ONE fused apply whose store "
                "epilogue also yields the norm; x and w updated in one pass; two scalar programs; > 1.0 of the per-primitive roofline comes from fusion",
        "final_alpha_beta": [af, bf], "norm_x": nxf, "rel_diff_norm_x_vs_unfused": abs(nxf - nx4) / nx4}
    del Gf, xf
    del G, A4, Bd, Sd, W4, rhs4, x4
    # config 3a: 64x64 dense 2048x2048 Float32 blocks (64 GiB of matrices generated on device), GEMV
    try:
        nb, k = 64, 2048
        Asp = B.JetSpace(T, k, k)
        mats = B.zeros(B.JetBSpace([Asp] * (nb * nb)))
        import ctypes as C
        B.check(B.lib.jets_buf_rand(mats._h, 3001, 0, 0))
        blocks = [[B.JopDense(B.getblock(mats, 1 + r + nb * c)) for c in range(nb)] for r in range(nb)]
        A = B.blockop(blocks)
        m, d = B.rand(B.domain(A), seed=3002), B.zeros(B.range_(A))
        m2 = B.zeros(B.domain(A))
        At = B.adjoint(A)
        ms_f = time_steps(torch, stream, lambda: B.mul_(d, A, m), 3, 1)
        ms_t = time_steps(torch, stream, lambda: B.mul_(m2, At, d), 3, 1)
        y = B.rand(B.range_(A), seed=3003)
        lhs, rhs = B.dot_product_test(A, m, y)
        by = nb * nb * k * k * 4
        out["config3a_dense_64x64_2048_f32_gemv"] = {
            "forward_ms": round(ms_f, 3), "adjoint_ms": round(ms_t, 3),
            "forward_gbs": round(by / ms_f / 1e6, 1), "adjoint_gbs": round(by / ms_t / 1e6, 1),
            "frac_of_hbm_peak": round(by / ((ms_f + ms_t) / 2) / 1e6 / peak, 4),
            "dot_product_test_rel": abs(lhs - rhs) / abs(lhs + rhs), "engine": B.plan_info(A)}
        del A, At, blocks, m, d, m2, y
        # config 3b: the same 64 GiB of matrices applied to 64 right-hand sides on the tcgen05/TMEM path
        # (split-TF32: 3 tensor-core products per element, 2.2 TFLOP algorithmic, 6.6 TFLOP issued)
        nrhs = 64
        blocks = [[B.JopDense(B.getblock(mats, 1 + r + nb * c), nrhs=nrhs) for c in range(nb)] for r in range(nb)]
        A = B.blockop(blocks)
        At = B.adjoint(A)
        m, d = B.rand(B.domain(A), seed=3004), B.zeros(B.range_(A))
        m2 = B.zeros(B.domain(A))
        ms_f = time_steps(torch, stream, lambda: B.mul_(d, A, m), 3, 1)
        ms_t = time_steps(torch, stream, lambda: B.mul_(m2, At, d), 3, 1)
        y = B.rand(B.range_(A), seed=3005)
        lhs, rhs = B.dot_product_test(A, m, y)
        by = nb * nb * k * k * 4 + 2 * nb * k * nrhs * 4
        fl = 2.0 * nb * nb * k * k * nrhs
        tf_peak = tensor_peak()
        out["config3b_dense_64x64_2048_f32_64rhs_tcgen05"] = {
            "forward_ms": round(ms_f, 3), "adjoint_ms": round(ms_t, 3),
            "forward_gbs": round(by / ms_f / 1e6, 1), "adjoint_gbs": round(by / ms_t / 1e6, 1),
            "frac_of_hbm_peak": round(by / ((ms_f + ms_t) / 2) / 1e6 / peak, 4),
            "algorithmic_tflops": round(fl / ((ms_f + ms_t) / 2) / 1e9, 1),
            # issued work in tf32-equivalent MMA time: the main term on kind::tf32 (1x) plus the two correction
            # terms on kind::f16/bf16 at twice the tf32 rate (2 x 0.5) = 2x the algorithmic flops
            # (JETS_B200_TC_MIXED=0 issues all three terms on tf32 = 3x)
            "issued_tf32_equiv_tflops": round(2 * fl / ((ms_f + ms_t) / 2) / 1e9, 1),
            "frac_of_tensor_peak_bf16_over_2": round(2 * fl / ((ms_f + ms_t) / 2) / 1e9 / (tf_peak / 2), 4),
            "tensor_peak_note": f"denominator = measured bf16 burst {tf_peak} TF/s / 2 (TF32 runs at half the bf16 rate); HBM is the binding roofline",
            "dot_product_test_rel": abs(lhs - rhs) / max(abs(lhs), abs(rhs)), "engine": B.plan_info(A)}
        del A, At, mats, blocks, m, d, m2, y
    except B.JetsError as e:  # e.g. not enough free memory on a shared box
        out.setdefault("config3a_dense_64x64_2048_f32_gemv", {"error": str(e)})
        out.setdefault("config3b_dense_64x64_2048_f32_64rhs_tcgen05", {"error": str(e)})
    try:
        out["reference_suite_shapes_latency"] = small_op_latency(B, torch, stream)
    except Exception as e:  # a diagnostic section must never cost the bench line
        out["reference_suite_shapes_latency"] = {"error": str(e)}
    return out


def small_op_latency(B, torch, stream):
    """The shapes of the reference's own PkgBenchmark suite (benchmark/benchmarks.jl:39-157: n = 100
    elements, a 4-stage composition, a 2x3 block operator) through the device path.  At this size a call is
    pure launch latency -- microseconds per call, where the CPU reference spends a fraction of one -- so this
    section documents the fixed cost per call of the device path; it is not a throughput claim."""
    import numpy as np
    n = 100
    g = np.random.default_rng(7)
    T = np.float64

    def per_call_us(fn, reps=2000):
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        t_host = (time.perf_counter() - t0) / reps * 1e6
        torch.cuda.synchronize()
        return {"device_us": round(a.elapsed_time(b) / reps * 1e3, 2), "host_issue_us": round(t_host, 2)}
    A = B.JopDiagonal(g.random(n))                       # JopFoo, benchmarks.jl:36-47
    F = B.JopPointwise(T, n, "square")                   # JopBar, :49-67
    m, d = B.rand(B.domain(A), seed=1), B.zeros(B.range_(A))
    out = {"what": "per-call cost at the reference suite's n=100 shapes (benchmark/benchmarks.jl); launch-latency bound"}
    out["JopLn mul!"] = per_call_us(lambda: B.mul_(d, A, m))
    At = A.T
    out["JopLn mul! adjoint"] = per_call_us(lambda: B.mul_(m, At, d))
    out["JopNl mul!"] = per_call_us(lambda: B.mul_(d, F, m))
    G = F @ A @ F @ A                                     # Composition, :69-82
    out["Composition mul!"] = per_call_us(lambda: B.mul_(d, G, m))
    J = B.jacobian(G, m)
    Jt = J.T
    out["Composition jacobian mul! adjoint"] = per_call_us(lambda: B.mul_(m, Jt, d))
    Fb = B.blockop([[B.JopPointwise(T, n, "square") for _ in range(3)] for _ in range(2)])   # Block, homogeneous, :84-124
    mb, db, eb = B.rand(B.domain(Fb), seed=2), B.zeros(B.range_(Fb)), B.rand(B.range_(Fb), seed=3)
    out["Block 2x3 mul!"] = per_call_us(lambda: B.mul_(db, Fb, mb))
    Jb = B.jacobian(Fb, mb)
    Jbt = Jb.T
    out["Block 2x3 jacobian mul! adjoint"] = per_call_us(lambda: B.mul_(mb, Jbt, db))
    fb = B.zeros(B.range_(Fb))
    out["BlockArray broadcast f .= d .+ e"] = per_call_us(lambda: B.lincomb_(fb, [(1.0, db), (1.0, eb)]))
    out["BlockArray dot (host result)"] = per_call_us(lambda: B.dot(db, eb), reps=500)
    out["BlockArray norm (host result)"] = per_call_us(lambda: B.norm(db), reps=500)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the block length (debugging only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs: device-resident timing only (prints a reduced line)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
