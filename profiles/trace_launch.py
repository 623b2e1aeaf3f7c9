"""Per-CTA timeline of back-to-back fused block-apply launches (JETS_B200_TRACE=1): where the microseconds of a short
launch go.  Config 1 (4x4 diagonal blocks, 1e6 Float64) forward/adjoint alternating over 6 operator sets, as bench.py
times it.  Output: per launch, medians over the CTAs of each phase, all relative to the launch's first CTA entry."""
import os, sys
os.environ["JETS_B200_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
import jets_b200 as B
B.init(0)
s = torch.cuda.current_stream()
B.check(B.lib.jets_stream_set(C.c_void_p(s.cuda_stream)))
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
if which == "c1":
    n, NS = 1_000_000, 6
    sp = B.JetSpace(np.float64, n)
    sets = []
    for i in range(NS):
        W = B.rand(B.JetBSpace([sp] * 16), seed=1001 + 10 * i)
        A = B.blockop([[B.JopDiagonal(B.getblock(W, 1 + r + 4 * c)) for c in range(4)] for r in range(4)])
        sets.append((A, B.adjoint(A), B.rand(B.domain(A), seed=2 + i), B.zeros(B.range_(A)), B.zeros(B.domain(A)), W))
    bytes_launch = 192e6
else:
    nb, n4, NS = 8, 1 << 20, 4
    sp = B.JetSpace(np.float64, n4)
    sets = []
    for i in range(NS):
        W = B.rand(B.JetBSpace([sp] * nb), seed=4001 + i)
        Bd = B.blockop([[B.JopDiagonal(B.getblock(W, a + 1)) if a == b else B.JopZeroBlock(sp, sp) for b in range(nb)] for a in range(nb)])
        Sd = B.blockop([[B.JopStencil(np.float64, n4, "lap") if a == b else B.JopZeroBlock(sp, sp) for b in range(nb)] for a in range(nb)])
        A = Bd - 0.5 * Sd
        sets.append((A, B.adjoint(A), B.rand(B.domain(A), seed=5), B.zeros(B.range_(A)), B.zeros(B.domain(A)), W))
    bytes_launch = 3 * nb * n4 * 8
cnt = [0]
def step():
    A_, At_, m_, d_, m2_, _ = sets[cnt[0] % NS]; cnt[0] += 1
    B.mul_(d_, A_, m_); B.mul_(m2_, At_, d_)
for _ in range(12): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(s)
for _ in range(24): step()
b.record(s); torch.cuda.synchronize()
pair_us = a.elapsed_time(b) / 24 * 1e3
L, Cn = 64, 160
buf = np.zeros(L * Cn * 8, dtype=np.uint64)
total = B.lib.jets_debug_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
T = buf.reshape(L, Cn, 8).astype(np.int64)
print(f"# {which}: {pair_us:.2f} us per fwd+adj pair (CUDA events, tracing on); {bytes_launch / 1e6:.0f} MB per launch = "
      f"{bytes_launch / 6457.4e3:.1f} us at the 6457 GB/s copy peak; {total} launches traced, the last 48 in launch order:")
print("# columns (us, median over CTAs unless noted): launch | gap from the previous launch's last exit to this launch's first entry | entry spread (last CTA entry) |"
      " barriers ready | past griddepcontrol.wait | first tile landed | first row stored | last group issued | end sentinel (min .. max) | state groups in flight before the wait")
order = [(total - 48 + i) % L for i in range(48)]
prev_end = None
rows = []
for li in order:
    t = T[li]
    live = t[:, 0] > 0
    t = t[live]
    t0 = t[:, 0].min()
    rel = lambda k: (t[:, k] - t0) / 1e3
    gap = (t0 - prev_end) / 1e3 if prev_end is not None else float("nan")
    prev_end = t[:, 6].max()
    rows.append((gap, rel(0).max(), np.median(rel(1)), np.median(rel(2)), np.median(rel(3)), np.median(rel(4)), np.median(rel(5)), rel(6).min(), rel(6).max(), np.median(t[:, 7])))
    print(f"{li:3d} | {gap:6.2f} | {rel(0).max():5.2f} | {np.median(rel(1)):5.2f} | {np.median(rel(2)):5.2f} | {np.median(rel(3)):5.2f} | {np.median(rel(4)):5.2f} | "
          f"{np.median(rel(5)):6.2f} | {rel(6).min():6.2f} .. {rel(6).max():6.2f} | {np.median(t[:, 7]):.0f}")
R = np.array(rows[1:])
print("# mean over launches: gap %.2f | entry spread %.2f | barriers %.2f | past wait %.2f | first tile %.2f | first store %.2f | last issue %.2f | end %.2f .. %.2f"
      % tuple(np.nanmean(R[:, :9], axis=0)))
print("# => per launch: first entry -> last exit %.2f us, + gap to the next launch %.2f us = %.2f us per launch (%.2f us per pair)"
      % (np.nanmean(R[:, 8]), np.nanmean(R[:, 0]), np.nanmean(R[:, 8]) + np.nanmean(R[:, 0]), 2 * (np.nanmean(R[:, 8]) + np.nanmean(R[:, 0]))))
