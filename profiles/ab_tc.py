"""A/B driver for the tcgen05 multi-RHS path (config 3b at 1/16 size): accuracy against a Float64
host product on a 4x4 grid of 2048x2048 blocks, then timing on the 16x16 grid (4 GiB of matrices).
Run once per JETS_B200_TC_MIXED setting (the switch is read once per process)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jets_b200 as B

T, k, nrhs = np.float32, 2048, 64
SKIP = os.environ.get("AB_SKIP_ACC") == "1"
tag = "ring=" + os.environ.get("JETS_B200_TC_RING", "0") + " mixed=" + os.environ.get("JETS_B200_TC_MIXED", "default") + " expt=" + os.environ.get("JETS_B200_TC_EXPT", "0")


def build(nb, seed):
    mats = B.zeros(B.JetBSpace([B.JetSpace(T, k, k)] * (nb * nb)))
    B.check(B.lib.jets_buf_rand(mats._h, seed, 0, int(os.environ.get("AB_DIST", "0"))))
    A = B.blockop([[B.JopDense(B.getblock(mats, 1 + r + nb * c), nrhs=nrhs) for c in range(nb)] for r in range(nb)])
    return mats, A


# ---- accuracy
if SKIP:
    nb = 1
else:
    nb = 4
mats, A = build(nb, 77)
mh = mats.to_host().astype(np.float64)
blocks = [mh[i * k * k:(i + 1) * k * k].reshape(k, k, order="F") for i in range(nb * nb)]
M = np.block([[blocks[r + nb * c] for c in range(nb)] for r in range(nb)])
dist = int(os.environ.get("AB_DIST", "0"))
m = (B.randn if dist else B.rand)(B.domain(A), seed=5)
x = m.to_host().astype(np.float64)
X = np.concatenate([x[b * k * nrhs:(b + 1) * k * nrhs].reshape(k, nrhs, order="F") for b in range(nb)], axis=0)
d = (A * m).to_host().astype(np.float64)
Dm = np.concatenate([d[b * k * nrhs:(b + 1) * k * nrhs].reshape(k, nrhs, order="F") for b in range(nb)], axis=0)
ref = M @ X
scale = np.abs(M) @ np.abs(X)
print(tag, "fwd: normwise rel err %.3e  max |err|/(|A||x|) %.3e" % (np.linalg.norm(Dm - ref) / np.linalg.norm(ref), np.max(np.abs(Dm - ref) / scale)))
y = (B.randn if dist else B.rand)(B.range_(A), seed=6)
yh = y.to_host().astype(np.float64)
Y = np.concatenate([yh[b * k * nrhs:(b + 1) * k * nrhs].reshape(k, nrhs, order="F") for b in range(nb)], axis=0)
t = (A.T * y).to_host().astype(np.float64)
Tm = np.concatenate([t[b * k * nrhs:(b + 1) * k * nrhs].reshape(k, nrhs, order="F") for b in range(nb)], axis=0)
ref = M.T @ Y
scale = np.abs(M.T) @ np.abs(Y)
print(tag, "adj: normwise rel err %.3e  max |err|/(|A||x|) %.3e" % (np.linalg.norm(Tm - ref) / np.linalg.norm(ref), np.max(np.abs(Tm - ref) / scale)))
print(tag, B.plan_info(A))
del mats, A, m, y

# ---- timing
nb = 16
mats, A = build(nb, 3001)
m = B.rand(B.domain(A), seed=3)
d = B.zeros(B.range_(A))
m2 = B.zeros(B.domain(A))
At = A.T
for _ in range(3):
    B.mul_(d, A, m)
    B.mul_(m2, At, d)
B.sync()
for nm, fn in (("fwd", lambda: B.mul_(d, A, m)), ("adj", lambda: B.mul_(m2, At, d))):
    best = 1e9
    for rep in range(3):
        B.sync()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        B.sync()
        best = min(best, (time.perf_counter() - t0) * 100)
    print(f"{tag} nrhs={nrhs} {nm}: {best:.3f} ms  {nb * nb * k * k * 4 / best / 1e6:.0f} GB/s of matrix bytes")
