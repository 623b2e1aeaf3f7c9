import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jets_b200 as B
np.set_printoptions(linewidth=200, precision=4, suppress=True)
T = np.float32
def run(rows, cols, nrhs, seed=0, kind="rand"):
    g = np.random.default_rng(seed)
    if kind == "rand":
        A = (g.random((rows, cols)) - 0.3).astype(T)
    elif kind == "rowid":
        A = np.zeros((rows, cols), T); A[:, 0] = np.arange(rows) + 1
    elif kind == "eye":
        A = np.eye(rows, cols, dtype=T)
    X = (g.random((cols, nrhs)) - 0.5).astype(T)
    Y = (g.random((rows, nrhs)) - 0.5).astype(T)
    if kind != "rand":
        X = np.ones((cols, nrhs), T) * (np.arange(nrhs) + 1)
        Y = np.ones((rows, nrhs), T) * (np.arange(nrhs) + 1)
    op = B.JopDense(A, nrhs=nrhs)
    f = (op * B.to_device(X, B.domain(op))).to_host()
    t = (op.T * B.to_device(Y, B.range_(op))).to_host()
    rf = A.astype(np.float64) @ X
    rt = A.astype(np.float64).T @ Y
    ef = np.linalg.norm(f - rf) / np.linalg.norm(rf)
    et = np.linalg.norm(t - rt) / np.linalg.norm(rt)
    print(f"{kind} {rows}x{cols} nrhs={nrhs}: fwd err {ef:.3e}  adj err {et:.3e}  plan {B.plan_info(op)}", flush=True)
    if ef > 1e-5:
        print(" fwd got\n", f[:6, :6], "\n ref\n", rf[:6, :6])
    if et > 1e-5:
        print(" adj got\n", t[:6, :6], "\n ref\n", rt[:6, :6])
for kind in ("eye", "rowid", "rand"):
    run(128, 32, 16, kind=kind)
run(128, 64, 16)
run(256, 128, 64)
run(300, 200, 5)
run(2048, 2048, 64)
