"""Builds libjets_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libjets_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177"]
# Elementwise / broadcast kernels must not contract a*b+c into an FMA: Julia's broadcast rounds
# each op separately and the oracle is bit-compared on these paths.
SOURCES = {
    "api.cu": [],
    "plan.cu": [],
    "kernels_fused.cu": ["-fmad=false"],
    "kernels_fused_fast.cu": ["-fmad=false"],
    "kernels_fused_bundle.cu": ["-fmad=false"],
    "kernels_vec.cu": ["-fmad=false"],
    "kernels_cplx.cu": ["-fmad=false"],
    "kernels_dense.cu": [],
    "kernels_gemm_tc.cu": [],
    "dist.cu": [],
    "dist_op.cu": [],
}


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stamp(src: str, flags: list[str]) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for f in [src] + sorted(
        os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".hpp", ".cuh", ".h"))
    ) + [os.path.join(HERE, "..", "include", "jets_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(name: str, extra: list[str], verbose: bool) -> str:
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    flags = ARCH + COMMON + extra
    stamp = _stamp(src, flags)
    stamp_file = obj + ".stamp"
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj
    cmd = [nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj


def build_lib(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = {k: v for k, v in SOURCES.items() if os.path.exists(os.path.join(CSRC, k))}
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda kv: _compile(kv[0], kv[1], verbose), srcs.items()))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build_lib(verbose="-v" in sys.argv, force="-f" in sys.argv))
