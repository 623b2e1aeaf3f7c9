"""A plain-C program (tests/c_abi/abi_smoke.c) consumes include/jets_b200.h and libjets_b200.so the way
a Julia `ccall` shim would -- no Python, no C++ in the host.  Compiling it with -std=c99 -pedantic -Wall
-Werror is the check that the header is a C ABI; on a CPU box it must report JETS_ERR_CUDA (no
fallback), on the B200 it must reproduce the 2x2 block-operator sums bit for bit."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c")
LIBDIR = os.path.join(ROOT, "jets.jl_b200")


def _build(tmp_path):
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), SRC,
                    "-o", exe, "-L", LIBDIR, "-ljets_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_is_plain_c_and_the_library_fails_loudly_without_a_gpu(tmp_path):
    import jets_b200  # noqa: F401  (makes sure the library is built)
    exe = _build(tmp_path)
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "JETS_ERR_CUDA" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_consumer_runs_the_block_operator_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi_smoke ok" in r.stdout
