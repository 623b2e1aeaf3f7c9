"""Replays the committed golden vectors (tests/golden/oracle_golden.npz, written by
tests/golden/make_golden.py from the numpy oracle -- see that file for what they pin and what they do
not): the oracle must still reproduce them on this machine's numpy (CPU suite), and the CUDA path must
reproduce them through the C ABI (GPU suite).  Elementwise / stencil / data-movement results are
bit-compared; keys ending in `_red` (reductions, dense products) to 1e-12 (Float64) / 1e-5 (Float32)."""
import os

import numpy as np
import pytest

from backends import OracleBackend, DeviceBackend
from golden_scenarios import SCENARIOS, run_scenario

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))


def compare(name, res):
    T = SCENARIOS[name][1]
    tol = 1e-5 if np.dtype(T) in (np.dtype(np.float32), np.dtype(np.complex64)) else 1e-12
    keys = [k.split("/", 1)[1] for k in GOLD.files if k.startswith(name + "/")]
    assert sorted(keys) == sorted(res), (sorted(keys), sorted(res))
    for k in keys:
        want, got = GOLD[f"{name}/{k}"], np.asarray(res[k])
        assert want.shape == got.shape, (name, k, want.shape, got.shape)
        if k.endswith("_red"):
            if np.iscomplexobj(want):      # complex results as (re, im) pairs
                want, got = np.ascontiguousarray(want).view(want.real.dtype), np.ascontiguousarray(got).view(got.real.dtype)
            w64, g64 = want.astype(np.float64).ravel(), got.astype(np.float64).ravel()
            fin = np.isfinite(w64)
            assert np.array_equal(fin, np.isfinite(g64))
            scale = np.linalg.norm(w64[fin]) if w64.size > 6 else np.abs(w64[fin])
            assert np.all(np.abs(w64[fin] - g64[fin]) <= tol * np.maximum(scale, 1e-300)), (name, k)
        else:
            assert want.dtype == got.dtype, (name, k, want.dtype, got.dtype)
            assert np.array_equal(want.view(np.uint8), got.view(np.uint8)), f"{name}/{k} is not bit-identical to the golden vector"


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_reproduces_the_golden_vectors(name):
    compare(name, run_scenario(OracleBackend(), name))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_device_reproduces_the_golden_vectors(name):
    compare(name, run_scenario(DeviceBackend(), name))
