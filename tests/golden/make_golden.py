"""Generates tests/golden/oracle_golden.npz: seeded inputs and the ORACLE's outputs for a set of small
hot-path scenarios.  Run from the repo root:  python tests/golden/make_golden.py

What these vectors are -- and are not.  The reference (ChevronETC/Jets.jl) is pure Julia and cannot run
in this image, and it ships no golden vectors of its own, so these are NOT reference outputs: they are the
numpy oracle's outputs frozen at the commit that generated them.  They pin (a) the oracle against silent
change (numpy upgrades, edits to oracle/jets_oracle.py) and (b) the device path against vectors that do
not depend on the numpy installed on the GPU box.  Parity with the reference itself remains pinned only by
the restated identities of test/runtests.jl (tests/test_oracle_reference_identities.py): "parity unpinned".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_scenarios import SCENARIOS, run_scenario  # noqa: E402
from backends import OracleBackend  # noqa: E402


def main():
    K = OracleBackend()
    out = {}
    for name in SCENARIOS:
        res = run_scenario(K, name)
        for k, v in res.items():
            out[f"{name}/{k}"] = np.asarray(v)
    path = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
