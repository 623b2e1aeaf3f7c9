"""CPU-side checks: the C-ABI library loads, exports every symbol include/jets_b200.h declares,
fails loudly without a GPU, and the host-side space bookkeeping matches the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jets_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jets_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import jets_b200 as B
    lib = ctypes.CDLL(B.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/jets_b200.h but not exported: {missing}"
    # and the ctypes binding covers the whole header, nothing more
    assert sorted(B.SIGNATURES) == names
    assert lib.jets_abi_version() == 1


def test_every_declaration_cites_the_reference():
    """Each group of entry points names the reference file:line it replaces."""
    src = open(HEADER).read()
    assert src.count("src/Jets.jl:") >= 25 and "test/runtests.jl:" in src


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import jets_b200 as B
    with pytest.raises(B.JetsError) as e:
        B.zeros(B.JetSpace(np.float64, 4))
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jets.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle/jets_oracle.py", "").replace(
                    "oracle's", "").replace("oracle is", ""), f"{f} references the oracle"


def test_block_space_indices_match_reference_rule():
    """JetBSpace ranges are cumulative, 1-based, inclusive (src/Jets.jl:742-748) in both the oracle
    and the device host mirror."""
    import jets_b200 as B
    from oracle import jets_oracle as J
    for mk in (B, J):
        R = mk.JetBSpace([mk.JetSpace(np.float64, 2), mk.JetSpace(np.float64, 2, 2), mk.JetSpace(np.float64, 2, 3)])
        assert R.indices == [(1, 2), (3, 6), (7, 12)]
        assert R.size() == (12,) and len(R) == 12 and mk.nblocks(R) == 3
        assert mk.indices(R, 2) == (3, 6) and mk.space(R, 3) == mk.JetSpace(np.float64, 2, 3)
        assert R == mk.JetBSpace([mk.JetSpace(np.float64, 2), mk.JetSpace(np.float64, 2, 2), mk.JetSpace(np.float64, 2, 3)])
        assert R.similar((0,)) == mk.JetSpace(np.float64, (0,))
