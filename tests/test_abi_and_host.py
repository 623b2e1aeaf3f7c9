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


def _indexmap(I):  # test/runtests.jl:219-225
    return tuple(I) if I[0] < 5 else (I[0] - 4, I[1])


def test_symmetric_space_bookkeeping_matches_the_oracle():
    """JetSSpace host logic (src/Jets.jl:408-484): storage offset / conjugation of every logical
    index and the multiplicity table behind norm() agree with the oracle's SymmetricArray."""
    import jets_b200 as B
    from oracle import jets_oracle as J
    Rb = B.JetSSpace(np.complex128, (8, 4), (4, 4), _indexmap)
    Ro = J.JetSSpace(np.complex128, (8, 4), (4, 4), _indexmap)
    assert Rb.size() == (8, 4) and len(Rb) == 32 and Rb._block_lens() == [16] and Rb.eltype == np.complex128
    assert Rb.similar((0, 0)).n == (0, 0) and Rb.similar((0, 0)).M == Rb.M
    g = np.random.default_rng(3)
    x = J.rand(Ro, g)
    P = x.A.reshape(-1, order="F")
    for i1 in range(1, 9):
        for i2 in range(1, 5):
            off, cj = Rb._parent_linear((i1, i2))
            assert (np.conj(P[off]) if cj else P[off]) == x[(i1, i2)]
    w = Rb.multiplicity()
    assert w.shape == (16,) and np.all(w == 2.0)
    for p in (1, 2, 3.5):
        assert np.isclose(np.sum(w * np.abs(P) ** p) ** (1 / p), J.norm(x, p), rtol=1e-14)


def test_complex_eltypes_are_in_the_abi():
    import jets_b200 as B
    from jets_b200 import _lib as L
    assert (L.F32, L.F64, L.C64, L.C128) == (0, 1, 2, 3)
    src = open(HEADER).read()
    assert "JETS_C64 = 2" in src and "JETS_C128 = 3" in src


def test_every_tuning_switch_is_documented():
    """Every environment switch the library, the host mirror or the bench reads is listed in DESIGN.md (section 7a)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "jets.jl_b200", "csrc", "*")):
        names |= set(re.findall(r'getenv\("(JETS_[A-Z0-9_]+)"\)', open(f).read()))
    for f in glob.glob(os.path.join(root, "jets.jl_b200", "*.py")) + [os.path.join(root, "bench.py")]:
        names |= set(re.findall(r'environ(?:\.get\(|\[)"(JETS_[A-Z0-9_]+)"', open(f).read()))
    design = open(os.path.join(root, "DESIGN.md")).read()
    missing = sorted(n for n in names if n not in design)
    assert names and not missing, missing
