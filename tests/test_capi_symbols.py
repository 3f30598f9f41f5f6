"""The C-ABI library loads and exports every symbol include/qmri.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qmri.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qmri_[a-z0-9_]+)\s*\(", src)) - {"qmri_denoise_fn"})


def test_library_exports_every_declared_symbol():
    import qmri_b200
    lib = qmri_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qmri.h but not exported"
    assert sorted(qmri_b200._capi.SIGNATURES) == names, "ctypes SIGNATURES and include/qmri.h diverge"
    assert lib.qmri_version() == 100


def test_no_cpu_fallback_without_gpu():
    import qmri_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(qmri_b200.QmriError, match="no CPU fallback"):
        qmri_b200.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "qmri-pnp-recon-poc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f"{f} references the oracle"


def test_params_struct_layout_matches_header():
    from qmri_b200._capi import AdmmParams
    # int, double, double, int, ptr, ptr, ptr, ptr, int  with natural alignment
    assert ctypes.sizeof(AdmmParams) == 72
    assert AdmmParams.gamma.offset == 8 and AdmmParams.noise_map.offset == 32 and AdmmParams.fn_space.offset == 64
