"""The MEX gateway a MATLAB maintainer would build compiles against the C ABI (mock mex.h; no MATLAB here)."""
import os
import shutil
import subprocess

import pytest

from conftest import PKG, ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_mex_gateway_compiles_against_header():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-DQMRI_MOCK_MEX", "-I" + os.path.join(PKG, "mex"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(PKG, "mex", "qmri_b200_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_matlab_wrappers_keep_reference_names():
    names = {"PnP_ADMM.m", "mrf_dtm_cpu.m", "setup_subsampling_spiralgrided.m", "setup_subsampling_epi.m",
             "denoiseImage_PnP_ADMM.m", "build_noise_map.m"}
    assert names <= set(os.listdir(os.path.join(PKG, "matlab")))
