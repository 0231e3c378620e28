"""The reference viewer's C++ API (cpp/hgmm_shim.h): symbol surface on CPU, a headless run on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG


def test_shim_exports_the_viewer_symbols():
    so = os.path.join(PKG, "hgmm_b200", "libhgmm_shim.so")
    assert os.path.exists(so), "run make in the package directory"
    syms = subprocess.run(["nm", "-DC", so], capture_output=True, text=True).stdout
    for name in ("GMM::solve(", "scanRegistration::initSimulation(", "scanRegistration::runSimulation(", "scanRegistration::copyBoidsToVBO(",
                 "scanRegistration::endSimulation(", "GMMRegistration::GMMRegistration(int)", "GMMRegistration::initSimulation(",
                 "GMMRegistration::pointCloudRegisterGPU(float)", "GMMRegistration::copyBoidsToVBO(", "GMMRegistration::endSimulation("):
        assert name in syms, name


def test_shim_compiles_against_the_reference_glm(tmp_path):
    glm = "/root/reference/external/include"
    if not os.path.isdir(glm):
        pytest.skip("reference checkout not present (GPU box)")
    out = tmp_path / "shim.o"
    subprocess.check_call(["g++", "-std=c++14", "-fPIC", "-I" + glm, "-c", os.path.join(PKG, "cpp", "hgmm_shim.cpp"), "-o", str(out)])
    syms = subprocess.run(["nm", "-C", str(out)], capture_output=True, text=True).stdout
    # the mangled names carry the viewer's own glm::tvec3<float, precision> type
    assert "GMM::solve(std::vector<glm::tvec3<float" in syms


@pytest.mark.gpu
def test_shim_headless_viewer_sequence():
    exe = os.path.join(PKG, "build", "shim_demo")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    flat = [float(v) for v in [l for l in out.stdout.splitlines() if l.startswith("FLAT")][0].split()[1:]]
    assert abs(flat[0] + 0.5) < 0.02 and abs(flat[3] - 0.5) < 0.02 and abs(flat[6] - 0.5) < 0.02      # two blobs at x = -+0.5
    reg = [float(v) for v in [l for l in out.stdout.splitlines() if l.startswith("REG")][0].split()[1:]]
    # target = Rz(0.1) source + t, so mapping the target back onto the source is ~Rz(-0.1)
    assert abs(reg[0] - np.cos(0.1)) < 0.02 and abs(reg[1] - np.sin(0.1)) < 0.03
