"""CPU: the C-ABI library loads and exports every symbol include/hgmm.h declares; host-side wrappers
fail loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hgmm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hgmm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from hgmm_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libhgmm.so missing: run python __graft_entry__.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libhgmm.so does not export %s" % s
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (set(syms) ^ set(_lib.SIGNATURES))


def test_version_and_node_count_need_no_gpu():
    from hgmm_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.hgmm_version()
    assert [lib.hgmm_tree_total_nodes(l) for l in range(1, 6)] == [8, 72, 584, 4680, 37448]


def test_config_struct_layout_matches_header(tmp_path):
    """the ctypes mirrors against the C compiler's view of include/hgmm.h: sizes and the offset of the last field"""
    import subprocess
    from hgmm_b200 import _lib
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hgmm.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(hgmm_flat_config), offsetof(hgmm_flat_config, reserved), sizeof(hgmm_tree_config), '
                   'offsetof(hgmm_tree_config, prune_min_points), sizeof(hgmm_reg_config), offsetof(hgmm_reg_config, lambda_c)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    c = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    py = [ctypes.sizeof(_lib.FlatConfig), _lib.FlatConfig.reserved.offset, ctypes.sizeof(_lib.TreeConfig),
          _lib.TreeConfig.prune_min_points.offset, ctypes.sizeof(_lib.RegConfig), _lib.RegConfig.lambda_c.offset]
    assert c == py, (c, py)
    assert c[0] == 32 and c[2] == 40 and c[4] == 16


def test_no_cpu_fallback():
    """without a CUDA device the product path must raise, not compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import hgmm_b200
    with pytest.raises(hgmm_b200.HgmmError):
        hgmm_b200.Engine(0)
    import numpy as np
    with pytest.raises(hgmm_b200.HgmmError):
        hgmm_b200.gmm_impl.train_gmm(np.zeros((10, 3), np.float32), 1, 0.0, np.zeros((2, 3)), np.ones((2, 3)), np.ones(2) / 2)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_reference_api_names_present():
    from hgmm_b200 import gmm_impl, gmm, hgmm
    for n in ("train_gmm", "predict", "init_gmm_params", "timer"):
        assert hasattr(gmm_impl, n)
    for n in ("Feature", "GMM_GPU", "GMM_GPU_Base", "GMM_CPU", "GMM_CPU_Base"):
        assert hasattr(gmm, n)
    for n in ("buildGMMTree", "GMMTree", "registration_gmmtree", "RigidTransformation", "EstepResult", "MstepResult", "child", "level"):
        assert hasattr(hgmm, n)
    assert hgmm.level(3) == 584 and hgmm.child(-1) == 0
    import numpy as np
    assert (hgmm.reference_init_indices(2) == np.random.RandomState(72).randint(72, size=72)).all()
