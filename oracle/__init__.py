"""CPU oracle for the hierarchical-GMM fit / register hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import, call, link or execute anything under `oracle/`.  The product path
(`gpu-accelerated-point-cloud-registration-using-hierarchical-gmm_b200/`) never imports it
and fails loudly when the CUDA library is missing.

Contents (each function cites the reference file:line it restates; paths are relative to
the reference checkout `/root/reference/`):

* `flat_gmm.py`      float64/float32 NumPy restatement of the flat EM variants
                     (`src/python/gmm_waymo/src/gmm_impl.py`, `src/c++/gmm_fit/gmm_kernels.cu`)
* `hgmm_tree.py`     vectorised restatement of the 8-ary tree build
                     (`src/python/hgmm/hgmm_cupy_cpu_working.py`, `hgmm_gpu.py`)
* `registration.py`  tree registration E-step / twist least-squares M-step / outer loop
                     (`src/python/hgmm/hgmm_gpu.py:550-577,620-664,729-768`), the
                     north-star weighted-Procrustes solve, the flat-mixture registration of BASELINE
                     configs[3] (no reference implementation: gmm_reg.cu:54-56 is empty), and the McAdams
                     3x3 SVD (`src/c++/common/svd3.h`)
* `l2reg.py`         L2-distance cost/gradient of the flat registration
                     (`src/python/gmmreg_gpu/cost_functions.py`, `so.py`, `transforms.py`)
* `synth.py`         alias of `hgmm_b200/synth.py` (the synthetic LiDAR / surface INPUT generators of SURVEY.md section 8d live
                     with the package: bench.py and the probes use them without touching this directory)
* `make_golden_lidar.py`  config-size tree fixtures: this oracle run once on the 100k / 50k LiDAR workloads
                     (`tests/golden/tree_build_lidar*`)
* `c/`               plain-C (OpenMP) restatement of the same E/M arithmetic used as the
                     timed CPU baseline; built into `oracle/_build/`
* `refshim/`, `make_golden.py`  harness that imports the *unmodified* reference Python
                     files in the build container, checks the restatement against them and
                     writes `tests/golden/*.npz`
* `build_ref.sh`     compiles the reference's own CUDA fitter from the sources where they
                     lie under `/root/reference` into `oracle/_ref/` (timing comparator)

Parity pinning: the reference ships no tests or golden vectors, so the
restatement is pinned against outputs of the reference itself run in the build container
(`make_golden.py`; fixtures committed under `tests/golden/`).  The C++ full-covariance
variant has no CPU implementation in the reference; its restatement is pinned only through
the reference CUDA binary run on the GPU box (`oracle/_ref`).
"""
