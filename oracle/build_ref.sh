#!/usr/bin/env bash
# Builds the REFERENCE's own CUDA flat fitter into oracle/_ref/ref_gmm_cuda (timing comparator for configs[1]).
# Sources are compiled where they lie under /root/reference; nothing is copied into the repo and the
# reference's own (broken, CUDA-10-era) CMake is not used.  The -include list pre-parses thrust/CCCL before
# common/utilities.h:24-25 #defines the macros E and G.  TEST INFRASTRUCTURE ONLY.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF=/root/reference
[ -d "$REF/src/c++/gmm_fit" ] || { echo "no reference checkout at $REF"; exit 0; }
OUT="$HERE/_ref"
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
INC="-I$REF/external/include -I$REF/src/c++"
PRE="-include thrust/reduce.h -include thrust/extrema.h -include thrust/execution_policy.h -include thrust/sort.h -include thrust/random.h -include thrust/device_vector.h -include chrono"
$NVCC -std=c++14 -O2 -gencode arch=compute_100a,code=sm_100a $INC $PRE -w -c "$REF/src/c++/gmm_fit/gmm_kernels.cu" -o "$OUT/gmm_kernels.o"
$NVCC -std=c++14 -O2 -gencode arch=compute_100a,code=sm_100a $INC -w -x cu -c "$REF/src/c++/common/utilities.cpp" -o "$OUT/utilities.o"
$NVCC -std=c++14 -O2 -gencode arch=compute_100a,code=sm_100a $INC $PRE -w -c "$HERE/ref_driver_main.cu" -o "$OUT/ref_driver_main.o"
$NVCC -gencode arch=compute_100a,code=sm_100a -o "$OUT/ref_gmm_cuda" "$OUT/ref_driver_main.o" "$OUT/gmm_kernels.o" "$OUT/utilities.o"
rm -f "$OUT"/*.o
echo "built $OUT/ref_gmm_cuda"
