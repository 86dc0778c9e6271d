#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's own serial ELAS from the sources where
# they lie under /root/reference (nothing is copied) into oracle/_ref/:
#   libelas_ref.so       strict IEEE  (-O2 -ffp-contract=off, no -ffast-math, no -march=native):
#                        the parity oracle == what `pip install .` of the reference builds
#                        (setup.py:26), SURVEY.md finding 2
#   libelas_ref_fast.so  the reference Makefile's flags (-O2 -ffast-math, Makefile:14,37):
#                        speed baseline only, results differ from the strict build
#   libelas_ref_omp.so   the reference's OpenMP variant (src/omp_includes/elas/elas.cpp, Makefile flags + -fopenmp):
#                        speed baseline only
# oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/serial_includes/elas" ]; then
    echo "build_ref.sh: $REF not present; keeping prebuilt files in $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
COMMON="$REF/src/common_includes/elas/descriptor.cpp $REF/src/common_includes/elas/filter.cpp $REF/src/common_includes/elas/matrix.cpp $REF/src/common_includes/elas/triangle.cpp"
STRICT="-O2 -std=c++17 -w -ffp-contract=off"
FAST="-O2 -std=c++17 -w -ffast-math"
g++ $STRICT -fPIC -shared -DORACLE_REF_FLAGS="\"g++ $STRICT\"" -include "$HERE/ref_prelude.h" -I"$REF/src" \
    "$HERE/ref_taps.cpp" $COMMON -o "$OUT/libelas_ref.so"
g++ $FAST -fPIC -shared -DORACLE_REF_FLAGS="\"g++ $FAST\"" -include "$HERE/ref_prelude.h" -I"$REF/src" \
    "$HERE/ref_taps.cpp" $COMMON -o "$OUT/libelas_ref_fast.so"
# the reference's OpenMP variant (src/omp_includes/elas, `make omp=1`: Makefile flags + -fopenmp): CPU timing baseline only
OMP="-O2 -std=c++17 -w -ffast-math -fopenmp"
g++ $OMP -fPIC -shared -DORACLE_REF_OMP -DORACLE_REF_FLAGS="\"g++ $OMP\"" -include "$HERE/ref_prelude.h" -I"$REF/src" \
    "$HERE/ref_taps.cpp" $COMMON -o "$OUT/libelas_ref_omp.so" || echo "build_ref.sh: the OpenMP variant did not build (baseline skipped)" >&2
# Row 19 (projectParallel): the reference's own CUDA kernel, cut out of its driver where it lies (the driver as a whole needs OpenCV /
# popt / GL and does not compile here) and compiled with the reference Makefile's nvcc flags for sm_100a.  Runs on the GPU box only.
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SV="$REF/src/parallel_includes/main/stereo_vision.cu"
if [ -x "$NVCC" ] && [ -f "$SV" ]; then
    awk '/^__global__ void projectParallel\(/{on=1} on{print} on&&/^}/{exit}' "$SV" > "$OUT/project_parallel_extract.cuh"
    if [ "$(grep -c 'make_double3' "$OUT/project_parallel_extract.cuh")" = "1" ]; then
        PFLAGS="-O2 -std=c++17 -w -gencode arch=compute_100a,code=sm_100a"
        "$NVCC" $PFLAGS -DORACLE_REF_PROJECT_FLAGS="\"nvcc $PFLAGS\"" -shared -Xcompiler -fPIC -I"$OUT" "$HERE/ref_project.cu" \
            -o "$OUT/libproject_ref.so" || echo "build_ref.sh: projectParallel did not build (row-19 reference oracle skipped)" >&2
    else
        echo "build_ref.sh: could not locate projectParallel in $SV" >&2
    fi
fi
echo "built $OUT/libelas_ref.so $OUT/libelas_ref_fast.so $OUT/libelas_ref_omp.so"
