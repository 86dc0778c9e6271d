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
echo "built $OUT/libelas_ref.so $OUT/libelas_ref_fast.so $OUT/libelas_ref_omp.so"
