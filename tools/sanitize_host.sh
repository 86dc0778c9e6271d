#!/usr/bin/env bash
# The host C++ of the library (host Delaunay stage, calibration, image readers, drop-in layer, synthetic generator) rebuilt with
# AddressSanitizer + UndefinedBehaviorSanitizer (or ThreadSanitizer: SAN=thread) next to the product build, linked with the product's
# CUDA objects, and the CPU tests that drive it run against that library.  No GPU needed.
#   tools/sanitize_host.sh            # ASan + UBSan
#   SAN=thread tools/sanitize_host.sh # TSan (the host Delaunay stage's worker threads)
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
PKG="$(ls -d "$ROOT"/low-cost-*_b200)"
SAN="${SAN:-address,undefined}"
OUT="$PKG/lib_san"
make -C "$PKG/csrc" -j8 >/dev/null
rm -rf "$OUT" && mkdir -p "$OUT/obj"
for f in host_delaunay synth calib dropin image_io; do
    g++ -O1 -g -std=c++17 -fPIC -Wall -ffp-contract=off -fsanitize="$SAN" -fno-omit-frame-pointer -c "$PKG/csrc/$f.cpp" -o "$OUT/obj/$f.o"
done
if [ "$SAN" = thread ]; then LIBS="-Xlinker -ltsan"; PRE="$(gcc -print-file-name=libtsan.so)"; else LIBS="-Xlinker -lasan -Xlinker -lubsan"; PRE="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)"; fi
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libelas_b200.so" "$PKG"/lib/obj/k_*.o "$PKG/lib/obj/pipeline.o" \
    "$PKG/lib/obj/band_split.o" "$OUT"/obj/*.o -lpthread -lz $LIBS
cp "$PKG/lib/libsvb_synth.so" "$OUT/"
cd "$ROOT"
SVB_LIB_DIR="$OUT" LD_PRELOAD="$PRE" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" \
    python -m pytest tests/test_cabi_host.py tests/test_calibration.py tests/test_image_io.py -x -q -p no:cacheprovider
rm -rf "$OUT"
