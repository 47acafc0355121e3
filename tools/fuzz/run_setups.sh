#!/bin/bash
# tools/fuzz/run_setups.sh [iterations] [seeds...] -- fuzzes every ...PlanSetup entry point of include/dali_b200.h under
# AddressSanitizer + UndefinedBehaviorSanitizer without a GPU: the whole kernel library is built with the sanitizers, the CUDA runtime
# calls the host side makes are replaced by cuda_stub.c, and fuzz_plan_setups.py draws the arguments (plausible values mixed with
# zero / negative / huge sizes, windows outside the image, NaN / infinite floats, invalid enum codes).  Each seed runs twice: the
# adversarial mix and a mild one (FUZZ_BAD_SCALE=0.05) whose set-ups reach the table / tiling code.  A finding is a sanitizer report.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(cd "$HERE/../.." && pwd)"
OUT="${FUZZ_DIR:-/tmp/dali_b200_fuzz}"; mkdir -p "$OUT/all"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
F="-gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer --expt-relaxed-constexpr -fmad=false -I$ROOT/include"
for f in "$ROOT"/dali_b200/csrc/*.cu; do $NVCC $F -c "$f" -o "$OUT/all/$(basename "$f").o" & done; wait
$NVCC -shared -o "$OUT/libfuzz_all.so" "$OUT"/all/*.o -gencode arch=compute_100a,code=sm_100a -lcudart -Xcompiler -fsanitize=address,-fsanitize=undefined
gcc -shared -fPIC -O1 -o "$OUT/cuda_stub.so" "$HERE/cuda_stub.c"
N="${1:-1500}"; shift || true
SEEDS="${*:-1 2 3}"
for s in $SEEDS; do
  for scale in 1 0.05; do
    DALIB200_LIB="$OUT/libfuzz_all.so" FUZZ_BAD_SCALE=$scale LD_PRELOAD="$(gcc -print-file-name=libasan.so) $OUT/cuda_stub.so" \
      ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1 python "$HERE/fuzz_plan_setups.py" "$s" "$N" | cut -c1-400
  done
done
