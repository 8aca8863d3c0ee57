#!/usr/bin/env bash
# Host code of libpcs_seq.so (flattener, sample groups, planner, C-ABI argument handling) under
# AddressSanitizer + UBSan, then ThreadSanitizer, on CPU: builds an instrumented copy of the library into a
# scratch tree and runs the "not gpu" host tests against it.  Not part of the pytest suite (it takes minutes).
#   usage: tests/run_host_sanitizers.sh [scratch dir]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-/tmp/pcs_sanitizers}"
rm -rf "$OUT" && mkdir -p "$OUT"
cp -r "$ROOT/process_b200" "$ROOT/tests" "$ROOT/oracle" "$ROOT/include" "$OUT/"
build() {  # $1 = sanitizer flags
  (cd "$ROOT/process_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 \
     -Xcompiler "-fPIC,-pthread,-fno-omit-frame-pointer,$1" -shared -o "$OUT/process_b200/libpcs_seq.so" \
     kernels.cu pcs_seq.cpp flatten.cpp -lpthread 2>/dev/null)
}
TESTS="tests/test_host_logic.py tests/test_api_mirror.py tests/test_capi_library.py"
echo "== AddressSanitizer + UndefinedBehaviorSanitizer"
build "-fsanitize=address,-fsanitize=undefined"
(cd "$OUT" && LD_PRELOAD="$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so)" \
   ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
   python -m pytest $TESTS -x -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -5)
echo "== ThreadSanitizer (8 host threads)"
build "-fsanitize=thread"
(cd "$OUT" && LD_PRELOAD="$(g++ -print-file-name=libtsan.so)" TSAN_OPTIONS="exitcode=0" PCS_HOST_THREADS=8 \
   python -m pytest tests/test_host_logic.py -x -q -m "not gpu" -p no:cacheprovider 2>&1 | \
   grep -E "SUMMARY|passed|failed|Location is" | sort | uniq -c)
