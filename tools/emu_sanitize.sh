#!/bin/bash
# The `-m gpu` tests on the CPU emulation of the kernels (tests/simt), built with AddressSanitizer and then with
# UndefinedBehaviorSanitizer: every access of a kernel to "device" memory is checked against the bounds of its cudaMalloc, every
# vector load against its alignment.  No GPU needed.   Usage: bash tools/emu_sanitize.sh [pytest -k expression]
set -u
K=${1:-"not back_to_back and not blob_adoption and not c2_c3_full_size and not million_triangles"}
run() {  # $1 = flags, $2 = runtime library, $3 = runtime options
  export RDN_SIMT_CXXFLAGS="$1"
  python tests/simt/build_emu.py > /dev/null || exit 1
  LD_PRELOAD=$(g++ -print-file-name=$2) $3 RDN_SIMT_EMU=1 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "$K" 2>&1 | grep -E "passed|failed|ERROR|Sanitizer|runtime error" | tail -5
}
echo "== AddressSanitizer"
run "-fsanitize=address -fno-omit-frame-pointer" libasan.so "env ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0"
echo "== UndefinedBehaviorSanitizer"
run "-fsanitize=undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer" libubsan.so "env UBSAN_OPTIONS=print_stacktrace=1"
