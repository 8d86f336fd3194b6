#!/bin/bash
# compute-sanitizer is closed on this GPU pool (profiles/r2_compute_sanitizer_closed.log).  Substitute for its memcheck on
# the arithmetic: the kernels' per-proof routines are __host__ __device__ code, so the host build (tests/hostsim) runs
# under AddressSanitizer + UndefinedBehaviorSanitizer through the whole hostsim test-suite (~16 min).
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -DBN254_COUNT_MULS -shared -fPIC \
    -o /tmp/_hostsim_asan.so tests/hostsim/hostsim.cpp
cp tests/hostsim/_hostsim.so /tmp/_hostsim_backup.so 2>/dev/null || true
cp /tmp/_hostsim_asan.so tests/hostsim/_hostsim.so && touch tests/hostsim/_hostsim.so
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_hostsim.py -x -q 2>&1 | tee profiles/r2_hostsim_asan_ubsan.log
cp /tmp/_hostsim_backup.so tests/hostsim/_hostsim.so 2>/dev/null && touch tests/hostsim/_hostsim.so
