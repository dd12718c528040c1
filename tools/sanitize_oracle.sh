#!/bin/bash
# ASan + UBSan over the CPU oracle ("-fsanitize=address,undefined on the C++ oracle"): builds an instrumented
# copy of oracle/libmrt_oracle.so under scratch/ and runs the oracle-only test files against it (MRT_ORACLE_SO).
# Output: profiles/r2_oracle_asan_ubsan.txt
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/san
g++ -O1 -g -march=x86-64-v3 -ffp-contract=off -fno-fast-math -std=c++17 -fPIC -pthread -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -fno-omit-frame-pointer -shared -o scratch/san/libmrt_oracle.so oracle/mrt_oracle.cpp
ASAN=$(g++ -print-file-name=libasan.so)
{
  echo "# ASan + UBSan run of the oracle: $(date -u +%FT%TZ), $(g++ --version | head -1)"
  echo "# g++ -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined oracle/mrt_oracle.cpp; LD_PRELOAD=$ASAN"
  MRT_ORACLE_SO=$PWD/scratch/san/libmrt_oracle.so LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest tests/test_oracle_golden.py tests/test_second_restatement.py tests/test_golden_vectors.py tests/test_fuzz_cpu.py -q -x -m "not gpu" -p no:cacheprovider 2>&1 | tail -15
} | tee profiles/r2_oracle_asan_ubsan.txt
