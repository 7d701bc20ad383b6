#!/bin/sh
# Builds the CPU oracle and its driver with AddressSanitizer + UndefinedBehaviorSanitizer and runs it.
# usage: oracle/san/run.sh [output-dir]    (default: a temporary directory)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${1:-$(mktemp -d)}
CC=$(test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)
$CC -O1 -g -std=c99 -Wall -Wextra -ffp-contract=off -fno-fast-math -mfma -mavx2 -fopenmp \
    -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer \
    "$HERE/sanitize_driver.c" "$HERE/../noc_oracle.c" -o "$OUT/oracle_san" -lm
ASAN_OPTIONS=detect_leaks=1:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 OMP_NUM_THREADS=2 "$OUT/oracle_san"
