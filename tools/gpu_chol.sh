#!/bin/bash
# Cholesky checks: sparse (cluster) path vs launch-per-operation path on banded and dense matrices
OUT=gpurun_out; mkdir -p $OUT
for args in "9993 90" "9993 0" "493 0" "493 20" "4000 0" "2000 130" "100 10" "64 0" "65 0"; do
  echo "== chol_bench $args"; timeout 120 tools/chol_bench $args 2>&1 | grep -E "iter 3|max \|A"
done
echo "== forced dense path 9993 90"; SRK_CHOL_PATH=dense timeout 120 tools/chol_bench 9993 90 | grep -E "iter 3|max \|A"
echo "== memcheck 1000 90"; timeout 300 compute-sanitizer --tool memcheck tools/chol_bench 1000 90 2>&1 | tail -4
echo "== racecheck 500 90"; timeout 300 compute-sanitizer --tool racecheck tools/chol_bench 500 90 2>&1 | tail -4
