#!/bin/bash
# ncu --set full captures: tools/gpu_ncu.sh <tag> <kernel-regex> <count> [launch-skip]
TAG=$1; RE=$2; CNT=${3:-6}; SKIP=${4:-0}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:$RE" -s $SKIP -c $CNT -o $OUT/prof_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_${TAG}.log 2>&1; echo "ncu rc=$?"
