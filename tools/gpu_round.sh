#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, ncu --set full of the streaming kernels.  Usage: tools/gpu_round.sh <tag> [noncu]
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
SRK_CHOL_PROFILE=1 timeout 120 tools/chol_bench 9993 90 > $OUT/chol_${TAG}_prof.log 2>&1; grep -E "trsv:|potrf|panel|syrk" $OUT/chol_${TAG}_prof.log | tail -6
timeout 120 tools/chol_bench 9993 90 > $OUT/chol_${TAG}.log 2>&1; echo "chol_bench rc=$?"; grep -E "iter 3|max" $OUT/chol_${TAG}.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_${TAG}.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"; tail -3 $OUT/bench_${TAG}.err
python - <<PY
import json
d=json.load(open('$OUT/bench_${TAG}.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(' ', k, round(v.get('avg_ms',0),3), round(v.get('frac',0),3))
PY
if [ "$2" != "noncu" ]; then
timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:k_jacobian|k_residual$|k_schur_mma|k_backsub|k_frame_blocks' -c 9 -o $OUT/prof_stream_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/ncu_stream_${TAG}.log 2>&1; echo "ncu rc=$?"
fi
