#!/bin/bash
# development aid: factor / substitution time of the c3 solve against the elimination-order switches
run() {
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-others 2>/dev/null > /tmp/sweep.json
  python - "$*" <<'P'
import json, sys
b = json.loads(open("/tmp/sweep.json").read().strip().splitlines()[-1]); k = b["kernels"]
print(sys.argv[1], round(b["ms_per_step"], 4), round(k["solve"]["avg_ms"], 4), round(k["solve_factor"]["avg_ms"], 4), round(k["solve_trsv"]["avg_ms"], 4), k["solve"]["work_counted"][100:], b["err_after_step"])
P
}
run SRK_SOLVE_LEVELS=2
run SRK_SOLVE_LEVELS=2 SRK_BAND_SMEM_PAD=0
run SRK_SOLVE_LEVELS=1
run SRK_SOLVE_LEVELS=1 SRK_BAND_SMEM_PAD=0
run SRK_SOLVE_LEVELS=2 SRK_SOLVE_MAX_PARTS=12
run SRK_SOLVE_LEVELS=2 SRK_SOLVE_MAX_PARTS=10
