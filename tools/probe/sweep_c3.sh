#!/bin/bash
# development aid: c3 step time against environment switches given as arguments (VAR=value ...)
run() {
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-others --no-cpu 2>/tmp/sweep.err > /tmp/sweep.json || tail -3 /tmp/sweep.err
  python - "$*" <<'P'
import json, sys
b = json.loads(open("/tmp/sweep.json").read().strip().splitlines()[-1]); k = b["kernels"]
print(sys.argv[1], round(b["ms_per_step"], 4), {n: round(v["avg_ms"], 3) for n, v in k.items() if "avg_ms" in v}, b["err_after_step"])
P
}
run SRK_X=0
for v in "$@"; do run $v; done
