#!/bin/bash
# final evidence of the round: launch list (c3) and a --set full capture of K2's final form
OUT=gpurun_out; mkdir -p $OUT
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/r02_launches_c3_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-others > $OUT/ncu_l3.log 2>&1; echo "launch list c3 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_schur_v3|k_jacobian|k_point_finish" -s 3 -c 3 -o $OUT/r02_prof_k2 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-others > $OUT/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
