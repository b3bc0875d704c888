#!/bin/bash
# final evidence of the round: launch list (c3) and a --set full capture of the solve kernels
OUT=gpurun_out; mkdir -p $OUT
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/r02_launches_c3_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-others > $OUT/ncu_l3.log 2>&1; echo "launch list c3 rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_band_chol|k_trsv_cluster|k_residual_dd_tiles|k_permute_tiles" -s 14 -c 14 -o $OUT/r02_prof_solve -f python bench.py --steps 1 --warmup 1 --no-cpu --no-others > $OUT/ncu_s.log 2>&1; echo "ncu solve rc=$?"
