#!/bin/bash
# final ncu evidence: launch list of the default bench command (primary workload) and full captures of the two BASELINE kernels
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ay_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/ay_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_wgroup --launch-skip 6 --launch-count 1 -o $O/ay_wgroup_cfg2 -f python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --no-e2e > $O/ay_ncu_full.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_postings_kernel --launch-skip 4 --launch-count 1 -o $O/ay_postings_cfg3 -f python bench.py --workload cfg3_3.1Gb_31kbins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ay_ncu_full3.log 2>&1
ls -la $O/ay_* | head
