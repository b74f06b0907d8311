#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_slots_sub --launch-skip 4 --launch-count 1 -o $O/aj_slots_sub_w32 -f python bench.py --workload w32_200Mb_2020bins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/aj_ncu32.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_postings_sub --launch-skip 4 --launch-count 1 -o $O/aj_postings_sub_k15 -f python bench.py --workload w16_k15 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/aj_ncu15.log 2>&1
tail -n 2 $O/aj_ncu32.log $O/aj_ncu15.log
