#!/bin/bash
# GPU run C: slot kernel v2 (position-level ring entries), grid waves, sharded call, README workload
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/c_pytest_gpu.log 2>&1; tail -3 $O/c_pytest_gpu.log
B="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B --workload cfg3_3.1Gb_31kbins > $O/c_cfg3_slots.json 2> $O/c_cfg3_slots.err
for w in 2 4 8; do RB_GRID_WAVES=$w timeout 200 $B --workload cfg3_3.1Gb_31kbins > $O/c_cfg3_waves$w.json 2>> $O/c_sweep.err; done
for ring in 1 3 4; do RB_SLOT_RING=$ring timeout 200 $B --workload cfg3_3.1Gb_31kbins > $O/c_cfg3_ring$ring.json 2>> $O/c_sweep.err; done
RB_SLOT_CTAS=1 timeout 200 $B --workload cfg3_3.1Gb_31kbins > $O/c_cfg3_ctas1.json 2>> $O/c_sweep.err
RB_SLOT_BYTES=1024 timeout 200 $B --workload cfg3_3.1Gb_31kbins > $O/c_cfg3_slot1024.json 2>> $O/c_sweep.err
for w in 1 2 4 8 16; do RB_GRID_WAVES=$w timeout 200 $B --workload cfg2_100x4Mb_100bins > $O/c_cfg2_waves$w.json 2>> $O/c_sweep.err; done
for w in 1 4 16; do RB_GRID_WAVES=$w timeout 200 $B --workload cfg2_k15 > $O/c_k15_waves$w.json 2>> $O/c_sweep.err; RB_GRID_WAVES=$w timeout 200 $B --workload cfg2_k17 > $O/c_k17_waves$w.json 2>> $O/c_sweep.err; done
timeout 400 python tools/readme_bench.py 100000 > $O/c_readme.json 2> $O/c_readme.err; tail -c 400 $O/c_readme.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:count_slots --launch-skip 4 --launch-count 1 -o $O/c_slots_cfg3 -f $B --workload cfg3_3.1Gb_31kbins > $O/c_ncu.log 2>&1
ls $O | grep "^c_" | wc -l
