#!/bin/bash
# round-2 evidence run: smoke, full tests, default bench, ncu launch list of the bench command, ncu --set full of the top kernel
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py smoke > $O/i_smoke.log 2>&1; tail -2 $O/i_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/i_pytest_gpu.log 2>&1; tail -3 $O/i_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/i_bench_default.json 2> $O/i_bench_default.err; tail -c 300 $O/i_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/i_launches_cfg2.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $O/i_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_wgroup --launch-skip 6 --launch-count 1 -o $O/i_wgroup_cfg2 -f python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --no-e2e > $O/i_ncu_full.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_postings --launch-skip 4 --launch-count 1 -o $O/i_postings_cfg3 -f python bench.py --workload cfg3_3.1Gb_31kbins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/i_ncu_full3.log 2>&1
ls -la $O | grep " i_"
