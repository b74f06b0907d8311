#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg3_3.1Gb_31kbins --no-cpu-baseline > gpurun_out/o_bench_cfg3.json 2> gpurun_out/o_bench_cfg3.err; cut -c1-330 gpurun_out/o_bench_cfg3.json; tail -3 gpurun_out/o_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:count_postings -s 3 -c 1 -o gpurun_out/o_postings_cfg3 -f \
  python bench.py --workload cfg3_3.1Gb_31kbins --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/o_ncu_cfg3.log 2>&1; tail -2 gpurun_out/o_ncu_cfg3.log | cut -c1-200
