#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/n_pytest_gpu.log
tail -4 gpurun_out/n_pytest_gpu.log
timeout 600 python bench.py --workload cfg3_3.1Gb_31kbins --no-cpu-baseline > gpurun_out/n_bench_cfg3.json 2> gpurun_out/n_bench_cfg3.err; cut -c1-400 gpurun_out/n_bench_cfg3.json; tail -3 gpurun_out/n_bench_cfg3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:count_postings -s 3 -c 1 -o gpurun_out/n_postings_cfg3 -f \
  python bench.py --workload cfg3_3.1Gb_31kbins --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/n_ncu_cfg3.log 2>&1; tail -3 gpurun_out/n_ncu_cfg3.log
