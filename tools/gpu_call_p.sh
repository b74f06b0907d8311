#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/p_pytest_gpu.log
tail -4 gpurun_out/p_pytest_gpu.log
timeout 300 python bench.py --workload mini5_40Mb_512bins_per_gpu --no-cpu-baseline > gpurun_out/p_bench_mini5.json 2> gpurun_out/p_bench_mini5.err; cut -c1-600 gpurun_out/p_bench_mini5.json; tail -3 gpurun_out/p_bench_mini5.err
timeout 600 python bench.py --workload cfg3_3.1Gb_31kbins --no-cpu-baseline > gpurun_out/p_bench_cfg3.json 2> gpurun_out/p_bench_cfg3.err; cut -c1-330 gpurun_out/p_bench_cfg3.json; tail -3 gpurun_out/p_bench_cfg3.err
timeout 900 python bench.py --workload cfg5_3.7Gb_37kbins_per_gpu --no-cpu-baseline > gpurun_out/p_bench_cfg5_1gpu.json 2> gpurun_out/p_bench_cfg5_1gpu.err; cut -c1-800 gpurun_out/p_bench_cfg5_1gpu.json; tail -3 gpurun_out/p_bench_cfg5_1gpu.err
