#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "postings or table or shard" > gpurun_out/s_pytest.log 2>&1; tail -2 gpurun_out/s_pytest.log
timeout 600 python bench.py --workload cfg3_3.1Gb_31kbins --no-cpu-baseline --no-e2e > gpurun_out/s_bench_cfg3.json 2> gpurun_out/s_bench_cfg3.err; cut -c1-330 gpurun_out/s_bench_cfg3.json; tail -3 gpurun_out/s_bench_cfg3.err
