#!/bin/bash
# one GPU-box call: parity tests, postings A/B on cfg3, default bench, ncu capture of the postings kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/m_pytest_gpu.log
tail -5 gpurun_out/m_pytest_gpu.log
timeout 600 python tools/postings_ab.py > gpurun_out/m_postings_ab.jsonl 2> gpurun_out/m_postings_ab.err; tail -5 gpurun_out/m_postings_ab.jsonl; tail -3 gpurun_out/m_postings_ab.err
timeout 600 python bench.py > gpurun_out/m_bench_cfg2.json 2> gpurun_out/m_bench_cfg2.err; cat gpurun_out/m_bench_cfg2.json | cut -c1-1800; tail -3 gpurun_out/m_bench_cfg2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:count_postings -s 3 -c 1 -o gpurun_out/m_postings_cfg3 -f \
  python bench.py --workload cfg3_3.1Gb_31kbins --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/m_ncu_cfg3.log 2>&1; tail -3 gpurun_out/m_ncu_cfg3.log
