#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/e_pytest_gpu.log 2>&1; tail -3 $O/e_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/e_bench_default.json 2> $O/e_bench_default.err; tail -c 300 $O/e_bench_default.err
python __graft_entry__.py smoke > $O/e_smoke.log 2>&1; tail -2 $O/e_smoke.log
