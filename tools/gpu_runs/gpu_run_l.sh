#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_random_shapes.py -m gpu -q --durations=5 > $O/l_pytest_random.log 2>&1; tail -12 $O/l_pytest_random.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/l_bench_default.json 2> $O/l_bench_default.err; tail -c 300 $O/l_bench_default.err
