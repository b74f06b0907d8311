#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 > gpurun_out/q_bench_cfg2_2gpu.json 2> gpurun_out/q_bench_cfg2_2gpu.err; grep '^{' gpurun_out/q_bench_cfg2_2gpu.json | cut -c1-200; tail -2 gpurun_out/q_bench_cfg2_2gpu.err
timeout 900 $TR bench.py --gpus 2 --workload cfg5_3.7Gb_37kbins_per_gpu > gpurun_out/q_bench_cfg5_2gpu.json 2> gpurun_out/q_bench_cfg5_2gpu.err; grep '^{' gpurun_out/q_bench_cfg5_2gpu.json | cut -c1-500; tail -2 gpurun_out/q_bench_cfg5_2gpu.err
