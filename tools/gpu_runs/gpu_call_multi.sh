#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
nproc > gpurun_out/r_host_${N}gpu.txt; free -g >> gpurun_out/r_host_${N}gpu.txt
timeout 600 $TR bench.py --gpus $N > gpurun_out/r_bench_cfg2_${N}gpu.json 2> gpurun_out/r_bench_cfg2_${N}gpu.err; grep '^{' gpurun_out/r_bench_cfg2_${N}gpu.json | cut -c1-200; tail -2 gpurun_out/r_bench_cfg2_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --workload cfg5_3.7Gb_37kbins_per_gpu > gpurun_out/r_bench_cfg5_${N}gpu.json 2> gpurun_out/r_bench_cfg5_${N}gpu.err; grep '^{' gpurun_out/r_bench_cfg5_${N}gpu.json | cut -c1-300; tail -2 gpurun_out/r_bench_cfg5_${N}gpu.err
