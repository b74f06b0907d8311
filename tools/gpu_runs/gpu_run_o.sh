#!/bin/bash
mkdir -p gpurun_out
RB_TRACE=2 RB_HOST_PACK=1 timeout 300 python bench.py --workload cfg2_100x4Mb_100bins --steps 3 --warmup 5 --no-cpu-baseline > gpurun_out/o_trace.json 2> gpurun_out/o_trace.err
grep -n "rb trace" gpurun_out/o_trace.err | tail -3
tail -40 gpurun_out/o_trace.err
