#!/bin/bash
# host packer speed on the GPU box's cores (no GPU work), then the default bench line
mkdir -p gpurun_out
g++ -std=c++17 -O2 -Ireadbouncer_b200/csrc tools/pack_bench.cpp readbouncer_b200/csrc/host_pack.o -lpthread -o /tmp/pack_bench
/tmp/pack_bench > gpurun_out/pack_bench.txt 2>&1
RB_HOST_THREADS=8 /tmp/pack_bench >> gpurun_out/pack_bench.txt 2>&1
cat gpurun_out/pack_bench.txt
