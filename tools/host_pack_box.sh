#!/bin/bash
# host packer on the GPU box's cores (no GPU work): thread scaling next to the pool's pure stream-read rate
mkdir -p gpurun_out
O=gpurun_out/f_pack_box.txt
lscpu | grep -E "Model name|Thread|Core|Socket|L3|NUMA node\(s\)|MHz" > $O
g++ -std=c++17 -O3 -Ireadbouncer_b200/csrc tools/pack_bench.cpp readbouncer_b200/csrc/host_pack.cpp -lpthread -o /tmp/pack_bench
for t in 1 4 8 12 16; do
  echo "threads=$t" >> $O
  RB_HOST_THREADS=$t /tmp/pack_bench | sed -n 2,7p >> $O
done
for t in 1 4 8 12 16; do RB_HOST_THREADS=$t python - >> $O <<'P'
import sys; sys.path.insert(0,'.')
import numpy as np, readbouncer_b200 as rb
a=np.ones(512<<20,np.uint8); print(rb.host_pack_info()['threads'], "threads read-only GB/s %.1f"%rb.capi.host_read_gbs(a))
P
done
cat $O
