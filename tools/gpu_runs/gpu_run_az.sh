#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for w in w5_30Mb_303bins w16_100Mb_1010bins w32_200Mb_2020bins; do
 for gw in 2 8 16 64; do
  RB_GRID_WAVES=$gw timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/az_${w}_gw$gw.json 2>> $O/az.err
  python - <<P
import json
d=json.loads(open('gpurun_out/az_${w}_gw$gw.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w waves=$gw value %.4g kernel_ms %.3f"%(d['value'],r['kernel_ms']))
P
 done
done
tail -n 2 $O/az.err
