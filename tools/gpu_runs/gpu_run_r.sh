#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for u in 1 2 4; do
  RB_TABLE_U=$u timeout 300 python bench.py --workload w4_200x2Mb_200bins --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r_w4_u$u.json 2>> $O/r.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r_w4_u$u.json').read().strip().splitlines()[-1]); r=d['roofline']
print("U=$u value %.4g kernel %s %.3f ms frac %.3f"%(d['value'],r['kernel'],r['kernel_ms'],r['frac']))
P
done
tail -3 $O/r.err
