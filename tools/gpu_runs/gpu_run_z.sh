#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for w in w4_200x2Mb_200bins w5_30Mb_303bins w16_100Mb_1010bins w32_200Mb_2020bins; do
 for pf in 0 1; do
  RB_CTABLE_PF=$pf timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/z_${w}_pf$pf.json 2>> $O/z.err
  python - <<P
import json
d=json.loads(open('gpurun_out/z_${w}_pf$pf.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w pf=$pf value %.4g kernel_ms %.3f frac %.3f req frac %.3f"%(d['value'],r['kernel_ms'],r['frac'],r.get('requests',{}).get('frac',0)))
P
 done
done
tail -n 3 $O/z.err
