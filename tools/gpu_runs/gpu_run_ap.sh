#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "medium_rows or broken_max or count_matches_oracle" 2>&1 | tail -4 > $O/ap_pytest.log
cat $O/ap_pytest.log
for w in w4_200x2Mb_200bins w5_30Mb_303bins w16_100Mb_1010bins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ap_${w}.json 2>> $O/ap.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ap_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g kernel %s kernel_ms %.3f frac %.3f req %.3f"%(d['value'],r['kernel'],r['kernel_ms'],r['frac'],r.get('requests',{}).get('frac',0)))
P
done
tail -n 3 $O/ap.err
