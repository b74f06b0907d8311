#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "count_matches_oracle or packed or two_threshold" 2>&1 | tail -4 > $O/ar_pytest.log
cat $O/ar_pytest.log
for w in cfg2_100x4Mb_100bins cfg1_5Mb_51bins cfg2_k14; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/ar_${w}.json 2>> $O/ar.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ar_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s kernel_ms %.3f frac %.3f req %.3f"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],r.get('requests',{}).get('frac',0)))
P
done
tail -n 3 $O/ar.err
