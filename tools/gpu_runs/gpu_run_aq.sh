#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "postings or count_matches_oracle or sharded_call" 2>&1 | tail -4 > $O/aq_pytest.log
cat $O/aq_pytest.log
for w in cfg3_3.1Gb_31kbins w32_200Mb_2020bins w16_k15 w64_400Mb_4040bins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/aq_${w}.json 2>> $O/aq.err
  python - <<P
import json
d=json.loads(open('gpurun_out/aq_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g kernel %s kernel_ms %.3f frac %.3f"%(d['value'],r['kernel'],r['kernel_ms'],r['frac']))
P
done
tail -n 3 $O/aq.err
