#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "postings or count_matches_oracle" 2>&1 | tail -3 > $O/at_pytest.log
cat $O/at_pytest.log
for w in w32_200Mb_2020bins w64_400Mb_4040bins w128_800Mb_8080bins w16_k15 cfg3_3.1Gb_31kbins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/at_${w}.json 2>> $O/at.err
  python - <<P
import json
d=json.loads(open('gpurun_out/at_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s kernel_ms %.3f frac %.3f"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac']))
P
done
tail -n 3 $O/at.err
