#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export RB_CTABLE=0
for ws in "w5_30Mb_303bins 2" "w32_200Mb_2020bins 2" "w32_200Mb_2020bins 4" "w64_400Mb_4040bins 4" "w64_400Mb_4040bins 8" "w128_800Mb_8080bins 4" "w128_800Mb_8080bins 8" "w256_1.6Gb_16160bins 8"; do
  set -- $ws; w=$1; sub=$2
  RB_POSTINGS_SUB=$sub timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ae_${w}_sub$sub.json 2>> $O/ae.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ae_${w}_sub$sub.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w sub=$sub value %.4g kernel_ms %.3f frac %.3f"%(d['value'],r['kernel_ms'],r['frac']))
P
done
tail -n 3 $O/ae.err
