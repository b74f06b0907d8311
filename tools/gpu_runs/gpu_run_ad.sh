#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
export RB_CTABLE=0
for w in w64_400Mb_4040bins w128_800Mb_8080bins; do
 for sub in 2 4; do
  RB_POSTINGS_SUB=$sub timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ab_${w}_sub$sub.json 2>> $O/ad.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ab_${w}_sub$sub.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w sub=$sub value %.4g kernel %s kernel_ms %.3f frac %.3f table %.2f GB"%(d['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_bytes',0)/1e9))
P
 done
done
tail -n 3 $O/ad.err
