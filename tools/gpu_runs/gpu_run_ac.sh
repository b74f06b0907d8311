#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "postings_long_lists or count_matches_oracle" 2>&1 | tail -6 > $O/ac_pytest.log
cat $O/ac_pytest.log
export RB_CTABLE=0
for w in w16_k15 w5_30Mb_303bins w16_100Mb_1010bins w32_200Mb_2020bins; do
 for sub in 0 2 4 8; do
  RB_POSTINGS_SUB=$sub timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ac_${w}_sub$sub.json 2>> $O/ac.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ac_${w}_sub$sub.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w sub=$sub value %.4g kernel %s kernel_ms %.3f frac %.3f table %.2f GB"%(d['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_bytes',0)/1e9))
P
 done
done
tail -n 3 $O/ac.err
