#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "postings_long_lists or count_matches_oracle or sharded_call" 2>&1 | tail -6 > $O/aa_pytest.log
cat $O/aa_pytest.log
for w in w64_400Mb_4040bins cfg3_3.1Gb_31kbins; do
 for sub in 0 8 16; do
  RB_POSTINGS_SUB=$sub timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/aa_${w}_sub$sub.json 2>> $O/aa.err
  python - <<P
import json
d=json.loads(open('gpurun_out/aa_${w}_sub$sub.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w sub=$sub value %.4g kernel_ms %.3f frac %.3f table %.2f GB"%(d['value'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_bytes',0)/1e9))
P
 done
done
tail -n 3 $O/aa.err
