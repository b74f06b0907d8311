#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "postings or count_matches_oracle or sharded_call" 2>&1 | tail -5 > $O/ak_pytest.log
cat $O/ak_pytest.log
for w in w32_200Mb_2020bins w64_400Mb_4040bins w128_800Mb_8080bins w16_k15; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ak_${w}.json 2>> $O/ak.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ak_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g kernel %s kernel_ms %.3f frac %.3f kind %s"%(d['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind')))
P
done
for w in w5_30Mb_303bins w32_200Mb_2020bins w64_400Mb_4040bins; do
  RB_CTABLE=0 RB_POSTINGS_LAYOUT=lists timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ak_${w}_lists.json 2>> $O/ak.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ak_${w}_lists.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w LISTS value %.4g kernel %s kernel_ms %.3f frac %.3f kind %s"%(d['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind')))
P
done
tail -n 3 $O/ak.err
