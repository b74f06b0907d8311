#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "postings_slots or count_matches_oracle" 2>&1 | tail -6 > $O/ah_pytest.log
cat $O/ah_pytest.log
export RB_CTABLE=0
export RB_POSTINGS_LAYOUT=slots
for ws in "w5_30Mb_303bins 128" "w16_100Mb_1010bins 128" "w32_200Mb_2020bins 128" "w64_400Mb_4040bins 128" "w64_400Mb_4040bins 256" "w128_800Mb_8080bins 256" "w16_k15 128"; do
  set -- $ws; w=$1; sb=$2
  RB_SLOT_BYTES=$sb timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/ah_${w}_slot$sb.json 2>> $O/ah.err
  python - <<P
import json
d=json.loads(open('gpurun_out/ah_${w}_slot$sb.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w slot=$sb value %.4g kernel %s kernel_ms %.3f frac %.3f table %.2f GB build %s"%(d['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_bytes',0)/1e9,d['config'].get('kmer_table_build_ms')))
P
done
tail -n 3 $O/ah.err
