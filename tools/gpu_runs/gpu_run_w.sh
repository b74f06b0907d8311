#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "medium_rows or count_matches_oracle or sharded_call" 2>&1 | tail -8 > $O/w_pytest.log
cat $O/w_pytest.log
for w in w4_200x2Mb_200bins w5_30Mb_303bins w16_100Mb_1010bins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/w_${w}.json 2>> $O/w.err
  python - <<P
import json
d=json.loads(open('gpurun_out/w_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s %.3f ms frac %.3f kind %s table %.2f GB req %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind'),d['config'].get('kmer_table_bytes',0)/1e9, {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.get('requests',{}).items() if k in ('peak_per_s','achieved_per_s','frac')}))
P
done
tail -3 $O/w.err
