#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for w in w5_30Mb_303bins w16_100Mb_1010bins w64_400Mb_4040bins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/t_$w.json 2>> $O/t.err
  python - <<P
import json
d=json.loads(open('gpurun_out/t_$w.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s %.3f ms frac %.3f kind %s table %.2f GB parity %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind'),d['config'].get('kmer_table_bytes',0)/1e9,str(d.get('parity') or d['config'].get('parity'))[:300]))
P
done
tail -3 $O/t.err
