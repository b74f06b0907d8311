#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for w in w32_200Mb_2020bins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/x_${w}.json 2>> $O/y.err
  python - <<P
import json
d=json.loads(open('gpurun_out/x_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s %.3f ms frac %.3f kind %s table %.2f GB build %s ms req %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind'),d['config'].get('kmer_table_bytes',0)/1e9, d['config'].get('kmer_table_build_ms'), {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.get('requests',{}).items() if k in ('peak_per_s','achieved_per_s','frac')}))
P
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_ctable_bs --launch-skip 4 --launch-count 1 -o $O/y_ctable_bs_w5 -f python bench.py --workload w5_30Mb_303bins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/y_ncu5.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_ctable_bs --launch-skip 4 --launch-count 1 -o $O/y_ctable_bs_w16 -f python bench.py --workload w16_100Mb_1010bins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/y_ncu16.log 2>&1
tail -n 3 $O/y.err
