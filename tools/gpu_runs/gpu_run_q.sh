#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for wl in cfg2_k14 cfg2_k16 w4_200x2Mb_200bins; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > $O/q_$wl.json 2>> $O/q.err
done
python - <<'P'
import json
for wl in ('cfg2_k14','cfg2_k16','w4_200x2Mb_200bins'):
    try: d=json.loads(open('gpurun_out/q_%s.json'%wl).read().strip().splitlines()[-1])
    except Exception as e: print(wl,'FAIL'); continue
    r=d['roofline']; c=d['config']
    print(wl, "value %.4g e2e %.4g kernel %s %.3f ms frac %.3f table %.3g GB span %d kind %d hit %.3f"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],c['kmer_table_bytes']/1e9,c['kmer_table_span'],c['kmer_table_kind'],c['hit_fraction']))
P
tail -3 $O/q.err
