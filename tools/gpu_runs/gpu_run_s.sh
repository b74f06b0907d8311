#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --workload cfg2_k16 --steps 10 --warmup 3 --no-cpu-baseline > $O/s_k16.json 2>> $O/s.err
python - <<P
import json
d=json.loads(open('gpurun_out/s_k16.json').read().strip().splitlines()[-1]); r=d['roofline']
print("k16 value %.4g e2e %.4g kernel %s %.3f ms frac %.3f span %s table %.1f GB build %.0f ms parity %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_span'),d['config'].get('kmer_table_bytes',0)/1e9,d['config'].get('table_build_ms') or -1,d['config'].get('parity')))
P
tail -3 $O/s.err
