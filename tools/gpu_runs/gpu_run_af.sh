#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/af_pytest.log
cat $O/af_pytest.log
for w in w32_200Mb_2020bins w64_400Mb_4040bins w128_800Mb_8080bins w16_k15; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/af_${w}.json 2>> $O/af.err
  python - <<P
import json
d=json.loads(open('gpurun_out/af_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s kernel_ms %.3f frac %.3f kind %s table %.2f GB"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind'),d['config'].get('kmer_table_bytes',0)/1e9))
P
done
tail -n 3 $O/af.err
