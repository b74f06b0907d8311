#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "medium_rows or count_matches_oracle or sharded_call or postings" 2>&1 | tail -6 > $O/x_pytest.log
cat $O/x_pytest.log
for w in w32_200Mb_2020bins w64_400Mb_4040bins; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/x_${w}.json 2>> $O/x.err
  python - <<P
import json
d=json.loads(open('gpurun_out/x_${w}.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$w value %.4g e2e %.4g kernel %s %.3f ms frac %.3f kind %s table %.2f GB build %s ms req %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('kmer_table_kind'),d['config'].get('kmer_table_bytes',0)/1e9, d['config'].get('kmer_table_build_ms'), {k:(round(v,3) if isinstance(v,float) else v) for k,v in r.get('requests',{}).items() if k in ('peak_per_s','achieved_per_s','frac')}))
P
done
RB_CTABLE=0 timeout 600 python bench.py --workload w32_200Mb_2020bins --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/x_w32_postings.json 2>> $O/x.err
python - <<P
import json
d=json.loads(open('gpurun_out/x_w32_postings.json').read().strip().splitlines()[-1]); r=d['roofline']
print("w32 postings value %.4g kernel %s %.3f ms frac %.3f"%(d['value'],r['kernel'],r['kernel_ms'],r['frac']))
P
tail -3 $O/x.err
