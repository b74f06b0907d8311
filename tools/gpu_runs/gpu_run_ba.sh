#!/bin/bash
# rehearsal of the driver's round-end sequence on the final tree: smoke, full GPU suite, reference arm, default bench line
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/ba_smoke.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $O/ba_pytest.log
cat $O/ba_pytest.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/ba_bench_reference.json 2> $O/ba_bench_reference.err; tail -c 400 $O/ba_bench_reference.json
timeout 600 python bench.py > $O/ba_bench_default.json 2> $O/ba_bench_default.err
python - <<P
import json
d=json.loads(open('gpurun_out/ba_bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print("primary value %.4g e2e %.4g kernel %s %.3f ms frac %.3f cpu %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['cpu_baseline']['value']))
for s in d['secondary']:
    if 'roofline' in s:
        r=s['roofline']; print(" ", s['config']['workload'], "%.4g"%s['value'], "e2e %.4g"%s['e2e']['value'], r['kernel'], round(r['kernel_ms'],3), round(r['frac'],3), 'oracle' in str(s.get('parity')))
    else: print(" ", s.get('workload'), s.get('wall_s'), s.get('error'))
P
tail -n 3 $O/ba_bench_default.err
