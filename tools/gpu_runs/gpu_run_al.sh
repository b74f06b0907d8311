#!/bin/bash
# final evidence run of round 2: smoke, full GPU suite, default bench line (1 GPU), launch list of the same command
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/al_smoke.log
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $O/al_pytest.log
cat $O/al_pytest.log
timeout 1200 python bench.py > $O/al_bench_default.json 2> $O/al_bench_default.err
python - <<P
import json
d=json.loads(open('gpurun_out/al_bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print("primary value %.4g e2e %.4g kernel %s %.3f ms frac %.3f cpu %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['cpu_baseline']['value']))
for s in d['secondary']:
    if 'roofline' in s:
        r=s['roofline']; print(" ", s['config']['workload'], "%.4g"%s['value'], "e2e %.4g"%s['e2e']['value'], r['kernel'], round(r['kernel_ms'],3), round(r['frac'],3), 'oracle' in str(s.get('parity')))
    else: print(" ", s.get('workload'), s.get('wall_s'), s.get('error'))
P
tail -n 3 $O/al_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/al_bench_reference.json 2>> $O/al_bench_default.err; tail -c 600 $O/al_bench_reference.json
