#!/bin/bash
# read mix: all-random (no chunk comes from the reference) and all-reference, final kernels, configs #2 and #3 (SURVEY 8d asks for both mixes)
mkdir -p gpurun_out
O=gpurun_out
for mix in 0.0 1.0; do
  timeout 100 python bench.py --no-secondary --no-cpu-baseline --from-ref $mix > $O/bc_bench_cfg2_fromref_$mix.json 2> $O/bc.err
  timeout 100 python bench.py --workload cfg3_3.1Gb_31kbins --no-secondary --no-cpu-baseline --from-ref $mix > $O/bc_bench_cfg3_fromref_$mix.json 2>> $O/bc.err
done
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/bc_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], "value %.4g e2e %.4g %s %.3f ms frac %.3f hit_fraction %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['config'].get('hit_fraction')))
    except Exception as e: print(f, 'failed', e)
P
tail -n 3 $O/bc.err
