#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for t in 1 0 1 0; do
  RB_TAPER=$t RB_HOST_PACK=1 timeout 300 python bench.py --workload cfg2_100x4Mb_100bins --steps 30 --warmup 5 --no-cpu-baseline > $O/p_taper_$t.json 2>> $O/p.err
  python - <<P
import json
d=json.loads(open('gpurun_out/p_taper_$t.json').read().strip().splitlines()[-1]); e=d['e2e']; print("taper $t: e2e %.4g  %.3f ms"%(e['value'],e['ms_per_step']))
P
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/p_pytest_gpu.log 2>&1; tail -3 $O/p_pytest_gpu.log
