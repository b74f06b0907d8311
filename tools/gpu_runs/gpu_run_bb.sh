#!/bin/bash
# 4-GPU line with the final kernels (completes 1 / 2 / 4 / 8)
mkdir -p gpurun_out
O=gpurun_out
N=${1:-4}
timeout 125 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > $O/bb_bench_default_${N}gpu.json 2> $O/bb_bench_default_${N}gpu.err
python - <<P
import json
d=json.loads(open('gpurun_out/bb_bench_default_${N}gpu.json').read().strip().splitlines()[-1]); r=d['roofline']
print("$N GPUs: value %.4g e2e %.4g kernel %s %.3f ms frac %.3f h2d ceiling %s"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac'],d['e2e'].get('h2d_ceiling_gbs')))
for s in d.get('secondary',[]):
    if 'roofline' in s:
        r=s['roofline']; print(" ", s['config']['workload'], "%.4g"%s['value'], r['kernel'], round(r['kernel_ms'],3), round(r['frac'],3), {k:v for k,v in s['config'].items() if 'collective' in k or 'allreduce' in k})
P
tail -n 2 $O/bb_bench_default_${N}gpu.err
