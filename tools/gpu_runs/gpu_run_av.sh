#!/bin/bash
# 2-GPU check of the default bench line (read-sharded config #2 + config #5 secondary with the NCCL key combine) and the world-size-2 GPU tests
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/av_bench_default_2gpu.json 2> $O/av_bench_default_2gpu.err
python - <<P
import json
d=json.loads(open('gpurun_out/av_bench_default_2gpu.json').read().strip().splitlines()[-1]); r=d['roofline']
print("2 GPUs: value %.4g e2e %.4g kernel %s %.3f ms frac %.3f"%(d['value'],d['e2e']['value'],r['kernel'],r['kernel_ms'],r['frac']))
for s in d.get('secondary',[]):
    if 'roofline' in s:
        r=s['roofline']; print(" ", s['config']['workload'], "%.4g"%s['value'], r['kernel'], round(r['kernel_ms'],3), round(r['frac'],3), s['config'].get('collective_ms'), str(s.get('parity'))[:200])
    else: print(" ", s)
P
tail -n 3 $O/av_bench_default_2gpu.err
timeout 900 python -m pytest tests -q -m gpu -k "two_devices or sharded or nccl or shard_combine or multi" 2>&1 | tail -3
