#!/bin/bash
# multi-GPU run (gpurun --gpus N): the tests that need >= 2 devices, then the default bench line at N ranks
# (primary: config #2 read-sharded; secondary: config #5's per-GPU workload with the NCCL all-reduce in the timed region)
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/m${N}_topo.txt 2>&1
nproc > $O/m${N}_host.txt; free -g >> $O/m${N}_host.txt
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "two_devices or sharded or shard_combine or dist" > $O/m${N}_pytest.log 2>&1; tail -3 $O/m${N}_pytest.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $O/m${N}_bench.json 2> $O/m${N}_bench.err
tail -c 600 $O/m${N}_bench.err; wc -c $O/m${N}_bench.json
