#!/bin/bash
# piece-size sweep of the host-buffer call (gpurun_out/)
mkdir -p gpurun_out
for mb in 4 8 16; do
  RB_TRACE=1 RB_PIECE_MB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/sw_trace_$mb.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('piece_mb', $mb, 'e2e_ms', d['e2e']['ms_per_step'], 'dev_ms', d['ms_per_step'], 'h2d', d['e2e']['h2d_bytes_per_step'])"
  tail -1 gpurun_out/sw_trace_$mb.err
done
