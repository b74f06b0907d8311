#!/bin/bash
# e2e: share of the pieces shipped as ASCII next to the packed ones (pinned input), now that packed pieces are 2 planes
mkdir -p gpurun_out
O=gpurun_out
for sh in 0 0.15 0.25 0.35 0.5; do
  RB_ASCII_SHARE=$sh RB_HOST_PACK=1 timeout 300 python bench.py --workload cfg2_100x4Mb_100bins --steps 20 --warmup 5 --no-cpu-baseline > $O/h_share_$sh.json 2>> $O/h.err
done
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/h_share_*.json')):
    try: d=json.loads(open(p).read().strip().splitlines()[-1])
    except Exception: print(p,'FAIL'); continue
    e=d['e2e']; print(p, "e2e %.4g ms %.3f h2d %.1f MB host_read %.0f"%(e['value'],e['ms_per_step'],e['h2d_bytes_per_step']/1e6,e['host_read_gbs_per_rank_min']))
P
