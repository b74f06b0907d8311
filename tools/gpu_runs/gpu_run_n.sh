#!/bin/bash
# e2e: piece size sweep now that packed pieces are 2 planes and the packer is cheaper
mkdir -p gpurun_out
O=gpurun_out
for mb in 2 4 8 16 32 64; do
  RB_PIECE_MB=$mb RB_HOST_PACK=1 timeout 300 python bench.py --workload cfg2_100x4Mb_100bins --steps 20 --warmup 5 --no-cpu-baseline > $O/n_piece_$mb.json 2>> $O/n.err
done
python - <<'P'
import json,glob
for mb in (2,4,8,16,32,64):
    try: d=json.loads(open('gpurun_out/n_piece_%d.json'%mb).read().strip().splitlines()[-1])
    except Exception: print(mb,'FAIL'); continue
    e=d['e2e']; print("piece %2d MB: e2e %.4g  %.3f ms  (device %.3f ms) host_read %.0f"%(mb,e['value'],e['ms_per_step'],d['roofline']['kernel_ms'],e['host_read_gbs_per_rank_min']))
P
