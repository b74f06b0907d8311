#!/bin/bash
# postings kernel: software-pipelined list groups (RB_POSTINGS_PIPE 0..3), parity then timing on config #3 and the config #5 shape
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "postings_long_lists" > $O/j_pytest.log 2>&1; tail -3 $O/j_pytest.log
B="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
for p in 0 1 2 3; do
  RB_POSTINGS_PIPE=$p timeout 200 $B --workload cfg3_3.1Gb_31kbins > $O/j_cfg3_pipe$p.json 2>> $O/j.err
done
for p in 0 1; do
  RB_POSTINGS_PIPE=$p timeout 300 $B --workload cfg5_3.7Gb_37kbins_per_gpu > $O/j_cfg5_pipe$p.json 2>> $O/j.err
done
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/j_cfg*_pipe*.json')):
    try: d=json.loads(open(p).read().strip().splitlines()[-1])
    except Exception: print(p,'FAIL'); continue
    print(p, "value %.4g kernel_ms %.3f frac %.3f"%(d['value'],d['roofline']['kernel_ms'],d['roofline']['frac']), d['clocks']['reasons'])
P
tail -5 $O/j.err
# e2e: cached plane stores + small staging rounds (planes served to the copy engine from the last-level cache?)
for cfg in "0 512" "1 512" "1 128" "1 64" "0 64"; do set -- $cfg
  RB_PACK_STORE=$1 RB_STAGE_MB=$2 RB_HOST_PACK=1 timeout 300 python bench.py --workload cfg2_100x4Mb_100bins --steps 20 --warmup 5 --no-cpu-baseline > $O/j_e2e_store$1_stage$2.json 2>> $O/j.err
done
python - <<'P'
import json,glob
for p in sorted(glob.glob('gpurun_out/j_e2e_*.json')):
    try: d=json.loads(open(p).read().strip().splitlines()[-1])
    except Exception: print(p,'FAIL'); continue
    e=d['e2e']; print(p, "e2e %.4g ms %.3f host_read %.0f"%(e['value'],e['ms_per_step'],e['host_read_gbs_per_rank_min']))
P
