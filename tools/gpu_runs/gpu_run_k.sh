#!/bin/bash
# usage=target shape: micro-batch latency through LiveClassifier with 3 target + 1 depletion filter, filters serial vs concurrent
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "drivers or live or shim or unblock" > $O/k_pytest.log 2>&1; tail -3 $O/k_pytest.log
L=readbouncer_b200/bin/rb_live_bench
timeout 300 $L 3 1 4000000 200 64 512 4096 32768 > $O/k_live_concurrent.jsonl 2> $O/k_live.err
RB_FILTERS_SERIAL=1 timeout 300 $L 3 1 4000000 200 64 512 4096 32768 > $O/k_live_serial.jsonl 2>> $O/k_live.err
timeout 300 $L 0 1 4000000 200 64 512 4096 > $O/k_live_deplete_only.jsonl 2>> $O/k_live.err
cat $O/k_live_concurrent.jsonl $O/k_live_serial.jsonl $O/k_live_deplete_only.jsonl; tail -3 $O/k_live.err
timeout 300 python tools/readme_bench.py 100000 > $O/k_readme_concurrent.json 2>> $O/k_live.err
RB_FILTERS_SERIAL=1 timeout 300 python tools/readme_bench.py 100000 > $O/k_readme_serial.json 2>> $O/k_live.err
python - <<'P'
import json
for f in ('k_readme_concurrent','k_readme_serial'):
    try: d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['s_per_read'], d['driver_wall_s'])
    except Exception as e: print(f,'FAIL',e)
P
