#!/bin/bash
# GPU run B: slot-layout postings -- parity, A/B against the list layout, ring / CTA sweep, ncu capture
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $O/b_pytest_gpu.log 2>&1; tail -3 $O/b_pytest_gpu.log
B="python bench.py --workload cfg3_3.1Gb_31kbins --steps 8 --warmup 3 --no-e2e"
timeout 300 $B > $O/b_cfg3_slots.json 2> $O/b_cfg3_slots.err
RB_POSTINGS_LAYOUT=lists timeout 300 $B > $O/b_cfg3_lists.json 2> $O/b_cfg3_lists.err
for ring in 1 2 3 4 6 8; do for ctas in 1 2; do
  RB_SLOT_RING=$ring RB_SLOT_CTAS=$ctas timeout 200 $B --no-cpu-baseline > $O/b_cfg3_ring${ring}_ctas${ctas}.json 2>> $O/b_sweep.err
done; done
for sb in 768 896 1024; do
  RB_SLOT_BYTES=$sb timeout 200 $B --no-cpu-baseline > $O/b_cfg3_slot${sb}.json 2>> $O/b_sweep.err
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:count_slots --launch-skip 4 --launch-count 1 -o $O/b_slots_cfg3 -f $B --no-cpu-baseline > $O/b_ncu.log 2>&1
timeout 300 python tools/exp_streams.py > $O/b_exp_streams.jsonl 2> $O/b_exp_streams.err
ls -la $O | tail -30
