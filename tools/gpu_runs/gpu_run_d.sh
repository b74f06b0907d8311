#!/bin/bash
# GPU run D: defaults (lists layout, 16 grid waves) re-verified, waves sweep, full default bench, sanitizer
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/d_pytest_gpu.log 2>&1; tail -3 $O/d_pytest_gpu.log
B="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline"
for w in 1 16 32 64 128; do
  for wl in cfg2_100x4Mb_100bins cfg2_k15 cfg2_k17 cfg3_3.1Gb_31kbins; do
    RB_GRID_WAVES=$w timeout 200 $B --workload $wl > $O/d_${wl}_waves$w.json 2>> $O/d_sweep.err
  done
done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/d_bench_default.json 2> $O/d_bench_default.err; tail -c 300 $O/d_bench_default.err
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "postings_slots and 9000-1 and 256 or sharded_call and 3-slots or sharded_call and 2-" > $O/d_racecheck.log 2>&1; tail -4 $O/d_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "postings_slots and 9000-1 and 128 or sharded_call and 3-lists" > $O/d_memcheck.log 2>&1; tail -4 $O/d_memcheck.log
ls $O | grep "^d_" | wc -l
