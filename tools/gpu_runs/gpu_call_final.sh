#!/bin/bash
# what the driver runs at round end, plus the launch list: GPU tests, smoke, bench (ours + reference arm), ncu launch list
mkdir -p gpurun_out
T=${1:-n}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err; cut -c1-250 gpurun_out/${T}_bench_cfg2.json; tail -2 gpurun_out/${T}_bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_cfg2_reference_arm.json 2> gpurun_out/${T}_ref.err; cut -c1-250 gpurun_out/${T}_bench_cfg2_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | cut -c1-120
timeout 600 python bench.py --workload cfg3_3.1Gb_31kbins > gpurun_out/${T}_bench_cfg3.json 2> gpurun_out/${T}_bench_cfg3.err; cut -c1-250 gpurun_out/${T}_bench_cfg3.json
