#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "medium_rows or count_matches_oracle or sharded_call" 2>&1 | tail -8 > $O/v_pytest.log
cat $O/v_pytest.log
RB_CTABLE_U=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_ctable --launch-skip 4 --launch-count 1 -o $O/v_ctable_w4 -f python bench.py --workload w4_200x2Mb_200bins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/v_ncu4.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:count_ctable --launch-skip 4 --launch-count 1 -o $O/v_ctable_w16 -f python bench.py --workload w16_100Mb_1010bins --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/v_ncu16.log 2>&1
tail -2 $O/v_ncu4.log $O/v_ncu16.log
