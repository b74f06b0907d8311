#!/bin/bash
# two more families of random filter / batch shapes through the default paths against the oracle
mkdir -p gpurun_out
O=gpurun_out
for base in 3000 7000; do
  RB_FUZZ_BASE=$base timeout 80 python -m pytest tests/test_gpu_random_shapes.py -q -m gpu 2>&1 | tail -4 > $O/bf_pytest_fuzz_$base.log; cat $O/bf_pytest_fuzz_$base.log
done
