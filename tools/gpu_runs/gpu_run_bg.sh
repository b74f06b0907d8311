#!/bin/bash
# ten more families of random filter / batch shapes (240 shapes) through the default paths against the oracle
mkdir -p gpurun_out
O=gpurun_out
: > $O/bg_pytest_fuzz.log
for base in 11000 12000 13000 14000 15000 16000 17000 18000 19000 20000; do
  echo "RB_FUZZ_BASE=$base" >> $O/bg_pytest_fuzz.log
  RB_FUZZ_BASE=$base timeout 30 python -m pytest tests/test_gpu_random_shapes.py -q -m gpu 2>&1 | tail -3 >> $O/bg_pytest_fuzz.log
done
grep -c passed $O/bg_pytest_fuzz.log; grep -i "fail\|error" $O/bg_pytest_fuzz.log | head
