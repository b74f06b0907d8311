#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $O/ag_pytest.log
cat $O/ag_pytest.log
