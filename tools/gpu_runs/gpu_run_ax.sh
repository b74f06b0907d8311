#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "edge_batches or broken_max" 2>&1 | tail -4 > $O/ax_pytest.log
cat $O/ax_pytest.log
