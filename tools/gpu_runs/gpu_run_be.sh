#!/bin/bash
# compute-sanitizer over the narrow-filter kernels, full workload list of tools/sanitize_narrow.py
mkdir -p gpurun_out
O=gpurun_out
timeout 85 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_narrow.py > $O/be_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 $O/be_memcheck.log
timeout 85 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_narrow.py > $O/be_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 $O/be_racecheck.log
