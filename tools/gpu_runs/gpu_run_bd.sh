#!/bin/bash
# compute-sanitizer over the narrow-filter kernels in their final form (tools/sanitize_narrow.py)
mkdir -p gpurun_out
O=gpurun_out
timeout 60 python tools/sanitize_narrow.py > $O/bd_plain.log 2>&1; echo "plain rc=$?"; tail -n 3 $O/bd_plain.log
timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_narrow.py quick > $O/bd_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 $O/bd_memcheck.log
timeout 75 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_narrow.py quick > $O/bd_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 $O/bd_racecheck.log
