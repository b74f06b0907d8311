#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/sanitize_lane_groups.py 2>&1 | tail -2
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_lane_groups.py > $O/am_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 $O/am_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_lane_groups.py > $O/am_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 $O/am_racecheck.log
