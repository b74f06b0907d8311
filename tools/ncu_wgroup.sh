#!/bin/bash
# ncu --set full capture of the two variants of the group-per-read kernel at cfg #2: ASCII input (device-resident bench path)
# and packed bit planes (host-buffer call).  Summaries: tools/ncu_summary.py gpurun_out/u_wgroup_*.ncu-rep
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 400 ncu $COMMON -k 'regex:count_wgroup_kernel<.*\(bool\)0>' -s 3 -c 1 -o gpurun_out/u_wgroup_ascii -f \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/u1.log 2>&1
grep -c "==PROF== Profiling" gpurun_out/u1.log
RB_HOST_PACK=1 timeout 400 ncu $COMMON -k 'regex:count_wgroup_kernel<.*\(bool\)1>' -s 45 -c 1 -o gpurun_out/u_wgroup_packed -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/u2.log 2>&1
grep -c "==PROF== Profiling" gpurun_out/u2.log
