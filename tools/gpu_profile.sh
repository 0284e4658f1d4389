#!/bin/bash
# ncu passes of the profiling recipe (B200_PROFILING.md); run under gpurun, outputs under gpurun_out/.
#   bash tools/gpu_profile.sh r1     then here:  python tools/summarise_profiles.py r1
set -x
R=${1:-r1}
cd "$(dirname "$0")/.."
NCU="ncu --clock-control none"
# every launch of a few steady-state frames with its device time (cold-cache, serialised: compare SHARES)
$NCU --metrics gpu__time_duration.sum -s 60 -c 120 --csv --log-file gpurun_out/launches_${R}.csv python tools/prof_target.py frame 2 > gpurun_out/launches_${R}.log 2>&1
# top kernels, full sets
$NCU --set full --import-source on -k regex:k_integrate -s 1 -c 1 -f -o gpurun_out/prof_integrate_${R} python tools/prof_target.py integrate 3 > gpurun_out/prof_integrate_${R}.log 2>&1
$NCU --set full --import-source on -k regex:k_gc$ -c 1 -f -o gpurun_out/prof_gc_${R} python tools/prof_target.py integrate 1 > gpurun_out/prof_gc_${R}.log 2>&1
$NCU --set full --import-source on -k regex:k_icp_iter -s 25 -c 1 -f -o gpurun_out/prof_icp_${R} python tools/prof_target.py icp 1 > gpurun_out/prof_icp_${R}.log 2>&1
$NCU --set full --import-source on -k regex:k_alloc -s 4 -c 1 -f -o gpurun_out/prof_alloc_${R} python tools/prof_target.py icp 1 > gpurun_out/prof_alloc_${R}.log 2>&1
$NCU --set full --import-source on -k regex:k_compact -s 4 -c 1 -f -o gpurun_out/prof_compact_${R} python tools/prof_target.py icp 1 > gpurun_out/prof_compact_${R}.log 2>&1
$NCU --set full --import-source on -k regex:k_raycast -s 4 -c 1 -f -o gpurun_out/prof_raycast_${R} python tools/prof_target.py raycast 1 > gpurun_out/prof_raycast_${R}.log 2>&1
ls -la gpurun_out/ | tail -20
