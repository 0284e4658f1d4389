#!/bin/bash
# GPU box: the ncu captures the round's profiles/ summaries are made from (tools/summarise_profiles.py reads them back here).
#   gpurun -- 'bash tools/gpu_profile.sh r2'
# One launch list of the frame loop (device time per launch) + one `--set full` capture per kernel of the path.
# A number printed by a run under ncu is never a bench value.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
EXTRA=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $out/launches_$tag.csv python tools/prof_target.py frame 3 > $out/launches_$tag.log 2>&1
cap() {  # name, kernel regex, skip, target args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $out/prof_${name}_$tag python tools/prof_target.py "$@" > $out/prof_${name}_$tag.log 2>&1
}
cap align k_icp_align 3 icp 1
cap alloc k_alloc 4 frame 1
cap allocinsert k_alloc 0 alloc
cap allocsteady k_alloc 1 alloc
cap compact k_compact 4 frame 1
cap preprocess k_preprocess 4 frame 1
cap integrate k_integrate 1 integrate 3
cap gc 'k_gc$' 0 integrate 1
cap raycast k_raycast 4 raycast 1
ls -la $out/*_$tag.ncu-rep
