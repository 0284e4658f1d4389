timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "alloc or fixed_sequence or overflow or capacity or fuzz or full_size or partition_union" 2>&1 | tail -4
for c in 144 140 136 132 128 120; do VH_ICP_CTAS=$c timeout 200 python bench.py --steps 300 --warmup 20 --repeats 5 --no-cpu --no-hbm --no-refexact > gpurun_out/r2_sw.json 2>gpurun_out/r2_sw.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_sw.json').read());print('ctas=$c', round(d['value']), round(d['e2e']['value']), [round(x/300*1000,1) for x in d['passes']['timed_ms']])"; done
EXTRA=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_red.sum
timeout 300 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k_alloc -s 0 -c 1 -f -o gpurun_out/prof_allocinsert_r2 python tools/prof_target.py alloc > gpurun_out/prof_allocinsert_r2.log 2>&1; tail -1 gpurun_out/prof_allocinsert_r2.log
