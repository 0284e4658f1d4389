set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size" 2>&1 | tail -5
EXTRA=lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_conflicts_pipe_lsu_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum
for pair in "allocinsert 0" "allocsteady 1"; do set -- $pair; timeout 300 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k_alloc -s $2 -c 1 -f -o gpurun_out/prof_$1_r2 python tools/prof_target.py alloc > gpurun_out/prof_$1_r2.log 2>&1; tail -2 gpurun_out/prof_$1_r2.log; done
timeout 300 python bench.py --workload C5 --steps 100 --warmup 10 --no-cpu --no-hbm --no-refexact > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_c5.json').read());print('C5', round(d['value']), round(d['e2e']['value']), d['results'])"
timeout 300 python bench.py --workload C3 --steps 100 --warmup 10 --no-cpu --no-hbm --no-refexact > gpurun_out/r2_c3.json 2> gpurun_out/r2_c3.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_c3.json').read());print('C3', round(d['value']), round(d['e2e']['value']), d['results'], d['stages_us'])"
