#!/bin/bash
# round-2 GPU call C: fused sp shells -- full GPU suite, bench, fused vs segmented A/B, launch list
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=10 ) > gpurun_out/r2c_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2c_pytest_gpu.log
for f in 1 0; do
  RCHEM_FUSE_SP=$f AB_NOBASE=1 AB_COMBOS=11 timeout 600 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2c_ab_fuse$f.txt 2>&1
  RCHEM_FUSE_SP=$f AB_NOBASE=1 AB_COMBOS=11 timeout 600 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2c_ab_fuse$f.txt 2>&1
  RCHEM_FUSE_SP=$f AB_NOBASE=1 AB_COMBOS=11 timeout 600 python tools/ab_jk.py 32 6-31G 1e-10 >> gpurun_out/r2c_ab_fuse$f.txt 2>&1
  cat gpurun_out/r2c_ab_fuse$f.txt
done
timeout 900 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2c_bench_n1.json'));print(d['ms_per_step'],d['value'],d['roofline']['frac'],d['parity'],d['exact_boys']['ms_per_step'])"
tail -3 gpurun_out/r2c_bench_n1.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2c_launches_h2o96_631g_ref.csv python tools/prof_jk.py 96 6-31G 1e-10 0 2 > gpurun_out/r2c_ncu_launches.log 2>&1
tail -2 gpurun_out/r2c_ncu_launches.log
ls -la gpurun_out/
