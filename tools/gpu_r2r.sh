#!/bin/bash
# round-2 GPU call R: launch list of (H2O)32/6-31G* on the current kernels
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2r_launches_h2o32_631gs_ref.csv python tools/prof_jk.py 32 '6-31G*' 1e-10 0 2 > gpurun_out/r2r_ncu_launches_631gs.log 2>&1
tail -2 gpurun_out/r2r_ncu_launches_631gs.log
