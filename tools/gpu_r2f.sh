#!/bin/bash
# round-2 GPU call F: dynamic chunk scheduling + direct reference-Boys rows + packed ket indices.
# GPU suite, A/B against the previous revision (librchem_b200_base.so) and single-change variants,
# heavy-threshold sweep, launch lists of the headline workload and of (H2O)32 / 6-31G*.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2f_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2f_pytest_gpu.log
V="nodirect|librchem_b200_nodirect.so|;boys_smem|librchem_b200_bsm.so|RCHEM_BOYS_SMEM=1;heavy1|| RCHEM_HEAVY_PASSES=1;heavy1.5||RCHEM_HEAVY_PASSES=1.5;heavy3||RCHEM_HEAVY_PASSES=3"
V=${V// /}
AB_COMBOS=11 AB_VARIANTS="$V" timeout 1500 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2f_ab.txt 2>&1
AB_COMBOS=11 timeout 600 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2f_ab.txt 2>&1
AB_COMBOS=11 timeout 600 python tools/ab_jk.py 32 6-31G 1e-10 >> gpurun_out/r2f_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 timeout 600 python tools/ab_jk.py 32 '6-31G*' 1e-10 >> gpurun_out/r2f_ab.txt 2>&1
cat gpurun_out/r2f_ab.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2f_launches_h2o96_631g_ref.csv python tools/prof_jk.py 96 6-31G 1e-10 0 2 > gpurun_out/r2f_ncu_launches.log 2>&1
tail -2 gpurun_out/r2f_ncu_launches.log
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2f_launches_h2o32_631gs_ref.csv python tools/prof_jk.py 32 '6-31G*' 1e-10 0 2 > gpurun_out/r2f_ncu_launches_631gs.log 2>&1
tail -2 gpurun_out/r2f_ncu_launches_631gs.log
ls -la gpurun_out/ | tail -8
