#!/bin/bash
# round-2 GPU call G: high-L classes emitted in parts (no ptxas fallback), highest classes launched first,
# medium light kernels at 3 blocks/SM (variant), nq histogram, ncu --set full of the new block / light kernels
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2g_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2g_pytest_gpu.log
V="order0||RCHEM_TASK_ORDER=0;lm3|librchem_b200_lm3.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 > gpurun_out/r2g_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 >> gpurun_out/r2g_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 6-31G 1e-10 >> gpurun_out/r2g_ab.txt 2>&1
cat gpurun_out/r2g_ab.txt
timeout 600 python tools/nq_hist.py 96 6-31G 1e-10 > gpurun_out/r2g_nq_hist.txt 2>&1
cat gpurun_out/r2g_nq_hist.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2g_launches_h2o32_631gs_ref.csv python tools/prof_jk.py 32 '6-31G*' 1e-10 0 2 > gpurun_out/r2g_ncu_launches_631gs.log 2>&1
tail -2 gpurun_out/r2g_ncu_launches_631gs.log
tools/ncu_full.sh r2g_sps_ss_block 'eri_jk_block_kernel<.*3, .*0, .*0, .*0, .*0>' 5 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2g_sps_ss_light 'eri_jk_light_multi_kernel<.*3, .*0, .*0, .*0, .*0>' 0 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2g_ss_ss_light 'eri_jk_light_multi_kernel<.*0, .*0, .*0, .*0, .*0>' 0 1 -- 96 6-31G 1e-10 0 1
ls -la gpurun_out/ | tail -8
