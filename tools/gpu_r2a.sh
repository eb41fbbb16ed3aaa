#!/bin/bash
# round-2 GPU call A: full GPU test suite, bench (both arms), ncu launch list with counters, one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/r2a_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 600 gpurun_out/r2a_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2a_launches_h2o96_631g_ref.csv python tools/prof_jk.py 96 6-31G 1e-10 0 2 > gpurun_out/r2a_ncu_launches.log 2>&1
tail -2 gpurun_out/r2a_ncu_launches.log
EXTRA=l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum,l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on --kernel-name-base demangled -k 'regex:eri_jk_block_kernel<.*1, .*0, .*0, .*0, .*0>' -s 27 -c 1 -f -o gpurun_out/r2a_psss_block python tools/prof_jk.py 96 6-31G 1e-10 0 1 > gpurun_out/r2a_ncu_full.log 2>&1
tail -2 gpurun_out/r2a_ncu_full.log
ls -la gpurun_out/
