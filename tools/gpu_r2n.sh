#!/bin/bash
# round-2 GPU call N: evidence for the final kernels -- GPU suite, bench lines (ours + reference arm), launch lists
# (two J/K builds for profiles/counters.json; the bench command itself), ncu --set full of the top launch
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2n_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2n_pytest_gpu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2n_launches_h2o96_631g_ref.csv python tools/prof_jk.py 96 6-31G 1e-10 0 2 > gpurun_out/r2n_ncu_launches.log 2>&1
tail -2 gpurun_out/r2n_ncu_launches.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 2500 --csv --log-file gpurun_out/r2n_launches_bench_cmd.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2n_ncu_bench_cmd.log 2>&1
tail -2 gpurun_out/r2n_ncu_bench_cmd.log
tools/ncu_full.sh r2n_sps_ss_block 'eri_jk_block_kernel<.*3, .*0, .*0, .*0, .*0>' 5 1 -- 96 6-31G 1e-10 0 1
timeout 900 python bench.py > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2n_bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
timeout 600 python bench.py --impl reference > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err
cut -c1-300 gpurun_out/r2n_bench_ref.json
for w in h2o32_631gs h2o32_631g h2o96_sto3g; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/r2n_bench_$w.json 2> gpurun_out/r2n_bench_$w.err
  python -c "import json;d=json.load(open('gpurun_out/r2n_bench_$w.json'));print('$w',d['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
done
timeout 600 python bench.py --mode tensor --workload h2o12_631gs > gpurun_out/r2n_bench_tensor_h2o12_631gs.json 2> gpurun_out/r2n_bench_tensor.err
python -c "import json;d=json.load(open('gpurun_out/r2n_bench_tensor_h2o12_631gs.json'));print(d['build_I_ms'],d['jk_inmem_ms'])"
ls -la gpurun_out | tail -5
