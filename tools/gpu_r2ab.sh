#!/bin/bash
# round-2 GPU call AB: evidence on the final revision (per-class far-field threshold): launch list -> counters.json,
# GPU suite, smoke, bench lines (ours, reference arm, other configs)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,launch__registers_per_thread,launch__grid_size,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 1200 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2ab_launches_h2o96_631g_ref.csv python tools/prof_jk.py 96 6-31G 1e-10 0 2 > gpurun_out/r2ab_ncu_launches.log 2>&1
tail -1 gpurun_out/r2ab_ncu_launches.log | cut -c1-200
python tools/counters_from_ncu.py gpurun_out/r2ab_launches_h2o96_631g_ref.csv h2o96_631g/reference/1 2 --source profiles/r02_launches_h2o96_631g_ref_ab.csv | tail -2
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2ab_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2ab_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1; tail -1 gpurun_out/r2ab_smoke.log | cut -c1-160
timeout 600 python bench.py --impl reference > gpurun_out/r2ab_bench_ref.json 2> gpurun_out/r2ab_bench_ref.err
python -c "import json;d=json.load(open('gpurun_out/r2ab_bench_ref.json'));print(d['impl'],d['value'],d['cpu_baseline']['cores'])"
timeout 900 python bench.py > gpurun_out/r2ab_bench_n1.json 2> gpurun_out/r2ab_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2ab_bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['roofline']['fp64_pipe_active'],d['roofline']['traffic'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'],d['exact_boys']['roofline_frac'],d['gpu_launches'])"
for w in h2o32_631gs h2o32_631g h2o96_sto3g; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/r2ab_bench_$w.json 2> gpurun_out/r2ab_bench_$w.err
  python -c "import json;d=json.load(open('gpurun_out/r2ab_bench_$w.json'));print('$w',d['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
done
cp profiles/counters.json gpurun_out/r2ab_counters.json
