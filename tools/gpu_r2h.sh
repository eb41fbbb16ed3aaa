#!/bin/bash
# round-2 GPU call H: reference-Boys rows with two slots per cell (thresholds + rows in one round trip),
# medium light kernels at 3 blocks/SM (default now) vs 4 (variant); GPU suite, sanitizers on the new
# block kernel, full bench line, dense-tensor bench + ncu of the HBM-bound kernels
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/r2h_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2h_pytest_gpu.log
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_jk.py ) > gpurun_out/r2h_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2h_sanitizer_memcheck.log
( time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_jk.py ) > gpurun_out/r2h_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2h_sanitizer_racecheck.log
AB_COMBOS=11 AB_VARIANTS="lm4|librchem_b200_lm4.so|" timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2h_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="lm4|librchem_b200_lm4.so|" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 >> gpurun_out/r2h_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 timeout 900 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2h_ab.txt 2>&1
cat gpurun_out/r2h_ab.txt
timeout 900 python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2h_bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
tail -3 gpurun_out/r2h_bench_n1.err
for w in h2o10_sto3g h2o12_631gs; do
  timeout 600 python bench.py --mode tensor --workload $w > gpurun_out/r2h_bench_tensor_$w.json 2> gpurun_out/r2h_bench_tensor_$w.err
  cat gpurun_out/r2h_bench_tensor_$w.json; tail -2 gpurun_out/r2h_bench_tensor_$w.err
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2h_launches_tensor_h2o12_631gs.csv python tools/time_tensor.py 12 '6-31G*' > gpurun_out/r2h_ncu_tensor.log 2>&1
tail -3 gpurun_out/r2h_ncu_tensor.log
ls -la gpurun_out/ | tail -12
