#!/bin/bash
# round-2 GPU call B: single-process multi-device path (oversubscribed on one GPU), sanitizers,
# bench lines of the other BASELINE configs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "several_devices or abi or tau_lowered" ) > gpurun_out/r2b_pytest_multi.log 2>&1
tail -5 gpurun_out/r2b_pytest_multi.log
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_jk.py ) > gpurun_out/r2b_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2b_sanitizer_memcheck.log
( time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_jk.py ) > gpurun_out/r2b_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2b_sanitizer_racecheck.log
for w in h2o32_631gs h2o32_631g h2o96_sto3g; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/r2b_bench_$w.json 2> gpurun_out/r2b_bench_$w.err
  python -c "import json;d=json.load(open('gpurun_out/r2b_bench_$w.json'));print('$w',d['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
done
ls -la gpurun_out/
