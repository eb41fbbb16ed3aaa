#!/bin/bash
# round-2 final check: what the driver runs at round end, on HEAD
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r2z_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log | cut -c1-200
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
python -c "import json;d=json.load(open('gpurun_out/r2z_bench_ref.json'));print(d['impl'],d['value'],d['cpu_baseline']['cores'])"
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2z_bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['roofline']['fp64_pipe_active'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'],d['gpu_launches'],d['clocks'])"
