#!/bin/bash
# round-2 GPU call S: bench lines of every workload on the final kernels (after the merged pack/finalise launches,
# the high-L chunk kernel at 4 blocks/SM and the fully unrolled digestion), GPU suite, smoke()
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2s_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2s_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; tail -2 gpurun_out/r2s_smoke.log
timeout 900 python bench.py > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2s_bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['roofline']['fp64_pipe_active'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'],d['gpu_launches'])"
for w in h2o32_631gs h2o32_631g h2o96_sto3g; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/r2s_bench_$w.json 2> gpurun_out/r2s_bench_$w.err
  python -c "import json;d=json.load(open('gpurun_out/r2s_bench_$w.json'));print('$w',d['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
done
for w in h2o10_sto3g h2o12_631gs; do
timeout 600 python bench.py --mode tensor --workload $w > gpurun_out/r2s_bench_tensor_$w.json 2> gpurun_out/r2s_bench_tensor_$w.err
python -c "import json;d=json.load(open('gpurun_out/r2s_bench_tensor_$w.json'));print('$w',d['build_I_ms'],d['jk_inmem_ms'],d['parity'])"
done
