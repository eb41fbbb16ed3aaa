#!/bin/bash
# round-2 GPU call T (8 GPUs): bench.py under torchrun at N = 8 (headline workload and configs[3])
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 10 --warmup 3 2> gpurun_out/r2t_bench_n8.err | grep '^{' > gpurun_out/r2t_bench_n8.json
python -c "import json;d=json.load(open('gpurun_out/r2t_bench_n8.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['checksum']['sum_J'],d['checksum']['sum_K'],d['parity']['max_abs_err'],d['exact_boys']['ms_per_step'])"
tail -3 gpurun_out/r2t_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --steps 10 --warmup 3 --workload h2o32_631gs 2> gpurun_out/r2t_bench_h2o32_631gs_n8.err | grep '^{' > gpurun_out/r2t_bench_h2o32_631gs_n8.json
python -c "import json;d=json.load(open('gpurun_out/r2t_bench_h2o32_631gs_n8.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'])"
