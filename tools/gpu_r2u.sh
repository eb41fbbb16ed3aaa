#!/bin/bash
# round-2 GPU call U (2 GPUs): in-library multi-GPU path (optimistic symmetry probe), tests + bench
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "several_devices or two_real_ranks or rank_partition or asymmetric" ) > gpurun_out/r2u_pytest_multi.log 2>&1
tail -3 gpurun_out/r2u_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r2u_bench_n2.err | grep '^{' > gpurun_out/r2u_bench_n2.json
python -c "import json;d=json.load(open('gpurun_out/r2u_bench_n2.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['checksum']['sum_J'],d['checksum']['sum_K'],d['parity']['max_abs_err'])"
