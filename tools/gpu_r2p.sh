#!/bin/bash
# round-2 GPU call P (4 GPUs): sharded reduction + D2H of the in-library multi-GPU path
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "several_devices or two_real_ranks or rank_partition" ) > gpurun_out/r2p_pytest_multi.log 2>&1
tail -3 gpurun_out/r2p_pytest_multi.log
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r2p_bench_n$n.err | grep '^{' > gpurun_out/r2p_bench_n$n.json
  python -c "import json;d=json.load(open('gpurun_out/r2p_bench_n$n.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['checksum']['sum_J'],d['checksum']['sum_K'],d['parity']['max_abs_err'])"
  tail -2 gpurun_out/r2p_bench_n$n.err
done
