#!/bin/bash
# round-2 GPU call Y (4 GPUs): configs[3] ((H2O)32 6-31G* and 6-31G) at N = 4 and 2
mkdir -p gpurun_out
for w in h2o32_631gs h2o32_631g; do for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 10 --warmup 3 --workload $w 2> gpurun_out/r2y_bench_${w}_n$n.err | grep '^{' > gpurun_out/r2y_bench_${w}_n$n.json
  python -c "import json;d=json.load(open('gpurun_out/r2y_bench_${w}_n$n.json'));print('$w',d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['parity']['max_abs_err'])"
done; done
