#!/bin/bash
# round-2 GPU call O (4 GPUs): scaling of the final kernels -- bench.py under torchrun at N = 4 (and N = 2, 1 on the same box)
mkdir -p gpurun_out
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r2o_bench_n$n.err | grep '^{' > gpurun_out/r2o_bench_n$n.json
  python -c "import json;d=json.load(open('gpurun_out/r2o_bench_n$n.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['checksum']['sum_J'],d['exact_boys']['ms_per_step'])"
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2o_bench_n1.json'));print(d['n_gpus'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['checksum']['sum_J'],d['exact_boys']['ms_per_step'])"
