#!/bin/bash
# round-2 GPU call I (2 GPUs): the multi-GPU paths on real devices -- bench.py under torchrun
# (device-timed NCCL path + e2e through ONE rchem_jk_direct call with RCHEM_OPT_NGPUS=2),
# the reference arm under torchrun, and the multi-device tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2i_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q -k "several_devices or two_real_ranks or rank_partition" ) > gpurun_out/r2i_pytest_multi.log 2>&1
tail -5 gpurun_out/r2i_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
python -c "import json;d=json.load(open('gpurun_out/r2i_bench_n2.json'));print(d['n_gpus'],d['ms_per_step'],d['value'],d['e2e'],d['roofline']['frac'],d['checksum'],d['parity'])"
tail -5 gpurun_out/r2i_bench_n2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2i_bench_n1.json'));print(d['n_gpus'],d['ms_per_step'],d['value'],d['e2e'],d['checksum'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2i_bench_ref_n2.json 2> gpurun_out/r2i_bench_ref_n2.err
cat gpurun_out/r2i_bench_ref_n2.json | cut -c1-600; tail -3 gpurun_out/r2i_bench_ref_n2.err
ls -la gpurun_out | tail -6
