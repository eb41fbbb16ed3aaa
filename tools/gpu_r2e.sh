#!/bin/bash
# round-2 GPU call E: bench lines of the fused-sp revision (ours + reference arm), ncu --set full of the
# three largest launches, Boys-grid-in-shared-memory A/B (librchem_b200_bsm.so = -DRCHEM_BOYS_SMEM=1)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2e_smi.txt
timeout 900 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2e_bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e'],d['roofline']['frac'],d['parity'],d['exact_boys']['ms_per_step'])"
tail -3 gpurun_out/r2e_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
cat gpurun_out/r2e_bench_ref.json
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="boys_smem|librchem_b200_bsm.so|" timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2e_ab_boys_smem.txt 2>&1
cat gpurun_out/r2e_ab_boys_smem.txt
tools/ncu_full.sh r2e_sps_ss_block 'eri_jk_block_kernel<.*3, .*0, .*0, .*0, .*0>' 25 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2e_spsp_sps_block 'eri_jk_block_kernel<.*3, .*3, .*3, .*0, .*0>' 9 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2e_sps_sps_block 'eri_jk_block_kernel<.*3, .*0, .*3, .*0, .*0>' 7 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2e_sps_sps_light 'eri_jk_light_multi_kernel<.*3, .*0, .*3, .*0, .*0>' 0 1 -- 96 6-31G 1e-10 0 1
ls -la gpurun_out/
