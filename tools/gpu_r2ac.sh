#!/bin/bash
# round-2 GPU call AC: ncu --set full of the largest launches of the final revision
mkdir -p gpurun_out
tools/ncu_full.sh r2ac_sps_ss_block 'eri_jk_block_kernel<.*3, .*0, .*0, .*0, .*0>' 5 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2ac_spsp_sps_block 'eri_jk_block_kernel<.*3, .*3, .*3, .*0, .*0>' 5 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2ac_sps_sps_block 'eri_jk_block_kernel<.*3, .*0, .*3, .*0, .*0>' 1 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2ac_sps_ss_light 'eri_jk_light_multi_kernel<.*3, .*0, .*0, .*0, .*0>' 0 1 -- 96 6-31G 1e-10 0 1
tools/ncu_full.sh r2ac_sps_sps_light 'eri_jk_light_multi_kernel<.*3, .*0, .*3, .*0, .*0>' 0 1 -- 96 6-31G 1e-10 0 1
ls -la gpurun_out | grep r2ac
