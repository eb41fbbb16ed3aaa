#!/bin/bash
# round-2 GPU call K: bra D block in shared memory (new) vs before (base); one 1024-thread block per SM
# for the small classes (t1024: L1 keeps ~120 kB instead of ~28 kB next to the (sp s| rows)
mkdir -p gpurun_out
V="t1024|librchem_b200_t1024.so|RCHEM_HEAVY_PASSES=1;t1024h2|librchem_b200_t1024.so|"
AB_COMBOS=11 AB_VARIANTS="$V" timeout 1500 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2k_ab.txt 2>&1
AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2k_ab.txt 2>&1
AB_COMBOS=11 timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 >> gpurun_out/r2k_ab.txt 2>&1
cat gpurun_out/r2k_ab.txt
