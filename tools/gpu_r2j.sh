#!/bin/bash
# round-2 GPU call J: block-size / kets-per-block variants of the small-class block kernel, chunk-kernel
# occupancy for the classes emitted in parts
mkdir -p gpurun_out
V="t448|librchem_b200_t448.so|;t384|librchem_b200_t384.so|;p16|librchem_b200_p16.so|;p4|librchem_b200_p4.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 1500 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2j_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2j_ab.txt 2>&1
cat gpurun_out/r2j_ab.txt
