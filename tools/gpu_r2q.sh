#!/bin/bash
# round-2 GPU call Q: resident blocks per SM of the chunk kernel for the classes emitted in parts
mkdir -p gpurun_out
V="cb4|librchem_b200_cb4.so|;cb6|librchem_b200_cb6.so|;cb8|librchem_b200_cb8.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 > gpurun_out/r2q_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 12 '6-31G*' 0 >> gpurun_out/r2q_ab.txt 2>&1
cat gpurun_out/r2q_ab.txt
