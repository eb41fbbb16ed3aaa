#!/bin/bash
# round-2 GPU call Q: resident blocks per SM of the chunk kernel for the classes emitted in parts
mkdir -p gpurun_out
V="u216|librchem_b200_u216.so|;u324|librchem_b200_u324.so|;u1296|librchem_b200_u1296.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 > gpurun_out/r2q_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 12 '6-31G*' 0 >> gpurun_out/r2q_ab.txt 2>&1
cat gpurun_out/r2q_ab.txt
