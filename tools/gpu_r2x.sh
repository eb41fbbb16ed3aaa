#!/bin/bash
# round-2 GPU call X: timing probe -- light kernel without its A-side global D loads / K atomics (results WRONG on purpose)
mkdir -p gpurun_out
V="allfar|librchem_b200_allfar.so|;nok|librchem_b200_nok.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2x_ab.txt 2>&1
cat gpurun_out/r2x_ab.txt
