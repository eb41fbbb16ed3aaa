#!/bin/bash
# round-2 GPU call V: unroll factor of the bra-primitive loop (the compiler's choice is x4 plus three remainder copies)
mkdir -p gpurun_out
V="kb1|librchem_b200_kb1.so|;kb2|librchem_b200_kb2.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2v_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2v_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 >> gpurun_out/r2v_ab.txt 2>&1
cat gpurun_out/r2v_ab.txt
