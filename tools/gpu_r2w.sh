#!/bin/bash
# round-2 GPU call W: stream count / heavy threshold re-tuned for a basis with d shells ((H2O)32/6-31G*)
mkdir -p gpurun_out
V="st4||RCHEM_STREAMS=4;st8||RCHEM_STREAMS=8;st12||RCHEM_STREAMS=12;st16||RCHEM_STREAMS=16;heavy1||RCHEM_HEAVY_PASSES=1;heavy3||RCHEM_HEAVY_PASSES=3;order0||RCHEM_TASK_ORDER=0"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 > gpurun_out/r2w_ab.txt 2>&1
cat gpurun_out/r2w_ab.txt
