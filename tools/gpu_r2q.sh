#!/bin/bash
# round-2 GPU call Q: block / light kernels (far-field scheduling, shared-memory rows) for the part classes up to 300 accumulators
mkdir -p gpurun_out
V="bm300|librchem_b200_bm300.so|"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 > gpurun_out/r2q_ab.txt 2>&1
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 12 '6-31G*' 0 >> gpurun_out/r2q_ab.txt 2>&1
cat gpurun_out/r2q_ab.txt
( RCHEM_B200_LIB=$PWD/rchem_b200/librchem_b200_bm300.so timeout 900 python -m pytest tests -m gpu -x -q -k "6-31G or d_shell or multicentre or all_classes or bse" ) > gpurun_out/r2q_pytest.log 2>&1
tail -3 gpurun_out/r2q_pytest.log
