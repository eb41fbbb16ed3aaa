#!/bin/bash
# round-2 GPU call Q: part-outer evaluation of the classes emitted in parts (accumulators of one part in registers)
mkdir -p gpurun_out
AB_COMBOS=11 timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 > gpurun_out/r2q_ab.txt 2>&1
AB_COMBOS=11 timeout 900 python tools/ab_jk.py 12 '6-31G*' 0 >> gpurun_out/r2q_ab.txt 2>&1
cat gpurun_out/r2q_ab.txt
( timeout 900 python -m pytest tests -m gpu -x -q -k "6-31G or d_shell or multicentre or all_classes or bse or tier1 or smoke" ) > gpurun_out/r2q_pytest.log 2>&1
tail -3 gpurun_out/r2q_pytest.log
timeout 600 python tools/time_tensor.py 12 '6-31G*' 2>&1 | head -3
