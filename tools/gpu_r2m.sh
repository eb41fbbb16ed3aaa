#!/bin/bash
# round-2 GPU call M: warp-uniform choice between the reference-Boys rows and the exact grid
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2m_pytest_gpu.log
AB_COMBOS=11 timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2m_ab.txt 2>&1
AB_COMBOS=11 timeout 900 python tools/ab_jk.py 96 STO-3G 1e-10 >> gpurun_out/r2m_ab.txt 2>&1
AB_COMBOS=11 timeout 900 python tools/ab_jk.py 32 '6-31G*' 1e-10 >> gpurun_out/r2m_ab.txt 2>&1
cat gpurun_out/r2m_ab.txt
