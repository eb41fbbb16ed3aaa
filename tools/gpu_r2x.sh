#!/bin/bash
# round-2 GPU call X: timing probes (results WRONG on purpose) -- everything through the far-field code in the light
# kernels only / in the block kernels only (librchem_b200_probe.so = -DRCHEM_PROBES)
mkdir -p gpurun_out
V="probe_none|librchem_b200_probe.so|;allfar_light|librchem_b200_probe.so|RCHEM_PROBE_ALLFAR_LIGHT=1;allfar_block|librchem_b200_probe.so|RCHEM_PROBE_ALLFAR_BLOCK=1"
AB_NOBASE=1 AB_COMBOS=11 AB_VARIANTS="$V" timeout 900 python tools/ab_jk.py 96 6-31G 1e-10 > gpurun_out/r2x_ab.txt 2>&1
cat gpurun_out/r2x_ab.txt
