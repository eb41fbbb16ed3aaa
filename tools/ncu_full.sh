#!/bin/bash
# usage: tools/ncu_full.sh <out-name> <la> <lb> <lc> <ld> <boys> <mode> <skip> <count> -- <python args for tools/prof_jk.py>
out=$1; la=$2; lb=$3; lc=$4; ld=$5; boys=$6; mode=$7; skip=$8; cnt=$9; shift 10
re="eri_kernel<\\(int\\)$la, \\(int\\)$lb, \\(int\\)$lc, \\(int\\)$ld, \\(int\\)$boys, \\(int\\)$mode>"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$re" -s "$skip" -c "$cnt" -f -o "gpurun_out/$out" python tools/prof_jk.py "$@" > "gpurun_out/$out.log" 2>&1
tail -2 "gpurun_out/$out.log"
ls -la gpurun_out/
