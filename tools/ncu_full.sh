#!/bin/bash
# usage: tools/ncu_full.sh <out-name> <kernel-regex (demangled)> <skip> <count> -- <python args for tools/prof_jk.py>
out=$1; re=$2; skip=$3; cnt=$4; shift 5
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$re" -s "$skip" -c "$cnt" -f -o "gpurun_out/$out" python tools/prof_jk.py "$@" > "gpurun_out/$out.log" 2>&1
tail -2 "gpurun_out/$out.log"
ls -la gpurun_out/
