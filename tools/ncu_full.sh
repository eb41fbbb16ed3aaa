#!/bin/bash
# usage: tools/ncu_full.sh <out-name> <kernel-regex> <skip> <count> -- <python args for tools/prof_jk.py>
# e.g.   tools/ncu_full.sh psss_block 'eri_jk_block_kernel<.*1, .*0, .*0, .*0, .*0>' 27 1 -- 96 6-31G 1e-10 0 1
# Keep <count> small: one --set full capture of a big launch is ~6 MB and ~20 s; gpurun copies
# back at most 64 MiB.  (The '.*' in the regex absorb ncu's "(int)" template-argument casts.)
# Summarise with: python tools/ncu_summary.py gpurun_out/<out-name>.ncu-rep
out=$1; re=$2; skip=$3; cnt=$4; shift 5
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:$re" -s "$skip" -c "$cnt" -f -o "gpurun_out/$out" python tools/prof_jk.py "$@" > "gpurun_out/$out.log" 2>&1
tail -2 "gpurun_out/$out.log"
ls -la gpurun_out/
