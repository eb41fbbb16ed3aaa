#!/bin/bash
# round-2 GPU call AA: compute-sanitizer on the final kernels, s/sp basis and a basis with d shells (part classes)
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_jk.py 3 6-31G ) > gpurun_out/r2aa_memcheck_631g.log 2>&1; echo "memcheck 6-31G rc=$?"; tail -3 gpurun_out/r2aa_memcheck_631g.log
( time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_jk.py 3 6-31G ) > gpurun_out/r2aa_racecheck_631g.log 2>&1; echo "racecheck 6-31G rc=$?"; tail -3 gpurun_out/r2aa_racecheck_631g.log
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_jk.py 2 '6-31G*' ) > gpurun_out/r2aa_memcheck_631gs.log 2>&1; echo "memcheck 6-31G* rc=$?"; tail -3 gpurun_out/r2aa_memcheck_631gs.log
( time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_jk.py 2 '6-31G*' ) > gpurun_out/r2aa_racecheck_631gs.log 2>&1; echo "racecheck 6-31G* rc=$?"; tail -3 gpurun_out/r2aa_racecheck_631gs.log
