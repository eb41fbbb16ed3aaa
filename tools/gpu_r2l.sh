#!/bin/bash
# round-2 GPU call L: retiled tensor_fill_kernel (tensor tests + bench + DRAM counters), more J/K variants
# (threads of the large wide-row block classes, small light kernels' blocks per SM, stream count)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "build_I or tensor or inmem or fixture or bse or consumer" ) > gpurun_out/r2l_pytest_tensor.log 2>&1
tail -3 gpurun_out/r2l_pytest_tensor.log
for w in h2o10_sto3g h2o12_631gs; do
  timeout 600 python bench.py --mode tensor --workload $w > gpurun_out/r2l_bench_tensor_$w.json 2> gpurun_out/r2l_bench_tensor_$w.err
  python -c "import json;d=json.load(open('gpurun_out/r2l_bench_tensor_$w.json'));print('$w',d['build_I_ms'],d['jk_inmem_ms'],d['roofline']['frac'],d['parity'])"; tail -2 gpurun_out/r2l_bench_tensor_$w.err
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:tensor_fill --csv --log-file gpurun_out/r2l_launches_tensor_fill.csv python tools/time_tensor.py 12 '6-31G*' > gpurun_out/r2l_ncu_tensor.log 2>&1
grep -c tensor_fill gpurun_out/r2l_launches_tensor_fill.csv; grep "tensor_fill" gpurun_out/r2l_launches_tensor_fill.csv | head -3 | cut -c1-250
V=""
AB_NOBASE=1 AB_COMBOS=11 timeout 600 python tools/ab_jk.py 10 STO-3G 0 > gpurun_out/r2l_ab.txt 2>&1
cat gpurun_out/r2l_ab.txt
