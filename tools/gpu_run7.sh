#!/bin/bash
mkdir -p gpurun_out
W=/tmp/c5; mkdir -p $W; cp oracle/_ref/inputs/* $W/; (cd $W && OMP_NUM_THREADS=$(nproc) /root/repo/oracle/_ref/ref_tool golden c5g7_2d.xml c5g7 --outers 0 --records "" > gen.log 2>&1; tail -1 gen.log)
ncu --set full --clock-control none --import-source on -k regex:sweep_cached_kernel -s 2 -c 2 -f -o gpurun_out/r1_cached_pergroup python tools/exp_ncu.py $W/c5g7.mocflat $W/c5g7.golden --batched 0 --kernel 3 --n-inner 4 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_cached_kernel -s 2 -c 2 -f -o gpurun_out/r1_cached_batched python tools/exp_ncu.py $W/c5g7.mocflat $W/c5g7.golden --batched 1 --kernel 3 --n-inner 4 > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
