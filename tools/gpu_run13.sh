#!/bin/bash
mkdir -p gpurun_out
K=${K:-4}
python tools/ncu_one.py --kernel $K --reps 3 2>&1 | tail -3
ncu --set full --import-source on --clock-control none -k regex:sweep_chunk -c 4 -f -o gpurun_out/chunk_k$K python tools/ncu_one.py --kernel $K --n-inner 1 > gpurun_out/ncu_k$K.log 2>&1; tail -3 gpurun_out/ncu_k$K.log
python tools/ncu_summary.py gpurun_out/chunk_k$K.ncu-rep > gpurun_out/chunk_k$K.summary.txt 2>&1; head -100 gpurun_out/chunk_k$K.summary.txt
