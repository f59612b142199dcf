#!/bin/bash
# first GPU pass: parity tests, micro-benchmarks, exploratory timing on C5G7-2D
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 ./profiles/microbench/ubench > gpurun_out/ubench.jsonl 2>&1
W=/tmp/c5; mkdir -p $W; cp oracle/_ref/inputs/* $W/; (cd $W && OMP_NUM_THREADS=$(nproc) /root/repo/oracle/_ref/ref_tool golden c5g7_2d.xml c5g7 --outers 1 --records "0:0:0,0:6:9" --cmfd > gen.log 2>&1; tail -2 gen.log)
timeout 900 python tools/exp_time.py $W/c5g7.mocflat $W/c5g7.golden > gpurun_out/exp_time.jsonl 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/ubench.jsonl; cat gpurun_out/exp_time.jsonl
