#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
W=/tmp/c5; mkdir -p $W; cp oracle/_ref/inputs/* $W/; (cd $W && OMP_NUM_THREADS=$(nproc) /root/repo/oracle/_ref/ref_tool golden c5g7_2d.xml c5g7 --outers 1 --records "0:0:0,0:6:9" --cmfd > gen.log 2>&1; tail -1 gen.log)
timeout 900 python tools/exp_time.py $W/c5g7.mocflat $W/c5g7.golden --kernels ${KERNELS:-3} > gpurun_out/exp_time.jsonl 2>&1
cat gpurun_out/exp_time.jsonl
