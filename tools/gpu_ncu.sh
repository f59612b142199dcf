#!/bin/bash
# ncu --set full capture of the chunk-kernel launches of one inner sweep (+ per-launch list of a short bench)
mkdir -p gpurun_out
TAG=${TAG:-x}
ncu --set full --import-source on --clock-control none -k regex:sweep_chunk -c 2 -f -o gpurun_out/chunk_$TAG python tools/ncu_one.py --kernel ${K:-4} --n-inner 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
python tools/ncu_summary.py gpurun_out/chunk_$TAG.ncu-rep > gpurun_out/chunk_$TAG.summary.txt 2>&1
grep -E "kernel:|time_duration|inst_executed.sum|issue_active|warps_active|l1tex__throughput|wavefronts_mem_shared|dram__bytes_read" gpurun_out/chunk_$TAG.summary.txt
