#!/bin/bash
# ncu --set full capture of the register-chunk kernel launches of one inner sweep (C5G7-2D, group 3)
mkdir -p gpurun_out
TAG=${TAG:-rc}
ncu --set full --import-source on --clock-control none -k regex:${KREGEX:-sweep_rchunk} -c ${COUNT:-2} -f -o gpurun_out/$TAG python tools/ncu_one.py --kernel ${K:-5} --n-inner ${NINNER:-1} --tally ${TALLY:-0} > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
python tools/ncu_summary.py gpurun_out/$TAG.ncu-rep > gpurun_out/$TAG.summary.txt 2>&1
grep -E "kernel:|time_duration|inst_executed.sum|issue_active|warps_active|l1tex__throughput|wavefronts_mem_shared|dram__bytes_read" gpurun_out/$TAG.summary.txt
