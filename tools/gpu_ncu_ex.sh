#!/bin/bash
mkdir -p gpurun_out
for ex in 0 1; do
MOCB200_CHUNK_EX=$ex ncu --set full --clock-control none -k regex:sweep_chunk -c 1 -f -o gpurun_out/chunk_ex$ex python tools/ncu_one.py --kernel 4 --n-inner 1 > gpurun_out/ncu_ex$ex.log 2>&1
ncu -i gpurun_out/chunk_ex$ex.ncu-rep --page raw --csv > gpurun_out/raw_ex$ex.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/raw_ex$ex.csv")))
hdr=rows[0]; v=rows[2]
print("EX=$ex")
for k in ("gpu__time_duration.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum","smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","smsp__inst_executed.sum"):
    if k in hdr: print("  ",k,v[hdr.index(k)])
PY
done
