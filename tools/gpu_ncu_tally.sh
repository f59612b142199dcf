#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:sweep_chunk -c 2 -f -o gpurun_out/chunk_tally python tools/ncu_one.py --kernel 4 --n-inner 1 --tally 1 > gpurun_out/ncu_tally.log 2>&1
python tools/ncu_lines.py gpurun_out/chunk_tally.ncu-rep sweep_chunk 0 sweep_chunk_kernelILi2ELi2ELi1E > gpurun_out/chunk_tally.lines.txt 2>&1
python tools/ncu_summary.py gpurun_out/chunk_tally.ncu-rep | head -40
