#!/bin/bash
# Round-1 evidence run: GPU tests, bench (both modes), reference arm, ncu launch list, ncu --set full of the
# dominant kernel (per-group chunk kernel) with per-source-line stall table. Outputs under gpurun_out/r1/.
O=gpurun_out/r1; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench_pergroup.json 2> $O/bench_pergroup.err; cut -c1-300 $O/bench_pergroup.json
timeout 600 python bench.py --mode batched --no-cpu-baseline > $O/bench_batched.json 2> $O/bench_batched.err; cut -c1-200 $O/bench_batched.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_pergroup.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:sweep_chunk -c 2 -f -o $O/r1_chunk_pergroup python tools/ncu_one.py --kernel 4 --n-inner 1 > $O/ncu_chunk.log 2>&1
python tools/ncu_summary.py $O/r1_chunk_pergroup.ncu-rep > $O/r1_chunk_pergroup.summary.txt 2>&1
python tools/ncu_lines.py $O/r1_chunk_pergroup.ncu-rep sweep_chunk 0 sweep_chunk_kernelILi2ELi2ELi0E > $O/r1_chunk_pergroup.lines.txt 2>&1
ncu --set full --clock-control none -k regex:sweep_chunk -c 2 -f -o $O/r1_chunk_tally python tools/ncu_one.py --kernel 4 --n-inner 1 --tally 1 > $O/ncu_chunk_tally.log 2>&1
python tools/ncu_summary.py $O/r1_chunk_tally.ncu-rep > $O/r1_chunk_tally.summary.txt 2>&1
grep -E "kernel:|time_duration|dram__bytes" $O/r1_chunk_pergroup.summary.txt $O/r1_chunk_tally.summary.txt
