#!/bin/bash
# bench (both modes) + ncu launch list of the bench command + one full capture of the dominant kernel
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_pergroup.json 2> gpurun_out/bench_pergroup.err; cat gpurun_out/bench_pergroup.json; tail -3 gpurun_out/bench_pergroup.err
python bench.py --steps 5 --warmup 3 --mode batched --no-cpu-baseline > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err; cat gpurun_out/bench_batched.json; tail -3 gpurun_out/bench_batched.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r1_launches_pergroup.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:sweep_cached_kernel -s 40 -c 4 -f -o gpurun_out/r1_bench_pergroup python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
tail -2 gpurun_out/ncu_bench2.log | cut -c1-300
