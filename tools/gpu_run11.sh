#!/bin/bash
# sanity run of the restored tree: GPU tests + 1-GPU bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r11.json 2>gpurun_out/bench_r11.err; cut -c1-1500 gpurun_out/bench_r11.json; tail -3 gpurun_out/bench_r11.err
