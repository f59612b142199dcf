#!/bin/bash
# chunk-kernel iteration loop: parity tests, bench (kernel 4), ncu summary of the chunk launches
mkdir -p gpurun_out
TAG=${TAG:-x}
timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 5 --warmup 3 --kernel ${K:-4} --no-cpu-baseline > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("$TAG value %.4g e2e %.4g ms/step %.3f roofline frac %.3f ms_per_launch %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
if [ -n "$NCU" ]; then
ncu --set full --import-source on --clock-control none -k regex:sweep_chunk -c 4 -f -o gpurun_out/chunk_$TAG python tools/ncu_one.py --kernel ${K:-4} --n-inner 1 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
python tools/ncu_summary.py gpurun_out/chunk_$TAG.ncu-rep > gpurun_out/chunk_$TAG.summary.txt 2>&1
grep -E "kernel:|time_duration|inst_executed.sum|issue_active|warps_active|long_scoreboard|short_scoreboard|stalled_wait|barrier|l1tex__throughput|wavefronts_mem_shared|dram__bytes_read" gpurun_out/chunk_$TAG.summary.txt
fi
