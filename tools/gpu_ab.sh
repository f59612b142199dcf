#!/bin/bash
# A/B of tuning knobs on the bench (kernel 4)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for nw in 1 2; do
  MOCB200_CHUNK_NW=$nw timeout 300 python bench.py --steps 5 --warmup 3 --kernel 4 --no-cpu-baseline > gpurun_out/bench_nw$nw.json 2>gpurun_out/bench_nw$nw.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_nw$nw.json"))
    print("nw=$nw value %.4g e2e %.4g ms/step %.3f roofline frac %.3f ms_per_launch %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_nw$nw.err").read()[-2000:])
PY
done
