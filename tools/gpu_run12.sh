#!/bin/bash
# chunk kernel: GPU parity tests, then bench A/B (kernel 3 = warp-block cached, 4 = chunk)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q 2>&1 | tail -8
for k in 4 3; do
  timeout 300 python bench.py --steps 5 --warmup 3 --kernel $k --no-cpu-baseline > gpurun_out/bench_k$k.json 2>gpurun_out/bench_k$k.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_k$k.json"))
    print("kernel $k value %.4g e2e %.4g ms/step %.3f roofline frac %.3f ms_per_launch %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"]))
except Exception as e:
    print("kernel $k failed", e); print(open("gpurun_out/bench_k$k.err").read()[-2000:])
PY
done
