#!/bin/bash
# SURVEY.md 8(d) config 3 "dense-B": C5G7-2D with cg 32x4, ray spacing 0.01 (S ~ 7.3e8 segments per group sweep)
mkdir -p gpurun_out
free -g | head -2; nproc
MEM=$(free -g | awk '/Mem:/{print $7}')
if [ "$MEM" -lt 150 ]; then echo "only $MEM GB of host memory available: not running dense-B"; exit 0; fi
timeout ${LIMIT:-1700} python bench.py --workload dense_b --max-polar 4 --steps 2 --warmup 3 --no-cpu-baseline \
  > gpurun_out/bench_dense_b.json 2> gpurun_out/bench_dense_b.err
tail -3 gpurun_out/bench_dense_b.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_dense_b.json"))
    print("dense-B: segments %.4g value %.4g e2e %.4g ms/step %.2f frac %.3f ms/inner %.4f" % (d["config"]["segments"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"]))
except Exception as e:
    print("dense-B failed", e)
PY
