#!/bin/bash
# A/B of tuning knobs on the bench (kernel 4): VARS="ENV=val ENV=val ..." one run per entry
mkdir -p gpurun_out
[ -n "$TESTS" ] && timeout 900 python -m pytest $TESTS -m gpu -x -q 2>&1 | tail -3
i=0
for v in ${VARS:-MOCB200_CHUNK_NW=2}; do
  i=$((i+1))
  env $(echo $v | tr ',' ' ') timeout 300 python bench.py --steps 5 --warmup 3 --kernel 4 --no-cpu-baseline > gpurun_out/bench_ab$i.json 2>gpurun_out/bench_ab$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_ab$i.json"))
    print("$v value %.4g e2e %.4g ms/step %.3f roofline frac %.3f ms_per_launch %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_ab$i.err").read()[-2000:])
PY
done
