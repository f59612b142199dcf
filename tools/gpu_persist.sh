#!/bin/bash
# Register-chunk sweep A/B on the bench workload: parity tests with the defaults, then one bench run per entry of
# RUNS ("label:ENV=val+ENV=val ...").
#   TESTS="tests/test_gpu_sweep.py" RUNS="p0:MOCB200_RC_PERSIST=0 p2:MOCB200_RC_PERSIST=2" tools/gpu_persist.sh
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout ${TEST_TIMEOUT:-1200} python -m pytest $TESTS -m gpu -x -q ${PYTEST_ARGS} 2>&1 | tail -15
fi
for r in ${RUNS:-p0:MOCB200_RC_PERSIST=0 p2:MOCB200_RC_PERSIST=2}; do
  lab=${r%%:*}; envs=${r#*:}
  env $(echo $envs | tr '+' ' ') timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BARGS > gpurun_out/ab_$lab.json 2> gpurun_out/ab_$lab.err
  python - "$lab" "$envs" <<'PY'
import json, sys
lab, envs = sys.argv[1:3]
try:
    d = json.load(open(f"gpurun_out/ab_{lab}.json"))
    print("%-10s %-44s value %.4g e2e %.4g ms/step %.3f frac %.3f ms/inner %.4f launches %d" % (lab, envs, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["gpu_launches"]))
except Exception as e:
    print(lab, "bench failed", e); print(open(f"gpurun_out/ab_{lab}.err").read()[-1500:])
PY
done
