#!/bin/bash
# Register-chunk kernel: parity tests, then an A/B of its launch geometries on the bench workload.
#   TESTS="tests/test_gpu_sweep.py" CFGS="11,4,4 13,4,4 ..." tools/gpu_rc.sh
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout ${TEST_TIMEOUT:-1500} python -m pytest $TESTS -m gpu -x -q ${PYTEST_ARGS} 2>&1 | tail -15
fi
run() { # label, env..., -- bench args
  local label="$1"; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $BARGS > gpurun_out/rc_$label.json 2> gpurun_out/rc_$label.err
  python - "$label" <<'PY'
import json, sys
lab = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/rc_{lab}.json"))
    print("%-12s value %.4g e2e %.4g ms/step %.3f frac %.3f ms/inner %.4f kernel %s" % (lab, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["arm"]["kernel"]))
except Exception as e:
    print(lab, "bench failed", e); print(open(f"gpurun_out/rc_{lab}.err").read()[-1500:])
PY
}
BARGS="--kernel 4" run chunk_r1 X=1
for c in ${CFGS:-11,4,4}; do
  BARGS="--kernel 5" run "rc_${c//,/_}" MOCB200_RC_CFG=$c
done
