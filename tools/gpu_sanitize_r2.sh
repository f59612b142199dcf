#!/bin/bash
# compute-sanitizer on the kernels that changed this round (small cases): persistent launch, tally walks, families, sources
O=gpurun_out/sanitize; mkdir -p $O
SEL='(persistent and mini2d and 4-1) or (sweep1g_matches and mini2d_gs and 5) or (corrections_match and 5) or (family and mini2d and 1) or (device_source and mini2d and False)'
for tool in racecheck memcheck; do
  timeout ${LIMIT:-110} compute-sanitizer --tool $tool --target-processes all python -m pytest tests/test_gpu_sweep.py -m gpu -q -x -k "$SEL" > $O/$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/$tool.log | tail -4
done
