#!/bin/bash
# Round-2 evidence on ONE GPU: the whole GPU suite, the bench line, the reference arm, the ncu launch list.
O=gpurun_out/final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
python bench.py > $O/bench_pergroup.json 2> $O/bench_pergroup.err; tail -c 600 $O/bench_pergroup.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --persistent 2 --no-cpu-baseline > $O/bench_persistent.json 2> $O/bench_persistent.err
python bench.py --mode batched --no-cpu-baseline > $O/bench_batched.json 2> $O/bench_batched.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_pergroup.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python - <<'PY'
import json
for n in ("bench_pergroup", "bench_persistent", "bench_batched", "bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/final/{n}.json") if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        print(n, "value %.4g e2e %.4g ms/step %.3f frac %s ms/inner %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("frac"), r.get("ms_per_launch")),
              "e2e_dev", (d.get("e2e_device_sources") or {}).get("value"), "ttc", d.get("time_to_converge_s"), "plugin", (d.get("e2e_plugin") or {}).get("value"))
    except Exception as e:
        print(n, "failed", e)
PY
