#!/bin/bash
# strong scaling of ONE C5G7-2D plane over angle families: N = 1 (whole sweep), 2, 4, 8 (what the box has)
O=gpurun_out/ascale; mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/ascale_n1.json 2> $O/ascale_n1.err
for n in ${NS:-2 4 8}; do
  [ $n -le $NG ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --shard angles --steps 5 --warmup 3 > $O/ascale_n$n.json 2> $O/ascale_n$n.err || tail -5 $O/ascale_n$n.err
done
python - <<'PY'
import json
base = base_e = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/ascale/ascale_n{n}.json") if l.startswith("{")][-1])
    except Exception:
        continue
    base = base or d["value"]; base_e = base_e or d["e2e"]["value"]
    comm = d.get("comm") or {}
    print(f"N={n} value {d['value']:.4g} x{d['value']/base:.2f} e2e {d['e2e']['value']:.4g} x{d['e2e']['value']/base_e:.2f} "
          f"ms/step {d['ms_per_step']:.3f} comm ms/step {comm.get('ms_per_step', 0):.3f} sweep ms/inner {d['roofline']['ms_per_launch']:.4f}")
PY
