#!/bin/bash
# weak-scaling bench over axial planes: N = 1, 2, 4 (8 when the box has them), one rank per GPU
O=gpurun_out/scale; mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 > $O/scale_n$n.json 2> $O/scale_n$n.err
done
python - <<'PY'
import json, glob
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f"gpurun_out/scale/scale_n{n}.json") if l.startswith("{")][-1])
    except Exception as e:
        continue
    base = base or d["value"]
    base_e = locals().get("base_e") or d["e2e"]["value"]
    comm = d.get("comm") or {}
    print(f"N={n} value {d['value']:.4g} x{d['value']/base:.2f} e2e {d['e2e']['value']:.4g} x{d['e2e']['value']/base_e:.2f} "
          f"ms/step {d['ms_per_step']:.3f} comm ms/step {comm.get('ms_per_step', 0):.3f}")
PY
