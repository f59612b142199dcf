#!/bin/bash
# Round-2 multi-GPU evidence on one 8-GPU box: weak scaling over axial planes (one C5G7-2D plane per rank, NCCL
# all-gather per sweep(group)), strong scaling of ONE plane over angle families (NCCL all-reduce per inner), and the
# multi-device plugin tests that a 1-GPU box skips.
O=gpurun_out/scale_r2; mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
run() { # n, label, extra args
  local n=$1 lab=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) \
      bench.py --gpus $n --steps 5 --warmup 3 "$@" > $O/${lab}_n$n.json 2> $O/${lab}_n$n.err || tail -3 $O/${lab}_n$n.err
}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/planes_n1.json 2> $O/planes_n1.err
cp $O/planes_n1.json $O/angles_n1.json
for n in ${PLANES_N:-2 4 8}; do [ $n -le $NG ] && run $n planes; done
for n in ${ANGLES_N:-2 4 8}; do [ $n -le $NG ] && run $n angles --shard angles; done
python - <<'PY'
import json
for lab in ("planes", "angles"):
    base = base_e = None
    for n in (1, 2, 4, 8):
        try:
            d = json.loads([l for l in open(f"gpurun_out/scale_r2/{lab}_n{n}.json") if l.startswith("{")][-1])
        except Exception:
            continue
        base = base or d["value"]; base_e = base_e or d["e2e"]["value"]
        comm = d.get("comm") or {}
        print(f"{lab} N={n} value {d['value']:.4g} x{d['value']/base:.2f} e2e {d['e2e']['value']:.4g} x{d['e2e']['value']/base_e:.2f} "
              f"ms/step {d['ms_per_step']:.3f} comm ms/step {comm.get('ms_per_step', 0):.3f} sweep ms/inner {d['roofline']['ms_per_launch']:.4f}")
PY
timeout 900 python -m pytest tests/test_gpu_plugin.py -m gpu -q -k "sharded or (subproblem and settled)" 2>&1 | tail -4 | tee $O/pytest_multi_gpu.txt
