#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_plugin.py -m gpu -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2>gpurun_out/scale_n1.err; cat gpurun_out/scale_n1.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_n2.json 2>gpurun_out/scale_n2.err; cat gpurun_out/scale_n2.json | cut -c1-400; tail -3 gpurun_out/scale_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
