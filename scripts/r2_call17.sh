#!/bin/bash
# Round 2, seventeenth GPU call (8 GPUs): the driver's scaling command at N = 8 — config 5, 3840x2160, 4096 spp split 8 ways
# (512 per GPU), torchrun + peer-memory reduce; then the reference arm the same way.
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "=== torchrun N=8, strong series"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_strong.json 2> gpurun_out/bench_n8_strong.err
tail -c 300 gpurun_out/bench_n8_strong.err; tail -1 gpurun_out/bench_n8_strong.json | cut -c1-500
echo "=== in-process group N=8, config 5 at 64 spp per GPU per step"
timeout -k 10 300 python bench.py --gpus 8 --launcher inproc --workload config5_combined --spp 64 --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_n8_inproc.json 2> gpurun_out/bench_n8_inproc.err
tail -c 300 gpurun_out/bench_n8_inproc.err; tail -1 gpurun_out/bench_n8_inproc.json | cut -c1-400
