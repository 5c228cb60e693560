#!/bin/bash
# Round 2, fourth GPU call (2 GPUs): the multi-GPU paths — 2-GPU gates (IPC peer reduce, in-process device group over
# real peer memory), the torchrun bench line of the strong-scaled config 5 series at N = 2 (few steps), the same through
# the in-process group, and the N = 1 line of the series as the driver's scaling run will ask for it.
mkdir -p gpurun_out
nvidia-smi -L
echo "=== 2-GPU gates"; timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -5
echo "=== torchrun N=2, strong series"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_strong.json 2> gpurun_out/bench_n2_strong.err
tail -c 400 gpurun_out/bench_n2_strong.err; tail -1 gpurun_out/bench_n2_strong.json | cut -c1-1200
echo "=== reference arm N=2"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 2>/dev/null | tail -1 | cut -c1-500
echo "=== in-process group N=2, config 5 at 64 spp per step"
timeout -k 10 600 python bench.py --gpus 2 --launcher inproc --workload config5_combined --spp 32 --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_n2_inproc.json 2> gpurun_out/bench_n2_inproc.err
tail -c 400 gpurun_out/bench_n2_inproc.err; tail -1 gpurun_out/bench_n2_inproc.json | cut -c1-900
echo "=== ranks N=2, same workload, for comparison"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload config5_combined --spp 32 --steps 3 --warmup 3 --no-cpu --no-extra 2>/dev/null | tail -1 | cut -c1-400
echo "=== N=1 of the series on this box (2 steps budget)"
timeout -k 10 600 python bench.py --gpus 1 --steps 3 --warmup 3 --max-seconds 70 --no-cpu > gpurun_out/bench_n1_strong.json 2> gpurun_out/bench_n1_strong.err
tail -c 400 gpurun_out/bench_n1_strong.err; tail -1 gpurun_out/bench_n1_strong.json | cut -c1-900
