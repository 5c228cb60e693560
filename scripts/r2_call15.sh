#!/bin/bash
# Round 2, fifteenth GPU call (2 GPUs): the multi-GPU paths with two wavefronts per device — 2-GPU gates (IPC peer reduce,
# in-process device group over real peer memory, the C consumer), the torchrun bench line of the strong-scaled config 5
# series at N = 2, the in-process group, and the N = 1 line of the series as the driver's scaling run asks for it.
mkdir -p gpurun_out
nvidia-smi -L
echo "=== 2-GPU gates"; timeout -k 10 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_host.py -x -q -m gpu 2>&1 | tail -5
echo "=== torchrun N=2, strong series"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2_strong.json 2> gpurun_out/bench_n2_strong.err
tail -c 300 gpurun_out/bench_n2_strong.err; tail -1 gpurun_out/bench_n2_strong.json | cut -c1-600
echo "=== in-process group N=2, config 5 at 64 spp per step"
timeout -k 10 600 python bench.py --gpus 2 --launcher inproc --workload config5_combined --spp 32 --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_n2_inproc.json 2> gpurun_out/bench_n2_inproc.err
tail -c 300 gpurun_out/bench_n2_inproc.err; tail -1 gpurun_out/bench_n2_inproc.json | cut -c1-400
echo "=== N=1 of the series on this box"
timeout -k 10 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_n1_strong.json 2> gpurun_out/bench_n1_strong.err
tail -c 300 gpurun_out/bench_n1_strong.err; tail -1 gpurun_out/bench_n1_strong.json | cut -c1-600
