#!/bin/bash
# 8-GPU session: data-parallel bench lines (ResNet-50 cfg, HRNet-w48 cfg); baselines skipped (they are 1-GPU quantities)
set +e
O=gpurun_out
mkdir -p $O
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err
[ "$2" = nohrnet ] || python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus $N --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench_${N}gpu_hrnet48.json 2> $O/r2_bench_${N}gpu_hrnet48.err
head -c 700 $O/r2_bench_${N}gpu.json; echo; tail -c 400 $O/r2_bench_${N}gpu.err; [ "$2" = nohrnet ] || { head -c 500 $O/r2_bench_${N}gpu_hrnet48.json; echo; tail -c 300 $O/r2_bench_${N}gpu_hrnet48.err; }
