#!/bin/bash
# 2-GPU session: NCCL equivalence test, data-parallel bench with the overlapped / single all-reduce
set +e
O=gpurun_out
mkdir -p $O
python tools/debug_bottleneck.py > $O/r2_dbg_bneck2.log 2>&1
python -m pytest tests/test_ops_gpu.py -m gpu -q -k bottleneck -p no:cacheprovider > $O/r2_pytest_bneck.log 2>&1; tail -3 $O/r2_pytest_bneck.log
python -m pytest tests/test_train_gpu.py -m gpu -q -s -k two_rank -p no:cacheprovider > $O/r2_pytest_2gpu.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_2gpu.json 2> $O/r2_bench_2gpu.err
RIH_OVERLAP_ALLREDUCE=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_2gpu_single.json 2> $O/r2_bench_2gpu_single.err
tail -15 $O/r2_pytest_2gpu.log
head -c 600 $O/r2_bench_2gpu.json; echo
tail -c 600 $O/r2_bench_2gpu.err
head -c 400 $O/r2_bench_2gpu_single.json; echo
tail -c 300 $O/r2_bench_2gpu_single.err
