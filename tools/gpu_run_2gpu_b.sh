#!/bin/bash
# 2-GPU session: the 2-rank NCCL equivalence test only
set +e
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -s -k two_rank -p no:cacheprovider > $O/r2_pytest_2gpu_b.log 2>&1
tail -8 $O/r2_pytest_2gpu_b.log
