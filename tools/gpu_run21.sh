#!/bin/bash
# session 21: grouped rank-3 TMA boxes for MN-major operands (RIH_TMA_GROUPED): tests with the switch on + A/B
set +e
O=gpurun_out
mkdir -p $O
RIH_TMA_GROUPED=1 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q -p no:cacheprovider > $O/r2_pytest21_grouped.log 2>&1; tail -3 $O/r2_pytest21_grouped.log
for v in 0 1 0 1; do
  RIH_TMA_GROUPED=$v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench21_grp$v.json 2> $O/r2_bench21_grp$v.err
  python -c "import json; d=json.loads(open('$O/r2_bench21_grp$v.json').read().strip().splitlines()[-1]); print('train grouped=$v', d['ms_per_step'])" | tee -a $O/r2_ab21.txt
done
for v in 0 1; do
  RIH_TMA_GROUPED=$v python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench21_hrnet48_grp$v.json 2> $O/r2_bench21_hrnet48_grp$v.err
  python -c "import json; d=json.loads(open('$O/r2_bench21_hrnet48_grp$v.json').read().strip().splitlines()[-1]); print('hrnet48 grouped=$v', d['ms_per_step'])" | tee -a $O/r2_ab21.txt
done
tail -c 600 $O/r2_bench21_grp1.err
