#!/bin/bash
# session 22: rank-5 grouped input boxes of the conv wgrad (RIH_TMA_GROUPED bit 1): tests with it on + A/B against bit 0 only
set +e
O=gpurun_out
mkdir -p $O
RIH_TMA_GROUPED=3 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py tests/test_hrnet_gpu.py -m gpu -q -p no:cacheprovider > $O/r2_pytest22_grouped3.log 2>&1; tail -3 $O/r2_pytest22_grouped3.log
for v in 1 3 1 3; do
  RIH_TMA_GROUPED=$v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench22_grp$v.json 2> $O/r2_bench22_grp$v.err
  python -c "import json; d=json.loads(open('$O/r2_bench22_grp$v.json').read().strip().splitlines()[-1]); print('train grouped=$v', d['ms_per_step'])" | tee -a $O/r2_ab22.txt
done
for v in 1 3; do
  RIH_TMA_GROUPED=$v python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench22_hrnet48_grp$v.json 2> $O/r2_bench22_hrnet48_grp$v.err
  python -c "import json; d=json.loads(open('$O/r2_bench22_hrnet48_grp$v.json').read().strip().splitlines()[-1]); print('hrnet48 grouped=$v', d['ms_per_step'])" | tee -a $O/r2_ab22.txt
done
tail -c 600 $O/r2_bench22_grp3.err
