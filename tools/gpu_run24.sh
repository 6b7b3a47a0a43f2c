#!/bin/bash
# session 24: grouped weight boxes in the conv dgrad (RIH_TMA_GROUPED bit 2): full tests with it on, A/B, smoke + bench of the tree
set +e
O=gpurun_out
mkdir -p $O
for v in 3 7 3 7; do
  RIH_TMA_GROUPED=$v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench24_grp$v.json 2> $O/r2_bench24_grp$v.err
  python -c "import json; d=json.loads(open('$O/r2_bench24_grp$v.json').read().strip().splitlines()[-1]); print('train grouped=$v', d['ms_per_step'])" | tee -a $O/r2_ab24.txt
done
RIH_TMA_GROUPED=7 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest24_grouped7.log 2>&1; tail -3 $O/r2_pytest24_grouped7.log
python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider > $O/r2_pytest24_default.log 2>&1; tail -2 $O/r2_pytest24_default.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke24.log 2>&1; tail -2 $O/r2_smoke24.log
RIH_TMA_GROUPED=7 python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench24_hrnet48_grp7.json 2> $O/r2_bench24_hrnet48_grp7.err
python -c "import json; d=json.loads(open('$O/r2_bench24_hrnet48_grp7.json').read().strip().splitlines()[-1]); print('hrnet48 grouped=7', d['ms_per_step'])" | tee -a $O/r2_ab24.txt
