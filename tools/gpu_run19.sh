#!/bin/bash
# session 19: warp-wide TMA issue in the producer warp (rih_set_epilogue_opt bit 2): tests + A/B on the train step, the trunk, HRNet-w48
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider > $O/r2_pytest19_ops.log 2>&1; tail -3 $O/r2_pytest19_ops.log
for v in 3 7; do
  RIH_EPI_OPT=$v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench19_opt$v.json 2> $O/r2_bench19_opt$v.err
  RIH_EPI_OPT=$v python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench19_hrnet48_opt$v.json 2> $O/r2_bench19_hrnet48_opt$v.err
  RIH_EPI_OPT=$v python bench.py --config forward --steps 20 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench19_fwd_opt$v.json 2> $O/r2_bench19_fwd_opt$v.err
done
for f in r2_bench19_opt3 r2_bench19_opt7 r2_bench19_hrnet48_opt3 r2_bench19_hrnet48_opt7 r2_bench19_fwd_opt3 r2_bench19_fwd_opt7; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms_per_step'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --deselect tests/test_ops_gpu.py > $O/r2_pytest19_rest.log 2>&1; tail -3 $O/r2_pytest19_rest.log
python tools/timeline.py --out $O/r2_timeline19.csv > $O/r2_timeline19.txt 2>&1
