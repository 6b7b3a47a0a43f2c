#!/bin/bash
# session 8: folded eval BatchNorm + residual-reading TMA epilogue: tests, train bench (+ slowest calls), forward bench A/B, timeline
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest8.log 2>&1
tail -4 $O/r2_pytest8.log
python bench.py --steps 10 --warmup 3 > $O/r2_bench8.json 2> $O/r2_bench8.err
RIH_TMA_RES=0 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench8_tmares0.json 2> $O/r2_bench8_tmares0.err
python bench.py --config forward --steps 20 --warmup 3 > $O/r2_bench8_fwd.json 2> $O/r2_bench8_fwd.err
RIH_BN_FOLD=0 python bench.py --config forward --steps 20 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench8_fwd_nofold.json 2> $O/r2_bench8_fwd_nofold.err
RIH_BN_FOLD=0 RIH_TMA_RES=0 python bench.py --config forward --steps 20 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench8_fwd_nofold_tmares0.json 2> $O/r2_bench8_fwd_nofold_tmares0.err
python tools/timeline.py --out $O/r2_timeline8.csv > $O/r2_timeline8.txt 2>&1
for f in r2_bench8 r2_bench8_tmares0 r2_bench8_fwd r2_bench8_fwd_nofold r2_bench8_fwd_nofold_tmares0; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
