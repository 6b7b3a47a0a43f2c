#!/bin/bash
# session 10: residual prefetch + shared-memory column vectors in the persistent epilogue: tests, train / forward bench with RIH_EPI_OPT A/B
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest10.log 2>&1
tail -4 $O/r2_pytest10.log
python bench.py --steps 10 --warmup 3 > $O/r2_bench10.json 2> $O/r2_bench10.err
for v in 0 1 2; do
  RIH_EPI_OPT=$v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench10_opt$v.json 2> $O/r2_bench10_opt$v.err
  RIH_EPI_OPT=$v python bench.py --config forward --steps 20 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench10_fwd_opt$v.json 2> $O/r2_bench10_fwd_opt$v.err
done
python bench.py --config forward --steps 20 --warmup 3 > $O/r2_bench10_fwd.json 2> $O/r2_bench10_fwd.err
python tools/timeline.py --forward --out $O/r2_timeline10_fwd.csv > $O/r2_timeline10_fwd.txt 2>&1
for f in r2_bench10 r2_bench10_opt0 r2_bench10_opt1 r2_bench10_opt2 r2_bench10_fwd r2_bench10_fwd_opt0 r2_bench10_fwd_opt1 r2_bench10_fwd_opt2; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'), (d.get('roofline') or {}).get('gemm_launches_by_path'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
