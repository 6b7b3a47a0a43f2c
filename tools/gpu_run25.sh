#!/bin/bash
# session 25: the final tree (all grouped TMA boxes on by default): smoke, bench train / forward
set +e
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1; tail -2 $O/r2_final_smoke.log
python bench.py --steps 10 --warmup 3 > $O/r2_final_bench_train.json 2> $O/r2_final_bench_train.err
python bench.py --config forward --steps 20 --warmup 3 --skip-cpu-baseline > $O/r2_final_bench_forward.json 2> $O/r2_final_bench_forward.err
python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider -k "conv2d or linear or gemm" > $O/r2_final_pytest_ops_subset.log 2>&1; tail -2 $O/r2_final_pytest_ops_subset.log
for f in r2_final_bench_train r2_final_bench_forward; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms_per_step'), d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
