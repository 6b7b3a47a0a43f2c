#!/bin/bash
# final session of round 2 (after the grouped TMA boxes): tests, smoke, bench train / forward / reference / hrnet48 with all baselines, train launch list
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_final_pytest_gpu.log 2>&1; tail -3 $O/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1; tail -3 $O/r2_final_smoke.log
python bench.py --steps 10 --warmup 3 > $O/r2_final_bench_train.json 2> $O/r2_final_bench_train.err
python bench.py --config forward --steps 20 --warmup 3 > $O/r2_final_bench_forward.json 2> $O/r2_final_bench_forward.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err
python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline > $O/r2_final_bench_hrnet48.json 2> $O/r2_final_bench_hrnet48.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_b64.csv python tools/profile_step.py --batch 64 --gemm-mode ref > $O/r2_prof_train.log 2>&1
python tools/summarize_launches.py $O/r2_launches_train_b64.csv > $O/r2_launches_train_b64_summary.txt 2>&1
for f in r2_final_bench_train r2_final_bench_forward r2_final_bench_reference r2_final_bench_hrnet48; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms_per_step'), d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
head -14 $O/r2_launches_train_b64_summary.txt
