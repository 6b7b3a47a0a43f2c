#!/bin/bash
# final session of round 2: tests, smoke, bench (train / forward / reference), ncu launch list of the eval forward
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_final_pytest_gpu.log 2>&1; tail -3 $O/r2_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1; tail -3 $O/r2_final_smoke.log
python bench.py --steps 10 --warmup 3 > $O/r2_final_bench_train.json 2> $O/r2_final_bench_train.err
python bench.py --config forward --steps 20 --warmup 3 > $O/r2_final_bench_forward.json 2> $O/r2_final_bench_forward.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err
for c in 56 96 0; do
  RIH_AUX_CTAS=$c python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_final_ab_auxctas$c.json 2> $O/r2_final_ab_auxctas$c.err
  python -c "import json; d=json.loads(open('$O/r2_final_ab_auxctas$c.json').read().strip().splitlines()[-1]); print('aux ctas $c', d['ms_per_step'])"
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_forward_b64.csv python tools/profile_step.py --batch 64 --gemm-mode ref --fwd-only > $O/r2_prof_fwd.log 2>&1
python tools/summarize_launches.py $O/r2_launches_forward_b64.csv > $O/r2_launches_forward_b64_summary.txt 2>&1
for f in r2_final_bench_train r2_final_bench_forward r2_final_bench_reference; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms_per_step'), d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'), (d.get('e2e') or {}).get('value'), d.get('parity'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
head -12 $O/r2_launches_forward_b64_summary.txt
