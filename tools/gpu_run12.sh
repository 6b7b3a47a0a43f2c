#!/bin/bash
# session 12: full validation of the tree (tests, smoke, bench train / forward / reference / hrnet48) + ncu launch lists + ncu --set full suite
# (everything written under gpurun_out/ must stay below 64 MiB: the .ncu-rep is exported to CSV and deleted)
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest12.log 2>&1; tail -3 $O/r2_pytest12.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke12.log 2>&1; tail -3 $O/r2_smoke12.log
python bench.py --steps 10 --warmup 3 > $O/r2_bench12.json 2> $O/r2_bench12.err
python bench.py --config forward --steps 20 --warmup 3 > $O/r2_bench12_fwd.json 2> $O/r2_bench12_fwd.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench12_reference.json 2> $O/r2_bench12_reference.err
python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline > $O/r2_bench12_hrnet48.json 2> $O/r2_bench12_hrnet48.err
RIH_EPI_OPT=7 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider -k "linear or folded or model or forward" > $O/r2_pytest12_opt7.log 2>&1; tail -2 $O/r2_pytest12_opt7.log
RIH_EPI_OPT=7 python bench.py --config forward --steps 20 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench12_fwd_opt7.json 2> $O/r2_bench12_fwd_opt7.err
RIH_EPI_OPT=7 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench12_opt7.json 2> $O/r2_bench12_opt7.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_b64.csv python tools/profile_step.py --batch 64 --gemm-mode ref > $O/r2_prof_train.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_forward_b64.csv python tools/profile_step.py --batch 64 --gemm-mode ref --fwd-only > $O/r2_prof_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_suite python tools/roofline_suite.py > $O/r2_prof_suite.log 2>&1
ncu -i /tmp/r2_suite.ncu-rep --page raw --csv > $O/r2_suite_raw.csv 2>/dev/null
python tools/summarize_ncu_suite.py $O/r2_suite_raw.csv > $O/r2_ncu_full_suite_b64.csv 2>$O/r2_suite_sum.err
python tools/summarize_launches.py $O/r2_launches_train_b64.csv > $O/r2_launches_train_b64_summary.txt 2>&1
python tools/summarize_launches.py $O/r2_launches_forward_b64.csv > $O/r2_launches_forward_b64_summary.txt 2>&1
for f in r2_bench12 r2_bench12_opt7 r2_bench12_fwd r2_bench12_fwd_opt7 r2_bench12_reference r2_bench12_hrnet48; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms_per_step'), d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'), (d.get('e2e') or {}).get('value'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
head -12 $O/r2_launches_forward_b64_summary.txt
du -sh $O
