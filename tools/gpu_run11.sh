#!/bin/bash
# session 11: wide wgrad tiles for padded channel counts (HRNet) tests + A/B, ncu launch lists (train step, forward), ncu --set full suite
set +e
O=gpurun_out
mkdir -p $O
RIH_WGRAD_WIDE=3 python -m pytest tests/test_ops_gpu.py tests/test_hrnet_gpu.py -m gpu -q -p no:cacheprovider -k "channels_multiple_of_16 or stride2 or hrnet or conv2d" > $O/r2_pytest11_wide3.log 2>&1; tail -3 $O/r2_pytest11_wide3.log
for v in 1 3; do
  RIH_WGRAD_WIDE=$v python bench.py --encoder hrnet48 --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench11_hrnet48_wide$v.json 2> $O/r2_bench11_hrnet48_wide$v.err
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_b64.csv python tools/profile_step.py --batch 64 --gemm-mode ref > $O/r2_prof_train.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_forward_b64.csv python tools/profile_step.py --batch 64 --gemm-mode ref --fwd-only > $O/r2_prof_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/r2_suite python tools/roofline_suite.py > $O/r2_prof_suite.log 2>&1
ncu -i $O/r2_suite.ncu-rep --page raw --csv > $O/r2_suite_raw.csv 2>/dev/null
python tools/summarize_ncu_suite.py $O/r2_suite_raw.csv > $O/r2_ncu_full_suite_b64.csv 2>$O/r2_suite_sum.err
python tools/summarize_launches.py $O/r2_launches_train_b64.csv > $O/r2_launches_train_b64_summary.txt 2>&1
python tools/summarize_launches.py $O/r2_launches_forward_b64.csv > $O/r2_launches_forward_b64_summary.txt 2>&1
for f in r2_bench11_hrnet48_wide1 r2_bench11_hrnet48_wide3; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
head -30 $O/r2_launches_train_b64_summary.txt; head -12 $O/r2_ncu_full_suite_b64.csv | cut -c1-400; tail -3 $O/r2_prof_suite.log
ls -la $O/r2_suite.ncu-rep
