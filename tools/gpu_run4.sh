#!/bin/bash
# GPU session 4: tests, A/B of direct stride-2 / bitmask, bench, timeline
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest4.log 2>&1
python -m pytest tests/test_train_gpu.py tests/test_parity_wide_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v Warning > $O/r2_pytest4_new.log
rm -f $O/r2_trunk_ab4.jsonl $O/r2_bench4_ab.jsonl
for v in "RIH_X=0" "RIH_S2_DIRECT=0"; do
  env $v python tools/trunk_bench.py >> $O/r2_trunk_ab4.jsonl 2>> $O/r2_trunk_ab4.err
done
python bench.py --steps 10 --warmup 3 > $O/r2_bench4.json 2> $O/r2_bench4.err
for v in "RIH_S2_DIRECT=0"; do
  env $v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager >> $O/r2_bench4_ab.jsonl 2>> $O/r2_bench4_ab.err
done
python tools/timeline.py --out $O/r2_timeline4.csv > $O/r2_timeline4.txt 2>&1
tail -3 $O/r2_pytest4.log
cat $O/r2_trunk_ab4.jsonl
head -c 300 $O/r2_bench4.json; echo
cut -c1-300 $O/r2_bench4_ab.jsonl
tail -c 300 $O/r2_bench4.err
