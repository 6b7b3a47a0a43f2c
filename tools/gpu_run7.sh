#!/bin/bash
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest7.log 2>&1
python -m pytest tests/test_train_gpu.py tests/test_parity_wide_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v Warning > $O/r2_pytest7_new.log
rm -f $O/r2_trunk_ab7.jsonl $O/r2_bench7_ab.jsonl
for v in "RIH_X=0"; do
  env $v python tools/trunk_bench.py >> $O/r2_trunk_ab7.jsonl 2>> $O/r2_trunk_ab7.err
done
python bench.py --steps 10 --warmup 3 > $O/r2_bench7.json 2> $O/r2_bench7.err
for v in "RIH_L2_HINTS=0"; do
  env $v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager >> $O/r2_bench7_ab.jsonl 2>> $O/r2_bench7_ab.err
done
python tools/timeline.py --out $O/r2_timeline7.csv > $O/r2_timeline7.txt 2>&1
tail -3 $O/r2_pytest7.log
cat $O/r2_trunk_ab7.jsonl
head -c 300 $O/r2_bench7.json; echo
cut -c1-300 $O/r2_bench7_ab.jsonl
tail -c 300 $O/r2_bench7.err
