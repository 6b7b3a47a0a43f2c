#!/bin/bash
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest6.log 2>&1
python -m pytest tests/test_train_gpu.py tests/test_parity_wide_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v Warning > $O/r2_pytest6_new.log
rm -f $O/r2_trunk_ab6.jsonl $O/r2_bench6_ab.jsonl
for v in "RIH_X=0"; do
  env $v python tools/trunk_bench.py >> $O/r2_trunk_ab6.jsonl 2>> $O/r2_trunk_ab6.err
done
python bench.py --steps 10 --warmup 3 > $O/r2_bench6.json 2> $O/r2_bench6.err
for v in "RIH_HAND_CTAS=48" "RIH_HAND_CTAS=72" "RIH_HAND_CTAS=100" "RIH_AUX_CTAS=48" "RIH_AUX_CTAS=110" "RIH_GRID_STREAMS=0"; do
  env $v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager >> $O/r2_bench6_ab.jsonl 2>> $O/r2_bench6_ab.err
done
python tools/timeline.py --out $O/r2_timeline6.csv > $O/r2_timeline6.txt 2>&1
tail -3 $O/r2_pytest6.log
cat $O/r2_trunk_ab6.jsonl
head -c 300 $O/r2_bench6.json; echo
cut -c1-300 $O/r2_bench6_ab.jsonl
tail -c 300 $O/r2_bench6.err
