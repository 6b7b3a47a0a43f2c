#!/bin/bash
# GPU session 2 of round 2: full GPU test suite, trunk-only A/B of tuning variants, full-step A/B, timeline.
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest2.log 2>&1
python -m pytest tests/test_train_gpu.py tests/test_parity_wide_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v Warning > $O/r2_pytest2_new.log
rm -f $O/r2_trunk_ab.jsonl $O/r2_bench2_ab.jsonl
for v in "RIH_X=0" "RIH_LIB_VARIANT=wg1" "RIH_LIB_VARIANT=deep" "RIH_LIB_VARIANT=wg1deep" "RIH_SERPENTINE=1" "RIH_L2_HINTS=1" "RIH_SERPENTINE=1 RIH_L2_HINTS=1" "RIH_PDL=1" "RIH_PDL=1 RIH_SERPENTINE=1 RIH_L2_HINTS=1 RIH_LIB_VARIANT=deep"; do
  env $v python tools/trunk_bench.py >> $O/r2_trunk_ab.jsonl 2>> $O/r2_trunk_ab.err
done
python bench.py --steps 10 --warmup 3 > $O/r2_bench2.json 2> $O/r2_bench2.err
for v in "RIH_PDL=1" "RIH_FUSED_QKV=0 RIH_RES_ALIAS=0 RIH_GRID_STREAMS=0" "RIH_NARROW_TILES=0" "RIH_GRID_STREAMS=0"; do
  env $v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager >> $O/r2_bench2_ab.jsonl 2>> $O/r2_bench2_ab.err
done
python tools/timeline.py --out $O/r2_timeline2.csv > $O/r2_timeline2.txt 2>&1
tail -3 $O/r2_pytest2.log
cat $O/r2_trunk_ab.jsonl
head -c 400 $O/r2_bench2.json; echo
cut -c1-330 $O/r2_bench2_ab.jsonl
tail -c 400 $O/r2_bench2.err
