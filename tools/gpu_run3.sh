#!/bin/bash
# GPU session 3 of round 2: tests, A/B of the stem / wgrad / grid-cap changes, determinism diagnosis, compute-sanitizer, timeline.
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > $O/r2_pytest3.log 2>&1
python -m pytest tests/test_train_gpu.py tests/test_parity_wide_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v Warning > $O/r2_pytest3_new.log
RIH_WGRAD_STREAM=0 RIH_HAND_STREAMS=0 python -m pytest tests/test_train_gpu.py -m gpu -q -s -k determinism -p no:cacheprovider 2>&1 | grep -v Warning > $O/r2_determinism_single_stream.log
rm -f $O/r2_trunk_ab3.jsonl $O/r2_bench3_ab.jsonl
for v in "RIH_X=0" "RIH_STEM_IGEMM=0" "RIH_WGRAD_WIDE=0" "RIH_L2_HINTS=0" "RIH_WGRAD_STREAM=0"; do
  env $v python tools/trunk_bench.py >> $O/r2_trunk_ab3.jsonl 2>> $O/r2_trunk_ab3.err
done
python bench.py --steps 10 --warmup 3 > $O/r2_bench3.json 2> $O/r2_bench3.err
for v in "RIH_EW_CAP=0" "RIH_GRID_STREAMS=0" "RIH_STEM_IGEMM=0" "RIH_WGRAD_WIDE=0" "RIH_AUX_CTAS=96"; do
  env $v python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager >> $O/r2_bench3_ab.jsonl 2>> $O/r2_bench3_ab.err
done
python tools/timeline.py --out $O/r2_timeline3.csv > $O/r2_timeline3.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file $O/r2_sanitizer_memcheck.log python tools/sanitize_step.py > $O/r2_sanitizer_memcheck.out 2>&1
echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.out
tail -3 $O/r2_pytest3.log
cat $O/r2_trunk_ab3.jsonl
head -c 300 $O/r2_bench3.json; echo
cut -c1-300 $O/r2_bench3_ab.jsonl
tail -c 300 $O/r2_bench3.err
tail -5 $O/r2_sanitizer_memcheck.log
tail -3 $O/r2_sanitizer_memcheck.out
