#!/bin/bash
# session 9: forward timeline, HRNet-w48 (tests with the folded eval path, bench train + forward, timeline), mano / graph / newgraph bench lines
set +e
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_hrnet_gpu.py tests/test_myhand_gpu.py -m gpu -q -p no:cacheprovider > $O/r2_pytest9_hrnet.log 2>&1; tail -3 $O/r2_pytest9_hrnet.log
python tools/timeline.py --forward --out $O/r2_timeline9_fwd.csv > $O/r2_timeline9_fwd.txt 2>&1
python bench.py --encoder hrnet48 --steps 10 --warmup 3 > $O/r2_bench9_hrnet48.json 2> $O/r2_bench9_hrnet48.err
python tools/timeline.py --encoder hrnet48 --batch 32 --out $O/r2_timeline9_hrnet48.csv > $O/r2_timeline9_hrnet48.txt 2>&1
python bench.py --config forward --encoder hrnet48 --steps 20 --warmup 3 --skip-cpu-baseline > $O/r2_bench9_fwd_hrnet48.json 2> $O/r2_bench9_fwd_hrnet48.err
python bench.py --config mano > $O/r2_bench9_mano.json 2> $O/r2_bench9_mano.err
python bench.py --encoder graph --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench9_graph.json 2> $O/r2_bench9_graph.err
python bench.py --encoder newgraph --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_bench9_newgraph.json 2> $O/r2_bench9_newgraph.err
for f in r2_bench9_hrnet48 r2_bench9_fwd_hrnet48 r2_bench9_mano r2_bench9_graph r2_bench9_newgraph; do echo $f; python - <<P
import json
try:
    d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]); print(d['value'], d['unit'], d['ms_per_step'], d.get('launches_per_step'), d.get('speedup_vs_gpu_eager'))
except Exception as e:
    print('ERR', e); print(open('$O/$f.err').read()[-1500:])
P
done
head -30 $O/r2_timeline9_fwd.txt
