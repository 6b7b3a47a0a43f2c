#!/bin/bash
# aux-stream CTA cap re-tuned after the trunk got faster (alternating order, same box)
set +e
O=gpurun_out
mkdir -p $O
for c in 72 96 120 96 72 120; do
  RIH_AUX_CTAS=$c python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-gpu-eager > $O/r2_ab15_$c.json 2> $O/r2_ab15_$c.err
  python -c "import json; d=json.loads(open('$O/r2_ab15_$c.json').read().strip().splitlines()[-1]); print('aux ctas $c', d['ms_per_step'])" | tee -a $O/r2_ab15_auxctas.txt
done
