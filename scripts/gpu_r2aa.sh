#!/bin/bash
OUT=gpurun_out/${1:-r2aa}; mkdir -p $OUT
for a in 2 3 4 6; do
ALLSET_MLP2_L2_PREFETCH=$a ALLSET_WGRAD_L2_PREFETCH=$a timeout 200 python scripts/mlp2_ab.py | tee -a $OUT/ab.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['env'].get('ALLSET_MLP2_L2_PREFETCH'), {k: round(v, 4) for k, v in d.items() if k != 'env'})"
done
