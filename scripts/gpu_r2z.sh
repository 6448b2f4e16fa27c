#!/bin/bash
OUT=gpurun_out/${1:-r2z}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=short -k "wgrad or autograd" > $OUT/pytest_wgrad.txt 2>&1; tail -3 $OUT/pytest_wgrad.txt
timeout 300 python scripts/linear_bench.py 2>&1 | tee $OUT/linear_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['d'], d['dtype'], {k: (v['ms'], v['frac'], v['cublas_ms']) for k, v in d.items() if isinstance(v, dict)})"
