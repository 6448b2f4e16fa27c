#!/bin/bash
OUT=gpurun_out/${1:-r2u}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 --tb=short -k "pma or PMA or transformer or setgnn or attention or alpha or full_size or powerlaw" > $OUT/pytest_pma.txt 2>&1; tail -4 $OUT/pytest_pma.txt
timeout 300 python scripts/kbench.py 2>&1 | grep '^{' | tee $OUT/kbench_default.json | cut -c1-600
KB_ONLY=pma_v2e,pma_e2v timeout 300 python scripts/kbench.py 10000000 2000000 30 64 2>&1 | grep '^{' | tee $OUT/kbench_d64.json | cut -c1-400
KB_ONLY=pma_v2e,pma_e2v KB_GRAPH=powerlaw timeout 300 python scripts/kbench.py 20000000 3200000 0 256 2>&1 | grep '^{' | tee $OUT/kbench_cfg5_scaled.json | cut -c1-400
