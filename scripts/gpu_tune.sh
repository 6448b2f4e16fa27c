#!/bin/bash
# parity of the aggregation ops with the default build, then kbench over all variant libraries
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${TESTS:-segment_reduce or layers or csr or powerlaw or config3 or pma or fused}" 2>&1 | tail -8 | tee $OUT/pytest.txt; fi
if [ -z "$SKIP_DEFAULT" ]; then
timeout 300 python scripts/kbench.py 2>&1 | tail -1 | tee $OUT/kbench_default.json
fi
for f in build/variants/lib_*.so; do
  [ -f "$f" ] || continue
  ALLSET_B200_LIB=$PWD/$f timeout 300 python scripts/kbench.py 2>&1 | tail -1 | tee $OUT/kbench_$(basename $f .so).json
done
