#!/bin/bash
# round 2, call T (1 GPU): source-level ncu capture of the PMA stream kernel, E->V direction
OUT=gpurun_out/${1:-r2t}; mkdir -p $OUT
T=/tmp/prof_pma
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pma_stream -s 1 -c 1 -o $T -f \
  env KB_ONLY=pma_e2v KB_ITERS=2 python scripts/kbench.py > $OUT/ncu_pma.log 2>&1; tail -2 $OUT/ncu_pma.log
ncu -i $T.ncu-rep --page raw --csv > $OUT/pma_raw.csv 2>/dev/null
ncu -i $T.ncu-rep --page source --csv > $OUT/pma_source.csv 2>/dev/null
ls -la $OUT
