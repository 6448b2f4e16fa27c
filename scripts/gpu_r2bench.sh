#!/bin/bash
OUT=gpurun_out/${1:-r2bench}; mkdir -p $OUT
timeout 500 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-200; tail -2 $OUT/bench.err
