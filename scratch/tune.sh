#!/bin/bash
for lib in scratch/variants/*.so; do
  echo "=== $lib"
  QMC4_B200_LIB=$PWD/$lib python scratch/eval_bench.py 2>&1 | tail -3
done
