#!/bin/bash
for lib in scratch/variants/*.so; do
  echo "=== $lib"
  for args in "2 16384 30 1000" "5 2368 10 1000" "3 4736 10 1000"; do
    QMC4_B200_LIB=$PWD/$lib python scratch/prof_cfg.py $args 2>&1 | tail -1
  done
done
