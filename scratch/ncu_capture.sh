#!/bin/bash
# ncu_capture.sh <name> <kernel-regex> <skip> -- <command...>
# One `ncu --set full` capture of one launch, exported on the box to CSV (raw + source pages) so that the
# 64 MiB return limit of gpurun_out/ is never hit; the .ncu-rep itself is dropped.
set -u
NAME=$1; KREGEX=$2; SKIP=$3; shift 4
OUT=gpurun_out/ncu
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c 1 -o $OUT/$NAME -f "$@" > $OUT/$NAME.log 2>&1
if [ -f $OUT/$NAME.ncu-rep ]; then
  ncu -i $OUT/$NAME.ncu-rep --page raw --csv > $OUT/$NAME.raw.csv 2>/dev/null
  ncu -i $OUT/$NAME.ncu-rep --page source --csv --print-source sass > $OUT/$NAME.sass.csv 2>/dev/null
  gzip -f $OUT/$NAME.sass.csv
  rm -f $OUT/$NAME.ncu-rep
else
  tail -5 $OUT/$NAME.log
fi
ls -la $OUT | tail -5
