#!/bin/bash
# Relative sizes of the three coding slices (first = highest stream priority). usage: bash tools/sweep_split.sh
for sp in "1,1,1" "5,4,3" "3,2,1" "3,4,5" "2,1,1" "4,4,3,1"; do
  g=$(echo $sp | tr ',' '\n' | wc -l)
  line=$(FPCC_GROUP_SPLIT=$sp python bench.py --steps 4 --warmup 2 --no-cpu-baseline --groups $g 2>&1 >/dev/null | grep "step_device" | tr '\n' ' ')
  echo "split=$sp: $line"
done
