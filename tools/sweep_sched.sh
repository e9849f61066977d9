#!/bin/bash
# Schedule sweep of the bench step on one GPU: per-group encode->decode pipelines vs phases, stream priorities, group counts.
# usage (GPU box): bash tools/sweep_sched.sh > gpurun_out/sweep_sched.txt
for cfg in "1 1 3" "0 0 3" "1 0 3" "0 1 3" "1 1 4" "1 1 6" "1 1 2" "0 0 1"; do
  set -- $cfg
  line=$(FPCC_BENCH_PIPE=$1 FPCC_STREAM_PRIO=$2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --groups $3 2>&1 >/dev/null | grep "step_device\|step_e2e\|Error\|error" | tr '\n' ' ')
  echo "pipe=$1 prio=$2 groups=$3: $line"
done
