#!/bin/bash
# Runs every `-m gpu` test in its OWN fresh process (first-use / ordering effects cannot hide behind earlier tests).
# usage: tools/gpu_isolated.sh [parallel jobs]   -> gpurun_out/isolated.log (one line per test) ; exit 1 on any failure
J=${1:-6}
mkdir -p gpurun_out
python -m pytest tests --collect-only -q -m gpu 2>/dev/null | grep '::' > gpurun_out/isolated_ids.txt
wc -l gpurun_out/isolated_ids.txt
run_one() {
  out=$(python -m pytest "$1" -x -q -m gpu 2>&1 | tail -25)
  if echo "$out" | grep -qE '^[0-9]+ (passed|skipped)|[0-9]+ passed|[0-9]+ skipped'; then
    if echo "$out" | grep -qE 'failed|error'; then echo "FAIL $1"; echo "$out"; else echo "ok   $1"; fi
  else echo "FAIL $1"; echo "$out"; fi
}
export -f run_one
cat gpurun_out/isolated_ids.txt | xargs -P "$J" -I{} bash -c 'run_one "$@"' _ {} > gpurun_out/isolated.log 2>&1
echo "isolated: $(grep -c '^ok' gpurun_out/isolated.log) ok, $(grep -c '^FAIL' gpurun_out/isolated.log) failed"
grep -A25 '^FAIL' gpurun_out/isolated.log | head -120
! grep -q '^FAIL' gpurun_out/isolated.log
