#!/bin/bash
# Run every GPU test in its own process (a trapped kernel poisons the CUDA context) with a timeout,
# collecting results under gpurun_out/.  Usage: scripts/gpu_tests_isolated.sh [pytest -k expression]
mkdir -p gpurun_out
OUT=gpurun_out/gpu_tests.log
: > $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $OUT 2>&1
ids=$(python -m pytest tests/test_gpu_parity.py -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
pass=0; failn=0
for id in $ids; do
  echo "=== $id" >> $OUT
  timeout 300 python -m pytest "$id" -x -q -m gpu -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|^$" | tail -25 >> $OUT
  rc=${PIPESTATUS[0]}
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id"; else failn=$((failn+1)); echo "FAIL($rc) $id"; fi
done
echo "passed=$pass failed=$failn" | tee -a $OUT
