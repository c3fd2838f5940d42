#!/bin/bash
# One gpurun call for the round's evidence: GPU test tier, bench + reference arm, ncu launch list of the bench command,
# ncu --set full of one forward (all 18 launches) for profiles/r02_ncu_full_int8_r18_pruned.json.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed" gpurun_out/${TAG}_gpu_tests.log | tail -2
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 100 python bench.py --steps 2 --warmup 1 --no-extra --no-latency --no-cpu-baseline > gpurun_out/${TAG}_bench_short.json 2>> gpurun_out/${TAG}_bench.err &&
IEVM_WAIT_LIMIT_MS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra --no-latency --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 100 python scripts/forward_once.py 256 1 int8 > gpurun_out/plain.log 2>&1 &&
IEVM_WAIT_LIMIT_MS=0 timeout 500 ncu --set full --clock-control none --import-source on -c 18 -o gpurun_out/${TAG}_ncu_full_int8 -f \
  python scripts/forward_once.py 256 1 int8 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/${TAG}_ncu_full_int8.ncu-rep
