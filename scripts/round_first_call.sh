#!/bin/bash
# First GPU call of a round, in one gpurun invocation (≈ 4 min of box time):
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash scripts/round_first_call.sh r02'
# 1. tcgen05.mma cost table with 1 / 2 / 4 interleaved accumulators (the open question of DESIGN.md section 7 step 0)
# 2. the GPU test tier in one process
# 3. bench.py (defaults) and its reference arm
# 4. ncu launch list of the bench command (profiles/<tag>_ncu_launches_bench.csv is what bench.py's roofline cites)
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
TAG=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
if [ -x scripts/ubench/mma_rates ]; then
  for il in 1 2 4; do timeout 30 scripts/ubench/mma_rates 1000 $il; done > gpurun_out/${TAG}_mma_rates.txt 2>&1
fi
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; grep -E "passed|failed" gpurun_out/${TAG}_gpu_tests.log | tail -2
timeout 120 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 90 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${TAG}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
# 5. A/B builds prepared at the end of round 1 (build them in the container first: python scripts/build_ab.py) (default-build SASS is unaffected by
#    them): parity first (golden logits, every tensor + accumulator, ragged batches, batch 64 vs the direct conv), then speed.
for v in interleave halfk interleave_halfk; do
  [ -f ab/lib_$v.so ] || continue
  extra=""; case $v in *halfk*) extra="IEVM_HALO_RB128=1";; esac
  env IEVM_LIB_PATH=ab/lib_$v.so $extra timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider \
    -k "golden or every_tensor or ragged or direct_conv or fp16_student" > gpurun_out/${TAG}_ab_${v}_tests.log 2>&1
  echo "A/B $v parity rc=$?"; grep -E "passed|failed" gpurun_out/${TAG}_ab_${v}_tests.log | tail -1
  env IEVM_LIB_PATH=ab/lib_$v.so $extra timeout 60 python scripts/layer_times.py 256 ${TAG}_ab_$v > gpurun_out/${TAG}_ab_${v}_layers.txt 2>&1
  tail -3 gpurun_out/${TAG}_ab_${v}_layers.txt
done
