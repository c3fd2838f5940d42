// Measured tensor-core peak of THIS device, for the roofline's denominator (bench.py): a stream of tcgen05.mma
// M128 x N256 x K32B instructions (weights-in-TMEM form, so only 64 B/clk of shared-memory operand traffic) from one
// CTA per SM, timed with CUDA events by the host.  MEASURED_PEAKS.json holds a cuBLAS bf16 figure only; kind::i8 has no
// library GEMM to measure against, so the engine measures its own instruction stream (profiles/r02_mma_rates2.txt:
// 128.0 clk per instruction = 8 189 MAC/clk/SM for kind::i8, 4 094 for kind::f16, independent of the accumulate chain).
#pragma once
#include "ptx.cuh"

namespace ievm {

template <int kKind>     // 0 = kind::i8, 1 = kind::f16
__global__ void __launch_bounds__(128) probe_mma_kernel(int iters, unsigned int* fail) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  constexpr int kBBytes = 256 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kBBytes);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < kBBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  const uint32_t idesc = kKind == 0 ? make_idesc_i8_u8s8(256, 128) : make_idesc_f16(256, 128);
  const uint32_t hi = smem_desc_hi(128);
  const uint32_t b_lo = smem_desc_lo(smem_u32(smem));
  if (threadIdx.x < 32) {
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j)     // operand A: TMEM columns 256.. (8 columns per 32-byte k-step), accumulator: columns 0..255
          umma_ts<kKind>(tmem, tmem + 256u + static_cast<uint32_t>(j & 3) * 8u, b_lo + 2u * (j & 3), hi, idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    const unsigned long long deadline = globaltimer_ns() + 4000000000ull;
    while (!mbar_try_wait(bar, 0)) {
      if (globaltimer_ns() > deadline) {
        *fail = 1;
        break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

}  // namespace ievm
