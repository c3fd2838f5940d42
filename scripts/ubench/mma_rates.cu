// Micro-benchmark for the decision DESIGN.md section 7 step (0) hinges on: what ONE tcgen05.mma (kind::i8, K = 32 bytes)
// costs on a B200 SM as a function of its shape and of where the M-side operand lives.
//
//   SS form: A (M rows x 32 B) and B (N rows x 32 B) both fetched from shared memory
//   TS form: A in tensor memory, only B crosses the shared-memory read port
//
// for M in {64, 128}, N in {64, 128, 160, 256}, operand rows of 64 and 128 bytes (64B / 128B swizzle), issued back to
// back by one thread the way the conv kernels do (k-steps advance the descriptor start by 32 B inside the swizzled row,
// `taps` different operand tiles in turn).  One CTA per SM on every SM (so that the number is the loaded-chip one),
// operands are zeros (timing only), accumulators alternate between two TMEM buffers.
//
// argv: [iterations = 2000] [interleave = 1]; interleave > 1 sends consecutive instructions to that many independent
// accumulators, which separates a per-instruction latency on a dependent chain from a throughput limit.
// Prints clk per MMA, the shared-memory operand bytes per MMA, bytes/clk and MACs/clk per SM.  Round-1 measurements
// that this generalises: SS M128 N64 ~96 clk (6 KB), TS M128 N64 ~73 clk (conv_wt.cuh).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../inference-efficient-vision-models_b200/csrc mma_rates.cu -o mma_rates
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace ievm;

struct Case {
  int ts;          // 0 = SS, 1 = TS
  int m, n;
  int row_bytes;   // 64 or 128
  int interleave;  // consecutive MMAs go round-robin to this many independent accumulators (1 = one dependent chain)
};

__global__ void __launch_bounds__(128)
mma_rate_kernel(Case c, int iters, unsigned long long* cycles, unsigned int* fail) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  constexpr int kTaps = 4;                               // distinct operand tiles visited in turn
  const int a_tile = 128 * c.row_bytes;                  // A tile: 128 rows (M = 64 uses the first 64)
  const int b_tile = 256 * c.row_bytes;                  // B tile: up to 256 rows
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_tile;                           // one A tile, kTaps B tiles would not fit at 128 B rows: B tiles
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * b_tile);   // alternate between two
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (a_tile + 2 * b_tile) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
  const uint32_t tmem = *slot;
  const uint32_t idesc = make_idesc_i8_u8s8(c.n, c.m);
  const uint32_t hi = smem_desc_hi(c.row_bytes);
  const uint32_t a_lo = smem_desc_lo(smem_u32(sA));
  const uint32_t b_lo = smem_desc_lo(smem_u32(sB));
  const int ksteps = c.row_bytes / 32;
  // accumulator buffers of 64 / 128 / 256 columns in columns 0..383 (the TS operand sits behind them)
  const uint32_t acc_cols = c.n <= 64 ? 64u : (c.n <= 128 ? 128u : 256u);
  const uint32_t max_buf = 384u / acc_cols;
  const uint32_t inter = static_cast<uint32_t>(c.interleave) < max_buf ? static_cast<uint32_t>(c.interleave) : max_buf;
  const uint32_t nbuf = inter > 1 ? inter : (max_buf > 1 ? 2u : 1u);
  const uint32_t a_col = 2u * 128u + 128u;               // TS operand: columns 384.. (8 columns per 32-byte k-step)
  unsigned long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    uint32_t buf = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t seq = 0;
      for (int tap = 0; tap < kTaps; ++tap) {
        const uint32_t boff = static_cast<uint32_t>((tap & 1) * b_tile) >> 4;
        for (int ks = 0; ks < ksteps; ++ks, ++seq) {
          // inter == 1: the whole iteration accumulates into one buffer (a tile's K loop), buffers alternate per iteration;
          // inter > 1: consecutive instructions go to different buffers (several tiles' K loops interleaved)
          const uint32_t b = inter > 1 ? seq % inter : buf;
          const uint32_t d = tmem + b * acc_cols;
          const uint32_t acc = (inter > 1 ? seq >= inter : seq > 0) ? 1u : 0u;
          if (c.ts) umma_ts<0>(d, tmem + a_col + static_cast<uint32_t>(ks) * 8u, b_lo + boff + 2u * ks, hi, idesc, acc);
          else umma_i8_lohi(d, a_lo + 2u * ks, b_lo + boff + 2u * ks, hi, idesc, acc);
        }
      }
      buf = (buf + 1) % nbuf;
    }
    umma_commit(bar);
    const unsigned long long deadline = globaltimer_ns() + 2000000000ull;
    while (!mbar_try_wait(bar, 0)) {
      if (globaltimer_ns() > deadline) {
        *fail = 1;
        break;
      }
    }
    t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  const int interleave = argc > 2 ? atoi(argv[2]) : 1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess || prop.major != 10) {
    fprintf(stderr, "needs an sm_100 device\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount;
  const int smem = 1024 + 128 * 128 + 2 * 256 * 128 + 64;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* d_cycles;
  unsigned int* d_fail;
  cudaMalloc(&d_cycles, sms * sizeof(unsigned long long));
  cudaMalloc(&d_fail, sizeof(unsigned int));
  std::vector<unsigned long long> h(sms);
  printf("interleave = %d independent accumulators\n", interleave);
  printf("%-4s %4s %4s %5s | %9s %10s %8s %9s\n", "form", "M", "N", "rowB", "clk/MMA", "smem B/MMA", "B/clk", "MAC/clk");
  for (int ts = 0; ts <= 1; ++ts)
    for (int m : {128, 64})
      for (int n : {64, 128, 160, 256})
        for (int rb : {64, 128}) {
          Case c{ts, m, n, rb, interleave};
          cudaMemset(d_fail, 0, sizeof(unsigned int));
          for (int rep = 0; rep < 2; ++rep)                 // first launch warms up
            mma_rate_kernel<<<sms, 128, smem>>>(c, iters, d_cycles, d_fail);
          const cudaError_t e = cudaDeviceSynchronize();
          unsigned int failed = 0;
          cudaMemcpy(&failed, d_fail, sizeof(failed), cudaMemcpyDeviceToHost);
          if (e != cudaSuccess || failed) {
            printf("%-4s %4d %4d %5d | failed (%s)\n", ts ? "TS" : "SS", m, n, rb, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            continue;
          }
          cudaMemcpy(h.data(), d_cycles, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
          double sum = 0;
          for (auto v : h) sum += static_cast<double>(v);
          const double mmas = static_cast<double>(iters) * 4 * (rb / 32);
          const double clk = sum / sms / mmas;
          const double bytes = (ts ? 0 : m * 32.0) + n * 32.0;
          printf("%-4s %4d %4d %5d | %9.1f %10.0f %8.1f %9.0f\n", ts ? "TS" : "SS", m, n, rb, clk, bytes, bytes / clk,
                 static_cast<double>(m) * n * 32 / clk);
        }
  return 0;
}
