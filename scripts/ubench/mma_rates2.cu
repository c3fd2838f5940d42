// Second tcgen05.mma cost table (round 2).  mma_rates.cu turned out to be bound by its own issue loop (runtime
// modulo / descriptor arithmetic per instruction: a flat ~156 clk for every shape, and 290 clk with the "interleave"
// option, which only added an integer division).  Here every shape is a template instance whose issue loop is 16
// back-to-back instructions with compile-time descriptor offsets, so what is measured is the tensor unit:
//
//   form SS : A (M rows x 32 B) and B (N rows x 32 B) from shared memory
//   form TS : A from tensor memory, B from shared memory
//   form P2 : cta_group::2, M = 256 over a CTA pair, each CTA supplies 128 A rows and N/2 B rows (SS)
//   kind i8 (K = 32 B) or f16 (K = 16 halves = 32 B); operand rows of 64 or 128 bytes (64B / 128B swizzle)
//   IL = number of independent accumulators consecutive instructions rotate over (1 = one dependent chain)
//
// One CTA (or pair) per SM on all SMs, operands are zeros.  Prints clk per instruction, bytes of shared-memory operand
// per instruction per CTA, bytes/clk and MAC/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../inference-efficient-vision-models_b200/csrc mma_rates2.cu -o mma_rates2
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace ievm;

enum { SS = 0, TS = 1, P2 = 2 };

__device__ __forceinline__ bool try_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int KIND, int FORM, int M, int N, int ROWB, int IL>
__global__ void __launch_bounds__(128) rate_kernel(int iters, unsigned long long* cycles, unsigned int* fail) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  constexpr int kATile = 128 * ROWB;
  constexpr int kBTile = 256 * ROWB + 4 * ROWB * 64;      // room for tap-shifted views
  uint8_t* sA = smem;
  uint8_t* sB = smem + kATile;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kBTile);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (kATile + kBTile) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (FORM == P2) {
      tmem_alloc_pair(slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(slot, 512);
      tmem_relinquish();
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (FORM == P2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  constexpr int kMmaM = FORM == P2 ? 256 : M;
  const uint32_t idesc = KIND == 0 ? make_idesc_i8_u8s8(N, kMmaM) : make_idesc_f16(N, kMmaM);
  const uint32_t hi = smem_desc_hi(ROWB);
  const uint32_t a_lo = smem_desc_lo(smem_u32(sA));
  const uint32_t b_lo = smem_desc_lo(smem_u32(sB));
  constexpr int kSteps = ROWB / 32;
  constexpr uint32_t kAccCols = N <= 64 ? 64u : (N <= 128 ? 128u : 256u);
  constexpr uint32_t kMaxBuf = 384u / kAccCols;
  constexpr uint32_t kIl = IL < (int)kMaxBuf ? IL : kMaxBuf;
  constexpr uint32_t kACol = 384u;
  const bool issuer_warp = threadIdx.x < 32 && (FORM != P2 || cluster_ctarank() == 0);
  if (issuer_warp) {
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int tap = j / kSteps;
          const int ks = j % kSteps;
          const uint32_t boff = static_cast<uint32_t>((tap % 4) * 3 * (ROWB / 16));     // row-shifted views, as the halo taps
          const uint32_t d = tmem + static_cast<uint32_t>(j % kIl) * kAccCols;
          if (FORM == TS) {
            if (KIND == 0) umma_ts<0>(d, tmem + kACol + static_cast<uint32_t>(ks) * 8u, b_lo + boff + 2u * ks, hi, idesc, 1u);
            else umma_ts<1>(d, tmem + kACol + static_cast<uint32_t>(ks) * 8u, b_lo + boff + 2u * ks, hi, idesc, 1u);
          } else if (FORM == P2) {
            if (KIND == 0) umma_i8_lohi_pair(d, a_lo + 2u * ks, b_lo + boff + 2u * ks, hi, idesc, 1u);
            else umma_f16_lohi_pair(d, a_lo + 2u * ks, b_lo + boff + 2u * ks, hi, idesc, 1u);
          } else {
            if (KIND == 0) umma_i8_lohi(d, a_lo + 2u * ks, b_lo + boff + 2u * ks, hi, idesc, 1u);
            else umma_f16_lohi(d, a_lo + 2u * ks, b_lo + boff + 2u * ks, hi, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if (FORM == P2) umma_commit_pair(bar);
      else umma_commit(bar);
    }
    __syncwarp();
    const unsigned long long deadline = globaltimer_ns() + 2000000000ull;
    while (!mbar_try_wait(bar, 0)) {
      if (globaltimer_ns() > deadline) {
        *fail = 1;
        break;
      }
    }
    if (lane_id() == 0) cycles[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 32 && FORM == P2 && cluster_ctarank() != 0) {
    cycles[blockIdx.x] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (FORM == P2) cluster_sync_all();
  if (threadIdx.x < 32) {
    if (FORM == P2) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

static int g_sms = 148;
static unsigned long long* d_cycles;
static unsigned int* d_fail;

template <int KIND, int FORM, int M, int N, int ROWB, int IL>
void run_case(int iters) {
  const int smem = 1024 + 128 * ROWB + 256 * ROWB + 4 * ROWB * 64 + 64;
  auto kern = rate_kernel<KIND, FORM, M, N, ROWB, IL>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d_fail, 0, sizeof(unsigned int));
  const int grid = FORM == P2 ? g_sms / 2 * 2 : g_sms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = FORM == P2 ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, iters, d_cycles, d_fail);
  }
  const cudaError_t e = cudaDeviceSynchronize();
  unsigned int failed = 0;
  cudaMemcpy(&failed, d_fail, sizeof(failed), cudaMemcpyDeviceToHost);
  const char* form = FORM == SS ? "SS" : (FORM == TS ? "TS" : "P2");
  if (e != cudaSuccess || failed) {
    printf("%-3s %-3s %4d %4d %5d %2d | failed (%s)\n", KIND ? "f16" : "i8", form, M, N, ROWB, IL, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    return;
  }
  std::vector<unsigned long long> h(grid);
  cudaMemcpy(h.data(), d_cycles, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  int cnt = 0;
  for (auto v : h)
    if (v) {
      sum += static_cast<double>(v);
      ++cnt;
    }
  const double clk = sum / cnt / (static_cast<double>(iters) * 16);
  const double a_bytes = FORM == TS ? 0 : 128 * 32.0 * (M == 64 ? 0.5 : 1.0);
  const double b_bytes = (FORM == P2 ? N / 2 : N) * 32.0;
  const double macs_per_sm = static_cast<double>(FORM == P2 ? 128 : M) * N * (KIND ? 16 : 32);
  printf("%-3s %-3s %4d %4d %5d %2d | %8.1f %9.0f %7.1f %8.0f\n", KIND ? "f16" : "i8", form, FORM == P2 ? 256 : M, N, ROWB, IL,
         clk, a_bytes + b_bytes, (a_bytes + b_bytes) / clk, macs_per_sm / clk);
  fflush(stdout);
}


// ---- the product's halo-mode issue loop, feature by feature ----
// A = 128-row view of a (rows x 58)-pixel patch starting at pixel (x0 + ky * 58 + kx), B = weight tile of tap (ky, kx);
// SHIFT = 0: every view starts at the patch origin (aligned), 1: as the product; COMMIT = 1: two tcgen05.commit per tile;
// ZACC = 1: the first instruction of a tile overwrites the accumulator; NBUF accumulators in rotation.
template <int N, int ROWB, int SHIFT, int COMMIT, int ZACC, int NBUF, int GAP, int WAITERS>
__global__ void __launch_bounds__(128 + 32 * WAITERS) convlike_kernel(int tiles, int fill, unsigned long long* cycles, unsigned int* fail) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  constexpr int kWp = 58;
  constexpr int kATile = 7 * kWp * ROWB + 1024;
  constexpr int kBTile = 9 * N * ROWB;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((kATile + 1023) / 1024) * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + kBTile);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  for (int i = threadIdx.x; i < (((kATile + 1023) / 1024) * 1024 + kBTile) / 4; i += blockDim.x) {
    uint32_t v = static_cast<uint32_t>(i) * 2654435761u + blockIdx.x * 40503u;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = fill == 0 ? 0u : (fill == 1 ? v : (v & 0x3f3f3f3fu));
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], 1);
    mbar_init(&bar[3], 1);
    mbar_init(&bar[4], 1);
    mbar_init(&bar[5], 1);
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
  const uint32_t idesc = make_idesc_i8_u8s8(N, 128);
  const uint32_t hi = smem_desc_hi(ROWB);
  const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA));
  const uint32_t b_lo0 = smem_desc_lo(smem_u32(sB));
  constexpr uint32_t row16 = ROWB / 16;
  constexpr uint32_t b_step = N * ROWB / 16;
  constexpr int kSteps = ROWB / 32;
  constexpr uint32_t kAccCols = N <= 64 ? 64u : (N <= 128 ? 128u : 256u);
  if (threadIdx.x < 32) {
    const unsigned long long t0 = clock64();
    int acc = 0;
    for (int tile = 0; tile < tiles; ++tile) {
      if (GAP == 1) {           // what the product's issuer does between tiles: two waits on (already complete) barriers + fence
        while (!mbar_try_wait(&bar[4], 1)) {}
        while (!mbar_try_wait(&bar[4], 1)) {}
        tc_fence_after();
      } else if (GAP == 2) {    // try_wait without the suspend-time hint
        while (!try_wait_nohint(&bar[4], 1)) {}
        while (!try_wait_nohint(&bar[4], 1)) {}
        tc_fence_after();
      } else if (GAP == 3) {    // test_wait (non-blocking)
        while (!test_wait(&bar[4], 1)) {}
        while (!test_wait(&bar[4], 1)) {}
        tc_fence_after();
      } else if (GAP == 4) {
        tc_fence_after();
      } else if (GAP == 5) {    // the two waits issued back to back, one branch
        const bool ok1 = try_wait_nohint(&bar[4], 1), ok2 = try_wait_nohint(&bar[5], 1);
        if (!(ok1 && ok2)) {
          while (!try_wait_nohint(&bar[4], 1)) {}
          while (!try_wait_nohint(&bar[5], 1)) {}
        }
      } else if (GAP == 6) {    // one wait, no fence
        while (!try_wait_nohint(&bar[4], 1)) {}
      }
      const int x0 = SHIFT ? (tile * 128) % kWp : 0;
      const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(x0) * row16;
      const uint32_t d = tmem + static_cast<uint32_t>(acc) * kAccCols;
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t toff = SHIFT ? static_cast<uint32_t>((tap / 3) * kWp + tap % 3) * row16 : 0u;
#pragma unroll
          for (int ks = 0; ks < kSteps; ++ks)
            umma_i8_lohi(d, a_lo + toff + 2u * ks, b_lo0 + tap * b_step + 2u * ks, hi, idesc, (ZACC && tap == 0 && ks == 0) ? 0u : 1u);
        }
        if (COMMIT) {
          umma_commit(&bar[1]);
          umma_commit(&bar[2]);
        }
      }
      __syncwarp();
      if (++acc == NBUF) acc = 0;
    }
    if (elect_one()) umma_commit(&bar[0]);
    __syncwarp();
    const unsigned long long deadline = globaltimer_ns() + 2000000000ull;
    while (!mbar_try_wait(&bar[0], 0)) {
      if (globaltimer_ns() > deadline) {
        *fail = 1;
        break;
      }
    }
    if (lane_id() == 0) {
      cycles[blockIdx.x] = clock64() - t0;
      mbar_arrive(&bar[3]);
    }
  } else if (threadIdx.x >= 128) {      // idle roles parked on a barrier, as the product's waiting warps are
    while (!mbar_try_wait(&bar[3], 0)) {}
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int ROWB, int SHIFT, int COMMIT, int ZACC, int NBUF, int GAP = 0, int WAITERS = 0>
void run_convlike(int tiles, int fill = 0) {
  const int smem = 1024 + ((7 * 58 * ROWB + 1024 + 1023) / 1024) * 1024 + 9 * N * ROWB + 128;
  auto kern = convlike_kernel<N, ROWB, SHIFT, COMMIT, ZACC, NBUF, GAP, WAITERS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d_fail, 0, sizeof(unsigned int));
  for (int rep = 0; rep < 2; ++rep) kern<<<g_sms, 128 + 32 * WAITERS, smem>>>(tiles, fill, d_cycles, d_fail);
  const cudaError_t e = cudaDeviceSynchronize();
  unsigned int failed = 0;
  cudaMemcpy(&failed, d_fail, sizeof(failed), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess || failed) {
    printf("convlike N=%d rowB=%d shift=%d commit=%d zacc=%d nbuf=%d | failed (%s)\n", N, ROWB, SHIFT, COMMIT, ZACC, NBUF, cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    return;
  }
  std::vector<unsigned long long> h(g_sms);
  cudaMemcpy(h.data(), d_cycles, g_sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (auto v : h) sum += static_cast<double>(v);
  const double per_tile = sum / g_sms / tiles;
  printf("fill=%d convlike N=%3d rowB=%3d shift=%d commit=%d zacc=%d nbuf=%d gap=%d waiters=%2d | %8.1f clk/tile %7.1f clk/MMA\n", fill, N, ROWB, SHIFT, COMMIT, ZACC,
         NBUF, GAP, WAITERS, per_tile, per_tile / (9 * (ROWB / 32)));
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 500;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess || prop.major != 10) {
    fprintf(stderr, "needs an sm_100 device\n");
    return 1;
  }
  g_sms = prop.multiProcessorCount;
  cudaMalloc(&d_cycles, g_sms * sizeof(unsigned long long));
  cudaMalloc(&d_fail, sizeof(unsigned int));
  if (argc > 2 && atoi(argv[2]) == 1) {
    const int tiles = 2000;
    for (int fill = 0; fill < 3; ++fill) {
      run_convlike<64, 64, 1, 1, 1, 8, 0, 0>(tiles, fill);
      run_convlike<128, 128, 1, 1, 1, 4, 0, 0>(tiles, fill);
    }
    return 0;
    run_convlike<64, 64, 1, 1, 1, 8, 1, 0>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8, 2, 0>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8, 3, 0>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8, 4, 0>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8, 5, 0>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8, 6, 0>(tiles);
    return 0;
    run_convlike<64, 64, 1, 1, 1, 8, 0, 16>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8, 1, 16>(tiles);
    run_convlike<128, 128, 1, 1, 1, 4, 1, 16>(tiles);
    run_convlike<64, 64, 0, 0, 0, 1>(tiles);
    run_convlike<64, 64, 0, 0, 0, 8>(tiles);
    run_convlike<64, 64, 1, 0, 0, 8>(tiles);
    run_convlike<64, 64, 0, 1, 0, 8>(tiles);
    run_convlike<64, 64, 0, 0, 1, 8>(tiles);
    run_convlike<64, 64, 1, 1, 1, 8>(tiles);
    run_convlike<64, 128, 0, 0, 0, 8>(tiles);
    run_convlike<64, 128, 1, 1, 1, 8>(tiles);
    run_convlike<128, 128, 0, 0, 0, 4>(tiles);
    run_convlike<128, 128, 1, 0, 0, 4>(tiles);
    run_convlike<128, 128, 0, 1, 0, 4>(tiles);
    run_convlike<128, 128, 0, 0, 1, 4>(tiles);
    run_convlike<128, 128, 1, 1, 1, 4>(tiles);
    return 0;
  }
  printf("%-3s %-3s %4s %4s %5s %2s | %8s %9s %7s %8s\n", "knd", "frm", "M", "N", "rowB", "IL", "clk/MMA", "smemB/MMA", "B/clk",
         "MAC/clk");
#define CASES_IL(K, F, M, N, R) run_case<K, F, M, N, R, 1>(iters); run_case<K, F, M, N, R, 2>(iters); run_case<K, F, M, N, R, 4>(iters);
#define CASES_N(K, F, M, R) CASES_IL(K, F, M, 64, R) CASES_IL(K, F, M, 128, R) run_case<K, F, M, 192, R, 1>(iters); run_case<K, F, M, 256, R, 1>(iters);
  CASES_N(0, SS, 128, 64)
  CASES_N(0, SS, 128, 128)
  CASES_N(0, SS, 64, 128)
  CASES_N(0, TS, 128, 128)
  CASES_N(0, TS, 64, 128)
  CASES_N(0, TS, 128, 64)
  CASES_N(0, P2, 256, 64)
  CASES_N(0, P2, 256, 128)
  run_case<0, TS, 128, 96, 128, 2>(iters);
  run_case<0, TS, 128, 112, 128, 2>(iters);
  run_case<0, TS, 128, 160, 128, 2>(iters);
  run_case<0, TS, 128, 224, 128, 1>(iters);
  CASES_N(1, SS, 128, 128)
  CASES_N(1, TS, 128, 128)
  CASES_N(1, P2, 256, 128)
  return 0;
}
