// TMEM -> register bandwidth (tcgen05.ld 32x32b.x32): how many bytes per clock an SM's epilogue warps can pull out of
// tensor memory.  W warps (W = 4, 8, 16; warp w reads lane quadrant w % 4), each issuing `iters` loads of 32 columns
// x 32 lanes x 4 B = 4 KB, one CTA per SM.  Decides whether the fused front end (64 KB of accumulators per pooled row)
// and the conv epilogues (32 KB per 128 x 64 tile) are bound by the TMEM read port.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../inference-efficient-vision-models_b200/csrc ldtm_rate.cu -o ldtm_rate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"

using namespace ievm;

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(512) ldtm_kernel(int iters, unsigned long long* cycles, unsigned int* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[32], w[32];
    ld32(base + ((it * 64) & 448), v);
    ld32(base + ((it * 64 + 32) & 480), w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) *sink = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess || prop.major != 10) {
    fprintf(stderr, "needs an sm_100 device\n");
    return 1;
  }
  const int sms = prop.multiProcessorCount;
  unsigned long long* d_cycles;
  unsigned int* d_sink;
  cudaMalloc(&d_cycles, sms * sizeof(unsigned long long));
  cudaMalloc(&d_sink, sizeof(unsigned int));
  std::vector<unsigned long long> h(sms);
  printf("%6s | %10s %12s\n", "warps", "clk/load", "B/clk/SM");
  for (int warps : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) ldtm_kernel<<<sms, 32 * warps>>>(iters, d_cycles, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("%6d | failed: %s\n", warps, cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cudaMemcpy(h.data(), d_cycles, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double sum = 0;
    for (auto v : h) sum += static_cast<double>(v);
    const double clk = sum / sms;
    const double bytes = static_cast<double>(warps) * iters * 2 * 4096;
    printf("%6d | %10.1f %12.1f\n", warps, clk / (iters * 2.0), bytes / clk);
  }
  return 0;
}
