// SURVEY 8(f)-4 -- calibration on the GPU.  The reference's PTQ calibration (quantization/engines.py:123-133,
// quantization/main.py:236-239) runs the observer-instrumented float model on the CPU; what an observer of the
// min/max family (MinMaxObserver, MovingAverageMinMaxObserver: quantization/main.py:196-207) keeps of every batch is
// the pair torch.aminmax(x) of each observed tensor.  Here the float forward runs on the FP16 engine with the block's
// residual add left un-fused (so the pre-add conv output exists) and these kernels reduce every observed tensor to
// that pair on the device; the host replays the pairs into the observers of the prepared module (ievm_b200.calibration).
//
// Histogram observers (HistogramObserver: the default fbgemm qconfig of quantization/engines.py:103) additionally keep
// torch.histc(x, 2048, min, max) over the observer's running range; observe_hist_multi_kernel produces those counts.
//
// The kernels are HBM-bound streaming passes over all observed tensors of a forward in ONE launch each: 16-byte loads,
// warp-shuffle + shared-memory block reduction, one atomicMin / atomicMax per block on an order-preserving integer
// encoding of the float, integer atomicAdd for histogram counts (all associative and commutative on integers, so the
// results are deterministic).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

namespace ievm {

// Order-preserving map float -> u32 (negative floats: all bits flipped; others: sign bit set).
__host__ __device__ __forceinline__ uint32_t obs_encode(float f) {
#ifdef __CUDA_ARCH__
  const uint32_t b = __float_as_uint(f);
#else
  uint32_t b;
  memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float obs_decode_dev(uint32_t e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}
inline float obs_decode(uint32_t e) {
  const uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
constexpr int kObsMaxRecords = 4096;         // ievm_observe calls between two clears of the log (observe = 1)
constexpr int kObsHistRecords = 64;          // ... with histograms (observe = 2): 8 KB per point and record
constexpr uint32_t kObsMinInit = 0xffffffffu;   // above every encoded value (decodes to NaN: "nothing observed")
constexpr uint32_t kObsMaxInit = 0u;

// enc[2 * i] = kObsMinInit, enc[2 * i + 1] = kObsMaxInit
__global__ void observe_init_kernel(uint32_t* __restrict__ enc, int pairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * pairs) enc[i] = (i & 1) ? kObsMaxInit : kObsMinInit;
}

template <typename T>
struct ObsVec;
template <>
struct ObsVec<float> {
  static constexpr int kLanes = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ float scalar(const float* p) { return *p; }
};
template <>
struct ObsVec<__half> {
  static constexpr int kLanes = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      f[2 * k] = p.x;
      f[2 * k + 1] = p.y;
    }
  }
  static __device__ __forceinline__ float scalar(const __half* p) { return __half2float(*p); }
};

// One observation point: the real channels of a [pixels][pitch] matrix (pitch a multiple of the vector width when
// c_real < pitch; a plain contiguous array is pitch == c_real == 1).  `count` = pixels * pitch.  All points of a forward
// are reduced by ONE launch: the grid is the concatenation of every point's blocks (first_block = prefix sum), sized so
// that a thread streams a few 16-byte vectors.
constexpr int kObsMaxPoints = 64;
constexpr int kObsBins = 2048;               // HistogramObserver's default (torch.ao.quantization.observer)
struct ObsPoint {
  const void* data;
  long long count;
  int pitch, c_real;
  int is_f32;
  int group;                                 // observer instance the point reports to (prepare_fx shares instances)
  int first_block;
  int reserved;
};
struct ObsTable {
  ObsPoint pt[kObsMaxPoints];
  int n;
  int total_blocks;
};

__device__ __forceinline__ int obs_find_point(const ObsTable& tab, int block, int& nblk) {
  int p = 0;
  while (p + 1 < tab.n && block >= tab.pt[p + 1].first_block) ++p;
  nblk = (p + 1 < tab.n ? tab.pt[p + 1].first_block : tab.total_blocks) - tab.pt[p].first_block;
  return p;
}

// Visit every valid (real-channel, non-NaN) element of the block's share of a point.
template <typename T, typename F>
__device__ __forceinline__ void obs_for_each(const ObsPoint& P, int b, int nblk, F&& f) {
  constexpr int V = ObsVec<T>::kLanes;
  const T* data = static_cast<const T*>(P.data);
  const long long nvec = P.count / V;
  const bool all_valid = P.c_real >= P.pitch;
  const long long stride = static_cast<long long>(nblk) * blockDim.x;
  for (long long i = static_cast<long long>(b) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(data) + i);
    float x[V];
    ObsVec<T>::unpack(v, x);
    int valid = V;
    if (!all_valid) valid = P.c_real - static_cast<int>((i * V) % P.pitch);     // lanes inside the real channels
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (j < valid && x[j] == x[j]) f(x[j]);
  }
  if (b == 0 && threadIdx.x == 0)                                               // tail of a contiguous array
    for (long long e = nvec * V; e < P.count; ++e) {
      const float x = ObsVec<T>::scalar(data + e);
      if ((all_valid || static_cast<int>(e % P.pitch) < P.c_real) && x == x) f(x);
    }
}

// rec[2 p], rec[2 p + 1] = encoded min / max of point p for this batch; run[2 g], run[2 g + 1] = running min / max of
// observer group g over all batches so far (never reset inside a calibration).
__global__ void __launch_bounds__(256)
observe_minmax_multi_kernel(const __grid_constant__ ObsTable tab, uint32_t* __restrict__ rec, uint32_t* __restrict__ run) {
  int nblk;
  const int p = obs_find_point(tab, blockIdx.x, nblk);
  const ObsPoint& P = tab.pt[p];
  const int b = blockIdx.x - P.first_block;
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);       // +inf, -inf
  auto upd = [&](float x) {
    lo = fminf(lo, x);
    hi = fmaxf(hi, x);
  };
  if (P.is_f32) obs_for_each<float>(P, b, nblk, upd);
  else obs_for_each<__half>(P, b, nblk, upd);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  __shared__ float s_lo[8], s_hi[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_lo[warp] = lo;
    s_hi[warp] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) {
      lo = fminf(lo, s_lo[w]);
      hi = fmaxf(hi, s_hi[w]);
    }
    if (lo <= hi) {                                                             // false when the block saw no value
      const uint32_t elo = obs_encode(lo), ehi = obs_encode(hi);
      atomicMin(rec + 2 * p, elo);
      atomicMax(rec + 2 * p + 1, ehi);
      atomicMin(run + 2 * P.group, elo);
      atomicMax(run + 2 * P.group + 1, ehi);
    }
  }
}

// torch.histc(x, 2048, min = lo, max = hi) of every point over its observer's running range [lo, hi] (what
// HistogramObserver.forward computes per batch), as exact integer counts.  The bin of an element restates ATen's CPU
// kernel (aten/src/ATen/native/cpu/HistogramKernel.cpp, linear interpolation): float32, in this order,
//   pos = int64(((x - lo) * bins) / (hi - lo)),  pos == bins -> bins - 1;   lo == hi -> lo -= 1, hi += 1
// (pinned against torch.histc by tests/test_calibration.py).  Shared-memory histogram per block; bin 0 -- where every
// zero of a post-ReLU tensor lands -- is counted in a register.
__global__ void __launch_bounds__(256)
observe_hist_multi_kernel(const __grid_constant__ ObsTable tab, const uint32_t* __restrict__ run, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kObsBins];
  int nblk;
  const int p = obs_find_point(tab, blockIdx.x, nblk);
  const ObsPoint& P = tab.pt[p];
  const int b = blockIdx.x - P.first_block;
  const uint32_t elo = run[2 * P.group], ehi = run[2 * P.group + 1];
  if (elo > ehi) return;                                                        // nothing observed (whole block exits)
  for (int i = threadIdx.x; i < kObsBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  float lo = obs_decode_dev(elo), hi = obs_decode_dev(ehi);
  if (lo == hi) {
    lo = __fadd_rn(lo, -1.0f);
    hi = __fadd_rn(hi, 1.0f);
  }
  const float range = __fsub_rn(hi, lo);
  uint32_t c0 = 0;
  auto upd = [&](float x) {
    int pos = static_cast<int>(__fdiv_rn(__fmul_rn(__fsub_rn(x, lo), static_cast<float>(kObsBins)), range));
    pos = min(pos, kObsBins - 1);
    if (pos <= 0) ++c0;
    else atomicAdd(&sh[pos], 1u);
  };
  if (P.is_f32) obs_for_each<float>(P, b, nblk, upd);
  else obs_for_each<__half>(P, b, nblk, upd);
  if (c0) atomicAdd(&sh[0], c0);
  __syncthreads();
  uint32_t* out = hist + static_cast<size_t>(p) * kObsBins;
  for (int i = threadIdx.x; i < kObsBins; i += blockDim.x)
    if (sh[i]) atomicAdd(out + i, sh[i]);
}

// out = relu(a + b) on f16 tensors of identical layout (the BasicBlock's `out += identity; out = relu(out)` left
// un-fused for calibration): fp32 add, one rounding.  nvec = elements / 8.
__global__ void __launch_bounds__(256)
add_relu_f16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, long long nvec) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const uint4 va = __ldg(a + i), vb = __ldg(b + i);
  const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
  uint32_t r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&wa[k]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&wb[k]));
    const __half2 h = __floats2half2_rn(fmaxf(__fadd_rn(fa.x, fb.x), 0.f), fmaxf(__fadd_rn(fa.y, fb.y), 0.f));
    r[k] = *reinterpret_cast<const uint32_t*>(&h);
  }
  out[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

}  // namespace ievm
