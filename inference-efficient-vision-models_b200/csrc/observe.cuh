// SURVEY 8(f)-4 -- calibration on the GPU.  The reference's PTQ calibration (quantization/engines.py:123-133,
// quantization/main.py:236-239) runs the observer-instrumented float model on the CPU; what an observer of the
// min/max family (MinMaxObserver, MovingAverageMinMaxObserver: quantization/main.py:196-207) keeps of every batch is
// the pair torch.aminmax(x) of each observed tensor.  Here the float forward runs on the FP16 engine with the block's
// residual add left un-fused (so the pre-add conv output exists) and these kernels reduce every observed tensor to
// that pair on the device; the host replays the pairs into the observers of the prepared module (ievm_b200.calibration).
//
// All three kernels are HBM-bound streaming passes: 16-byte loads, warp-shuffle + shared-memory block reduction, one
// atomicMin / atomicMax per block on an order-preserving integer encoding of the float (min / max are associative and
// commutative, so the atomics are deterministic).
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
inline float obs_decode(uint32_t e) {
  const uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  float f;
  memcpy(&f, &b, 4);
  return f;
}
constexpr int kObsMaxRecords = 4096;         // ievm_observe calls between two resets of the log
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

// min / max over the real channels of a [pixels][pitch] matrix (pitch a multiple of the vector width when
// c_real < pitch; a plain contiguous array is pitch == c_real == 1).  NaNs are skipped.  `count` = pixels * pitch.
template <typename T>
__global__ void __launch_bounds__(256)
observe_minmax_kernel(const T* __restrict__ data, long long count, int pitch, int c_real, uint32_t* __restrict__ enc2) {
  constexpr int V = ObsVec<T>::kLanes;
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);   // +inf, -inf
  const long long nvec = count / V;
  const bool all_valid = c_real >= pitch;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(data) + i);
    float f[V];
    ObsVec<T>::unpack(v, f);
    int valid = V;
    if (!all_valid) valid = c_real - static_cast<int>((i * V) % pitch);     // lanes of this vector inside the real channels
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (j < valid && f[j] == f[j]) {
        lo = fminf(lo, f[j]);
        hi = fmaxf(hi, f[j]);
      }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)                                  // tail of a contiguous array
    for (long long e = nvec * V; e < count; ++e) {
      const float x = ObsVec<T>::scalar(data + e);
      if ((all_valid || static_cast<int>(e % pitch) < c_real) && x == x) {
        lo = fminf(lo, x);
        hi = fmaxf(hi, x);
      }
    }
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
    if (lo <= hi) {                                                         // false when the block saw no value
      atomicMin(enc2, obs_encode(lo));
      atomicMax(enc2 + 1, obs_encode(hi));
    }
  }
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
