// CUDA-core kernels around the tensor-core convolutions: input quantisation, the 3-channel stem,
// max-pool, the avgpool+fc+dequantize head, and a plain direct convolution that the parity tests use
// to cross-check the tcgen05 kernel on the GPU at sizes the CPU oracle cannot reach.
// All INT8 arithmetic follows the formulas in oracle/int8_forward.py (float32, RNE, no contraction).
#pragma once
#include <cuda_fp16.h>

#include "conv_tc.cuh"

namespace ievm {

// --------------------------------------------------------------------------------------------
// quantize_per_tensor (graph node 1): f32 NCHW [n,3,h,w] -> u8 NHWC4 (4th byte = zp) stored with a
// constant border of zero-point pixels, [n][h + 6][w + 8][4] with the image at (row 3, column 4):
// the 7x7/2 stem then reads every window without a single bounds check (the border is written once
// at engine creation and never touched again).
// One thread = 4 consecutive pixels of a row: three coalesced float4 loads, one 16-byte store.
// --------------------------------------------------------------------------------------------
constexpr int kInPadTop = 3;
constexpr int kInPadLeft = 4;
constexpr int kInPadH = 6;     // rows added (3 above, 3 below)
constexpr int kInPadW = 8;     // columns added (4 left, 4 right)

__device__ __forceinline__ uint32_t quant_u8(float x, float inv_scale, int zp) {
  const int q = __float2int_rn(__fmul_rn(x, inv_scale)) + zp;
  return static_cast<uint32_t>(min(max(q, 0), 255));
}

__global__ void __launch_bounds__(256)
quantize_nchw3_to_nhwc4_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, long long n_quads, int h, int w,
                               float inv_scale, int zp) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const int qpr = w >> 2;                         // quads per row
  const long long rowid = i / qpr;                // img * h + y
  const int xq = static_cast<int>(i - rowid * qpr);
  const long long img = rowid / h;
  const int y = static_cast<int>(rowid - img * h);
  const long long plane = static_cast<long long>(h) * w;
  const float* base = x + img * 3 * plane + static_cast<long long>(y) * w + 4 * xq;
  const float4 r = __ldg(reinterpret_cast<const float4*>(base));
  const float4 g = __ldg(reinterpret_cast<const float4*>(base + plane));
  const float4 b = __ldg(reinterpret_cast<const float4*>(base + 2 * plane));
  const uint32_t z = static_cast<uint32_t>(zp) << 24;
  uint4 o;
  o.x = quant_u8(r.x, inv_scale, zp) | (quant_u8(g.x, inv_scale, zp) << 8) | (quant_u8(b.x, inv_scale, zp) << 16) | z;
  o.y = quant_u8(r.y, inv_scale, zp) | (quant_u8(g.y, inv_scale, zp) << 8) | (quant_u8(b.y, inv_scale, zp) << 16) | z;
  o.z = quant_u8(r.z, inv_scale, zp) | (quant_u8(g.z, inv_scale, zp) << 8) | (quant_u8(b.z, inv_scale, zp) << 16) | z;
  o.w = quant_u8(r.w, inv_scale, zp) | (quant_u8(g.w, inv_scale, zp) << 8) | (quant_u8(b.w, inv_scale, zp) << 16) | z;
  const long long orow = (img * (h + kInPadH) + y + kInPadTop) * (w + kInPadW) + kInPadLeft + 4 * xq;
  *reinterpret_cast<uint4*>(out + orow * 4) = o;
}

__device__ __forceinline__ int dp4a_u8s8(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
  return d;
}

// --------------------------------------------------------------------------------------------
// Stem: 7x7 stride-2 pad-3 conv over u8 NHWC4 + requant + ReLU -> u8 NHWC(cpad).
// One thread = one output pixel; the 49 input pixels sit in registers, weights are broadcast from
// shared memory as [tap][cout] s8x4 words.  sum (xq - zp) w = sum xq w - zp * sum w because
// out-of-image taps read the zero-point border of the padded input tensor.
// --------------------------------------------------------------------------------------------
struct StemParams {
  int n, h, w, ho, wo;
  int cpad;                  // padded cout (multiple of 16)
  int in_zp;
  const uint32_t* w4;        // [49][cpad] s8x4 (c0,c1,c2,0)
  const int* wsum;           // [cpad]
  const float* bdiv;
  const float* mult;
  int out_zp, out_lo;
};

__global__ void __launch_bounds__(128)
stem_conv7x7_simt_kernel(const uint8_t* __restrict__ xq, uint8_t* __restrict__ out, const StemParams p) {
  extern __shared__ uint32_t s_w[];    // [49][cpad]
  for (int i = threadIdx.x; i < 49 * p.cpad; i += blockDim.x) s_w[i] = p.w4[i];
  __syncthreads();
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long m_total = static_cast<long long>(p.n) * p.ho * p.wo;
  if (m >= m_total) return;
  const int hw = p.ho * p.wo;
  const int img = static_cast<int>(m / hw);
  const int rem = static_cast<int>(m - static_cast<long long>(img) * hw);
  const int oy = rem / p.wo, ox = rem - (rem / p.wo) * p.wo;
  const int wp = p.w + kInPadW;
  const uint32_t* in32 = reinterpret_cast<const uint32_t*>(xq) + static_cast<long long>(img) * (p.h + kInPadH) * wp;
  uint32_t px[49];
#pragma unroll
  for (int ky = 0; ky < 7; ++ky) {
    const int iy = oy * 2 + ky;                  // padded row of input row 2*oy - 3 + ky
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) px[ky * 7 + kx] = __ldg(in32 + iy * wp + ox * 2 + kx + 1);
  }
  uint8_t* orow = out + m * p.cpad;
  for (int c0 = 0; c0 < p.cpad; c0 += 16) {
    int acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0;
#pragma unroll
    for (int t = 0; t < 49; ++t) {
      const uint4* wr = reinterpret_cast<const uint4*>(s_w + t * p.cpad + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 w = wr[j];
        acc[4 * j] = dp4a_u8s8(px[t], w.x, acc[4 * j]);
        acc[4 * j + 1] = dp4a_u8s8(px[t], w.y, acc[4 * j + 1]);
        acc[4 * j + 2] = dp4a_u8s8(px[t], w.z, acc[4 * j + 2]);
        acc[4 * j + 3] = dp4a_u8s8(px[t], w.w, acc[4 * j + 3]);
      }
    }
    uint32_t packed[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t wd = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = c0 + 4 * j + b;
        const int a = acc[4 * j + b] - p.in_zp * __ldg(p.wsum + c);
        wd |= static_cast<uint32_t>(requant_i8(a, __ldg(p.bdiv + c), __ldg(p.mult + c), p.out_zp, p.out_lo)) << (8 * b);
      }
      packed[j] = wd;
    }
    *reinterpret_cast<uint4*>(orow + c0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// --------------------------------------------------------------------------------------------
// MaxPool2d(3, 2, 1) on u8 NHWC (graph node `maxpool`): padding ignored, qparams pass through.
// One thread = one output pixel x 16 channels.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool3x3s2_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int h, int w, int ho,
                       int wo, int cpad) {
  const int groups = cpad / 16;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * ho * wo * groups;
  if (i >= total) return;
  const int g = static_cast<int>(i % groups);
  const long long pix = i / groups;
  const int ox = static_cast<int>(pix % wo);
  const int oy = static_cast<int>((pix / wo) % ho);
  const int img = static_cast<int>(pix / (static_cast<long long>(wo) * ho));
  uint4 best = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(
          in + ((static_cast<long long>(img) * h + iy) * w + ix) * cpad + g * 16));
      best.x = __vmaxu4(best.x, v.x);
      best.y = __vmaxu4(best.y, v.y);
      best.z = __vmaxu4(best.z, v.z);
      best.w = __vmaxu4(best.w, v.w);
    }
  }
  *reinterpret_cast<uint4*>(out + pix * cpad + g * 16) = best;
}

// FP16 variant: one thread = one output pixel x 8 channels.
__global__ void __launch_bounds__(256)
maxpool3x3s2_f16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int h, int w, int ho,
                        int wo, int cpad) {
  const int groups = cpad / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * ho * wo * groups;
  if (i >= total) return;
  const int g = static_cast<int>(i % groups);
  const long long pix = i / groups;
  const int ox = static_cast<int>(pix % wo);
  const int oy = static_cast<int>((pix / wo) % ho);
  const int img = static_cast<int>(pix / (static_cast<long long>(wo) * ho));
  __half2 best[4];
  bool first = true;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(
          in + ((static_cast<long long>(img) * h + iy) * w + ix) * cpad + g * 8));
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) best[j] = first ? hv[j] : __hmax2(best[j], hv[j]);
      first = false;
    }
  }
  *reinterpret_cast<uint4*>(out + pix * cpad + g * 8) = *reinterpret_cast<const uint4*>(best);
}

// --------------------------------------------------------------------------------------------
// Head: AdaptiveAvgPool2d(1) + flatten + quantized Linear + dequantize, one CTA per image.
//   pooled = clamp(rne(float(sum) / count), 0, 255)  (same qparams as the input, zp == in_zp)
//   acc[o] = sum_c (pooled[c] - in_zp) * w[o][c];  q = requant(acc);  logit = (q - fc_zp) * fc_scale
// The kernel is a latency chain, not a bandwidth problem (23 KB per image): inside the graph step it cost 15 us at batch
// 256 and 9 us at batch 1 when every thread walked its pixels one dependent load at a time and fetched the fc weights
// from L2 inside the channel loop.  Now the fc weights are copied to shared memory BEFORE griddepcontrol.wait (they do
// not depend on the previous kernel, so the copy hides behind its tail), every thread issues all of its pixel loads
// before the first add, and the channel loop reads weights from shared memory.
// --------------------------------------------------------------------------------------------
struct HeadParams {
  int hw;            // pixels per image (7*7)
  int c;             // real channels
  int cpad;          // channel pitch
  int classes;
  int in_zp;
  const int8_t* w;   // [classes][cpad]
  const float* bdiv;
  const float* mult;
  int fc_zp;
  float fc_scale;
  int w_smem;        // 1: the launch carries classes * cpad bytes of dynamic shared memory for the weights
};

constexpr int kHeadThreads = 256;
constexpr int kMaxClasses = 16;
constexpr int kHeadLoadsInFlight = 8;         // 16-byte loads a thread issues before it starts adding
constexpr int kHeadPackMax = 256;             // pixels a 16-bit lane of the INT8 head's packed sums can take (256 * 255 < 2^16)
constexpr int kHeadWeightSmemMax = 30 * 1024; // dynamic shared memory the heads may use for the fc weights

constexpr int kHeadSumWords = 4096;          // shared partial channel sums: pixel phases x channel pitch

// Phase 1 (both heads): per-channel sums over the image's pixels with 16-byte loads.  Thread (ph, g) walks pixels
// ph, ph + PH, ... of 16-byte channel group g; the PH partial sums per channel are combined in a fixed order in
// phase 2, so the result does not depend on scheduling (and the integer sums are exact anyway).
__device__ __forceinline__ void head_phases(int groups, int& g_stride, int& ph_count) {
  ph_count = groups >= kHeadThreads ? 1 : kHeadThreads / groups;
  g_stride = groups >= kHeadThreads ? kHeadThreads : groups;
}

// fc weights -> shared memory (16-byte copies; `bytes` is a multiple of 16 because the channel pitch is)
__device__ __forceinline__ void head_stage_weights(const void* w, void* smem, int bytes) {
  const uint4* src = static_cast<const uint4*>(w);
  uint4* dst = static_cast<uint4*>(smem);
  for (int i = threadIdx.x; i < bytes / 16; i += kHeadThreads) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(kHeadThreads)
head_i8_kernel(const uint8_t* __restrict__ in, float* __restrict__ logits, uint8_t* __restrict__ pooled_dbg,
               const HeadParams p) {
  extern __shared__ __align__(16) uint8_t head_w_smem[];
  __shared__ int s_sum[kHeadSumWords];
  __shared__ int s_part[kHeadThreads / 32][kMaxClasses];
  griddep_launch_dependents();
  const int8_t* w = p.w;
  if (p.w_smem) {
    head_stage_weights(p.w, head_w_smem, p.classes * p.cpad);
    w = reinterpret_cast<const int8_t*>(head_w_smem);
  }
  griddep_wait();
  const int img = blockIdx.x;
  const uint8_t* base = in + static_cast<long long>(img) * p.hw * p.cpad;
  const int groups = p.cpad >> 4;
  int g_stride, ph_count;
  head_phases(groups, g_stride, ph_count);
  if (ph_count * p.cpad > kHeadSumWords) ph_count = kHeadSumWords / p.cpad;      // very wide heads: fewer phases
  const int ph = threadIdx.x / g_stride, g0 = threadIdx.x - ph * g_stride;
  if (ph < ph_count) {
    for (int g = g0; g < groups; g += g_stride) {
      // two channels per 32-bit accumulator (bytes 0 / 2 and bytes 1 / 3 of every word as 16-bit lanes: one AND or one
      // byte permute plus one add per two channels); a lane holds at most kHeadPackMax pixels x 255 < 2^16 before it is
      // flushed into the 32-bit sums
      int sum[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) sum[j] = 0;
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = 0u;
      int packed = 0;
      auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sum[4 * k] += pk[2 * k] & 0xffffu;
          sum[4 * k + 2] += pk[2 * k] >> 16;
          sum[4 * k + 1] += pk[2 * k + 1] & 0xffffu;
          sum[4 * k + 3] += pk[2 * k + 1] >> 16;
          pk[2 * k] = pk[2 * k + 1] = 0u;
        }
        packed = 0;
      };
      for (int px0 = ph; px0 < p.hw; px0 += kHeadLoadsInFlight * ph_count) {
        uint4 v[kHeadLoadsInFlight];
#pragma unroll
        for (int it = 0; it < kHeadLoadsInFlight; ++it) {
          const int px = px0 + it * ph_count;
          v[it] = px < p.hw ? __ldg(reinterpret_cast<const uint4*>(base + px * p.cpad + g * 16)) : make_uint4(0u, 0u, 0u, 0u);
        }
        if (packed + kHeadLoadsInFlight > kHeadPackMax) flush();
#pragma unroll
        for (int it = 0; it < kHeadLoadsInFlight; ++it) {
          const uint32_t wv[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            pk[2 * k] += wv[k] & 0x00ff00ffu;                          // bytes 0 and 2
            pk[2 * k + 1] += __byte_perm(wv[k], 0u, 0x4341);           // bytes 1 and 3 (selector 4 = a zero byte)
          }
        }
        packed += kHeadLoadsInFlight;
      }
      flush();
#pragma unroll
      for (int j = 0; j < 16; ++j) s_sum[ph * p.cpad + g * 16 + j] = sum[j];
    }
  }
  __syncthreads();                      // partial sums and the staged weights
  int acc[kMaxClasses];
#pragma unroll
  for (int o = 0; o < kMaxClasses; ++o) acc[o] = 0;
  const float cnt = static_cast<float>(p.hw);
  for (int c = threadIdx.x; c < p.c; c += kHeadThreads) {
    int sc = 0;
    for (int k = 0; k < ph_count; ++k) sc += s_sum[k * p.cpad + c];
    int q = __float2int_rn(__fdiv_rn(__int2float_rn(sc), cnt));
    q = min(max(q, 0), 255);
    if (pooled_dbg) pooled_dbg[static_cast<long long>(img) * p.c + c] = static_cast<uint8_t>(q);
    const int xv = q - p.in_zp;
#pragma unroll
    for (int o = 0; o < kMaxClasses; ++o)
      if (o < p.classes) acc[o] += xv * static_cast<int>(w[o * p.cpad + c]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < kMaxClasses; ++o) {
    if (o < p.classes) {
      int v = acc[o];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) s_part[warp][o] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < p.classes) {
    const int o = threadIdx.x;
    int a = 0;
    for (int wv = 0; wv < kHeadThreads / 32; ++wv) a += s_part[wv][o];
    const int q = requant_i8(a, p.bdiv[o], p.mult[o], p.fc_zp, 0);
    logits[static_cast<long long>(img) * p.classes + o] = __fmul_rn(__int2float_rn(q - p.fc_zp), p.fc_scale);
  }
}

// FP16 head: avgpool (fp32 accumulate, rounded to f16 like the reference's pooled tensor) + fc.
struct HeadF16Params {
  int hw, c, cpad, classes;
  const __half* w;   // [classes][cpad]
  const float* bias;
  __half* pooled = nullptr;   // calibration only (observe.cuh): the avgpool output, [n][c]
  int w_smem = 0;             // as HeadParams::w_smem (2 * classes * cpad bytes)
};

__global__ void __launch_bounds__(kHeadThreads)
head_f16_kernel(const __half* __restrict__ in, __half* __restrict__ logits, const HeadF16Params p) {
  extern __shared__ __align__(16) uint8_t head_w_smem[];
  __shared__ float s_sum[kHeadSumWords];
  __shared__ float s_part[kHeadThreads / 32][kMaxClasses];
  griddep_launch_dependents();
  const __half* w = p.w;
  if (p.w_smem) {
    head_stage_weights(p.w, head_w_smem, 2 * p.classes * p.cpad);
    w = reinterpret_cast<const __half*>(head_w_smem);
  }
  griddep_wait();
  const int img = blockIdx.x;
  const __half* base = in + static_cast<long long>(img) * p.hw * p.cpad;
  const int groups = p.cpad >> 3;                  // 8 halves = 16 bytes
  int g_stride, ph_count;
  head_phases(groups, g_stride, ph_count);
  if (ph_count * p.cpad > kHeadSumWords) ph_count = kHeadSumWords / p.cpad;
  const int ph = threadIdx.x / g_stride, g0 = threadIdx.x - ph * g_stride;
  if (ph < ph_count) {
    for (int g = g0; g < groups; g += g_stride) {
      float sum[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] = 0.f;
      // pixels are added in the fixed order ph, ph + PH, ...: the result does not depend on the batch or the schedule
      for (int px0 = ph; px0 < p.hw; px0 += kHeadLoadsInFlight * ph_count) {
        uint4 v[kHeadLoadsInFlight];
#pragma unroll
        for (int it = 0; it < kHeadLoadsInFlight; ++it) {
          const int px = px0 + it * ph_count;
          v[it] = px < p.hw ? __ldg(reinterpret_cast<const uint4*>(base + px * p.cpad + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int it = 0; it < kHeadLoadsInFlight; ++it) {
          const uint32_t wv[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wv[k]));
            sum[2 * k] += f.x;           // a pixel past the image adds +0.0: the sum is unchanged
            sum[2 * k + 1] += f.y;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_sum[ph * p.cpad + g * 8 + j] = sum[j];
    }
  }
  __syncthreads();
  float acc[kMaxClasses];
#pragma unroll
  for (int o = 0; o < kMaxClasses; ++o) acc[o] = 0.f;
  const float inv = 1.0f / static_cast<float>(p.hw);
  for (int c = threadIdx.x; c < p.c; c += kHeadThreads) {
    float sc = 0.f;
    for (int k = 0; k < ph_count; ++k) sc += s_sum[k * p.cpad + c];
    const __half mh = __float2half_rn(sc * inv);
    if (p.pooled) p.pooled[static_cast<long long>(img) * p.c + c] = mh;
    const float m = __half2float(mh);
#pragma unroll
    for (int o = 0; o < kMaxClasses; ++o)
      if (o < p.classes) acc[o] += m * __half2float(w[o * p.cpad + c]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < kMaxClasses; ++o) {
    if (o < p.classes) {
      float v = acc[o];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) s_part[warp][o] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < p.classes) {
    const int o = threadIdx.x;
    float a = p.bias[o];
    for (int wv = 0; wv < kHeadThreads / 32; ++wv) a += s_part[wv][o];
    logits[static_cast<long long>(img) * p.classes + o] = __float2half_rn(a);
  }
}

// --------------------------------------------------------------------------------------------
// Direct convolution over the same packed operands as the tensor-core kernel (one thread = one
// output pixel x one output channel).  Test / cross-check path only: never used by forward().
// --------------------------------------------------------------------------------------------
struct ConvDirectParams {
  int n, h, w, ho, wo;
  int cin_pitch;       // input elements per pixel
  int cin_w;           // packed weight channels per tap (kchunks * kc_elems)
  int cin_real;
  int cout_pad;
  int ksize, stride, pad;
  int in_zp;           // zero point of the input tensor (0 for every post-ReLU tensor)
};

__global__ void __launch_bounds__(128)
conv_direct_i8_kernel(const uint8_t* __restrict__ in, const int8_t* __restrict__ wp, uint8_t* __restrict__ out,
                      const ConvDirectParams g, const ConvTcParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.m_total) * g.cout_pad;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % g.cout_pad);
  const long long m = idx / g.cout_pad;
  const int hw = g.ho * g.wo;
  const int img = static_cast<int>(m / hw);
  const int rem = static_cast<int>(m - static_cast<long long>(img) * hw);
  const int oy = rem / g.wo, ox = rem - (rem / g.wo) * g.wo;
  int acc = 0;
  const int taps = g.ksize * g.ksize;
  for (int ky = 0; ky < g.ksize; ++ky) {
    const int iy = oy * g.stride - g.pad + ky;
    if (iy < 0 || iy >= g.h) continue;
    for (int kx = 0; kx < g.ksize; ++kx) {
      const int ix = ox * g.stride - g.pad + kx;
      if (ix < 0 || ix >= g.w) continue;
      const uint8_t* ip = in + ((static_cast<long long>(img) * g.h + iy) * g.w + ix) * g.cin_pitch;
      const int8_t* wr = wp + (static_cast<long long>(co) * taps + ky * g.ksize + kx) * g.cin_w;
      for (int ci = 0; ci < g.cin_real; ++ci) acc += (static_cast<int>(ip[ci]) - g.in_zp) * static_cast<int>(wr[ci]);
    }
  }
  if (p.dump_acc) p.dump_acc[m * p.dump_pitch + co] = acc;
  int q = requant_i8(acc, p.ep0[co], p.ep1[co], p.out_zp, p.out_lo);
  if (p.res != nullptr) {
    const int r = static_cast<const uint8_t*>(p.res)[m * p.res_pitch + co];
    q = add_relu_i8(q, p.out_zp, p.a_scale, r, p.res_zp, p.res_scale, p.inv_add_scale, p.add_zp);
  }
  static_cast<uint8_t*>(p.out)[m * p.out_pitch + co] = static_cast<uint8_t>(q);
}

__global__ void __launch_bounds__(128)
conv_direct_f16_kernel(const __half* __restrict__ in, const __half* __restrict__ wp, const ConvDirectParams g,
                       const ConvTcParams p) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.m_total) * g.cout_pad;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % g.cout_pad);
  const long long m = idx / g.cout_pad;
  const int hw = g.ho * g.wo;
  const int img = static_cast<int>(m / hw);
  const int rem = static_cast<int>(m - static_cast<long long>(img) * hw);
  const int oy = rem / g.wo, ox = rem - (rem / g.wo) * g.wo;
  float acc = 0.f;
  const int taps = g.ksize * g.ksize;
  for (int ky = 0; ky < g.ksize; ++ky) {
    const int iy = oy * g.stride - g.pad + ky;
    if (iy < 0 || iy >= g.h) continue;
    for (int kx = 0; kx < g.ksize; ++kx) {
      const int ix = ox * g.stride - g.pad + kx;
      if (ix < 0 || ix >= g.w) continue;
      const __half* ip = in + ((static_cast<long long>(img) * g.h + iy) * g.w + ix) * g.cin_pitch;
      const __half* wr = wp + (static_cast<long long>(co) * taps + ky * g.ksize + kx) * g.cin_w;
      for (int ci = 0; ci < g.cin_real; ++ci) acc += __half2float(ip[ci]) * __half2float(wr[ci]);
    }
  }
  float f = acc + p.ep0[co];
  if (p.res != nullptr) f += __half2float(static_cast<const __half*>(p.res)[m * p.res_pitch + co]);
  if (p.relu) f = fmaxf(f, 0.f);
  static_cast<__half*>(p.out)[m * p.out_pitch + co] = __float2half_rn(f);
}

// --------------------------------------------------------------------------------------------
// FP16 stem: 7x7/2 conv over f16 NCHW input (3 channels) + folded-BN bias + ReLU -> f16 NHWC(cpad).
// One thread = one output pixel x 8 output channels; weights [tap*3 + c][cpad] f16 in shared memory.
// --------------------------------------------------------------------------------------------
struct StemF16Params {
  int n, h, w, ho, wo, cpad;
  const __half* wt;      // [147][cpad]
  const float* bias;     // [cpad]
};

__global__ void __launch_bounds__(128)
stem_conv7x7_f16_kernel(const __half* __restrict__ x, __half* __restrict__ out, const StemF16Params p) {
  extern __shared__ uint32_t s_raw[];
  __half* s_w = reinterpret_cast<__half*>(s_raw);
  for (int i = threadIdx.x; i < 147 * p.cpad; i += blockDim.x) s_w[i] = p.wt[i];
  __syncthreads();
  const int groups = p.cpad / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(p.n) * p.ho * p.wo * groups;
  if (idx >= total) return;
  // pixel-major thread order inside a warp keeps the NCHW reads coalesced
  const long long m = idx % (static_cast<long long>(p.n) * p.ho * p.wo);
  const int g = static_cast<int>(idx / (static_cast<long long>(p.n) * p.ho * p.wo));
  const int hw = p.ho * p.wo;
  const int img = static_cast<int>(m / hw);
  const int rem = static_cast<int>(m - static_cast<long long>(img) * hw);
  const int oy = rem / p.wo, ox = rem - (rem / p.wo) * p.wo;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const long long plane = static_cast<long long>(p.h) * p.w;
  const __half* xin = x + static_cast<long long>(img) * 3 * plane;
  for (int ky = 0; ky < 7; ++ky) {
    const int iy = oy * 2 - 3 + ky;
    if (iy < 0 || iy >= p.h) continue;
    for (int kx = 0; kx < 7; ++kx) {
      const int ix = ox * 2 - 3 + kx;
      if (ix < 0 || ix >= p.w) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float xv = __half2float(__ldg(xin + c * plane + static_cast<long long>(iy) * p.w + ix));
        const uint4 wv = *reinterpret_cast<const uint4*>(s_w + ((ky * 7 + kx) * 3 + c) * p.cpad + g * 8);
        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 w2 = __half22float2(wh[j]);
          acc[2 * j] = fmaf(xv, w2.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(xv, w2.y, acc[2 * j + 1]);
        }
      }
    }
  }
  __half2 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = fmaxf(acc[2 * j] + p.bias[g * 8 + 2 * j], 0.f);
    const float b = fmaxf(acc[2 * j + 1] + p.bias[g * 8 + 2 * j + 1], 0.f);
    o[j] = __floats2half2_rn(a, b);
  }
  *reinterpret_cast<uint4*>(out + m * p.cpad + g * 8) = *reinterpret_cast<const uint4*>(o);
}

// --------------------------------------------------------------------------------------------
// KD evaluation loss (knowledge_distillation/train.py:47-57): per-row CE and T^2 * KL, fp32.
// out3 += {sum_i CE_i, sum_i KL_i * T^2, correct_count}; the caller divides by the batch size.
// --------------------------------------------------------------------------------------------
__global__ void kd_loss_kernel(const float* __restrict__ s, const float* __restrict__ t,
                               const long long* __restrict__ y, int n, int classes, float temp, float* out3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float ce = 0.f, kl = 0.f, correct = 0.f;
  if (i < n) {
    const float* sr = s + static_cast<long long>(i) * classes;
    const float* tr = t + static_cast<long long>(i) * classes;
    float ms = -INFINITY, mst = -INFINITY, mtt = -INFINITY;
    int arg = 0;
    for (int c = 0; c < classes; ++c) {
      if (sr[c] > ms) { ms = sr[c]; arg = c; }
      mst = fmaxf(mst, sr[c] / temp);
      mtt = fmaxf(mtt, tr[c] / temp);
    }
    float zs = 0.f, zst = 0.f, ztt = 0.f;
    for (int c = 0; c < classes; ++c) {
      zs += expf(sr[c] - ms);
      zst += expf(sr[c] / temp - mst);
      ztt += expf(tr[c] / temp - mtt);
    }
    const float lzs = logf(zs), lzst = logf(zst), lztt = logf(ztt);
    const long long label64 = y[i];
    const bool label_ok = label64 >= 0 && label64 < classes;       // an out-of-range label poisons the loss (NaN) instead
    const int label = label_ok ? static_cast<int>(label64) : 0;    // of reading outside the row
    ce = label_ok ? -(sr[label] - ms - lzs) : __int_as_float(0x7fc00000);
    for (int c = 0; c < classes; ++c) {
      const float logp_t = tr[c] / temp - mtt - lztt;
      const float logp_s = sr[c] / temp - mst - lzst;
      kl += expf(logp_t) * (logp_t - logp_s);
    }
    kl *= temp * temp;
    correct = (label_ok && arg == label) ? 1.f : 0.f;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ce += __shfl_xor_sync(0xffffffffu, ce, d);
    kl += __shfl_xor_sync(0xffffffffu, kl, d);
    correct += __shfl_xor_sync(0xffffffffu, correct, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out3 + 0, ce);
    atomicAdd(out3 + 1, kl);
    atomicAdd(out3 + 2, correct);
  }
}

__global__ void kd_finalize_kernel(float* out3, float inv_n) {
  if (threadIdx.x < 2) out3[threadIdx.x] *= inv_n;     // CE and KL are batch means; [2] stays a count
}

// --------------------------------------------------------------------------------------------
// evaluate_accuracy (quantization/engines.py:59-63) without a host round trip per batch:
// counters[0] += #(argmax(logits) == label), counters[1] += n.  torch.max returns the lowest index on ties.
// --------------------------------------------------------------------------------------------
template <typename T>
__global__ void count_correct_kernel(const T* __restrict__ logits, const long long* __restrict__ labels, int n, int classes,
                                     unsigned long long* counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int hit = 0;
  if (i < n) {
    const T* r = logits + static_cast<long long>(i) * classes;
    float best = static_cast<float>(r[0]);
    int arg = 0;
    for (int c = 1; c < classes; ++c) {
      const float v = static_cast<float>(r[c]);
      if (v > best) {
        best = v;
        arg = c;
      }
    }
    hit = static_cast<long long>(arg) == labels[i] ? 1u : 0u;      // an out-of-range label never matches
  }
  const unsigned int warp_hits = __popc(__ballot_sync(0xffffffffu, hit != 0));
  if ((threadIdx.x & 31) == 0 && warp_hits) atomicAdd(counters, static_cast<unsigned long long>(warp_hits));
  if (i == 0) atomicAdd(counters + 1, static_cast<unsigned long long>(n));
}

// --------------------------------------------------------------------------------------------
// T.Resize((224, 224)) of the reference's dataset transform (quantization/dataset.py:15) on decoded 8-bit images:
// Pillow's ImagingResample (bilinear), bit for bit -- a horizontal then a vertical pass, each a weighted sum with
// 22-bit fixed-point coefficients (precomputed on the host exactly as Resample.c:precompute_coeffs does), a rounding
// bias of half an LSB and a clip to [0, 255]; the intermediate image is rounded to 8 bits between the passes.
// bounds[o] = (first input index, taps), kk[o][ksize] = coefficients.  One thread per output pixel (3 channels).
// --------------------------------------------------------------------------------------------
constexpr int kResizePrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t resize_clip8(int acc) { return static_cast<uint8_t>(min(max(acc >> kResizePrecisionBits, 0), 255)); }

// in [rows][in_w][3] -> out [rows][out_w][3]
__global__ void __launch_bounds__(256)
resize_rows_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long rows, int in_w, int out_w,
                      const int2* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * out_w) return;
  const long long r = i / out_w;
  const int xx = static_cast<int>(i - r * out_w);
  const int2 b = __ldg(bounds + xx);
  const uint8_t* src = in + (r * in_w + b.x) * 3;
  int a0 = 1 << (kResizePrecisionBits - 1), a1 = a0, a2 = a0;
  for (int x = 0; x < b.y; ++x) {
    const int k = __ldg(kk + xx * ksize + x);
    a0 += static_cast<int>(src[3 * x]) * k;
    a1 += static_cast<int>(src[3 * x + 1]) * k;
    a2 += static_cast<int>(src[3 * x + 2]) * k;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = resize_clip8(a0);
  dst[1] = resize_clip8(a1);
  dst[2] = resize_clip8(a2);
}

// in [n][in_h][w][3] -> out [n][out_h][w][3]
__global__ void __launch_bounds__(256)
resize_cols_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int in_h, int out_h, int w,
                      const int2* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_img = static_cast<long long>(out_h) * w;
  if (i >= per_img * n) return;
  const long long img = i / per_img;
  const int rem = static_cast<int>(i - img * per_img);
  const int yy = rem / w, x = rem - yy * w;
  const int2 b = __ldg(bounds + yy);
  const uint8_t* src = in + ((img * in_h + b.x) * w + x) * 3;
  const long long row_stride = static_cast<long long>(w) * 3;
  int a0 = 1 << (kResizePrecisionBits - 1), a1 = a0, a2 = a0;
  for (int y = 0; y < b.y; ++y) {
    const int k = __ldg(kk + yy * ksize + y);
    a0 += static_cast<int>(src[y * row_stride]) * k;
    a1 += static_cast<int>(src[y * row_stride + 1]) * k;
    a2 += static_cast<int>(src[y * row_stride + 2]) * k;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = resize_clip8(a0);
  dst[1] = resize_clip8(a1);
  dst[2] = resize_clip8(a2);
}

}  // namespace ievm
