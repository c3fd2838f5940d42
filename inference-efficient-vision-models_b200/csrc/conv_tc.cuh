// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel, (tap, cin)] * W[cout, (tap, cin)]
//
// * A tiles (128 consecutive output pixels in (n, oy, ox) order x one channel chunk of one filter tap)
//   are gathered straight from the NHWC activation tensor by TMA in im2col mode; spatial padding and
//   the ragged last tile are zero-filled by the TMA unit (zero == the real value 0 because every
//   tensor-core layer's input zero-point is 0, checked at engine creation).
// * W tiles come from a pre-packed K-major [cout_pad][taps * cin_chunks * kc] matrix by tiled TMA.
// * Both land in shared memory in the 64B/128B-swizzled K-major layout tcgen05.mma reads directly.
// * Accumulators (int32 for kind::i8, fp32 for kind::f16) live in TMEM, double buffered so that the
//   epilogue of tile i overlaps the MMAs of tile i+1; the kernel is persistent (grid = #SMs).
// * The epilogue is fused: INT8 requantisation (+ReLU clamp) and, for a block's last conv, the whole
//   quantized::add_relu with the residual; FP16 bias (+residual) (+ReLU).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace ievm {

constexpr int kTileM = 128;
constexpr int kConvThreads = 192;
constexpr int kMaxStages = 12;

enum : int { kDtypeI8 = 0, kDtypeF16 = 1 };

struct ConvTcParams {
  // implicit-GEMM geometry
  int m_total;         // n * ho * wo
  int ho, wo;
  int stride, pad, ksize;
  int kchunks;         // channel chunks per filter tap
  int kc_bytes;        // bytes per smem row == bytes per channel chunk (64 or 128)
  int kc_elems;
  int bn;              // UMMA N (multiple of 16, <= 256)
  int n_tiles, m_tiles;
  int stages;
  int tmem_cols;       // power of two >= 32 covering two accumulator buffers
  int acc_stride;      // column offset of the second accumulator buffer
  uint32_t idesc;
  // epilogue
  void* out;
  int out_pitch;       // elements per output pixel row (== padded cout)
  const void* res;     // residual tensor (same pixel indexing), or nullptr
  int res_pitch;
  const float* ep0;    // i8: bias / (x_scale * w_scale[c]);  f16: folded bias
  const float* ep1;    // i8: (x_scale * w_scale[c]) / out_scale
  int out_zp, out_lo;  // i8: zero point of the conv's own output and its lower clamp (zp if ReLU else 0)
  float a_scale;       // i8 residual path: scale of the conv's own output
  float res_scale;
  int res_zp;
  float inv_add_scale; // 1 / scale of quantized::add_relu's output
  int add_zp;
  int relu;            // f16: apply ReLU at the end
  int32_t* dump_acc;   // debug: raw accumulators [m_total][dump_pitch] (bit pattern for f16)
  int dump_pitch;
  unsigned int* stuck_flag;   // mapped host word; written before a bounded wait gives up
};

__device__ __forceinline__ void wait_or_die(uint64_t* bar, uint32_t parity, uint32_t code, unsigned int* flag) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > IEVM_WAIT_LIMIT_NS) {
      if (flag) {
        *reinterpret_cast<volatile unsigned int*>(flag) = code;
        __threadfence_system();
      }
      __trap();
    }
  }
}

// ---- INT8 epilogue arithmetic: float32, no FMA contraction, round-half-even (== fbgemm/ATen) ----
__device__ __forceinline__ int requant_i8(int acc, float bdiv, float mult, int zp, int lo) {
  const float v = __fmul_rn(__fadd_rn(__int2float_rn(acc), bdiv), mult);
  const int q = __float2int_rn(v) + zp;
  return min(max(q, lo), 255);
}
__device__ __forceinline__ int add_relu_i8(int aq, int a_zp, float a_scale, int rq, int r_zp, float r_scale,
                                           float inv_scale, int zp) {
  const float a = __fmul_rn(__int2float_rn(aq - a_zp), a_scale);
  const float b = __fmul_rn(__int2float_rn(rq - r_zp), r_scale);
  const float s = fmaxf(__fadd_rn(a, b), 0.0f);
  const int q = __float2int_rn(__fmul_rn(s, inv_scale)) + zp;
  return min(max(q, 0), 255);
}

template <int kDtype>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (required by the 128B swizzle atoms) in the shared address space.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int a_bytes = kTileM * p.kc_bytes;
  const int b_bytes = p.bn * p.kc_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + p.stages * b_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;       // warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);        // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int num_kb = p.ksize * p.ksize * p.kchunks;
  const int hw = p.ho * p.wo;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        const int m0 = m_tile * kTileM;
        const int img = m0 / hw;
        const int rem = m0 - img * hw;
        const int oy = rem / p.wo;
        const int ox = rem - oy * p.wo;
        const int base_w = ox * p.stride - p.pad;
        const int base_h = oy * p.stride - p.pad;
        int kb = 0;
        for (int ty = 0; ty < p.ksize; ++ty) {
          for (int tx = 0; tx < p.ksize; ++tx) {
            for (int ch = 0; ch < p.kchunks; ++ch, ++kb) {
              wait_or_die(&empty_bar[stage], phase ^ 1u, 0x100u | stage, p.stuck_flag);
              mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(a_bytes + b_bytes));
              tma_load_im2col_4d(sA + stage * a_bytes, &tmap_a, &full_bar[stage], ch * p.kc_elems, base_w, base_h,
                                 img, static_cast<uint16_t>(tx), static_cast<uint16_t>(ty));
              tma_load_2d(sB + stage * b_bytes, &tmap_b, &full_bar[stage], kb * p.kc_elems, n_tile * p.bn);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int ksteps = p.kc_bytes / 32;      // one tcgen05.mma consumes 32 bytes of K per row
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      wait_or_die(&tempty_bar[acc], acc_phase ^ 1u, 0x200u | acc, p.stuck_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
      for (int kb = 0; kb < num_kb; ++kb) {
        wait_or_die(&full_bar[stage], phase, 0x300u | stage, p.stuck_flag);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(sA + stage * a_bytes);
          const uint32_t b_addr = smem_u32(sB + stage * b_bytes);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adesc = make_smem_desc(a_addr + k * 32, p.kc_bytes);
            const uint64_t bdesc = make_smem_desc(b_addr + k * 32, p.kc_bytes);
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            if (kDtype == kDtypeI8) umma_i8(d_tmem, adesc, bdesc, p.idesc, accum);
            else umma_f16(d_tmem, adesc, bdesc, p.idesc, accum);
          }
          umma_commit(&empty_bar[stage]);                 // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const int m = m_tile * kTileM + row;
      const bool valid = m < p.m_total;
      const int n0 = n_tile * p.bn;
      wait_or_die(&tfull_bar[acc], acc_phase, 0x400u | acc, p.stuck_flag);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * p.acc_stride);
      for (int c0 = 0; c0 < p.bn; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>(c0), v);
        tmem_ld_wait();
        const int ch = n0 + c0;
        if (valid && p.dump_acc != nullptr) {
          int4* d = reinterpret_cast<int4*>(p.dump_acc + static_cast<size_t>(m) * p.dump_pitch + ch);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            d[j] = make_int4(static_cast<int>(v[4 * j]), static_cast<int>(v[4 * j + 1]),
                             static_cast<int>(v[4 * j + 2]), static_cast<int>(v[4 * j + 3]));
        }
        if (kDtype == kDtypeI8) {
          float bd[16], mu[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ep0 + ch) + j);
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(p.ep1 + ch) + j);
            bd[4 * j] = b4.x; bd[4 * j + 1] = b4.y; bd[4 * j + 2] = b4.z; bd[4 * j + 3] = b4.w;
            mu[4 * j] = m4.x; mu[4 * j + 1] = m4.y; mu[4 * j + 2] = m4.z; mu[4 * j + 3] = m4.w;
          }
          uint32_t rq[4] = {0u, 0u, 0u, 0u};
          const bool has_res = p.res != nullptr;
          if (has_res && valid) {
            const uint4 r4 = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.res) +
                                                             static_cast<size_t>(m) * p.res_pitch + ch);
            rq[0] = r4.x; rq[1] = r4.y; rq[2] = r4.z; rq[3] = r4.w;
          }
          uint32_t packed[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int i = 4 * j + b;
              int q = requant_i8(static_cast<int>(v[i]), bd[i], mu[i], p.out_zp, p.out_lo);
              if (has_res) {
                const int r = static_cast<int>((rq[j] >> (8 * b)) & 0xffu);
                q = add_relu_i8(q, p.out_zp, p.a_scale, r, p.res_zp, p.res_scale, p.inv_add_scale, p.add_zp);
              }
              w |= static_cast<uint32_t>(q) << (8 * b);
            }
            packed[j] = w;
          }
          if (valid) {
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out) + static_cast<size_t>(m) * p.out_pitch + ch) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
        } else {
          float f[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ep0 + ch) + j);
            f[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
          if (p.res != nullptr && valid) {
            const __half* rp = static_cast<const __half*>(p.res) + static_cast<size_t>(m) * p.res_pitch + ch;
            const uint4 r0 = *reinterpret_cast<const uint4*>(rp);
            const uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(&rw[j]));
              f[2 * j] += r2.x;
              f[2 * j + 1] += r2.y;
            }
          }
          uint32_t hw2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float x = f[2 * j], y = f[2 * j + 1];
            if (p.relu) {
              x = fmaxf(x, 0.0f);
              y = fmaxf(y, 0.0f);
            }
            const __half2 h = __floats2half2_rn(x, y);
            hw2[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          if (valid) {
            __half* op = static_cast<__half*>(p.out) + static_cast<size_t>(m) * p.out_pitch + ch;
            *reinterpret_cast<uint4*>(op) = make_uint4(hw2[0], hw2[1], hw2[2], hw2[3]);
            *reinterpret_cast<uint4*>(op + 8) = make_uint4(hw2[4], hw2[5], hw2[6], hw2[7]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace ievm
