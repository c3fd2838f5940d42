// INT8 stem (7x7 stride-2 pad-3 conv over the 3-channel quantised input + requant + ReLU) on tcgen05.
//
// The input tensor is u8 NHWC4 (4th byte = zero point) with a zero-point border (simt_kernels.cuh).  One output pixel needs, for each of the 7
// filter rows, the 7 input pixels ix = 2*ox-3 .. 2*ox+3; widening the window by one pixel on the left
// (ix = 2*ox-4, weight 0) makes it 8 pixels * 4 B = one 8-byte-aligned 32-byte segment.  The GEMM is
//
//   M = 128 output pixels, K = 8 segments * 32 B (7 filter rows + 1 zero segment) = 2 k-blocks of
//   128 B, N = padded cout (64 for the pruned student)
//
// The A operand is too narrow per pixel (4 B) for TMA im2col (16 B minimum), so four builder warps
// gather it with 8-byte loads and write it to shared memory in the 128B-swizzled K-major layout that
// tcgen05.mma expects; out-of-image pixels read the zero-point border, and the epilogue subtracts
// zp * sum(w) (exact in int32) before the usual float requantisation.
// Weights (cout_pad x 256 B) stay resident in shared memory; accumulators are double buffered in TMEM.
//
// Warp roles (800 threads): warp 0 = weights TMA + TMEM owner + MMA issuer, warps 1..8 = A builders
// (thread = output pixel x k-block half, all its loads in flight at once), warps 9..24 = epilogue.
#pragma once
#include "conv_tc.cuh"
#include "simt_kernels.cuh"

namespace ievm {

constexpr int kStemBuildWarps = 8;          // 256 threads: thread = (tile row, k-block half)
constexpr int kStemThreads = 32 * (1 + kStemBuildWarps + kEpiWarps);
constexpr int kStemKBytes = 256;          // 8 segments x 32 B
constexpr int kStemStages = 4;

struct StemTcParams {
  int n, h, w, ho, wo;
  int m_total, m_tiles;
  int cpad;                 // UMMA N
  int in_zp;
  int tmem_cols, acc_stride;
  uint32_t idesc;
  const uint8_t* xq;        // [n][h + 6][w + 8][4], image at (3, 4), zero-point border
  uint8_t* out;             // [n][ho][wo][cpad]
  const float* bdiv;
  const float* mult;
  const int* zwsum;         // in_zp * sum_k w[c][k]
  int out_zp, out_lo;
  int32_t* dump_acc;        // debug: corrected accumulators [m_total][cpad]
  unsigned int* stuck_flag;
};

__global__ void __launch_bounds__(kStemThreads, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const StemTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  constexpr int kABlock = kTileM * 128;            // one k-block of A: 128 rows x 128 B
  constexpr int kAStage = 2 * kABlock;
  const int b_block = p.cpad * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStemStages * kAStage;
  float* s_bd = reinterpret_cast<float*>(sB + 2 * b_block);
  float* s_mu = s_bd + p.cpad;
  int* s_zw = reinterpret_cast<int*>(s_mu + p.cpad);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_zw + p.cpad);
  uint64_t* empty_bar = full_bar + kStemStages;
  uint64_t* tfull_bar = empty_bar + kStemStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.cpad; i += kStemThreads) {
    s_bd[i] = p.bdiv[i];
    s_mu[i] = p.mult[i];
    s_zw[i] = p.zwsum[i];
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_w);
      for (int i = 0; i < kStemStages; ++i) {
        mbar_init(&full_bar[i], kStemBuildWarps * 32);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], kEpiWarps);
      }
      mbar_init(w_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const int hw = p.ho * p.wo;

  if (warp == 0) {
    // ================================ weights + MMA issuer ================================
    if (lane == 0) {
      mbar_expect_tx(w_bar, static_cast<uint32_t>(2 * b_block));
      tma_load_2d(sB, &tmap_w, w_bar, 0, 0);
      tma_load_2d(sB + b_block, &tmap_w, w_bar, 128, 0);
    }
    wait_or_die(w_bar, 0, 0x600u, p.stuck_flag);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      wait_or_die(&tempty_bar[acc], acc_phase ^ 1u, 0x610u | acc, p.stuck_flag);
      wait_or_die(&full_bar[stage], phase, 0x620u | stage, p.stuck_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
      const uint32_t hi = smem_desc_hi(128);
      const uint32_t a_lo = smem_desc_lo(smem_u32(sA)) + static_cast<uint32_t>(stage) * (kAStage >> 4);
      const uint32_t b_lo = smem_desc_lo(smem_u32(sB));
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_i8_lohi(d_tmem, a_lo + kb * (kABlock >> 4) + 2 * k, b_lo + kb * (static_cast<uint32_t>(b_block) >> 4) + 2 * k,
                         hi, p.idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == kStemStages) {
        stage = 0;
        phase ^= 1u;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp <= kStemBuildWarps) {
    // ================================ A builders ================================
    // 256 threads per tile: thread = (output pixel r, k-block kbh).  k-block 0 holds filter rows 0..3,
    // k-block 1 rows 4..6 (+ the zero-weight segment, left unwritten).
    const int bt = (warp - 1) * 32 + lane;
    const int r = bt & 127;
    const int kbh = bt >> 7;
    const int nseg = kbh == 0 ? 4 : 3;
    const uint32_t sw = static_cast<uint32_t>(r & 7);         // 128B-swizzle phase of this row
    const uint2* in_pairs = reinterpret_cast<const uint2*>(p.xq);
    const int hp = p.h + kInPadH;
    const int wp2 = (p.w + kInPadW) >> 1;                    // pixel pairs per padded row
    uint32_t chunk_off[8];                                   // swizzled 16-byte chunk offsets of this row
#pragma unroll
    for (int c = 0; c < 8; ++c) chunk_off[c] = ((static_cast<uint32_t>(c) ^ sw) << 4);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const int m = min(tile * kTileM + r, p.m_total - 1);      // rows past the end rebuild the last pixel (discarded)
      const int img = m / hw;
      const int rem = m - img * hw;
      const int oy = rem / p.wo;
      const int ox = rem - oy * p.wo;
      // Padded input (see quantize kernel): filter row ky of this pixel starts at padded row 2*oy + ky,
      // padded column 2*ox (= input column 2*ox - 4); 32 bytes = 4 aligned pixel pairs.
      const uint2* src = in_pairs + (static_cast<size_t>(img) * hp + 2 * oy + 4 * kbh) * wp2 + ox;
      uint2 q[4][4];
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        if (sg < nseg) {
#pragma unroll
          for (int j = 0; j < 4; ++j) q[sg][j] = __ldg(src + sg * wp2 + j);
        }
      }
      wait_or_die(&empty_bar[stage], phase ^ 1u, 0x630u | stage, p.stuck_flag);
      uint8_t* blk = sA + stage * kAStage + kbh * kABlock + r * 128;
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        if (sg < nseg) {
          *reinterpret_cast<uint4*>(blk + chunk_off[sg * 2]) = make_uint4(q[sg][0].x, q[sg][0].y, q[sg][1].x, q[sg][1].y);
          *reinterpret_cast<uint4*>(blk + chunk_off[sg * 2 + 1]) = make_uint4(q[sg][2].x, q[sg][2].y, q[sg][3].x, q[sg][3].y);
        }
      }
      fence_proxy_async_smem();                                // generic-proxy writes -> visible to the MMA
      mbar_arrive(&full_bar[stage]);
      if (++stage == kStemStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int half = (warp - 1 - kStemBuildWarps) >> 2;   // quadrant = warp % 4, kEpiSub warps each
    const int row = quad * 32 + lane;
    const int nchunks = p.cpad >> 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      const int m = tile * kTileM + row;
      const bool valid = m < p.m_total;
      wait_or_die(&tfull_bar[acc], acc_phase, 0x640u | acc, p.stuck_flag);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * p.acc_stride);
      for (int c = half; c < nchunks; c += kEpiSub) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>(c * 16), v);
        tmem_ld_wait();
        const int ch = c * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int4 z = *reinterpret_cast<const int4*>(s_zw + ch + 4 * j);
          v[4 * j] -= static_cast<uint32_t>(z.x);
          v[4 * j + 1] -= static_cast<uint32_t>(z.y);
          v[4 * j + 2] -= static_cast<uint32_t>(z.z);
          v[4 * j + 3] -= static_cast<uint32_t>(z.w);
        }
        if (valid && p.dump_acc != nullptr) {
          int4* d = reinterpret_cast<int4*>(p.dump_acc + static_cast<size_t>(m) * p.cpad + ch);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            d[j] = make_int4(static_cast<int>(v[4 * j]), static_cast<int>(v[4 * j + 1]),
                             static_cast<int>(v[4 * j + 2]), static_cast<int>(v[4 * j + 3]));
        }
        const uint4 o = epilogue16_i8<false>(v, s_bd + ch, s_mu + ch, p.out_zp, p.out_lo);
        if (valid) *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(m) * p.cpad + ch) = o;
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
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace ievm
