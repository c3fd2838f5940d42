// Weights-stationary 3x3 / stride-1 convolution for narrow INT8 layers (<= 64 input and output channels: layer1 of
// the pruned ResNet-18), with the requantisation (+ the whole quantized::add_relu) fused.
//
// Why a second conv kernel.  tcgen05.mma fetches SHARED-memory operands at ~64 B/clk/SM (measured),
// so the pixels-as-M formulation of conv_tc.cuh, which reads 128 x 32 B of activations AND 64 x 32 B of weights per
// M128 x N64 x K32 instruction, needs ~100 clk per MMA where the tensor pipe needs 32: a 64-channel 56x56 layer
// takes 43 us for 15 us of math.  Here the roles are swapped, as in the fused front end (frontend_v2.cuh):
//
//   * the WEIGHTS are the M-side operand and live in TENSOR MEMORY for the whole kernel (TS-form MMA): nothing is
//     re-read for them;
//   * M = 128 = two output rows x 64 channels: lanes 64..127 hold the same filters shifted down by one input row,
//     so that one activation operand yields output rows 2t and 2t+1 (K = 4 input rows x 3 columns x 64 channels,
//     a quarter of it zero weights);
//   * the ACTIVATIONS are the N-side operand: one tiled TMA box brings the 4 x (W+2) x 64 B input patch of a row
//     pair, and the 12 taps are 12 row-shifted views of it (64B-swizzled K-major, the halo trick of conv_tc.cuh).
//     Per MMA the shared-memory port delivers 64 x 32 B = 32 clk of traffic for 32 clk of math.
//
// TMEM lane = (output row parity, channel), TMEM column = x: an epilogue thread owns ONE channel of a row, so the
// per-channel requantisation constants are registers, and a warp's store instruction writes the 32 consecutive
// channel bytes of one pixel (one full 32-byte sector).  The same holds for the residual bytes it reads.
//
// STATUS (round 1): bit-exact (all parity tests pass with IEVM_WT=1) but NOT used by default.  Measured on B200 at
// batch 256, 56x56x64 layer: the MMA stream alone takes 43 us here as well -- a TS-form M128 x N64 x K32 instruction
// costs ~73 clk, not the 32 clk of its math, so small-N instructions are bound per instruction, not per byte -- and
// the byte-granular epilogue traffic costs +29 us (stores) / +80 us more (residual loads): 72 / 152 us against
// 48 / 64 us for conv_tc.cuh.  Kept as the record of that experiment (profiles/r01_timing_experiments.md).
#pragma once
#include "conv_tc.cuh"

namespace ievm {

constexpr int kWtEpiWarps = 16;                 // 4 per TMEM lane quadrant, each taking 14 of the 56 columns
constexpr int kWtThreads = 64 + 32 * kWtEpiWarps;
constexpr int kWtAcc = 5;                       // accumulator buffers of 64 columns
constexpr int kWtACol0 = kWtAcc * 64;           // weight operand: 192 columns behind them
constexpr int kWtKSteps = 24;                   // 12 taps x 2 k-steps of 32 bytes
constexpr int kWtMaxStages = 10;

struct ConvWtParams {
  int n, h, w;               // images, height (even), width (<= 62)
  int wp;                    // w + 2
  int tiles;                 // n * h / 2
  int tiles_per_img;         // h / 2
  uint32_t tpi_magic;
  int stages, stage_bytes, tx_bytes;
  int cols_per_warp;         // ceil(w / 4)
  uint32_t idesc;
  const uint8_t* wpack;      // [128][768] row-major: lane (r, co), K = (dy * 3 + kx) * 64 + ci
  uint8_t* out;              // [n][h][w][64]
  const uint8_t* res;        // residual, same layout, or nullptr
  const float* ep0;          // bias / (x_s * w_s[c])
  const float* ep1;          // (x_s * w_s[c]) / out_s
  int out_zp, out_lo;
  float a_scale, res_scale, inv_add_scale;
  int res_zp, add_zp;
  int fast_round;
  unsigned int* stuck_flag;
};

template <bool kHasRes>
__global__ void __launch_bounds__(kWtThreads, 1)
conv_wt_kernel(const __grid_constant__ CUtensorMap tmap_x, const ConvWtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;                                            // activation patches, p.stages x p.stage_bytes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kWtMaxStages;
  uint64_t* tfull_bar = empty_bar + kWtMaxStages;
  uint64_t* tempty_bar = tfull_bar + kWtAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kWtAcc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kWtAcc; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kWtEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  griddep_launch_dependents();

  // weight operand -> tensor memory (independent of the previous kernel): warps 2..5 cover the four lane quadrants
  if (warp >= 2 && warp < 6) {
    const int quad = warp & 3;
    const uint4* src = reinterpret_cast<const uint4*>(p.wpack + static_cast<size_t>(quad * 32 + lane) * (kWtKSteps * 32));
    const uint32_t a_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kWtACol0;
#pragma unroll 1
    for (int c = 0; c < kWtKSteps * 8 / 16; ++c) {
      uint32_t wv[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 q = __ldg(src + 4 * c + j);
        wv[4 * j] = q.x;
        wv[4 * j + 1] = q.y;
        wv[4 * j + 2] = q.z;
        wv[4 * j + 3] = q.w;
      }
      tmem_st_32x32b_x16(a_addr + static_cast<uint32_t>(16 * c), wv);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ================================ TMA producer: one input patch per row pair ================================
    griddep_wait();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int img = fast_div(tile, p.tiles_per_img, p.tpi_magic);
      const int oy0 = 2 * (tile - img * p.tiles_per_img);
      wait_or_die(&empty_bar[stage], phase ^ 1u, 0xA00u | stage, p.stuck_flag);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.tx_bytes));
        tma_load_4d(sA + stage * p.stage_bytes, &tmap_x, &full_bar[stage], 0, -1, oy0 - 1, img);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t hi = smem_desc_hi(64u);
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(sA));
    const uint32_t stage_step = static_cast<uint32_t>(p.stage_bytes) >> 4;
    uint32_t tap_off[12];                          // patch rows (16-byte units: 4 per 64-byte row) of tap (dy, kx)
#pragma unroll
    for (int tp = 0; tp < 12; ++tp) tap_off[tp] = static_cast<uint32_t>((tp / 3) * p.wp + (tp % 3)) * 4u;
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      wait_or_die(&tempty_bar[acc], acc_phase ^ 1u, 0xA20u | acc, p.stuck_flag);
      wait_or_die(&full_bar[stage], phase, 0xA30u | stage, p.stuck_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 64);
      const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(stage) * stage_step;
      if (elect_one()) {
#ifndef IEVM_EXP_WT_NOMMA
#pragma unroll
        for (int tp = 0; tp < 12; ++tp) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_ts<0>(d_tmem, tmem_base + kWtACol0 + static_cast<uint32_t>(tp * 2 + ks) * 8u, b_lo + tap_off[tp] + 2u * ks, hi,
                       p.idesc, (tp | ks) != 0 ? 1u : 0u);
        }
#endif
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
      if (++acc == kWtAcc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int sub = (warp - 2) >> 2;               // which quarter of the row's columns
    const int r = quad >> 1;                       // output row parity held by this lane quadrant
    const int co = (quad & 1) * 32 + lane;         // channel
    const int x0 = sub * p.cols_per_warp;          // first column of this warp (<= 16 columns)
    const int ncols = min(p.cols_per_warp, p.w - x0);
    const float bd = p.ep0[co], mu = p.ep1[co];
    griddep_wait();                                // before the first residual read / output store
    // fused add_relu constants (see conv_tc.cuh)
    const float lo_f = static_cast<float>(p.out_lo - p.out_zp), hi_f = static_cast<float>(255 - p.out_zp);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int img = fast_div(tile, p.tiles_per_img, p.tpi_magic);
      const int oy = 2 * (tile - img * p.tiles_per_img) + r;
      const size_t pix0 = (static_cast<size_t>(img) * p.h + oy) * p.w + x0;
      uint8_t* op = p.out + pix0 * 64 + co;
      uint32_t rq[16];
      if (kHasRes) {
        const uint8_t* rp = p.res + pix0 * 64 + co;
#pragma unroll
#ifdef IEVM_EXP_WT_NOSTORE
        for (int j = 0; j < 16; ++j) rq[j] = static_cast<uint32_t>(j + lane);
#else
        for (int j = 0; j < 16; ++j) rq[j] = j < ncols ? __ldg(rp + j * 64) : 0u;     // does not depend on the MMA
#endif
      }
      wait_or_die(&tfull_bar[acc], acc_phase, 0xA40u | acc, p.stuck_flag);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * 64 + x0), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);              // accumulator words are in registers
      if (++acc == kWtAcc) {
        acc = 0;
        acc_phase ^= 1u;
      }
#ifdef IEVM_EXP_WT_NOEPI
      continue;
#endif
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < ncols) {
          const float t0 = __fmul_rn(__fadd_rn(__int2float_rn(static_cast<int>(v[j])), bd), mu);
          int q;
          if (!kHasRes) {
            q = p.fast_round ? round_add<true>(t0, p.out_zp) : round_add<false>(t0, p.out_zp);
            q = min(max(q, p.out_lo), 255);
          } else {
            float t = fminf(fmaxf(t0, lo_f), hi_f);
            t = __fadd_rn(__fadd_rn(t, kRoundMagic), -kRoundMagic);                  // == float(q2 - zp2)
            // ATen's fused dequantisation, as in conv_tc.cuh: fma(scale, float(q), fl(scale * -zp))
            const float a = __fmaf_rn(p.a_scale, __fadd_rn(t, static_cast<float>(p.out_zp)),
                                      __fmul_rn(p.a_scale, -static_cast<float>(p.out_zp)));
            const float rb = __fmaf_rn(p.res_scale, __uint2float_rn(rq[j]), __fmul_rn(p.res_scale, -static_cast<float>(p.res_zp)));
            const float s = fmaxf(__fadd_rn(a, rb), 0.0f);
            const float u = __fmul_rn(s, p.inv_add_scale);
            q = p.fast_round ? round_add<true>(u, p.add_zp) : round_add<false>(u, p.add_zp);
            q = min(max(q, 0), 255);
          }
#ifdef IEVM_EXP_WT_NOSTORE
          if (q == 0x7fffffff) op[j * 64] = static_cast<uint8_t>(q);
#else
          op[j * 64] = static_cast<uint8_t>(q);
#endif
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace ievm
