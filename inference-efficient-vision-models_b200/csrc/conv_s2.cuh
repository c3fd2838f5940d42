// A BasicBlock's first 3x3 STRIDE-2 conv and its 1x1 stride-2 downsample over 64-byte pixels (INT8 layer 2.0), as
// four PHASE PATCHES -- the stride-2 counterpart of conv_tc.cuh's halo mode, with conv_dual.cuh's two outputs.
//
// The layer is bound by the TMA unit's row rate (~3 cycles per row whatever its length): per-tap im2col
// loads are nine (pixel-pair form: six) 128-row loads per 128 output pixels for ~1 300 cycles of MMAs.  A stride-2 conv
// reads input pixel (2 oy - 1 + ky, 2 ox - 1 + kx): split the input by the parity of its row and column into four
// phase images P[py][px][y'][x'] = in[2 y' + py][2 x' + px].  Tap (ky, kx) then reads phase (ky != 1, kx != 1) at
// (oy - [ky == 0], ox - [kx == 0]) -- a STRIDE-1 shifted view.  TMA's traversal strides (elementStrides 2 x 2) deliver
// each phase of a patch densely, so one sub-tile of R output rows needs four boxes of (R + 1) x (W_out + 1) pixels:
// 580 rows per 112 output pixels instead of 896 per 128, every input pixel fetched once, and the nine taps are
// row-shifted UMMA descriptor views of the four buffers (as in halo mode: the 64B swizzle acts on absolute
// shared-memory address bits, buffers are 1 KB aligned).  The downsample reads pixel (2 oy, 2 ox) = phase (0, 0) at
// (oy, ox): the view of tap (1, 1), a second accumulator with its own weights, tables and output tensor.
//
// Positions: MMA row p of a sub-tile = r * wp + c, wp = W_out + 1, output (oy0 + r, c) for c < W_out (c = W_out is a junk
// lane); a phase buffer holds rows y' = oy0 - 1 .. oy0 + R - 1 and columns x' = -1 .. W_out - 1, so tap (ky, kx) starts
// ([ky != 0] * wp + [kx != 0]) rows into its phase's buffer.  Accumulators: ring of two PAIRS (conv, downsample).
// Warp roles as in conv_tc.cuh; epilogue group g drains class g & 1 of every second sub-tile.
#pragma once
#include "conv_dual.cuh"

namespace ievm {

__global__ void __launch_bounds__(kConvThreads, 1)
conv_s2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_b2, const __grid_constant__ ConvTcParams p,
               const __grid_constant__ ConvDualParams x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  constexpr int kRowBytes = 64;
  constexpr int kBlocks = 10;                               // nine taps + the downsample
  const int pb = p.phase_bytes;                             // one phase buffer (multiple of 1 KB)
  const int b_bytes = p.bn * kRowBytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.stages * p.a_stage_bytes + 1024;   // tap views of the last rows run < 1 KB past a stage
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + kBlocks * b_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;
  uint64_t* bres_bar = tempty_bar + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_b2);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);          // the four warps of the group that drains the buffer
    }
    mbar_init(bres_bar, 1);
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

  // contiguous range of sub-tiles (img, s) = (t / T, t % T) per CTA, sizes differ by at most one
  int t_begin, t_end;
  {
    const int q = p.total_subs / static_cast<int>(gridDim.x), r = p.total_subs - q * static_cast<int>(gridDim.x);
    const int b = static_cast<int>(blockIdx.x);
    t_begin = b * q + (b < r ? b : r);
    t_end = t_begin + q + (b < r ? 1 : 0);
  }

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      mbar_expect_tx(bres_bar, static_cast<uint32_t>(kBlocks * b_bytes));
      for (int kb = 0; kb < 9; ++kb) tma_load_2d(sB + kb * b_bytes, &tmap_b, bres_bar, kb * kRowBytes, 0);
      tma_load_2d(sB + 9 * b_bytes, &tmap_b2, bres_bar, 0, 0);
    }
    __syncwarp();
    griddep_wait_conv();
    int stage = 0;
    uint32_t phase = 0;
    int img = fast_div(t_begin, p.subs_per_img, p.spi_magic);
    int s = t_begin - img * p.subs_per_img;
    for (int t = t_begin; t < t_end; ++t) {
      wait_or_die(&empty_bar[stage], phase ^ 1u, 0x100u | stage, p.stuck_flag);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.a_tx_bytes));
        uint8_t* dst = sA + stage * p.a_stage_bytes;
        const int h0 = 2 * s * p.sub_rows - 2;             // input row of phase row y' = oy0 - 1, phase 0
#pragma unroll
        for (int q = 0; q < 4; ++q) tma_load_4d(dst + q * pb, &tmap_a, &full_bar[stage], 0, (q & 1) - 2, h0 + (q >> 1), img);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
      if (++s == p.subs_per_img) {
        s = 0;
        ++img;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t hi = smem_desc_hi(kRowBytes);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA));
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(sB));
    const uint32_t a_step = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
    const uint32_t b_step = static_cast<uint32_t>(b_bytes) >> 4;
    const uint32_t pb16 = static_cast<uint32_t>(pb) >> 4;
    const uint32_t wp4 = static_cast<uint32_t>(p.wp) * (kRowBytes / 16);
    const uint32_t idesc = p.idesc;
    wait_or_die(bres_bar, 0, 0x500u, p.stuck_flag);
    for (int t = t_begin; t < t_end; ++t) {
      const int i = t - t_begin;
      const int acc0 = 2 * (i & 1);
      const uint32_t acc_par = (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u;
      // one batched test: the patch has landed, both accumulators of the pair are drained
      if (!mbar_try_wait5(&full_bar[stage], phase, &tempty_bar[acc0], acc_par, &tempty_bar[acc0 + 1], acc_par,
                          &tempty_bar[acc0 + 1], acc_par, &tempty_bar[acc0 + 1], acc_par)) {
        wait_or_die(&full_bar[stage], phase, 0x300u | stage, p.stuck_flag);
        wait_or_die(&tempty_bar[acc0], acc_par, 0x200u | acc0, p.stuck_flag);
        wait_or_die(&tempty_bar[acc0 + 1], acc_par, 0x200u | (acc0 + 1), p.stuck_flag);
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d0 = tmem_base + static_cast<uint32_t>(acc0 * p.acc_stride);
        const uint32_t d1 = d0 + static_cast<uint32_t>(p.acc_stride);
        const uint32_t a_base = a_lo0 + static_cast<uint32_t>(stage) * a_step;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          const uint32_t ao = a_base + static_cast<uint32_t>((ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0)) * pb16 +
                              (ky != 0 ? wp4 : 0u) + (kx != 0 ? static_cast<uint32_t>(kRowBytes / 16) : 0u);
          const uint32_t bo = b_lo0 + static_cast<uint32_t>(tap) * b_step;
          umma_i8_lohi(d0, ao, bo, hi, idesc, tap != 0 ? 1u : 0u);
          umma_i8_lohi(d0, ao + 2, bo + 2, hi, idesc, 1u);
        }
        {
          // the downsample: pixel (2 oy, 2 ox) = phase (0, 0) at (oy, ox) = the view of tap (1, 1)
          const uint32_t ao = a_base + wp4 + static_cast<uint32_t>(kRowBytes / 16);
          const uint32_t bo = b_lo0 + 9u * b_step;
          umma_i8_lohi(d1, ao, bo, hi, idesc, 0u);
          umma_i8_lohi(d1, ao + 2, bo + 2, hi, idesc, 1u);
        }
        umma_commit(&tfull_bar[acc0]);
        umma_commit(&tfull_bar[acc0 + 1]);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int group = ((warp - 2) >> 2) & 3;
    const bool cls1 = (group & 1) != 0;
    const int row = quad * 32 + lane;
    const int r = fast_div(row, p.wp, p.wp_magic);
    const int c = row - r * p.wp;
    const bool pos_ok = row < p.sub_pos && c < p.wo;
    const int nchunks = p.bn >> 4;
    griddep_wait_conv();
    uint8_t* out_base = static_cast<uint8_t*>(cls1 ? x.out : p.out);
    const int zp = cls1 ? x.out_zp : p.out_zp;
    const int lo = cls1 ? x.out_lo : p.out_lo;
    const bool fast = (cls1 ? x.fast_round : p.fast_round) != 0;
    int32_t* dump = cls1 ? x.dump_acc : p.dump_acc;
    // e0 / e1: this class's per-channel tables in the kernel-parameter block (constant bank, uniform index)
    // fast_sel: -1 = rounding form chosen per chunk and the debug dump compiled in, 0 / 1 = straight-line code (see
    // epilogue_chunk in conv_tc.cuh)
    auto run = [&](auto fast_sel, const float* e0, const float* e1) __attribute__((always_inline)) {
      constexpr int kFs = decltype(fast_sel)::value;
      for (int i = group >> 1; t_begin + i < t_end; i += 2) {
        const int t = t_begin + i;
        const int acc = 2 * (i & 1) + (cls1 ? 1 : 0);
        const uint32_t acc_phase = static_cast<uint32_t>(i >> 1) & 1u;
        const int img = fast_div(t, p.subs_per_img, p.spi_magic);
        const int oy = (t - img * p.subs_per_img) * p.sub_rows + r;
        const bool valid = pos_ok && oy < p.ho;
        const int m = (img * p.ho + oy) * p.wo + c;
        uint8_t* out_row = out_base + static_cast<size_t>(m) * p.out_pitch;
        wait_or_die(&tfull_bar[acc], acc_phase, 0x400u | acc, p.stuck_flag);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * p.acc_stride);
        uint32_t va[16], vb[16];
        auto chunk = [&](const uint32_t (&v)[16], int ch) {
          if (kFs < 0 && valid && dump != nullptr) {
            int4* d = reinterpret_cast<int4*>(dump + static_cast<size_t>(m) * p.dump_pitch + ch);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              d[j] = make_int4(static_cast<int>(v[4 * j]), static_cast<int>(v[4 * j + 1]), static_cast<int>(v[4 * j + 2]),
                               static_cast<int>(v[4 * j + 3]));
          }
          uint4 o;
          if (kFs >= 0) o = epilogue16_i8<kFs == 1>(v, e0 + ch, e1 + ch, zp, lo);
          else o = fast ? epilogue16_i8<true>(v, e0 + ch, e1 + ch, zp, lo) : epilogue16_i8<false>(v, e0 + ch, e1 + ch, zp, lo);
          if (valid) *reinterpret_cast<uint4*>(out_row + ch) = o;
        };
        tmem_ld_32x32b_x16(t_row, va);
#pragma unroll 1
        for (int cc = 0; cc < nchunks; cc += 2) {
          tmem_ld_wait();
          if (cc + 1 < nchunks) tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>((cc + 1) * 16), vb);
          chunk(va, cc * 16);
          if (cc + 1 >= nchunks) break;
          tmem_ld_wait();
          if (cc + 2 < nchunks) tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>((cc + 2) * 16), va);
          chunk(vb, (cc + 1) * 16);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
    };
    auto dispatch = [&](const float* e0, const float* e1) {
      if (fast && dump == nullptr) run(std::integral_constant<int, 1>{}, e0, e1);      // the product's form
      else run(std::integral_constant<int, -1>{}, e0, e1);
    };
    if (cls1) dispatch(x.epc0, x.epc1);
    else dispatch(p.epc0, p.epc1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace ievm
