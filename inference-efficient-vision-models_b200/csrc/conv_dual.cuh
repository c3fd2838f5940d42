// Two convolutions of one residual block in ONE launch (the block's first 3x3 conv and its 1x1 downsample,
// torchvision BasicBlock.forward as traced by convert_fx in quantization/engines.py:118).
//
// Both read the block input, with the same stride s, and the 1x1's operand tile IS the 3x3's centre-tap tile: the 3x3
// (pad 1) reads pixel (s*oy - 1 + ky, s*ox - 1 + kx), the 1x1 (pad 0) reads (s*oy, s*ox) = tap (1, 1).  The launch runs
// two TILE CLASSES over the 3x3's own im2col tensor map:
//   class 0: the 3x3 -- nine taps x channel chunks, weights `tmap_b`, epilogue / output of `p`
//   class 1: the downsample -- the centre tap only, weights `tmap_b2`, epilogue tables / output tensor of `x`
// Same tile shape (128 pixels x bn), same accumulator ring, same pipeline; class-1 tiles follow the class-0 tiles in the
// persistent schedule, so the light tiles fill the ragged last round.  What it buys: the downsample kernels were 12-16 us
// each for 3-8 us of work (launch ramp, prologue, pipeline fill and drain on an almost idle GPU); as a tile class they
// cost their steady-state share only, and the forward has three launches fewer.
//
// Pixel-pair form (INT8 layer 2.0: 3x3 stride 2 over 64-byte pixels).  An im2col TMA delivers ROWS, ~3 cycles each
// whatever their length, and nine 128-row x 64-byte loads per tile are 3 450 cycles for 1 150 cycles of
// MMAs.  Two adjacent pixels are one 128-byte "wide pixel"; output column ox reads pixels 2ox-1, 2ox, 2ox+1 = the second
// half of wide pixel ox-1 and both halves of wide pixel ox, so the conv is a 3 (rows, stride 2) x 2 (wide columns,
// stride 1, left pad 1) conv over 128-byte pixels with zero weights on the unused half: six loads of 128 rows instead
// of nine, and the 1x1 downsample is the first half of wide tap (1, 1).  ConvTcParams::kw / stride_w / pad_w carry the
// W-side geometry; the weights are packed accordingly by the host (ievm.cu: upload_conv_operands).
//
// Warp roles as in conv_tc.cuh (im2col mode): warp 0 TMA producer, warp 1 MMA issuer, 16 epilogue warps in groups.
#pragma once
#include "conv_tc.cuh"

namespace ievm {

struct ConvDualParams {
  void* out;             // the downsample's output tensor, [m_total][cout_pad]
  const float* ep0;      // its per-channel tables (as ConvTcParams::ep0 / ep1)
  const float* ep1;
  int out_zp, out_lo;
  int fast_round;
  int relu;              // f16
  int32_t* dump_acc;     // debug: raw accumulators of class 1
  // the same tables inside the kernel-parameter block (filled when cout_pad <= kEpConst; conv_s2.cuh reads its tables as
  // constant-bank operands through a uniform index, as the wide halo kernels do: no shared-memory loads in the epilogue)
  float epc0[kEpConst];
  float epc1[kEpConst];
};

template <int kDtype, int kCluster>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_dual_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_b2, const __grid_constant__ ConvTcParams p,
                 const __grid_constant__ ConvDualParams x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int kw = p.kw != 0 ? p.kw : p.ksize;              // taps along W (pixel-pair form: 2 wide taps)
  const int stride_w = p.kw != 0 ? p.stride_w : p.stride;
  const int pad_w = p.kw != 0 ? p.pad_w : p.pad;
  const int num_kb = p.ksize * kw * p.kchunks;            // class 0; class 1 has p.kchunks k-blocks
  const int a_bytes = p.a_stage_bytes;
  const int b_bytes = (p.bn / kCluster) * p.kc_bytes;
  const int G = p.kb_group;
  const int slots = p.stages * G;
  const int b_slots = p.resident_b ? num_kb + p.kchunks : slots;     // resident: class 1's blocks sit behind class 0's
  uint8_t* sA = smem;
  uint8_t* sB = smem + slots * a_bytes;
  float* s_ep = reinterpret_cast<float*>(sB + b_slots * b_bytes);    // [class][table][cout_pad]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ep + 4 * p.cout_pad);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint64_t* tempty_bar = tfull_bar + kMaxAcc;
  uint64_t* bres_bar = tempty_bar + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kGroupsMax = kEpiWarps / 4;

  for (int i = threadIdx.x; i < p.cout_pad; i += kConvThreads) {
    s_ep[i] = p.ep0[i];
    s_ep[p.cout_pad + i] = p.ep1[i];
    s_ep[2 * p.cout_pad + i] = x.ep0[i];
    s_ep[3 * p.cout_pad + i] = x.ep1[i];
  }
  const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_b2);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < p.nacc; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kCluster * kEpiWarps / (p.nacc < kGroupsMax ? p.nacc : kGroupsMax));
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCluster > 1) {
      tmem_alloc_pair(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  griddep_launch_dependents();

  // Schedule: work item t of stride `tile_step`; t < class_tiles is a class-0 tile, the rest are class-1 tiles of the
  // same (m, n) grid.
  const int class_tiles = kCluster > 1 ? ((p.m_tiles + kCluster - 1) / kCluster) * p.n_tiles : p.m_tiles * p.n_tiles;
  const int total_tiles = 2 * class_tiles;
  const int tile_first = kCluster > 1 ? static_cast<int>(blockIdx.x) / kCluster : static_cast<int>(blockIdx.x);
  const int tile_step = kCluster > 1 ? static_cast<int>(gridDim.x) / kCluster : static_cast<int>(gridDim.x);
  const int hw = p.ho * p.wo;
  const int ctr_y = p.ksize >> 1, ctr_x = kw >> 1;       // the tap whose tile is the downsample's operand

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (p.resident_b && elect_one()) {
      mbar_expect_tx(bres_bar, static_cast<uint32_t>((num_kb + p.kchunks) * b_bytes));
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(sB + kb * b_bytes, &tmap_b, bres_bar, kb * p.kc_elems, 0);
      for (int ch = 0; ch < p.kchunks; ++ch)
        tma_load_2d(sB + (num_kb + ch) * b_bytes, &tmap_b2, bres_bar, ch * p.kc_elems, 0);
    }
    __syncwarp();
    griddep_wait_conv();
    const uint32_t tx_bytes = static_cast<uint32_t>(p.resident_b ? p.a_tx_bytes : p.a_tx_bytes + b_bytes);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const bool cls1 = tile >= class_tiles;
      const int ct = cls1 ? tile - class_tiles : tile;
      const int m_group = ct / p.n_tiles;
      const int n_tile = ct - m_group * p.n_tiles;
      const int m_tile = kCluster > 1 ? m_group * kCluster + static_cast<int>(crank) : m_group;
      const int m0 = m_tile * kTileM;
      const int img = fast_div(m0, hw, p.hw_magic);
      const int rem = m0 - img * hw;
      const int oy = fast_div(rem, p.wo, p.wo_magic);
      const int ox = rem - oy * p.wo;
      const int base_w = ox * stride_w - pad_w;
      const int base_h = oy * p.stride - p.pad;
      const int ty_lo = cls1 ? ctr_y : 0, ty_hi = cls1 ? ctr_y + 1 : p.ksize;
      const int tx_lo = cls1 ? ctr_x : 0, tx_hi = cls1 ? ctr_x + 1 : kw;
      const int nkb = cls1 ? p.kchunks : num_kb;
      const CUtensorMap* wmap = cls1 ? &tmap_b2 : &tmap_b;
      int kb = 0, g = 0;
      for (int ty = ty_lo; ty < ty_hi; ++ty) {
        for (int tx = tx_lo; tx < tx_hi; ++tx) {
          for (int ch = 0; ch < p.kchunks; ++ch) {
            if (g == 0) wait_or_die(&empty_bar[stage], phase ^ 1u, 0x100u | stage, p.stuck_flag);
            const int slot = stage * G + g;
            const uint32_t cnt = static_cast<uint32_t>(min(G, nkb - kb));        // k-blocks of this barrier group (g == 0)
            if (elect_one()) {
              if (kCluster > 1) {
                if (leader && g == 0) mbar_expect_tx(&full_bar[stage], 2u * tx_bytes * cnt);
                tma_load_im2col_4d_pair(sA + slot * a_bytes, &tmap_a, &full_bar[stage], ch * p.kc_elems, base_w, base_h, img,
                                        static_cast<uint16_t>(tx), static_cast<uint16_t>(ty));
                tma_load_2d_pair(sB + slot * b_bytes, wmap, &full_bar[stage], kb * p.kc_elems,
                                 n_tile * p.bn + static_cast<int>(crank) * (p.bn / kCluster));
              } else {
                if (g == 0) mbar_expect_tx(&full_bar[stage], tx_bytes * cnt);
                tma_load_im2col_4d(sA + slot * a_bytes, &tmap_a, &full_bar[stage], ch * p.kc_elems, base_w, base_h, img,
                                   static_cast<uint16_t>(tx), static_cast<uint16_t>(ty));
                if (!p.resident_b) tma_load_2d(sB + slot * b_bytes, wmap, &full_bar[stage], kb * p.kc_elems, n_tile * p.bn);
              }
            }
            __syncwarp();
            ++kb;
            if (++g == G || kb == nkb) {
              g = 0;
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
    const bool wide = p.kc_bytes == 128;
    const uint32_t hi = smem_desc_hi(static_cast<uint32_t>(p.kc_bytes));
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA));
    const uint32_t b_lo0 = smem_desc_lo(smem_u32(sB));
    const uint32_t a_step = static_cast<uint32_t>(a_bytes) >> 4;
    const uint32_t b_step = static_cast<uint32_t>(b_bytes) >> 4;
    const uint32_t idesc = p.idesc;
    auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t accum) {
      if (kCluster > 1) {
        if (kDtype == kDtypeI8) umma_i8_lohi_pair(d, a_lo, b_lo, hi, idesc, accum);
        else umma_f16_lohi_pair(d, a_lo, b_lo, hi, idesc, accum);
      } else {
        if (kDtype == kDtypeI8) umma_i8_lohi(d, a_lo, b_lo, hi, idesc, accum);
        else umma_f16_lohi(d, a_lo, b_lo, hi, idesc, accum);
      }
    };
    if (p.resident_b) wait_or_die(bres_bar, 0, 0x500u, p.stuck_flag);
    for (int tile = tile_first; leader && tile < total_tiles; tile += tile_step) {
      const bool cls1 = tile >= class_tiles;
      const int nkb = cls1 ? p.kchunks : num_kb;
      const int b_first = cls1 ? num_kb : 0;             // resident weights: first block of this class
      wait_or_die(&tempty_bar[acc], acc_phase ^ 1u, 0x200u | acc, p.stuck_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
      for (int kb0 = 0; kb0 < nkb; kb0 += G) {
        const int cnt = min(G, nkb - kb0);
        wait_or_die(&full_bar[stage], phase, 0x300u | stage, p.stuck_flag);
        tc_fence_after();
        if (elect_one()) {
          for (int g = 0; g < cnt; ++g) {
            const int kb = kb0 + g;
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage * G + g) * a_step;
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(p.resident_b ? b_first + kb : stage * G + g) * b_step;
            mma(d_tmem, a_lo, b_lo, kb != 0 ? 1u : 0u);
            mma(d_tmem, a_lo + 2, b_lo + 2, 1u);
            if (wide) {
              mma(d_tmem, a_lo + 4, b_lo + 4, 1u);
              mma(d_tmem, a_lo + 6, b_lo + 6, 1u);
            }
          }
          const bool last = kb0 + cnt == nkb;
          if (kCluster > 1) {
            umma_commit_pair(&empty_bar[stage]);
            if (last) umma_commit_pair(&tfull_bar[acc]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (last) umma_commit(&tfull_bar[acc]);
          }
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++acc == p.nacc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int groups = p.nacc < kGroupsMax ? p.nacc : kGroupsMax;
    const int group = ((warp - 2) >> 2) & (groups - 1);
    const int sub = ((warp - 2) >> 2) / groups;
    const int csub = kGroupsMax / groups;
    const int row = quad * 32 + lane;
    const int nchunks = p.bn >> 4;
    griddep_wait_conv();
    int acc_next = 0, seq = 0;
    uint32_t acc_phase_next = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++seq) {
      const int acc = acc_next;
      const uint32_t acc_phase = acc_phase_next;
      if (++acc_next == p.nacc) {
        acc_next = 0;
        acc_phase_next ^= 1u;
      }
      if ((seq & (groups - 1)) != group) continue;
      const bool cls1 = tile >= class_tiles;
      const int ct = cls1 ? tile - class_tiles : tile;
      const int m_group = ct / p.n_tiles;
      const int n_tile = ct - m_group * p.n_tiles;
      const int m_tile = kCluster > 1 ? m_group * kCluster + static_cast<int>(crank) : m_group;
      const int m = m_tile * kTileM + row;
      const bool valid = m < p.m_total;
      const int n0 = n_tile * p.bn;
      // this tile's class: output tensor, tables, clamp
      uint8_t* out_row = static_cast<uint8_t*>(cls1 ? x.out : p.out) +
                         static_cast<size_t>(m) * p.out_pitch * (kDtype == kDtypeI8 ? 1 : 2);
      const float* e0 = s_ep + (cls1 ? 2 * p.cout_pad : 0);
      const float* e1 = e0 + p.cout_pad;
      const int zp = cls1 ? x.out_zp : p.out_zp;
      const int lo = cls1 ? x.out_lo : p.out_lo;
      const bool fast = (cls1 ? x.fast_round : p.fast_round) != 0;
      const bool relu = (cls1 ? x.relu : p.relu) != 0;
      int32_t* dump = cls1 ? x.dump_acc : p.dump_acc;

      wait_or_die(&tfull_bar[acc], acc_phase, 0x400u | acc, p.stuck_flag);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * p.acc_stride);
      uint32_t va[16], vb[16];
      auto chunk = [&](const uint32_t (&v)[16], int c) {
        const int ch = n0 + c * 16;
        if (valid && dump != nullptr) {
          int4* d = reinterpret_cast<int4*>(dump + static_cast<size_t>(m) * p.dump_pitch + ch);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            d[j] = make_int4(static_cast<int>(v[4 * j]), static_cast<int>(v[4 * j + 1]), static_cast<int>(v[4 * j + 2]),
                             static_cast<int>(v[4 * j + 3]));
        }
        if (kDtype == kDtypeI8) {
          const uint4 o = fast ? epilogue16_i8<true>(v, e0 + ch, e1 + ch, zp, lo) : epilogue16_i8<false>(v, e0 + ch, e1 + ch, zp, lo);
          if (valid) *reinterpret_cast<uint4*>(out_row + ch) = o;
        } else {
          __half* op = reinterpret_cast<__half*>(out_row) + ch;
          if (relu) epilogue16_f16<false, true>(v, e0 + ch, op, op, valid);
          else epilogue16_f16<false, false>(v, e0 + ch, op, op, valid);
        }
      };
      int c = sub;
      if (c < nchunks) tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>(c * 16), va);
      while (c < nchunks) {
        tmem_ld_wait();
        if (c + csub < nchunks) tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>((c + csub) * 16), vb);
        chunk(va, c);
        c += csub;
        if (c >= nchunks) break;
        tmem_ld_wait();
        if (c + csub < nchunks) tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>((c + csub) * 16), va);
        chunk(vb, c);
        c += csub;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCluster > 1 && !leader) mbar_arrive_remote(&tempty_bar[acc], 0u);
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (kCluster > 1) tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace ievm
