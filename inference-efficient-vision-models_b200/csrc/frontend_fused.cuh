// Fused INT8 front end: quantize_per_tensor -> 7x7/2 stem conv (+folded BN) -> requant + ReLU ->
// MaxPool2d(3,2,1), one kernel, f32 NCHW in, u8 NHWC pooled tensor out.  Neither the quantised image nor
// the 112x112xC stem output (the largest tensor of the network) ever touches HBM.
//
// Work decomposition.  A work unit is (image, strip): a strip is 16 stem-output columns wide and the
// CTA walks down it in tiles of 8 stem rows (= one 128-row MMA tile, row r <-> (ry = r / 16, rx = r % 16)).
// Strips start at stem column 14*s - 1, so a strip owns pooled columns 7*s .. 7*s + 6 outright (their
// 3-wide windows need stem columns 14*s - 1 .. 14*s + 13); neighbouring strips recompute two stem
// columns (14 % overhead) and nothing else.  Vertically nothing is recomputed: tile t owns pooled rows
// 4*t .. 4*t + 3, whose windows need stem rows 8*t - 1 .. 8*t + 7 -- the tile's own eight rows plus the
// last row of the previous tile, which is still in shared memory.
//
// Pipeline inside a CTA (persistent, grid = #SMs):
//   builder warps : cp.async the f32 input rows of the tile three steps ahead into a raw staging area
//                   (no registers held while HBM answers), build this tile's A operand from the u8 ring
//                   (LDS -> 128B-swizzled STS), then quantise the rows of the next tile from the raw
//                   staging area into the ring.  Every input pixel is quantised exactly once per strip.
//   MMA warp      : 8 x tcgen05.mma kind::i8 (M128 x N x K32) per tile, accumulators double buffered in TMEM.
//   epilogue warps: TMEM -> int32 accumulator tile in shared memory -> 3x3/2 max on the accumulators ->
//                   requant (+ReLU clamp) of the pooled values only -> stores of the pooled tensor.
// GEMM layout (shared with stem_tc.cuh): K = 8 segments x 32 B, segment ky = filter row ky, inside a
// segment byte j*4 + c <-> input pixel 2*ox - 4 + j, channel c (j = 0 and c = 3 have zero weights);
// out-of-image pixels hold the zero point and the epilogue subtracts zp * sum(w).
#pragma once
#include "conv_tc.cuh"
#include "stem_tc.cuh"

#include <climits>

namespace ievm {

constexpr int kFeBuildWarps = 8;                 // 256 threads: (tile row, k-block half)
constexpr int kFeEpiWarps = 8;
constexpr int kFeThreads = 32 * (1 + kFeBuildWarps + kFeEpiWarps);
constexpr int kFeStages = 2;
constexpr int kFeStripCols = 16;                 // stem columns per strip (tile is 8 x 16)
constexpr int kFeStripStep = 14;                 // new stem columns per strip (7 pooled columns)
constexpr int kFeRing = 64;                      // u8 ring: input rows (power of two >= 2 x 21)
constexpr int kFePairs = 20;                     // pixel pairs per staging row (38 pixels needed)
constexpr int kFeRowBytes = kFePairs * 8;        // 160 B
constexpr int kFeDepth = 3;                      // raw (f32) chunks in flight
constexpr int kFeChunkRows = 21;                 // rows per raw chunk (first tile of a strip: 21, later tiles: 16)
constexpr int kFeRawBytes = kFeChunkRows * 3 * kFePairs * 8;   // one raw chunk: [row][plane][pair] float2

struct FrontendParams {
  int n, h, w;               // input
  int ho, wo;                // stem output
  int ph, pw;                // pooled output
  int strips, tiles_per_strip;
  int cpad;                  // UMMA N == pooled channel pitch
  int in_zp;
  float inv_scale;
  int tmem_cols, acc_stride;
  uint32_t idesc;
  const float* x;            // [n][3][h][w]
  uint8_t* out;              // [n][ph][pw][cpad]
  const float* bdiv;
  const float* mult;
  const int* zwsum;
  int out_zp, out_lo;
  unsigned int* stuck_flag;
};

__global__ void __launch_bounds__(kFeThreads, 1)
frontend_fused_kernel(const __grid_constant__ CUtensorMap tmap_w, const FrontendParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  constexpr int kABlock = kTileM * 128;
  constexpr int kAStage = 2 * kABlock;
  const int b_block = p.cpad * 128;
  const int tile_row_bytes = 16 * p.cpad * 4;                   // one stem row of the int32 accumulator tile
  const int tile_bytes = 9 * tile_row_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = sA + kFeStages * kAStage;
  uint8_t* sRing = sB + 2 * b_block;                            // [kFeRing][kFeRowBytes]
  uint8_t* sRaw = sRing + kFeRing * kFeRowBytes;                // [kFeDepth][kFeRawBytes]
  uint8_t* sTile = sRaw + kFeDepth * kFeRawBytes;               // int32 [2][9][16][cpad] + one row of INT_MIN
  uint8_t* sNeutralRow = sTile + 2 * tile_bytes;
  float* s_bd = reinterpret_cast<float*>(sNeutralRow + tile_row_bytes);
  float* s_mu = s_bd + p.cpad;
  int* s_zw = reinterpret_cast<int*>(s_mu + p.cpad);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_zw + p.cpad);
  uint64_t* empty_bar = full_bar + kFeStages;
  uint64_t* tfull_bar = empty_bar + kFeStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.cpad; i += kFeThreads) {
    s_bd[i] = p.bdiv[i];
    s_mu[i] = p.mult[i];
    s_zw[i] = p.zwsum[i];
  }
  for (int i = threadIdx.x; i < tile_row_bytes / 4; i += kFeThreads) reinterpret_cast<int*>(sNeutralRow)[i] = INT_MIN;
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_w);
      for (int i = 0; i < kFeStages; ++i) {
        mbar_init(&full_bar[i], kFeBuildWarps * 32);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], kFeEpiWarps);
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
  griddep_launch_dependents();

  const int units = p.n * p.strips;
  const int tps = p.tiles_per_strip;

  if (warp == 0) {
    // ================================ weights + MMA issuer ================================
    if (elect_one()) {
      mbar_expect_tx(w_bar, static_cast<uint32_t>(2 * b_block));
      tma_load_2d(sB, &tmap_w, w_bar, 0, 0);
      tma_load_2d(sB + b_block, &tmap_w, w_bar, 128, 0);
    }
    __syncwarp();
    wait_or_die(w_bar, 0, 0x700u, p.stuck_flag);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    const uint32_t hi = smem_desc_hi(128);
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA));
    const uint32_t b_lo = smem_desc_lo(smem_u32(sB));
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      for (int t = 0; t < tps; ++t) {
        wait_or_die(&tempty_bar[acc], acc_phase ^ 1u, 0x710u | acc, p.stuck_flag);
        wait_or_die(&full_bar[stage], phase, 0x720u | stage, p.stuck_flag);
        tc_fence_after();
        fence_proxy_async_smem();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_stride);
        const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage) * (kAStage >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_i8_lohi(d_tmem, a_lo + kb * (kABlock >> 4) + 2 * k,
                           b_lo + kb * (static_cast<uint32_t>(b_block) >> 4) + 2 * k, hi, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == kFeStages) {
          stage = 0;
          phase ^= 1u;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp <= kFeBuildWarps) {
    // ================================ loaders / A builders ================================
    const int bt = (warp - 1) * 32 + lane;                  // 0 .. 255
    const int r = bt & 127;                                 // tile row
    const int kbh = bt >> 7;                                // k-block half
    const int ry = r >> 4, rx = r & 15;
    const int nseg = kbh == 0 ? 4 : 3;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint32_t chunk_off[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) chunk_off[c] = ((static_cast<uint32_t>(c) ^ sw) << 4);
    const size_t plane = static_cast<size_t>(p.h) * p.w;
    const uint32_t zp32 = static_cast<uint32_t>(p.in_zp) * 0x01010101u;
    constexpr int kBuilders = kFeBuildWarps * 32;
    const int my_units = blockIdx.x < units ? (units - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const int total_steps = my_units * tps;                 // this CTA's tiles, in order
    // This thread's (at most two) items of a raw chunk: item i <-> (local row i / 20, pixel pair i % 20).
    const int rl_a = bt / kFePairs, pair_a = bt - rl_a * kFePairs;
    const int rl_b = (bt + kBuilders) / kFePairs, pair_b = (bt + kBuilders) - rl_b * kFePairs;
    const int raw_off_a = (rl_a * 3 * kFePairs + pair_a) * 8, raw_off_b = (rl_b * 3 * kFePairs + pair_b) * 8;

    // A cursor walks this CTA's (unit, tile) sequence; divisions only happen at strip boundaries.
    // The raw chunk of a step holds the input rows its tile adds to the ring: rows -3 .. 17 for t = 0
    // (21 rows), rows 16 t + 2 .. 16 t + 17 otherwise.
    struct Cursor {
      int step, t, ubase, col0;
      const float* img;
    };
    auto cursor_unit = [&](Cursor& c, int k) {
      const int unit = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
      const int img_i = unit / p.strips;
      const int strip = unit - img_i * p.strips;
      c.col0 = 2 * (kFeStripStep * strip - 1) - 4;          // first input column held in the ring (even)
      c.img = p.x + static_cast<size_t>(img_i) * 3 * plane;
      c.ubase = (k * (16 * tps + 5)) & (kFeRing - 1);        // ring position of row -3 of this unit
    };
    auto cursor_init = [&](Cursor& c, int step) {
      c.step = step;
      const int k = step / tps;
      c.t = step - k * tps;
      if (step < total_steps) cursor_unit(c, k);
    };
    auto cursor_next = [&](Cursor& c) {
      ++c.step;
      if (++c.t == tps) {
        c.t = 0;
        if (c.step < total_steps) cursor_unit(c, c.step / tps);
      }
    };
    auto issue_item = [&](const Cursor& c, uint8_t* raw, int row, int pair, int off) {
      const int ix = c.col0 + 2 * pair;
      if (row >= 0 && row < p.h && ix >= 0 && ix < p.w) {   // out-of-image pairs become the zero point later
        const float* src = c.img + static_cast<size_t>(row) * p.w + ix;
        cp_async_8(raw + off, src);
        cp_async_8(raw + off + kFePairs * 8, src + plane);
        cp_async_8(raw + off + 2 * kFePairs * 8, src + 2 * plane);
      }
    };
    auto issue_chunk = [&](const Cursor& c) {
      if (c.step < total_steps) {
        uint8_t* raw = sRaw + (c.step % kFeDepth) * kFeRawBytes;
        const int row_lo = c.t == 0 ? -3 : 16 * c.t + 2;
        const int nrows = c.t == 0 ? 21 : 16;
        if (rl_a < nrows) issue_item(c, raw, row_lo + rl_a, pair_a, raw_off_a);
        if (rl_b < nrows) issue_item(c, raw, row_lo + rl_b, pair_b, raw_off_b);
      }
      cp_async_commit();                                     // one group per step, even when empty
    };
    auto quantize_item = [&](const Cursor& c, const uint8_t* raw, int row, int pair, int off) {
      const int ix = c.col0 + 2 * pair;
      uint2 q = make_uint2(zp32, zp32);
      if (row >= 0 && row < p.h && ix >= 0 && ix < p.w) {
        const float2 fr = *reinterpret_cast<const float2*>(raw + off);
        const float2 fg = *reinterpret_cast<const float2*>(raw + off + kFePairs * 8);
        const float2 fb = *reinterpret_cast<const float2*>(raw + off + 2 * kFePairs * 8);
        // clamp(rne(x / s) + zp, 0, 255) with the clamp done by the saturating byte pack
        auto qi = [&](float v) { return __float2int_rn(__fmul_rn(v, p.inv_scale)) + p.in_zp; };
        q.x = pack_sat_u8(qi(fg.x), qi(fr.x), pack_sat_u8(p.in_zp, qi(fb.x), 0u));
        q.y = pack_sat_u8(qi(fg.y), qi(fr.y), pack_sat_u8(p.in_zp, qi(fb.y), 0u));
      }
      *reinterpret_cast<uint2*>(sRing + ((c.ubase + row + 3) & (kFeRing - 1)) * kFeRowBytes + pair * 8) = q;
    };
    // quantise a landed chunk (this thread's own copies: same item -> thread mapping as issue_chunk)
    auto quantize_chunk = [&](const Cursor& c) {
      const uint8_t* raw = sRaw + (c.step % kFeDepth) * kFeRawBytes;
      const int row_lo = c.t == 0 ? -3 : 16 * c.t + 2;
      const int nrows = c.t == 0 ? 21 : 16;
      if (rl_a < nrows) quantize_item(c, raw, row_lo + rl_a, pair_a, raw_off_a);
      if (rl_b < nrows) quantize_item(c, raw, row_lo + rl_b, pair_b, raw_off_b);
    };

    // prologue: chunks 0 .. kFeDepth-1 in flight, chunk 0 quantised
    Cursor c_issue, c_quant, c_build;
    cursor_init(c_issue, 0);
#pragma unroll
    for (int g = 0; g < kFeDepth; ++g) {
      issue_chunk(c_issue);
      cursor_next(c_issue);
    }
    cursor_init(c_quant, 0);
    cursor_init(c_build, 0);
    if (total_steps > 0) {
      cp_async_wait<kFeDepth - 1>();
      quantize_chunk(c_quant);
    }
    cursor_next(c_quant);
    named_bar_sync(2, kBuilders);

    int stage = 0;
    uint32_t phase = 0;
    for (int g = 0; g < total_steps; ++g) {
      // (1) build this tile's A operand from the ring
      uint2 q[4][4];
      const int pos0 = c_build.ubase + 2 * (8 * c_build.t + ry) + 4 * kbh;   // ring position of filter row 4*kbh
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        if (sg < nseg) {
          const uint2* src = reinterpret_cast<const uint2*>(sRing + ((pos0 + sg) & (kFeRing - 1)) * kFeRowBytes) + rx;
#pragma unroll
          for (int j = 0; j < 4; ++j) q[sg][j] = src[j];
        }
      }
      wait_or_die(&empty_bar[stage], phase ^ 1u, 0x730u | stage, p.stuck_flag);
      uint8_t* blk = sA + stage * kAStage + kbh * kABlock + r * 128;
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        if (sg < nseg) {
          *reinterpret_cast<uint4*>(blk + chunk_off[sg * 2]) = make_uint4(q[sg][0].x, q[sg][0].y, q[sg][1].x, q[sg][1].y);
          *reinterpret_cast<uint4*>(blk + chunk_off[sg * 2 + 1]) = make_uint4(q[sg][2].x, q[sg][2].y, q[sg][3].x, q[sg][3].y);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&full_bar[stage]);
      if (++stage == kFeStages) {
        stage = 0;
        phase ^= 1u;
      }
      cursor_next(c_build);
      // (2) the chunk of tile g+1 has landed (issued kFeDepth steps ago): quantise it into the ring, then
      //     reuse the raw slot that tile g's chunk occupied for the chunk kFeDepth steps ahead.
      if (g + 1 < total_steps) {
        cp_async_wait<kFeDepth - 2>();
        quantize_chunk(c_quant);
      }
      cursor_next(c_quant);
      issue_chunk(c_issue);
      cursor_next(c_issue);
      named_bar_sync(2, kBuilders);                          // ring rows of tile g+1 visible to every builder
    }
    cp_async_wait<0>();
  } else {
    // ================================ epilogue: max-pool on the accumulators, then requantise ==========
    // requant(x) = clamp(rne((float(x) + b) * m) + zp) is non-decreasing in x (m > 0), so
    // max over the 3x3 window of requant(acc) == requant(max over the window of acc), bit for bit:
    // pooling the raw int32 accumulators first cuts the float work 4.6x.  Out-of-map positions hold
    // INT_MIN (max-pool padding is ignored, as in the reference).
    const int ew = warp - 1 - kFeBuildWarps;                // 0 .. kFeEpiWarps-1
    const int quad = warp & 3;
    const int sub = ew >> 2;                                // which of the quadrant's warps
    constexpr int kSub = kFeEpiWarps / 4;
    constexpr int kEpiThreads = kFeEpiWarps * 32;
    const int et = ew * 32 + lane;                          // epilogue thread id (pool work item)
    const int row = quad * 32 + lane;
    const int ry = row >> 4, rx = row & 15;
    const int nchunks = p.cpad >> 4;
    const int units4 = p.cpad >> 2;                         // 16-byte units (4 int32 channels) per pixel
    const int pix_bytes = p.cpad * 4;
    const int trow_bytes = 16 * pix_bytes;                  // one stem row of the int32 tile
    const int tile_i32_bytes = 9 * trow_bytes;
    const int4 kNeutral = make_int4(INT_MIN, INT_MIN, INT_MIN, INT_MIN);
    // pooling work item of this thread (fixed): 2 row pairs x 7 columns x units4 channel units
    const int my_u = et % units4;
    const int pool_pcr = et / units4;                       // hp * 7 + pc
    const int pool_hp = pool_pcr / 7, pool_pc = pool_pcr - 7 * (pool_pcr / 7);
    const bool pool_active = pool_pcr < 14;
    int pool_off[3];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int cx = 2 * pool_pc + dx;
      pool_off[dx] = cx * pix_bytes + ((my_u ^ cx) << 4);
    }
    const float4 my_bd = *reinterpret_cast<const float4*>(s_bd + 4 * my_u);
    const float4 my_mu = *reinterpret_cast<const float4*>(s_mu + 4 * my_u);
    const int4 my_zw = *reinterpret_cast<const int4*>(s_zw + 4 * my_u);
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
      const int img_i = unit / p.strips;
      const int strip = unit - img_i * p.strips;
      const int c0 = kFeStripStep * strip - 1;
      const int ox = c0 + rx;
      for (int t = 0; t < tps; ++t) {
        const int oy = 8 * t + ry;
        const bool inrange = ox >= 0 && ox < p.wo && oy < p.ho;
        uint8_t* tile = sTile + buf * tile_i32_bytes;
        wait_or_die(&tfull_bar[acc], acc_phase, 0x740u | acc, p.stuck_flag);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(acc * p.acc_stride);
        uint8_t* pix = tile + (ry + 1) * trow_bytes + rx * pix_bytes;
        for (int c = sub; c < nchunks; c += kSub) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>(c * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int4 o = make_int4(static_cast<int>(v[4 * j]), static_cast<int>(v[4 * j + 1]), static_cast<int>(v[4 * j + 2]),
                               static_cast<int>(v[4 * j + 3]));
            if (!inrange) o = kNeutral;
            *reinterpret_cast<int4*>(pix + (((4 * c + j) ^ rx) << 4)) = o;     // unit index XOR column: conflict-free
          }
        }
        // accumulator drained: hand the TMEM buffer back before pooling
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
        named_bar_sync(1, kEpiThreads);                      // whole tile of accumulators is in shared memory
        // 3x3 stride-2 max on the accumulators.  Thread <-> (pooled-row pair hp, pooled column pc, 4-channel unit
        // u), fixed for the whole kernel: its three column offsets are precomputed, the two pooled rows share
        // the horizontal maximum of the middle stem row (5 rows x 3 loads instead of 2 x 9).
        if (pool_active) {
          const uint8_t* prev_row = t == 0 ? sNeutralRow : sTile + (buf ^ 1) * tile_i32_bytes + 8 * trow_bytes;
          int4 hm[5];
#pragma unroll
          for (int rr = 0; rr < 5; ++rr) {
            const int trow = 4 * pool_hp + rr;               // tile row 0 = last row of the previous tile
            const uint8_t* rowp = trow == 0 ? prev_row : tile + trow * trow_bytes;
            const int4 a = *reinterpret_cast<const int4*>(rowp + pool_off[0]);
            const int4 b4 = *reinterpret_cast<const int4*>(rowp + pool_off[1]);
            const int4 c4 = *reinterpret_cast<const int4*>(rowp + pool_off[2]);
            hm[rr] = make_int4(max(max(a.x, b4.x), c4.x), max(max(a.y, b4.y), c4.y), max(max(a.z, b4.z), c4.z),
                               max(max(a.w, b4.w), c4.w));
          }
          const int px = 7 * strip + pool_pc;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int py = 4 * t + 2 * pool_hp + j;
            if (py < p.ph && px < p.pw) {
              const int4 &r0 = hm[2 * j], &r1 = hm[2 * j + 1], &r2 = hm[2 * j + 2];
              const int q0 = requant_i8(max(max(r0.x, r1.x), r2.x) - my_zw.x, my_bd.x, my_mu.x, p.out_zp, p.out_lo);
              const int q1 = requant_i8(max(max(r0.y, r1.y), r2.y) - my_zw.y, my_bd.y, my_mu.y, p.out_zp, p.out_lo);
              const int q2 = requant_i8(max(max(r0.z, r1.z), r2.z) - my_zw.z, my_bd.z, my_mu.z, p.out_zp, p.out_lo);
              const int q3 = requant_i8(max(max(r0.w, r1.w), r2.w) - my_zw.w, my_bd.w, my_mu.w, p.out_zp, p.out_lo);
              *reinterpret_cast<uint32_t*>(p.out + ((static_cast<size_t>(img_i) * p.ph + py) * p.pw + px) * p.cpad + 4 * my_u) =
                  static_cast<uint32_t>(q0) | (static_cast<uint32_t>(q1) << 8) | (static_cast<uint32_t>(q2) << 16) |
                  (static_cast<uint32_t>(q3) << 24);
            }
          }
        }
        // the next tile overwrites the other buffer, whose last row this tile's windows were still reading
        named_bar_sync(1, kEpiThreads);
        buf ^= 1;
      }
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
