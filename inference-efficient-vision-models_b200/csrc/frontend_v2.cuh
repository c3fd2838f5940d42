// Fused front end, second generation: [quantize_per_tensor ->] 7x7/2 stem conv (+folded BN) -> requant/bias + ReLU
// -> MaxPool2d(3,2,1) in one kernel, for 224-wide inputs and stems of <= 64 channels (INT8 and FP16).
//
// Three ideas carry it:
//
// 1. No im2col build.  The (quantised) input is written to shared memory exactly once, as 16-byte RECORDS:
//      INT8: record (Y, p) = input rows 2Y, 2Y+1 x columns 2p, 2p+1 x 3 channels (12 bytes + 4 zero-weight bytes)
//      FP16: record (y, p) = input row y x columns 2p, 2p+1 x 3 channels        (6 halves + 2 zero halves)
//    A LINE is one row (pair) of 116 records (p = -2 .. 113, out-of-image records hold the zero point / 0).
//    The stem window of output column ox is the four consecutive records ox .. ox+3 of a line, and the window
//    of ox+1 is the same run shifted by ONE record.  That is precisely the no-swizzle K-major UMMA operand
//    layout -- rows 16 bytes apart, eight-row groups 128 bytes apart (SBO), the second 16-byte K chunk at
//    LBO -- with LBO = 16 bytes: the K chunks of neighbouring rows overlap in memory, which a read-only
//    operand does not mind.  One tcgen05.mma therefore reads a line in place as a [112 pixels x 32 B] operand.
//
// 2. Channels are the M dimension, pixels the N dimension: D[m][ox], m = channel.  TMEM lane = channel,
//    TMEM column = stem column, so one epilogue thread holds a whole stem row of ONE channel: the 3-wide
//    horizontal max of the pool runs on registers, per-channel requantisation constants are scalars, and
//    nothing about the pool goes through shared memory except one 28-value exchange (below).
//
// 3. M = 128 holds TWO stem rows: rows 0..63 of the weight operand compute stem row 2T, rows 64..127 hold the
//    same filters shifted down by one row pair (two input rows) and compute stem row 2T+1 from the same
//    B operand.  A tile = pooled row T = 5 (INT8) / 9 (FP16) lines x 2 k-steps = 10 / 18 MMAs of
//    M128 x N112.  The weight operand is copied once per CTA into TENSOR MEMORY (TS-form MMA): tcgen05.mma fetches
//    shared-memory operands at ~64 B/clk/SM, so keeping the 128-row operand out of shared memory is worth 12 %.
//    Pooled row T needs stem rows 2T-1, 2T, 2T+1: thread (lane 64+c) keeps the horizontal maxima of
//    row 2T-1 in registers from the previous tile, thread (lane c) has row 2T; the two exchange half of
//    their maxima through shared memory and each finishes half of the pooled columns.
//
// requant(x) is non-decreasing in x (and so is half(relu(x + b))), hence pooling the raw accumulators first is
// bit-exact and only the pooled values are requantised.
//
// Pipeline (persistent CTA, 14 warps): TMA producer (raw f32/f16 rows, zero-filled outside the image: q(0.0) is
// the zero point) -> 4 quantiser warps (raw -> records) -> MMA warp -> 8 epilogue warps (two warpgroups, one per
// half of the pooled columns).  Work unit = (image, range of pooled rows); a unit starts with one warm-up tile
// whose only purpose is stem row 2*T0-1.
#pragma once
#include <cuda_fp16.h>

#include <climits>

#include "conv_tc.cuh"

namespace ievm {

constexpr int kF2W = 224;                        // input width this kernel is specialised for
constexpr int kF2Wo = 112;                       // stem columns == UMMA N
constexpr int kF2Pw = 56;                        // pooled columns
constexpr int kF2PxPerWg = 28;                   // pooled columns per epilogue warpgroup
constexpr int kF2Rec = 116;                      // records per line (column pairs -2 .. 113)
constexpr int kF2LinePitch = kF2Rec * 16;        // 1856 B
constexpr int kF2RawSlots = 6;
constexpr int kF2ChunkSlots = 8;                 // chunk slots in the line ring
constexpr int kF2TmemSlotCols = 128;              // accumulator slots; the weight operand sits behind them in TMEM
constexpr int kF2EpiWarps = 8;
constexpr int kF2QuantWarps = 4;
constexpr int kF2QuantThreads = 32 * kF2QuantWarps;
constexpr int kF2MmaWarp = kF2EpiWarps;
constexpr int kF2TmaWarp = kF2EpiWarps + 1;
constexpr int kF2FirstQuantWarp = kF2EpiWarps + 2;
constexpr int kF2Threads = 32 * (kF2EpiWarps + 2 + kF2QuantWarps);
constexpr int kF2LowerPx = 16;                   // of a warpgroup's 28 pooled columns, the lane-c thread finishes 16
constexpr int kF2XRows = 7;                      // uint4 rows of one exchange buffer (4 upper->lower, 3 lower->upper)

// kIn: 0 = the path's native input (f32 NCHW for INT8, f16 NCHW for FP16); 1 = decoded 8-bit images, u8 NHWC (RGB
// interleaved) mapped through a 3 x 256 lookup table that composes ToTensor, Normalize and quantize_per_tensor
// (INT8 only; SURVEY 8(f)-1: the input pipeline in front of the hot path, 4x fewer input bytes).
template <int kDtype, int kIn = 0>
struct F2Cfg {
  static constexpr int kLpc = kDtype == kDtypeI8 ? 2 : 4;             // lines per chunk (a chunk = 4 input rows)
  static constexpr int kSegs = 2 * kLpc + 1;                          // lines one tile reads
  static constexpr int kRowOff = kDtype == kDtypeI8 ? 4 : 3;          // chunk ci = input rows 4ci - kRowOff ..
  static constexpr int kARowBytes = kSegs * 64;                       // weight operand: 128 rows x kSegs x 64 B, kept
  static constexpr int kABytes = kARowBytes * 128;                    //   in TENSOR MEMORY (lane = row, 4 bytes per column)
  static constexpr int kACols = kARowBytes / 4;                       // 80 / 144 columns
  static constexpr int kSlots = (512 - kACols) / kF2TmemSlotCols;     // accumulator ring: 3 (INT8) / 2 (FP16) slots
  static constexpr int kACol0 = kSlots * kF2TmemSlotCols;
  static constexpr int kInElem = kDtype == kDtypeI8 ? 4 : 2;          // f32 / f16 input
  // raw rows start kBoxX0 columns left of the image so that the box begins on a 16-byte boundary of the row
  // (TMA faults on an 8-byte-aligned box start); they cover the 232 columns -4 .. 227 the records need
  static constexpr int kBoxX0 = kDtype == kDtypeI8 ? 4 : 8;
  static constexpr int kBoxW = 2 * kF2Rec + (kBoxX0 - 4) + (kDtype == kDtypeI8 ? 0 : 4);   // 232 / 240
  // u8 NHWC rows are described to TMA as 32-bit words (box <= 256 elements): 176 words = 16 bytes before the
  // image + 228 pixels x 3 bytes, rounded up to 16 bytes
  static constexpr int kU8RowBytes = 704;
  static constexpr int kRawTx = kIn == 1 ? 4 * kU8RowBytes : 3 * 4 * kBoxW * kInElem;   // bytes one chunk's TMA delivers
  static constexpr int kRawStride = (kRawTx + 127) / 128 * 128;
  static constexpr int kLines = kF2ChunkSlots * kLpc;                 // power of two
  static constexpr int kRecsPerChunk = kLpc * kF2Rec;
  static constexpr size_t kSmemBytes = 1024 + static_cast<size_t>(kLines) * kF2LinePitch +
                                       static_cast<size_t>(kF2RawSlots) * kRawStride +
                                       2 * 2 * kF2XRows * 64 * 16 + 64 * 8 + 32 + (kIn == 1 ? 768 : 0);
};

struct Frontend2Params {
  int n, h;                  // images, input height (multiple of 4); the width is kF2W
  int ho, ph;                // stem rows (h / 2), pooled rows (h / 4)
  int tpu;                   // cap on the pooled rows of one work unit (0 = none; diagnostics: IEVM_FRONT_TPU)
  int in_zp;
  float inv_scale;
  uint32_t idesc;
  const uint8_t* wpack;      // weight operand, row-major [128][kARowBytes] (upload_front2_weights)
  void* out;                 // [n][ph][56][64] u8 / f16
  const float* bdiv;         // i8: bias / (x_s * w_s[c]);  f16: folded bias
  const float* mult;         // i8: (x_s * w_s[c]) / out_s
  const int* zwsum;          // i8: in_zp * sum(w[c])
  const uint8_t* lut;        // u8 input mode: [3][256] quantised value of every 8-bit level per channel
  int out_zp, out_lo;
  int fast_round;            // i8: see ConvTcParams::fast_round
  int32_t* dump_acc;         // debug: corrected stem accumulators [n][ho][112][64] (f16: float bit patterns)
  unsigned int* stuck_flag;
};

// Work distribution: the n * ph pooled rows of the batch are cut into gridDim.x CONTIGUOUS ranges whose sizes differ by at
// most one row; a CTA walks its range as units = runs of rows inside one image (a unit starts with one warm-up tile for
// the stem row above it).  Round 1 cut every image into equal units and dealt them round-robin: 103.8 tiles per CTA at
// batch 256 for 96.9 rows of work; contiguous ranges need 96.9 + one warm-up per image touched (~ 2.7).
struct F2Unit {
  int img, t0, nt;
};
struct F2Walk {
  int row, row_end;
};
__device__ __forceinline__ F2Walk f2_walk_begin(const Frontend2Params& p) {
  const int total = p.n * p.ph, g = static_cast<int>(gridDim.x), b = static_cast<int>(blockIdx.x);
  const int q = total / g, r = total - q * g;
  F2Walk w;
  w.row = b * q + min(b, r);
  w.row_end = w.row + q + (b < r ? 1 : 0);
  return w;
}
__device__ __forceinline__ bool f2_next_unit(const Frontend2Params& p, F2Walk& w, F2Unit& un) {
  if (w.row >= w.row_end) return false;
  un.img = w.row / p.ph;
  un.t0 = w.row - un.img * p.ph;
  un.nt = min(p.ph - un.t0, w.row_end - w.row);
  if (p.tpu > 0) un.nt = min(un.nt, p.tpu);
  w.row += un.nt;
  return true;
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
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

// Accumulator words are int32 (INT8) or float bit patterns (FP16); "lowest" is the pool's padding value.
template <int kDtype>
__device__ __forceinline__ uint32_t f2_lowest() {
  return kDtype == kDtypeI8 ? 0x80000000u : 0xFF800000u;   // INT_MIN / -inf
}
template <int kDtype>
__device__ __forceinline__ uint32_t f2_max(uint32_t a, uint32_t b) {
  if (kDtype == kDtypeI8) return static_cast<uint32_t>(max(static_cast<int>(a), static_cast<int>(b)));
  return __float_as_uint(fmaxf(__uint_as_float(a), __uint_as_float(b)));
}
template <int kDtype>
__device__ __forceinline__ uint32_t f2_max3(uint32_t a, uint32_t b, uint32_t c) {
  if (kDtype == kDtypeI8)
    return static_cast<uint32_t>(__vimax3_s32(static_cast<int>(a), static_cast<int>(b), static_cast<int>(c)));
  return __float_as_uint(fmaxf(fmaxf(__uint_as_float(a), __uint_as_float(b)), __uint_as_float(c)));
}

// Epilogue of one warpgroup (kWg = 0: pooled columns 0..27, kWg = 1: 28..55).
// kDump: the debug instantiation that also writes the raw stem accumulators (ievm_debug_frontend); the product's epilogue
// loop carries none of that code (64 byte-stride stores and their addressing per tile, skipped by a branch, were a quarter
// of the loop's instruction footprint).
template <int kDtype, int kWg, int kSlots, bool kFast, bool kDump>
__device__ __forceinline__ void f2_epilogue(const Frontend2Params& p, uint32_t tmem_base, uint64_t* tmem_full,
                                            uint64_t* tmem_empty, uint4* s_x) {
  constexpr int kOff = kWg == 0 ? -1 : 7;        // register index of a pooled column's first stem column: 2k + kOff
  constexpr int kCol0 = kWg == 0 ? 0 : 48;       // first TMEM column this warpgroup loads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3;
  const bool upper = quad >= 2;                  // lanes 64..127: stem row 2T+1
  const int c = (quad & 1) * 32 + lane;          // channel
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kCol0;
  const uint32_t kLow = f2_lowest<kDtype>();
  const float bd = p.bdiv[c];
  const float mu = kDtype == kDtypeI8 ? p.mult[c] : 0.f;
  const int zw = kDtype == kDtypeI8 ? p.zwsum[c] : 0;
  const int out_elem = kDtype == kDtypeI8 ? 1 : 2;
  uint4* xs = s_x + kWg * 2 * kF2XRows * 64;

#ifdef IEVM_EXP_TIMING
  long long f2_tm_wait = 0;
  const long long f2_tm_t0 = clock64();
#endif
  uint32_t prev[kF2PxPerWg];                     // upper threads: horizontal maxima of the previous tile's row
#pragma unroll
  for (int k = 0; k < kF2PxPerWg; ++k) prev[k] = kLow;
  int t = 0, ts_next = 0;
  uint32_t ph_next = 0;
  F2Walk walk = f2_walk_begin(p);
  F2Unit un;
  while (f2_next_unit(p, walk, un)) {
    for (int i = 0; i <= un.nt; ++i, ++t) {
      const int T = un.t0 - 1 + i;               // pooled row of this tile (the warm-up tile i == 0 only feeds `prev`)
      const int ts = ts_next;
      const uint32_t ph = ph_next;
      if (++ts_next == kSlots) {
        ts_next = 0;
        ph_next ^= 1u;
      }
#ifdef IEVM_EXP_TIMING
      const long long tw0 = clock64();
#endif
      wait_or_die_ool(&tmem_full[ts], ph, 0x840u | ts, p.stuck_flag);
#ifdef IEVM_EXP_TIMING
      f2_tm_wait += clock64() - tw0;
#endif
      tc_fence_after();
      const uint32_t taddr = lane_addr + static_cast<uint32_t>(ts * kF2TmemSlotCols);
      const bool real = i > 0;
      uint32_t hm[kF2PxPerWg];
      if (real || upper) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        if (kDump && p.dump_acc != nullptr && T >= 0) {
          const int oy = 2 * T + (upper ? 1 : 0);
          int32_t* d = p.dump_acc + ((static_cast<size_t>(un.img) * p.ho + oy) * kF2Wo + kCol0) * 64 + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) d[j * 64] = static_cast<int32_t>(v[j]) - zw;
        }
#pragma unroll
        for (int k = 0; k < kF2PxPerWg; ++k) {
          const int i0 = 2 * k + kOff;
          if (i0 + 2 <= 31) hm[k] = f2_max3<kDtype>(i0 < 0 ? kLow : v[i0 < 0 ? 0 : i0], v[i0 + 1], v[i0 + 2]);
        }
        const uint32_t c30 = v[30], c31 = v[31];
        tmem_ld_32x32b_x32(taddr + 32, v);
        tmem_ld_wait();
        if (kDump && p.dump_acc != nullptr && T >= 0) {
          const int oy = 2 * T + (upper ? 1 : 0);
          int32_t* d = p.dump_acc + ((static_cast<size_t>(un.img) * p.ho + oy) * kF2Wo + kCol0 + 32) * 64 + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) d[j * 64] = static_cast<int32_t>(v[j]) - zw;
        }
#pragma unroll
        for (int k = 0; k < kF2PxPerWg; ++k) {
          const int i0 = 2 * k + kOff;
          if (i0 + 2 >= 32) {
            const uint32_t a = i0 >= 32 ? v[i0 >= 32 ? i0 - 32 : 0] : (i0 == 30 ? c30 : c31);
            const uint32_t b = i0 + 1 >= 32 ? v[i0 + 1 >= 32 ? i0 + 1 - 32 : 0] : c31;
            hm[k] = f2_max3<kDtype>(a, b, v[i0 + 2 - 32]);
          }
        }
      }
      tc_fence_before();
      uint4* xb = xs + (t & 1) * kF2XRows * 64;
      if (upper) {
        // stem row 2T+1 joins the row 2T-1 kept from the previous tile; rows above the image do not exist
#pragma unroll
        for (int k = 0; k < kF2PxPerWg; ++k) {
          const uint32_t cur = T < 0 ? kLow : hm[k];
          hm[k] = f2_max<kDtype>(prev[k], cur);
          prev[k] = cur;
        }
        if (real) {
#pragma unroll
          for (int j = 0; j < 4; ++j) xb[j * 64 + c] = make_uint4(hm[4 * j], hm[4 * j + 1], hm[4 * j + 2], hm[4 * j + 3]);
        }
      } else if (real) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
          xb[(4 + j) * 64 + c] = make_uint4(hm[kF2LowerPx + 4 * j], hm[kF2LowerPx + 4 * j + 1], hm[kF2LowerPx + 4 * j + 2],
                                            hm[kF2LowerPx + 4 * j + 3]);
      }
      named_bar_sync(1 + kWg, 128);
      if ((threadIdx.x & 127) == 0) mbar_arrive(&tmem_empty[ts]);       // this warpgroup's TMEM reads of the slot are done
      if (!real) continue;
      uint8_t* orow = static_cast<uint8_t*>(p.out) +
                      (((static_cast<size_t>(un.img) * p.ph + T) * kF2Pw + kWg * kF2PxPerWg) * 64 + c) * out_elem;
      auto finish = [&](int k, uint32_t m) {
        if (kDtype == kDtypeI8) {
          if (kFast) {
            // magic-number rounding (|value| < 2^21 for every possible input, checked at engine creation): FADD instead
            // of the 8-cycle F2I; the zero-point add and BOTH clamps are one VIADDMNMX.RELU -- the host enables this
            // instantiation only when the lower clamp is 0 (a ReLU stem whose output zero point is 0: every fbgemm qconfig)
            const float v = __fmul_rn(__fadd_rn(__int2float_rn(static_cast<int>(m) - zw), bd), mu);
            orow[k * 64] = static_cast<uint8_t>(
                __viaddmin_s32_relu(__float_as_int(__fadd_rn(v, kRoundMagic)), p.out_zp - kRoundMagicBits, 255));
          } else {
            orow[k * 64] = static_cast<uint8_t>(requant_i8(static_cast<int>(m) - zw, bd, mu, p.out_zp, p.out_lo));
          }
        } else {
          reinterpret_cast<__half*>(orow)[k * 64] = __float2half_rn(fmaxf(__uint_as_float(m) + bd, 0.f));
        }
      };
      if (!upper) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 o = xb[j * 64 + c];
          finish(4 * j + 0, f2_max<kDtype>(hm[4 * j + 0], o.x));
          finish(4 * j + 1, f2_max<kDtype>(hm[4 * j + 1], o.y));
          finish(4 * j + 2, f2_max<kDtype>(hm[4 * j + 2], o.z));
          finish(4 * j + 3, f2_max<kDtype>(hm[4 * j + 3], o.w));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint4 o = xb[(4 + j) * 64 + c];
          finish(kF2LowerPx + 4 * j + 0, f2_max<kDtype>(hm[kF2LowerPx + 4 * j + 0], o.x));
          finish(kF2LowerPx + 4 * j + 1, f2_max<kDtype>(hm[kF2LowerPx + 4 * j + 1], o.y));
          finish(kF2LowerPx + 4 * j + 2, f2_max<kDtype>(hm[kF2LowerPx + 4 * j + 2], o.z));
          finish(kF2LowerPx + 4 * j + 3, f2_max<kDtype>(hm[kF2LowerPx + 4 * j + 3], o.w));
        }
      }
    }
  }
#ifdef IEVM_EXP_TIMING
  if (threadIdx.x == 0 && blockIdx.x < kExpCtas) {
    unsigned long long* o = g_exp_timing + (static_cast<size_t>(31) * kExpCtas + blockIdx.x) * kExpWords;
    o[6] = clock64() - f2_tm_t0;
    o[7] = f2_tm_wait;
  }
#endif
}

template <int kDtype, int kIn, bool kFast = false, bool kDump = false>
__global__ void __launch_bounds__(kF2Threads, 1)
frontend2_kernel(const __grid_constant__ CUtensorMap tmap_x, const Frontend2Params p) {
  using Cfg = F2Cfg<kDtype, kIn>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sLines = smem;
  uint8_t* sRaw = sLines + Cfg::kLines * kF2LinePitch;
  uint4* sX = reinterpret_cast<uint4*>(sRaw + kF2RawSlots * Cfg::kRawStride);
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(sX + 2 * 2 * kF2XRows * 64);
  uint64_t* raw_empty = raw_full + kF2RawSlots;
  uint64_t* line_full = raw_empty + kF2RawSlots;
  uint64_t* line_empty = line_full + kF2ChunkSlots;
  uint64_t* tmem_full = line_empty + kF2ChunkSlots;
  uint64_t* tmem_empty = tmem_full + Cfg::kSlots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + Cfg::kSlots);
  uint8_t* sLut = reinterpret_cast<uint8_t*>(tmem_slot + 4);      // u8 input mode: [3][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (kIn == 1)
    for (int i = threadIdx.x; i < 768 / 4; i += kF2Threads)
      reinterpret_cast<uint32_t*>(sLut)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.lut) + i);
  if (warp == kF2TmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < kF2RawSlots; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], kF2QuantThreads);
    }
    for (int i = 0; i < kF2ChunkSlots; ++i) {
      mbar_init(&line_full[i], kF2QuantThreads);
      mbar_init(&line_empty[i], 1);
    }
    for (int i = 0; i < Cfg::kSlots; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2);              // one arrival per epilogue warpgroup
    }
    fence_barrier_init();
  }
  if (warp == kF2MmaWarp) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  griddep_launch_dependents();

  // weight operand -> tensor memory, once per CTA: thread (quadrant q, lane l) of warps 0..3 owns row 32 q + l
  if (warp < 4) {
    const uint4* src = reinterpret_cast<const uint4*>(p.wpack + static_cast<size_t>(warp * 32 + lane) * Cfg::kARowBytes);
    const uint32_t a_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + Cfg::kACol0;
#pragma unroll 1
    for (int c = 0; c < Cfg::kACols / 16; ++c) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 q = __ldg(src + 4 * c + j);
        w[4 * j] = q.x;
        w[4 * j + 1] = q.y;
        w[4 * j + 2] = q.z;
        w[4 * j + 3] = q.w;
      }
      tmem_st_32x32b_x16(a_addr + static_cast<uint32_t>(16 * c), w);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();


  if (warp < kF2EpiWarps) {
    if (warp < 4) f2_epilogue<kDtype, 0, Cfg::kSlots, kFast, kDump>(p, tmem_base, tmem_full, tmem_empty, sX);
    else f2_epilogue<kDtype, 1, Cfg::kSlots, kFast, kDump>(p, tmem_base, tmem_full, tmem_empty, sX);
  } else if (warp == kF2MmaWarp) {
    // ================================ MMA issuer ================================
    const uint32_t hi = (128u >> 4) | (1u << 14);                                  // SBO = 128 B, version 1, no swizzle
    const uint32_t a_tmem0 = tmem_base + Cfg::kACol0;                                  // 8 columns (32 bytes of K) per k-step
    const uint32_t b_lo0 = ((smem_u32(sLines) & 0x3FFFFu) >> 4) | (1u << 16);          // LBO = 16 B (next record)
    int q0 = 0, t = 0, ts = 0;
    uint32_t ph = 0;
#ifdef IEVM_EXP_TIMING
    long long f2_wait = 0, f2_wait_line = 0;
    const long long f2_t0 = clock64();
#endif
    F2Walk walk = f2_walk_begin(p);
    F2Unit un;
    while (f2_next_unit(p, walk, un)) {
      for (int i = 0; i <= un.nt; ++i, ++t) {
        const int qn = q0 + i + 2;               // newest chunk this tile reads (chunks complete in order)
        {
          // both barrier tests in ONE round trip: every mbarrier test by the issuing thread is ~150 cycles of idle tensor
          // pipe (conv_tc.cuh), and a tile is only ten instructions long
          uint64_t* b0 = &line_full[qn % kF2ChunkSlots];
          uint64_t* b1 = &tmem_empty[ts];
          const uint32_t p0 = (qn / kF2ChunkSlots) & 1u, p1 = ph ^ 1u;
#ifdef IEVM_EXP_TIMING
          const long long tw0 = clock64();
#endif
          if (!mbar_try_wait5(b0, p0, b1, p1, b1, p1, b1, p1, b1, p1)) {
#ifdef IEVM_EXP_TIMING
            const long long tw1 = clock64();
            wait_or_die_ool(b0, p0, 0x820u | (qn % kF2ChunkSlots), p.stuck_flag);
            f2_wait_line += clock64() - tw1;
#else
            wait_or_die_ool(b0, p0, 0x820u | (qn % kF2ChunkSlots), p.stuck_flag);
#endif
            wait_or_die_ool(b1, p1, 0x830u | ts, p.stuck_flag);
          }
#ifdef IEVM_EXP_TIMING
          f2_wait += clock64() - tw0;
#endif
        }
        tc_fence_after();
        fence_proxy_async_smem();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ts * kF2TmemSlotCols);
        const int line0 = (q0 + i) * Cfg::kLpc;
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < Cfg::kSegs; ++s) {
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>((line0 + s) & (Cfg::kLines - 1)) * (kF2LinePitch >> 4);
#pragma unroll
            for (int jh = 0; jh < 2; ++jh)
              umma_ts<kDtype>(d_tmem, a_tmem0 + static_cast<uint32_t>(2 * s + jh) * 8u, b_lo + 2u * jh, hi, p.idesc,
                              (s | jh) != 0 ? 1u : 0u);
          }
          umma_commit(&tmem_full[ts]);
          umma_commit(&line_empty[(q0 + i) % kF2ChunkSlots]);      // tile i is the last reader of chunk i
          if (i == un.nt) {
            umma_commit(&line_empty[(q0 + i + 1) % kF2ChunkSlots]);
            umma_commit(&line_empty[(q0 + i + 2) % kF2ChunkSlots]);
          }
        }
        __syncwarp();
        if (++ts == Cfg::kSlots) {
          ts = 0;
          ph ^= 1u;
        }
      }
      q0 += un.nt + 3;
    }
#ifdef IEVM_EXP_TIMING
    if (lane == 0 && blockIdx.x < kExpCtas) {
      unsigned long long* o = g_exp_timing + (static_cast<size_t>(31) * kExpCtas + blockIdx.x) * kExpWords;
      o[0] = clock64() - f2_t0;
      o[1] = f2_wait_line;
      o[2] = f2_wait;
      o[8] = t;
      o[9] = clock64() - f2_t0;
      o[10] = 1;
    }
#endif
  } else if (warp == kF2TmaWarp) {
    // ================================ TMA producer: raw input rows ================================
    int q = 0;
#ifdef IEVM_EXP_TIMING
    long long f2_wait = 0;
    const long long f2_t0 = clock64();
#endif
    F2Walk walk = f2_walk_begin(p);
    F2Unit un;
    while (f2_next_unit(p, walk, un)) {
      for (int j = 0; j < un.nt + 3; ++j, ++q) {
        const int slot = q % kF2RawSlots;
#ifdef IEVM_EXP_TIMING
        const long long tw0 = clock64();
#endif
        wait_or_die_ool(&raw_empty[slot], ((q / kF2RawSlots) & 1u) ^ 1u, 0x800u | slot, p.stuck_flag);
#ifdef IEVM_EXP_TIMING
        f2_wait += clock64() - tw0;
#endif
        if (elect_one()) {
          mbar_expect_tx(&raw_full[slot], static_cast<uint32_t>(Cfg::kRawTx));
          if (kIn == 1)     // (32-bit words of a row, row, image): 4 words = 16 bytes before the first pixel
            tma_load_3d(sRaw + slot * Cfg::kRawStride, &tmap_x, &raw_full[slot], -4, 4 * (un.t0 - 1 + j) - Cfg::kRowOff,
                        un.img);
          else
            tma_load_3d(sRaw + slot * Cfg::kRawStride, &tmap_x, &raw_full[slot], -Cfg::kBoxX0, 4 * (un.t0 - 1 + j) - Cfg::kRowOff,
                        3 * un.img);
        }
        __syncwarp();
      }
    }
#ifdef IEVM_EXP_TIMING
    if (lane == 0 && blockIdx.x < kExpCtas) {
      unsigned long long* o = g_exp_timing + (static_cast<size_t>(31) * kExpCtas + blockIdx.x) * kExpWords;
      o[4] = clock64() - f2_t0;
      o[5] = f2_wait;
    }
#endif
  } else {
    // ================================ quantisers: raw rows -> records ================================
    const int qt = threadIdx.x - 32 * kF2FirstQuantWarp;
    int q = 0;
#ifdef IEVM_EXP_TIMING
    long long f2_wait_raw = 0, f2_wait_line = 0;
    const long long f2_t0 = clock64();
#endif
    F2Walk walk = f2_walk_begin(p);
    F2Unit un;
    while (f2_next_unit(p, walk, un)) {
      for (int j = 0; j < un.nt + 3; ++j, ++q) {
        const int slot = q % kF2RawSlots, cs = q % kF2ChunkSlots;
#ifdef IEVM_EXP_TIMING
        const long long tw0 = clock64();
#endif
        wait_or_die_ool(&raw_full[slot], (q / kF2RawSlots) & 1u, 0x810u | slot, p.stuck_flag);
#ifdef IEVM_EXP_TIMING
        const long long tw1 = clock64();
        f2_wait_raw += tw1 - tw0;
#endif
        wait_or_die_ool(&line_empty[cs], ((q / kF2ChunkSlots) & 1u) ^ 1u, 0x818u | cs, p.stuck_flag);
#ifdef IEVM_EXP_TIMING
        f2_wait_line += clock64() - tw1;
#endif
        const uint8_t* raw = sRaw + slot * Cfg::kRawStride;
        uint8_t* lines = sLines + cs * Cfg::kLpc * kF2LinePitch;
        if (kIn == 1) {
          // u8 NHWC rows: pixel column x, channel c at byte 16 + 3x + c of the raw row.  Record (l, ri) covers columns
          // 2p, 2p+1 (p = ri - 2): six consecutive bytes per row, already in record order (cp*3 + c); each goes
          // through the lookup table.  Out-of-image records hold the zero point.
          const int row0 = 4 * (un.t0 - 1 + j) - Cfg::kRowOff;
          const uint32_t zp4 = static_cast<uint32_t>(p.in_zp) * 0x01010101u;
#pragma unroll
          for (int it = 0; it < (Cfg::kRecsPerChunk + kF2QuantThreads - 1) / kF2QuantThreads; ++it) {
            const int idx = qt + it * kF2QuantThreads;
            if (idx < Cfg::kRecsPerChunk) {
              const int l = idx >= kF2Rec ? 1 : 0;
              const int ri = idx - l * kF2Rec;
              const int pcol = ri - 2;
              const int row = row0 + 2 * l;
              uint4 o = make_uint4(zp4, zp4, zp4, 0u);
              if (pcol >= 0 && pcol < kF2W / 2 && row >= 0 && row < p.h) {
                uint32_t b[12];
#pragma unroll
                for (int rp = 0; rp < 2; ++rp) {
                  const uint16_t* src = reinterpret_cast<const uint16_t*>(raw + (2 * l + rp) * Cfg::kU8RowBytes + 16 + 6 * pcol);
                  const uint32_t h0 = src[0], h1 = src[1], h2 = src[2];      // bytes (c0 c1) (c2 c0') (c1' c2')
                  b[rp * 6 + 0] = sLut[h0 & 0xffu];
                  b[rp * 6 + 1] = sLut[256 + (h0 >> 8)];
                  b[rp * 6 + 2] = sLut[512 + (h1 & 0xffu)];
                  b[rp * 6 + 3] = sLut[h1 >> 8];
                  b[rp * 6 + 4] = sLut[256 + (h2 & 0xffu)];
                  b[rp * 6 + 5] = sLut[512 + (h2 >> 8)];
                }
                o.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
                o.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
                o.z = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
              }
              *reinterpret_cast<uint4*>(lines + l * kF2LinePitch + ri * 16) = o;
            }
          }
        } else if (kDtype == kDtypeI8) {
          // record (l, ri): rows 2l, 2l+1 of the chunk, raw columns 2ri, 2ri+1; byte = rp*6 + cp*3 + c
          const float* rf = reinterpret_cast<const float*>(raw);
#pragma unroll
          for (int it = 0; it < (Cfg::kRecsPerChunk + kF2QuantThreads - 1) / kF2QuantThreads; ++it) {
            const int idx = qt + it * kF2QuantThreads;
            if (idx < Cfg::kRecsPerChunk) {
              const int l = idx >= kF2Rec ? 1 : 0;
              const int ri = idx - l * kF2Rec;
              int qv[3][2][2];                   // [c][rp][cp]
#pragma unroll
              for (int ch = 0; ch < 3; ++ch)
#pragma unroll
                for (int rp = 0; rp < 2; ++rp) {
                  const float2 f = *reinterpret_cast<const float2*>(rf + (ch * 4 + 2 * l + rp) * Cfg::kBoxW + 2 * ri);
                  qv[ch][rp][0] = __float2int_rn(__fmul_rn(f.x, p.inv_scale)) + p.in_zp;
                  qv[ch][rp][1] = __float2int_rn(__fmul_rn(f.y, p.inv_scale)) + p.in_zp;
                }
              uint4 o;
              o.x = pack4_sat_u8(qv[0][0][0], qv[1][0][0], qv[2][0][0], qv[0][0][1]);
              o.y = pack4_sat_u8(qv[1][0][1], qv[2][0][1], qv[0][1][0], qv[1][1][0]);
              o.z = pack4_sat_u8(qv[2][1][0], qv[0][1][1], qv[1][1][1], qv[2][1][1]);
              o.w = 0u;
              *reinterpret_cast<uint4*>(lines + l * kF2LinePitch + ri * 16) = o;
            }
          }
        } else {
          // record (l, ri): row l of the chunk, raw columns 2ri, 2ri+1; half e = cp*3 + c
          const uint32_t* rh = reinterpret_cast<const uint32_t*>(raw);
#pragma unroll
          for (int it = 0; it < (Cfg::kRecsPerChunk + kF2QuantThreads - 1) / kF2QuantThreads; ++it) {
            const int idx = qt + it * kF2QuantThreads;
            if (idx < Cfg::kRecsPerChunk) {
              const int l = idx / kF2Rec;
              const int ri = idx - l * kF2Rec;
              constexpr int kSkip = (Cfg::kBoxX0 - 4) / 2;      // column pairs between the box start and record 0
              const uint32_t p0 = rh[((0 * 4 + l) * Cfg::kBoxW >> 1) + ri + kSkip];
              const uint32_t p1 = rh[((1 * 4 + l) * Cfg::kBoxW >> 1) + ri + kSkip];
              const uint32_t p2 = rh[((2 * 4 + l) * Cfg::kBoxW >> 1) + ri + kSkip];
              uint4 o;
              o.x = __byte_perm(p0, p1, 0x5410);   // (cp0,c0) (cp0,c1)
              o.y = __byte_perm(p2, p0, 0x7610);   // (cp0,c2) (cp1,c0)
              o.z = __byte_perm(p1, p2, 0x7632);   // (cp1,c1) (cp1,c2)
              o.w = 0u;
              *reinterpret_cast<uint4*>(lines + l * kF2LinePitch + ri * 16) = o;
            }
          }
        }
        mbar_arrive(&raw_empty[slot]);           // after the reads above (release)
        fence_proxy_async_smem();                // records visible to the tensor core's async-proxy reads
        mbar_arrive(&line_full[cs]);
      }
    }
#ifdef IEVM_EXP_TIMING
    if (qt == 0 && blockIdx.x < kExpCtas) {
      unsigned long long* o = g_exp_timing + (static_cast<size_t>(31) * kExpCtas + blockIdx.x) * kExpWords;
      o[11] = clock64() - f2_t0;
      o[12] = f2_wait_raw;
      o[13] = f2_wait_line;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kF2MmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace ievm
