// comb_kernels.cuh -- field-difference / combing metric, the full-frame streaming pass (HBM-read bound).
//
// Spec (integer, order independent => bit-exact by construction):
//   comb(y,x) = | p[y-2] + 4 p[y] + p[y+2] - 3 (p[y-1] + p[y+1]) |            2 <= y < H-2
//   shima[f] += comb >= thS ; lshima[f] += comb >= thL                         f = y & 1 (top / bottom field)
//   move[f]  += | p_n[y][x] - p_{n-1}[y][x] | >= thM                           every row; prev(first) = itself
//   per frame: int32[12] = [Y, C(=U+V)] x [top, bottom] x [move, shima, lshima]
//
// Design for sm_100a:
//   * work unit = (plane tile of 128 px x 8R rows (R = 15..17 picked so that tiles cover the plane without partial
//     bands: 1080 = 8 x 136 - 8), run of consecutive frames).  A CTA streams the tile of frame
//     n, n+1, ... through a 3-stage shared-memory ring filled by TMA (cp.async.bulk.tensor.3d, one instruction
//     per tile incl. the +-2 row halo, out-of-frame rows/cols zero-filled by the TMA unit).  The tile of frame
//     n-1 is still in the ring when frame n is processed, so the inter-frame difference costs no second HBM read:
//     every frame byte is fetched from HBM once (plus 4/136 halo rows).
//   * each thread owns an 8-pixel-wide column strip and walks R rows with a 5-row sliding window held in
//     registers as fp16x2.  Bytes zero-extended into 16-bit lanes ARE exact fp16 values (subnormals, k*2^-24),
//     so PRMT is the whole u8->f16 conversion, the 5-tap response (|.| <= 1530 < 2048) is exact in fp16, and
//     HSET2.GE with |x| does the threshold on two pixels per instruction.  The inter-frame difference runs
//     4 pixels per instruction (VABSDIFF4 + SWAR compare + IDP.4A count).  Tensor cores are not used: nothing
//     here is a contraction.
//   * counters: per-thread pair-coded mask sums -> REDUX per warp -> plain stores to shared -> six writer threads
//     decode and issue <= 6 global RED per tile-frame.
//   * static equal-share partition of all (tile, frame) pairs over 148 x occupancy CTAs (host side): no tail.
#pragma once
#include <cuda_fp16.h>
#include "amtk_internal.h"
#include "tma_utils.cuh"

namespace amtk {

constexpr int kCombTW = 128;            // tile width in bytes (= pixels for u8)

// Compile-time shape of one kernel variant: R rows per run (tile height 8R), STRIP pixels per thread-row,
// STAGES ring slots.
// SYNC selects the end-of-step synchronisation: 0 = one block barrier per tile-frame; 1 = "release" mode: every
// warp arrives on a per-slot mbarrier when it is done with the previous-frame slot and runs ahead (by at most one
// step), only warp 0 waits for all arrivals, flushes the counters and refills the slot (measured 9 % slower; tune-only).
template <int R_, int STRIP_, int STAGES_, int SYNC_ = 0, int RUNS_ = 8>
struct CombCfg {
  static constexpr int R = R_, STRIP = STRIP_, STAGES = STAGES_, SYNC = SYNC_, RUNS = RUNS_;   // RUNS vertical runs per tile
  static constexpr int TH = RUNS * R;                      // output rows per tile
  static constexpr int BOXH = TH + 4;                      // with +-2 halo rows
  static constexpr int STAGE_BYTES = kCombTW * BOXH;
  static constexpr int THREADS = (kCombTW / STRIP) * RUNS;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 128;  // + alignment slack
  static constexpr int NQ = STRIP / 2;                     // half2 per thread-row
};

struct CombPlane {
  int W, H;                 // plane size in samples
  int tilesX, tilesY;
  int tile0;                // first tile id of this plane
  int cls;                  // 0 = Y, 1 = C
  unsigned thM, thS, thL;   // u8: thM = (0x80-thM)*0x01010101, thS/thL = fp16x2 bit patterns; u16: (0x8000-thM)*0x00010001, fp32 bits
};

struct CombSegment {        // a run of frames of one tile, processed by one CTA
  int tile;                 // global tile id
  int fbegin, fend;         // frame indices inside the device window
};

struct CombArgs {
  CUtensorMap map[3];       // 3-D (x, y, frame) u8 views of the Y, U, V planes of the device window (box 128 x BOXH)
  CUtensorMap map_half[2];  // U and V again with a 64-byte wide box: the two 64-pixel remainders share one tile
  CombPlane plane[4];       // [3] = pseudo plane of the merged U|V remainder tiles (tilesX = 1), when used
  int half_x;               // x of the remainder columns in the chroma planes
  const CombSegment* segs;
  const int* seg_start;     // [gridDim.x + 1]
  int* counts;              // [nframes_out][12]
  int out_frame0;           // counts row = frame - out_frame0
};

// ---------------------------------------------------------------------------------------------------------
// per-thread tile-frame body
// ---------------------------------------------------------------------------------------------------------
template <int NQ> struct HRow { __half2 v[NQ]; };    // STRIP pixels of one row as fp16x2

template <int STRIP> struct RawRow;
template <> struct RawRow<8> {
  uint2 r;
  __device__ __forceinline__ void load(const uint8_t* p) { r = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ uint32_t word(int i) const { return i ? r.y : r.x; }
};
template <> struct RawRow<4> {
  uint32_t r;
  __device__ __forceinline__ void load(const uint8_t* p) { r = *reinterpret_cast<const uint32_t*>(p); }
  __device__ __forceinline__ uint32_t word(int) const { return r; }
};

template <int STRIP>
__device__ __forceinline__ HRow<STRIP / 2> bytes_to_half(const RawRow<STRIP>& raw) {
  HRow<STRIP / 2> r;
#pragma unroll
  for (int i = 0; i < STRIP / 4; ++i) {
    const uint32_t w = raw.word(i);
    uint32_t a = __byte_perm(w, 0, 0x4140), b = __byte_perm(w, 0, 0x4342);   // zero-extended bytes = exact fp16 subnormals
    r.v[2 * i] = *reinterpret_cast<__half2*>(&a); r.v[2 * i + 1] = *reinterpret_cast<__half2*>(&b);
  }
  return r;
}

// bytes of d that are >= th (1 <= th <= 128): bit 7 of each byte of the result.  kM = (0x80 - th) * 0x01010101.
__device__ __forceinline__ uint32_t bytes_ge(uint32_t d, uint32_t kM) {
  return (((d & 0x7F7F7F7Fu) + kM) | d) & 0x80808080u;
}

// Raw per-thread counters of one tile-frame.  slot = row parity RELATIVE to the thread's first row (j & 1); the
// caller maps slots to fields.  S/L: 8-bit path = pair-coded mask sums (decode_pair after any number of integer
// additions), 16-bit path = plain counts.  M: 128 per hit.
struct RawCounts { uint32_t S[2], L[2], M[2]; };

template <typename Cfg, bool EDGE, int PITCH>
__device__ __forceinline__ RawCounts comb_tile_rows(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev,
                                                    const uint32_t* __restrict__ th_rows /* EDGE: [2][R] thresholds of this run */,
                                                    uint32_t kM, uint32_t thS_bits, uint32_t thL_bits) {
  // cur/prev point at this thread's strip in smem row (run*R) of the box, i.e. global row y_first-2.
  constexpr int R = Cfg::R, NQ = Cfg::NQ, STRIP = Cfg::STRIP;
  const __half2 thS = *reinterpret_cast<const __half2*>(&thS_bits);
  const __half2 thL = *reinterpret_cast<const __half2*>(&thL_bits);
  const __half2 k4 = __float2half2_rn(4.0f), km3 = __float2half2_rn(-3.0f);
  RawCounts c = { { 0u, 0u }, { 0u, 0u }, { 0u, 0u } };

  RawRow<STRIP> raw_c, raw_n, rtmp;
  rtmp.load(cur);                 HRow<NQ> h0 = bytes_to_half<STRIP>(rtmp);
  rtmp.load(cur + PITCH);       HRow<NQ> h1 = bytes_to_half<STRIP>(rtmp);
  raw_c.load(cur + 2 * PITCH);  HRow<NQ> h2 = bytes_to_half<STRIP>(raw_c);     // centre row of j=0
  raw_n.load(cur + 3 * PITCH);  HRow<NQ> h3 = bytes_to_half<STRIP>(raw_n);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    RawRow<STRIP> raw_nn; raw_nn.load(cur + (j + 4) * PITCH);
    const HRow<NQ> h4 = bytes_to_half<STRIP>(raw_nn);
    RawRow<STRIP> pv; pv.load(prev + (j + 2) * PITCH);
    const int f = j & 1;          // accumulator slot
    // rows y < 2 and y >= H-2 have no comb response (spec): edge tiles read per-row thresholds (infinite there)
    // from a small shared table built once per segment -- two LDS instead of compare/select on the ALU pipe
    __half2 tS = thS, tL = thL;
    if (EDGE) {
      tS = *reinterpret_cast<const __half2*>(&th_rows[j]);
      tL = *reinterpret_cast<const __half2*>(&th_rows[Cfg::R + j]);
    }
    // inter-frame difference of the centre row first (4 pixels per op): putting this ALU-pipe work ahead of the
    // FMA-pipe stencil of the same row measured 2 % faster than the reverse order (tools/tune_comb.py history)
#pragma unroll
    for (int i = 0; i < STRIP / 4; ++i)
      c.M[f] = __dp4a(bytes_ge(__vabsdiffu4(raw_c.word(i), pv.word(i)), kM), 0x01010101u, c.M[f]);
#pragma unroll
    for (int q = 0; q < NQ; q += 2) {
      uint32_t mS[2], mL[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        __half2 t = __hadd2(h0.v[q + e], h4.v[q + e]);
        t = __hfma2(k4, h2.v[q + e], t);
        const __half2 u = __hadd2(h1.v[q + e], h3.v[q + e]);
        const __half2 r = __habs2(__hfma2(km3, u, t));
        mS[e] = __hge2_mask(r, tS);
        mL[e] = __hge2_mask(r, tL);
      }
      // 0xFFFF-per-lane masks are subtracted as plain 32-bit integers; decode_pair() undoes the lane coupling
      c.S[f] = c.S[f] - mS[0] - mS[1];
      c.L[f] = c.L[f] - mL[0] - mL[1];
    }
    h0 = h1; h1 = h2; h2 = h3; h3 = h4; raw_c = raw_n; raw_n = raw_nn;
  }
  return c;
}
// acc = cl + 65536*(ch - cl) (mod 2^32) for lane counts cl, ch, and sums of such values keep that form as long as the
// totals stay below 65536  =>  cl + ch = hi16 + 2*lo16.  Applied ONCE per tile-frame, after the warp/CTA reduction.
__device__ __forceinline__ uint32_t decode_pair(uint32_t a) { return ((a >> 16) + 2u * (a & 0xFFFFu)) & 0xFFFFu; }

// ---------------------------------------------------------------------------------------------------------
// 16-bit samples (YUV420P10/12/16): same tile machinery, 4 pixels (8 bytes) per thread-row.  The 5-tap response of
// 16-bit samples (<= 6*65535) is exact in fp32, so the stencil runs as FADD/FFMA on floats made from the u16 lanes
// with one PRMT (0x4B00 high half = 2^23 + x) and one FADD (-2^23) each; FSET.BF.GE with |x| thresholds, float
// counters (exact below 2^24); the inter-frame difference stays integer: VIMNMX3.U16x2 max/min, SWAR compare, IDP.4A.
// kM = (0x8000 - thM) * 0x00010001 (1 <= thM <= 32768); thS/thL are float bit patterns.
// ---------------------------------------------------------------------------------------------------------
struct F4 { float v[4]; };
__device__ __forceinline__ F4 u16x4_to_float(uint2 raw) {
  F4 r;
  r.v[0] = __uint_as_float(__byte_perm(raw.x, 0x4B000000u, 0x7410)) - 8388608.0f;
  r.v[1] = __uint_as_float(__byte_perm(raw.x, 0x4B000000u, 0x7432)) - 8388608.0f;
  r.v[2] = __uint_as_float(__byte_perm(raw.y, 0x4B000000u, 0x7410)) - 8388608.0f;
  r.v[3] = __uint_as_float(__byte_perm(raw.y, 0x4B000000u, 0x7432)) - 8388608.0f;
  return r;
}
__device__ __forceinline__ uint32_t halves_ge(uint32_t a, uint32_t b, uint32_t kM) {   // |a-b| >= thM per 16-bit lane -> bit 15/31
  const uint32_t d = __vimax3_u16x2(a, b, 0u) - __vimin3_u16x2(a, b, 0xFFFFFFFFu);
  return (((d & 0x7FFF7FFFu) + kM) | d) & 0x80008000u;
}

template <typename Cfg, bool EDGE, int PITCH>
__device__ __forceinline__ RawCounts comb_tile_rows_u16(const uint8_t* __restrict__ cur, const uint8_t* __restrict__ prev,
                                                        const uint32_t* __restrict__ th_rows,
                                                        uint32_t kM, uint32_t thS_bits, uint32_t thL_bits) {
  constexpr int R = Cfg::R;
  const float thS = __uint_as_float(thS_bits), thL = __uint_as_float(thL_bits);
  float fS[2] = { 0.0f, 0.0f }, fL[2] = { 0.0f, 0.0f };
  uint32_t accM[2] = { 0u, 0u };
  uint2 raw_c = *reinterpret_cast<const uint2*>(cur + 2 * PITCH), raw_n = *reinterpret_cast<const uint2*>(cur + 3 * PITCH);
  F4 h0 = u16x4_to_float(*reinterpret_cast<const uint2*>(cur));
  F4 h1 = u16x4_to_float(*reinterpret_cast<const uint2*>(cur + PITCH));
  F4 h2 = u16x4_to_float(raw_c);
  F4 h3 = u16x4_to_float(raw_n);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const uint2 raw_nn = *reinterpret_cast<const uint2*>(cur + (j + 4) * PITCH);
    const F4 h4 = u16x4_to_float(raw_nn);
    const uint2 pv = *reinterpret_cast<const uint2*>(prev + (j + 2) * PITCH);
    const int f = j & 1;
    float tS = thS, tL = thL;
    if (EDGE) { tS = __uint_as_float(th_rows[j]); tL = __uint_as_float(th_rows[R + j]); }
    accM[f] = __dp4a(halves_ge(raw_c.x, pv.x, kM), 0x01010101u, accM[f]);    // difference first, as in the 8-bit path
    accM[f] = __dp4a(halves_ge(raw_c.y, pv.y, kM), 0x01010101u, accM[f]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t = h0.v[q] + h4.v[q];
      t = __fmaf_rn(4.0f, h2.v[q], t);                      // exact: all operands are integers below 2^24
      const float u = h1.v[q] + h3.v[q];
      const float r = fabsf(__fmaf_rn(-3.0f, u, t));
      fS[f] += (r >= tS) ? 1.0f : 0.0f;
      fL[f] += (r >= tL) ? 1.0f : 0.0f;
    }
    h0 = h1; h1 = h2; h2 = h3; h3 = h4; raw_c = raw_n; raw_n = raw_nn;
  }
  RawCounts c;
  c.S[0] = (uint32_t)fS[0]; c.S[1] = (uint32_t)fS[1]; c.L[0] = (uint32_t)fL[0]; c.L[1] = (uint32_t)fL[1];
  c.M[0] = accM[0]; c.M[1] = accM[1];
  return c;
}

template <typename Cfg, int BPS>
// minBlocks is spelled out: ptxas schedules this kernel measurably better with (THREADS, 1) than with (THREADS) alone
// (1.28 vs 1.31 ms per 1800 frames; occupancy is set by shared memory either way).
__global__ void __launch_bounds__(Cfg::THREADS, 1) comb_tma_kernel(const __grid_constant__ CombArgs a) {
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128-byte aligned ring base; pointer arithmetic stays on the __shared__ array so loads compile to LDS
  uint8_t* tiles = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t full_bar[S];
  __shared__ __align__(8) uint64_t empty_bar[S];
  constexpr bool kRelease = Cfg::SYNC == 1;
  uint32_t ephase = 0;                                       // warp 0: parity to wait for on each empty_bar, one bit per slot
  __shared__ unsigned int red[2][Cfg::THREADS / 32][8];       // [buffer][warp][field*3 + metric]: raw per-warp sums, plain stores
  __shared__ uint32_t th_tab[Cfg::RUNS][2][Cfg::R];          // EDGE tiles: per-row thresholds of every run

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  constexpr int TPR = kCombTW / Cfg::STRIP;       // threads per row
  static_assert(TPR == 16 && Cfg::RUNS % 4 == 0 && Cfg::THREADS == 16 * Cfg::RUNS, "lane mapping below assumes 16 threads per row");
  // A warp holds two runs whose first rows have the SAME parity (run, run + RUNS/2: their distance RUNS/2 * R is even),
  // so "slot 0 / slot 1" of the raw counters means the same field for all 32 lanes and one full-warp REDUX sums them.
  const int strip = lane & 15, run = (tid >> 5) + (lane >> 4) * (Cfg::RUNS / 2);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], Cfg::THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t gload = 0;      // loads consumed so far by this CTA (ring position of L_0 of the current segment)
  uint32_t gstep = 0;      // tile-frames processed so far (selects the red[] buffer)
  const int seg_lo = a.seg_start[blockIdx.x], seg_hi = a.seg_start[blockIdx.x + 1];
  for (int si = seg_lo; si < seg_hi; ++si) {
    if (kRelease && si > seg_lo) __syncthreads();    // th_tab / ring slots of the previous segment are no longer read
    const CombSegment seg = a.segs[si];
    const int pl = (seg.tile >= a.plane[3].tile0) ? 3 : (seg.tile >= a.plane[2].tile0) ? 2 : (seg.tile >= a.plane[1].tile0) ? 1 : 0;
    const CombPlane& P = a.plane[pl];
    const bool merged = pl == 3;                     // strips 0..7 <- U remainder, strips 8..15 <- V remainder
    const CUtensorMap* map = &a.map[merged ? 0 : pl];
    const int lt = seg.tile - P.tile0;
    const int ty = lt / P.tilesX, tx = lt - ty * P.tilesX;
    const int x0 = tx * (kCombTW / BPS), y0 = ty * Cfg::TH;      // TMA coordinates are in samples
    const int nf = seg.fend - seg.fbegin;
    const int nloads = nf + 1;                       // L_0 = previous frame, L_k = frame fbegin+k-1
    const int fprev = seg.fbegin > 0 ? seg.fbegin - 1 : seg.fbegin;
    const bool edge = (y0 < 2) || (y0 + Cfg::TH + 2 > P.H);
    const int y_first = y0 + run * Cfg::R;
    const bool rows_live = y_first < P.H;            // thread's run intersects the plane

    auto issue_at = [&](int j, int st) {             // thread 0 only; st = (gload + j) % S
      const int fr = (j == 0) ? fprev : seg.fbegin + j - 1;
      mbar_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
      if (!merged) {
        tma_load_3d(tiles + st * Cfg::STAGE_BYTES, map, &full_bar[st], x0, y0 - 2, fr);
      } else {                                       // two half-width boxes, one behind the other in the stage
        tma_load_3d(tiles + st * Cfg::STAGE_BYTES, &a.map_half[0], &full_bar[st], a.half_x, y0 - 2, fr);
        tma_load_3d(tiles + st * Cfg::STAGE_BYTES + Cfg::STAGE_BYTES / 2, &a.map_half[1], &full_bar[st], a.half_x, y0 - 2, fr);
      }
    };
    if (tid == 0) {
      const int pro = nloads < S ? nloads : S;
      for (int j = 0; j < pro; ++j) issue_at(j, (int)((gload + (uint32_t)j) % S));
    }
    if (edge || merged) {                            // (re)build the per-row threshold table of this tile
      for (int i = tid; i < Cfg::RUNS * Cfg::R; i += Cfg::THREADS) {
        const int y = y0 + i;
        const bool ok = y >= 2 && y < P.H - 2;
        const uint32_t inf = BPS == 1 ? 0x7C007C00u : 0x7F800000u;     // +inf as fp16x2 / fp32
        th_tab[i / Cfg::R][0][i % Cfg::R] = ok ? P.thS : inf;
        th_tab[i / Cfg::R][1][i % Cfg::R] = ok ? P.thL : inf;
      }
      __syncthreads();
    }
    int stp = (int)(gload % S);                             // ring slot / phase of load g-1, advanced without div/mod
    uint32_t php = (gload / S) & 1u;
    mbar_wait(&full_bar[stp], php);                         // L_0
    const int flip = y_first & 1;                           // slot 0 of this thread's run holds rows of this parity
    for (int k = 1; k <= nf; ++k) {
      int st = stp + 1; uint32_t ph = php;
      if (st == S) { st = 0; ph ^= 1u; }
      mbar_wait(&full_bar[st], ph);
      RawCounts c = { { 0u, 0u }, { 0u, 0u }, { 0u, 0u } };
      if (!merged) {
        const uint8_t* cur = tiles + st * Cfg::STAGE_BYTES + (run * Cfg::R) * kCombTW + strip * Cfg::STRIP;
        const uint8_t* prv = tiles + stp * Cfg::STAGE_BYTES + (run * Cfg::R) * kCombTW + strip * Cfg::STRIP;
        if (!edge) {
          if (BPS == 1) c = comb_tile_rows<Cfg, false, kCombTW>(cur, prv, nullptr, P.thM, P.thS, P.thL);
          else c = comb_tile_rows_u16<Cfg, false, kCombTW>(cur, prv, nullptr, P.thM, P.thS, P.thL);
        } else if (rows_live) {
          if (BPS == 1) c = comb_tile_rows<Cfg, true, kCombTW>(cur, prv, &th_tab[run][0][0], P.thM, P.thS, P.thL);
          else c = comb_tile_rows_u16<Cfg, true, kCombTW>(cur, prv, &th_tab[run][0][0], P.thM, P.thS, P.thL);
        }
      } else if (rows_live) {                        // half-width sub-tiles: row pitch 64 bytes
        constexpr int HP = kCombTW / 2, HTPR = HP / Cfg::STRIP;
        const int off = (strip / HTPR) * (Cfg::STAGE_BYTES / 2) + (run * Cfg::R) * HP + (strip % HTPR) * Cfg::STRIP;
        if (BPS == 1) c = comb_tile_rows<Cfg, true, HP>(tiles + st * Cfg::STAGE_BYTES + off, tiles + stp * Cfg::STAGE_BYTES + off,
                                                        &th_tab[run][0][0], P.thM, P.thS, P.thL);
        else c = comb_tile_rows_u16<Cfg, true, HP>(tiles + st * Cfg::STAGE_BYTES + off, tiles + stp * Cfg::STAGE_BYTES + off,
                                                   &th_tab[run][0][0], P.thM, P.thS, P.thL);
      }
      // Raw (still pair-coded) counters are summed per warp (REDUX), stored per warp in shared memory and added up,
      // decoded and sent to global memory by six writer threads once per tile-frame.
      const uint32_t rM0 = __reduce_add_sync(0xFFFFFFFFu, c.M[0]), rM1 = __reduce_add_sync(0xFFFFFFFFu, c.M[1]);
      const uint32_t rS0 = __reduce_add_sync(0xFFFFFFFFu, c.S[0]), rS1 = __reduce_add_sync(0xFFFFFFFFu, c.S[1]);
      const uint32_t rL0 = __reduce_add_sync(0xFFFFFFFFu, c.L[0]), rL1 = __reduce_add_sync(0xFFFFFFFFu, c.L[1]);
      const int rb = gstep & 1;
      if (lane < 6) {                                // lane = field*3 + metric; slot 0 holds field `flip`
        const int fld = lane >= 3, met = lane - 3 * fld, slot = fld ^ flip;
        const uint32_t v = met == 0 ? (slot ? rM1 : rM0) : met == 1 ? (slot ? rS1 : rS0) : (slot ? rL1 : rL0);
        red[rb][tid >> 5][lane] = v;                 // plain store: no shared-memory atomics in front of the barrier
      }
      auto flush = [&]() {                           // tid = field*3 + metric = the counts[] layout of one class
        unsigned v = 0;
#pragma unroll
        for (int w = 0; w < Cfg::THREADS / 32; ++w) v += red[rb][w][tid];
        const int metric = tid >= 3 ? tid - 3 : tid;
        v = metric == 0 ? (v >> 7) : (BPS == 1 ? decode_pair(v) : v);
        if (v) atomicAdd(a.counts + (size_t)(seg.fbegin + k - 1 - a.out_frame0) * 12 + P.cls * 6 + tid, (int)v);
      };
      if (!kRelease) {
        __syncthreads();                             // all reads of stage stp done; red[rb] complete
        if (tid < 6) flush();
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stp]); // this warp: done with slot stp, its sums are in red[rb]
        if (tid < 32) {
          mbar_wait(&empty_bar[stp], (ephase >> stp) & 1u);
          ephase ^= 1u << stp;
          if (tid < 6) flush();
          __syncwarp();
        }
      }
      if (tid == 0 && (k - 1 + S) < nloads) issue_at(k - 1 + S, stp);     // refill the slot that was just released
      ++gstep;
      stp = st; php = ph;
    }
    gload += (uint32_t)nloads;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Generic fallback: any sample size (8/10/12/16-bit) and any pitch.  One thread per pixel, plain cached loads,
// ballot/popc counting.  Used for YUV420P10..16 clips and for 8-bit layouts TMA cannot describe (pitch or offsets
// not multiples of 16 bytes).  Same integer spec, same results; roughly an order of magnitude below the streaming
// kernel (it re-reads the previous frame and the vertical neighbours through L1/L2).
// ---------------------------------------------------------------------------------------------------------
struct CombGenericArgs {
  const uint8_t* base; long long frame_stride; long long off[3];
  int pitch[3];            // ELEMENTS
  int W[3], H[3];
  int thM[3], thS[3], thL[3];
  int first_frame;         // window-relative index of the first frame to analyse
  int prev_of_first;       // window-relative index of its predecessor (== first_frame when there is none)
  int nframes;
  int* counts;             // row 0 = first_frame
};

// block = 8 warps, tile = 128 columns x 128 rows of one plane of one frame: warp w walks rows [16w, 16w+16) of the
// tile, lane l owns columns l, l+32, l+64, l+96 with a 5-row sliding window in registers (2 loads per pixel: current
// and previous frame).  Counters: registers -> REDUX -> shared -> 6 global atomics per block.
constexpr int kGenTW = 128, kGenTH = 128;
template <typename pixel_t>
__global__ void __launch_bounds__(256) comb_generic_kernel(const CombGenericArgs a) {
  __shared__ int blk[6];
  const int pl = blockIdx.z % 3, f = blockIdx.z / 3;
  const int W = a.W[pl], H = a.H[pl], pitch = a.pitch[pl];
  const int x0 = blockIdx.x * kGenTW, y0 = blockIdx.y * kGenTH;
  if (x0 >= W || y0 >= H) return;                              // block-uniform
  if (threadIdx.x < 6) blk[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cur_f = a.first_frame + f, prev_f = f == 0 ? a.prev_of_first : cur_f - 1;
  const pixel_t* cur = reinterpret_cast<const pixel_t*>(a.base + (long long)cur_f * a.frame_stride + a.off[pl]);
  const pixel_t* prv = reinterpret_cast<const pixel_t*>(a.base + (long long)prev_f * a.frame_stride + a.off[pl]);
  const int thM = a.thM[pl], thS = a.thS[pl], thL = a.thL[pl];
  int cnt[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } };                // [field][move, shima, lshima]
  const int ya = y0 + warp * 16, yb = min(ya + 16, H);
#pragma unroll
  for (int cg = 0; cg < 4; ++cg) {
    const int x = x0 + lane + 32 * cg;
    if (x >= W || ya >= H) continue;
    auto px = [&](int y) -> int { return (y >= 0 && y < H) ? (int)cur[x + (long long)y * pitch] : 0; };
    int r0 = px(ya - 2), r1 = px(ya - 1), r2 = px(ya), r3 = px(ya + 1);
    for (int y = ya; y < yb; ++y) {
      const int r4 = px(y + 2);
      int d = r2 - (int)prv[x + (long long)y * pitch]; d = d < 0 ? -d : d;
      const int fld = y & 1;
      cnt[fld][0] += d >= thM;
      if (y >= 2 && y < H - 2) {
        int v = r0 + 4 * r2 + r4 - 3 * (r1 + r3); v = v < 0 ? -v : v;
        cnt[fld][1] += v >= thS; cnt[fld][2] += v >= thL;
      }
      r0 = r1; r1 = r2; r2 = r3; r3 = r4;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int v = __reduce_add_sync(0xFFFFFFFFu, cnt[k / 3][k % 3]);
    if (lane == 0 && v) atomicAdd(&blk[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 6 && blk[threadIdx.x])
    atomicAdd(a.counts + (size_t)f * 12 + (pl ? 6 : 0) + threadIdx.x, blk[threadIdx.x]);    // [field][metric] order = k
}

}  // namespace amtk
