// comb_mma.cuh -- field-difference / combing metric, streaming pass, tensor-core variant (8-bit samples).
//
// Why tensor cores in an HBM-streaming kernel: the SIMT kernels (comb_kernels.cuh, comb_stream.cuh) are NOT memory bound.
// They sit at 0.70 of the measured HBM peak because the 5-tap vertical stencil costs ~6 issue slots per pixel
// (u8 -> fp16 conversion 0.5, stencil 2.0, thresholds 1.0, mask sums 0.5, inter-frame difference 1.25, loads/overhead)
// and the ALU pipe (HSET2/PRMT/LOP3/IADD3) saturates first (profiles/r02_*).  The stencil is a product with a banded
// Toeplitz matrix -- exact in u8 x u8 -> s32 -- and the tensor pipe is otherwise idle, so here it does the stencil:
//
//     D[x][n] = sum_r  tile[r][x] * band[r][n]          M = 128 pixels of a tile row, K = 64 box rows, N = 128
//        n in [0,60):    pos(y0+n) = p[y-2] + 4 p[y] + p[y+2]        (band entries 1, 4, 1)
//        n in [64,124):  neg(y0+n) = 3 (p[y-1] + p[y+1])             (band entries 3, 3)
//
// as two tcgen05.mma.kind::i8 (M128 N128 K32) per tile-frame, issued by one thread, with the TMA-staged tile itself as the
// MN-major A operand (128-byte swizzle: the tile is never copied or converted) and the accumulator in tensor memory.
// Both halves are < 2048, so their low 16 bits are exact fp16 bit patterns (k * 2^-24): tcgen05.ld ... .pack::16b hands
// every thread (= one pixel column) its 60 responses as 2 x 30 packed registers, and the rest is
//     r = HADD2(pos, -neg);  HSET2.GE(|r|, thS);  HSET2.GE(|r|, thL);  3-input adds of the masks
// = 2 issue slots per pixel for the comb response instead of 4.5.  The two 16-bit lanes of a register are rows y, y+1 =
// the two fields, so one pair-coded accumulator carries both field counters.  The inter-frame difference stays on the
// SIMT side (VABSDIFF4 + SWAR compare + IDP.4A on the raw bytes of the current and the previous slot).
//
// Same integer spec, same counters, bit-identical results (integer adds commute).  Measured SLOWER than the warp-stream
// kernel (1.5 ms vs 1.23 ms per 1800 1080p frames): it is an opt-in experiment (AMTK_COMB_MMA=1|2), not the
// product path.  Every device-side wait has a watchdog (mm_wait): a protocol error fails the call, it cannot hang the GPU.
#pragma once
#include <cuda_fp16.h>
#include "amtk_internal.h"
#include "tma_utils.cuh"
#include "comb_stream.cuh"       // WsArgs / WsClass / CombSegment / bytes_ge

namespace amtk {

constexpr int kMmTW = 128;                       // tile width in bytes = MMA M
constexpr int kMmTH = 60;                        // output rows per tile
constexpr int kMmBoxH = 64;                      // + 2 halo rows above and below = MMA K
constexpr int kMmSlot = kMmTW * kMmBoxH;         // 8192 bytes
constexpr int kMmStages = 4;                     // previous frame, current frame, two frames in flight (power of two)
constexpr int kMmConsumerWarps = 4;
constexpr int kMmThreads = 32 * (kMmConsumerWarps + 1);   // + the producer warp
constexpr int kMmBandBytes = 128 * kMmBoxH;      // N x K u8
constexpr int kMmSmemPerStream = kMmStages * kMmSlot;    // + kMmBandBytes + 1024 once per CTA   // + slack for the 1024-byte alignment the swizzle needs
constexpr int kMmBandLBO = 2048, kMmBandSBO = 128;                   // band matrix: K-major, no swizzle, 8x16-byte core matrices

// ---- tcgen05 wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], u8 x u8 -> s32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// 32 lanes x 32 columns, low 16 bits of each column, two adjacent columns per register (even column in the low half)
__device__ __forceinline__ void tc_ld_pack16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t tc_ld1(uint32_t taddr) {        // one 32-bit column of this thread's lane
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}

// Watchdog wait.  The kernel is an experiment with a known rare hang (see the header of this file / DESIGN.md 3.1b), so no
// wait may spin forever: after ~1 s a thread records (code, step, CTA, thread) in dbg[] and raises dbg[0]; from then on
// every wait of every CTA returns at once, the launch drains with garbage results and the host reports the failure.
constexpr long long kMmWaitLimit = 2000000000ll;           // clock64 ticks (~1 s)
__device__ __noinline__ void mm_wait_slow(uint64_t* bar, uint32_t parity, int* dbg, int code, int step, bool sleepy) {
  const long long t0 = clock64();
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
    if (sleepy) __nanosleep(64);
    if ((spin & 15u) == 15u) {
      if (*reinterpret_cast<volatile int*>(dbg)) return;
      if (clock64() - t0 > kMmWaitLimit) {
        if (atomicCAS(dbg, 0, 1) == 0) { dbg[1] = code; dbg[2] = step; dbg[3] = (int)blockIdx.x; dbg[4] = (int)threadIdx.x; dbg[5] = (int)parity; __threadfence(); }
        return;
      }
    }
  }
}
__device__ __forceinline__ void mm_wait(uint64_t* bar, uint32_t parity, int* dbg, int code, int step, bool sleepy = false) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (!done) mm_wait_slow(bar, parity, dbg, code, step, sleepy);
}

// shared-memory matrix descriptor (sm_100 format: version 1 at bit 46)
__device__ __forceinline__ uint64_t mm_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor: D = s32, A = u8 MN-major, B = u8 K-major, M = 128, N = 128, dense, no saturation
constexpr uint32_t kMmIdesc = (2u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (0u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// band[n][k], k = box row (global row y0 - 2 + k), n = output column of D
__device__ __forceinline__ uint32_t mm_band(int n, int k) {
  if (n < kMmTH) { const int d = k - n; return (d == 0 || d == 4) ? 1u : (d == 2 ? 4u : 0u); }
  if (n >= 64 && n < 64 + kMmTH) { const int d = k - (n - 64); return (d == 1 || d == 3) ? 3u : 0u; }
  return 0u;
}

// Roles: warps 0..3 = consumers (warp w owns tensor-memory lanes 32w..32w+31 = 32 pixel columns of each tile, and 15 of
// the 60 strip rows of the inter-frame difference); warp 4 = producer (TMA loads, MMA issue, counter flush).
// A CTA streams TWO tiles at once (work items come in pairs with the same frame range): every step loads, multiplies and
// thresholds frame k of both tiles, so the fixed per-step costs (two mbarrier waits, the accumulator hand-over, loop and
// ring bookkeeping) are paid once per two tile-frames, and a consumer warp has ~500 independent instructions between
// handing the accumulators back and needing the next ones -- that is what hides the ~800-cycle MMA round trip.
// No block barrier in the frame loop: consumers wait only for data (full_bar, mma_bar); the producer waits for free_bar,
// on which every consumer warp arrives once both accumulators of step k are in its registers and it is done with the
// slots of step k-1.  2 CTAs per SM (2 x 256 tensor-memory columns = all 512).
template <int NS>
__global__ void __launch_bounds__(kMmThreads, 4 / NS) comb_mma_kernel(const __grid_constant__ WsArgs a) {
  constexpr int kStageBytes = NS * kMmSlot;
  constexpr int kTmemCols = 128 * NS;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMmStages];
  __shared__ __align__(8) uint64_t mma_bar, free_bar;
  __shared__ uint32_t tmem_holder;
  __shared__ int item_s;
  __shared__ uint32_t red[4][16];                            // [step & 3][stream * 8 + counter]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == kMmConsumerWarps;
  uint8_t* slots = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* band = slots + kMmStages * kStageBytes;

  // ---- one-time setup: tensor memory, band matrix, barriers ----
  if (producer) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int o = tid * 4; o < kMmBandBytes; o += kMmThreads * 4) {       // 4 consecutive k of one row n per store
    const int kc = o / kMmBandLBO, rem = o - kc * kMmBandLBO;
    const int ng = rem / kMmBandSBO, i = (rem % kMmBandSBO) >> 4, kk = rem & 15;
    const int n = ng * 8 + i, k = kc * 16 + kk;
    *reinterpret_cast<uint32_t*>(band + o) = mm_band(n, k) | (mm_band(n, k + 1) << 8) | (mm_band(n, k + 2) << 16) | (mm_band(n, k + 3) << 24);
  }
  if (tid < 64) red[tid >> 4][tid & 15] = 0u;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kMmStages; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(&mma_bar, 1);
    mbar_init(&free_bar, kMmConsumerWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // band matrix: generic-proxy stores -> async-proxy (MMA) reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  int* const dbg = a.queue + 16;                             // watchdog record (64 bytes after the queue counter, zeroed per launch)
  const uint32_t tmem_base = tmem_holder;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 lanes
  const uint32_t slots_u32 = smem_u32(slots);
  const uint64_t descB0 = mm_desc(smem_u32(band), kMmBandLBO, kMmBandSBO, 0u);
  const uint64_t descB1 = mm_desc(smem_u32(band) + 2 * kMmBandLBO, kMmBandLBO, kMmBandSBO, 0u);

  // inter-frame difference: consumer thread t owns strip (t & 7) of the box rows b0, b0+8, b0+16, b0+24 with
  // b0 = 2 + (t >> 3) for t < 64 and 34 + ((t - 64) >> 3) for the rest (rows 2..61 = the 60 tile rows; the fourth row of
  // the last four row groups would be 62..65 and is skipped).  Rows 8 apart share (row & 7), so the swizzled 16-byte chunk
  // (chunk index ^ (row & 7)) is the same in all four: one address + immediates.  They also share the row parity = the field.
  const int mv_b0 = (tid < 64 ? 2 : 34) + ((tid & 63) >> 3);
  const int mv_toff = mv_b0 * kMmTW + (((tid & 7) ^ (mv_b0 & 7)) << 4);
  const bool mv_row3 = mv_b0 + 24 <= kMmTH + 1;
  const int mv_shift = (mv_b0 & 1) ? 16 : 0;                 // tile row = box row - 2: same parity; odd rows count in the high half
  // publishing lane i < 6 of a consumer warp owns counter i of counts[]' [field][move, shima, lshima] layout
  const uint32_t pub_selM = (lane == 0 || lane == 3) ? 0xFFFFFFFFu : 0u, pub_selS = (lane == 1 || lane == 4) ? 0xFFFFFFFFu : 0u;
  const uint32_t pub_selL = (lane == 2 || lane == 5) ? 0xFFFFFFFFu : 0u;
  const int pub_shift = (lane >= 3 && lane < 6) ? 16 : 0;
  uint32_t* const pub_red = &red[0][lane & 7];

  uint32_t gload = 0;          // stage loads consumed so far (ring position)
  uint32_t nmma = 0;           // MMA batches committed so far (phase of mma_bar)
  uint32_t nfree = 0;          // completed phases of free_bar (producer)
  int* pend_crow = nullptr;    // producer: counters of the last frame of the previous item still sit in red[]
  int pend_k = 0;
  for (;;) {
    if (tid == kMmConsumerWarps * 32) item_s = atomicAdd(a.queue, 1);
    __syncthreads();
    // Every consumer has published the last frame of the previous item before it reached this barrier: flush it now.
    // (A closing arrival on free_bar instead would let the consumers complete TWO phases -- last step, closing -- without
    // the producer in between; a parity wait that is two phases late never returns.  That was the rare hang of the first
    // version: watchdog record "wait 2, step nf, producer".)
    if (producer && pend_k > 0) {
      if (lane < 8 * NS) {
        const uint32_t v = red[pend_k & 3][lane];
        red[pend_k & 3][lane] = 0u;
        if (pend_crow && v) atomicAdd(pend_crow + (size_t)pend_k * 12, (int)v);
      }
      pend_k = 0;
    }
    // warp-uniform by construction; the reduction tells the compiler so (uniform registers for all the bookkeeping)
    const int pair = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)item_s);
    if (NS * pair >= a.nitems) break;
    int tileS[NS], fb = 0, fe = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const CombSegment seg = a.segs[NS * pair + s];
      tileS[s] = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(seg.tile + 0x40000000)) - 0x40000000;   // < 0: filler stream, results dropped
      if (s == 0) { fb = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)seg.fbegin); fe = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)seg.fend); }
    }
    const int nf = fe - fb;
    const int nloads = nf + 1;                               // L_0 = previous frame, L_k = frame fb+k-1
    const int fprev = fb > 0 ? fb - 1 : fb;
    int ciS[NS], txS[NS], y0S[NS]; bool dropS[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      dropS[s] = tileS[s] < 0;
      const int tile = dropS[s] ? ~tileS[s] : tileS[s];
      int ci = 0;
#pragma unroll
      for (int k = 1; k < kWsMaxClasses; ++k) if (k < a.nclasses && tile >= a.cl[k].tile0) ci = k;
      const int lt = tile - a.cl[ci].tile0;
      const int ty = lt / a.cl[ci].tilesX;
      ciS[s] = ci; txS[s] = lt - ty * a.cl[ci].tilesX; y0S[s] = ty * kMmTH;
    }

    if (producer) {
      // =========================== producer warp ===========================
      auto issue_load = [&](int j) {                         // lane 0 only: both tiles of load j into one stage
        const uint32_t gl = gload + (uint32_t)j;
        const int st = (int)(gl & (kMmStages - 1));
        const int fr = (j == 0) ? fprev : fb + j - 1;
        mbar_expect_tx(&full_bar[st], kStageBytes);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          tma_load_3d(slots + st * kStageBytes + s * kMmSlot, &a.map[a.cl[ciS[s]].map], &full_bar[st], txS[s] * kMmTW, y0S[s] - 2, fr);
      };
      auto issue_mma = [&](int j) {                          // all lanes wait for the stage, lane 0 issues
        const uint32_t gl = gload + (uint32_t)j;
        const int st = (int)(gl & (kMmStages - 1));
        mm_wait(&full_bar[st], (gl / kMmStages) & 1u, dbg, 1, j);
        if (lane == 0) {
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const uint32_t sa = slots_u32 + st * kStageBytes + s * kMmSlot;
            tc_mma_i8(tmem_base + s * 128u, mm_desc(sa, 0u, 1024u, 2u), descB0, kMmIdesc, 0u);                 // box rows 0..31
            tc_mma_i8(tmem_base + s * 128u, mm_desc(sa + 32 * kMmTW, 0u, 1024u, 2u), descB1, kMmIdesc, 1u);    // box rows 32..63
          }
          tc_commit(&mma_bar);
        }
        __syncwarp();
      };
      const int fl_s = lane >> 3, fl_i = lane & 7;           // flush lane = (stream, counter)
      int* const crow = (lane < 8 * NS && fl_i < 6 && !dropS[fl_s & (NS - 1)])
                            ? a.counts + a.cl[ciS[fl_s & (NS - 1)]].cls * 6 + fl_i + ((long long)fb - 1 - a.out_frame0) * 12 : nullptr;
      auto flush = [&](int k) {                              // counters of frame k: shared -> global, slot cleared for reuse
        if (lane < 8 * NS) {
          const uint32_t v = red[k & 3][lane];
          red[k & 3][lane] = 0u;
          if (crow && v) atomicAdd(crow + (size_t)k * 12, (int)v);
        }
      };
      if (lane == 0) {
        const int pro = nloads < kMmStages ? nloads : kMmStages;
        for (int j = 0; j < pro; ++j) issue_load(j);
      }
      __syncwarp();
      issue_mma(1);
      for (int k = 1; k <= nf; ++k) {
        mm_wait(&free_bar, nfree & 1u, dbg, 2, k, true); ++nfree;   // accumulators of step k are in registers; stage of load k-1 is free
        if (lane == 0 && (k + kMmStages - 1) < nloads) issue_load(k + kMmStages - 1);
        if (k < nf) issue_mma(k + 1);
        if (k > 1) flush(k - 1);                             // every consumer published frame k-1 before it arrived for step k
      }
      pend_crow = crow; pend_k = nf;                         // frame nf is flushed after the next block barrier
      nmma += (uint32_t)nf;
    } else {
      // =========================== consumer warps ===========================
      // rows of a tile the spec excludes although the plain pass counts them: y < 2 and H-2 <= y < H+2 (rows >= H are
      // zero-filled by TMA, but the windows of H, H+1 still see the last two real rows).  Bit n = tile row n.
      unsigned long long fixS[NS];
      uint32_t kMS[NS], tSS[NS], tLS[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const WsClass& C = a.cl[ciS[s]];
        unsigned long long fix = 0ull;
        if (y0S[s] == 0) fix |= 3ull;
        const int lo = max(C.H - 2 - y0S[s], 0), hi = min(C.H + 2 - y0S[s], kMmTH);
        if (hi > lo) fix |= ((1ull << hi) - 1ull) & ~((1ull << lo) - 1ull);
        fixS[s] = fix; kMS[s] = C.thM; tSS[s] = C.thS; tLS[s] = C.thL;
      }
      const bool any_fix = (fixS[0] | fixS[NS - 1]) != 0ull;

      // inter-frame difference of load j against load j-1, one tile: packed hits, low half = even rows, high = odd rows
      auto move_tile = [&](const uint8_t* cur, const uint8_t* prv, uint32_t kM) -> uint32_t {
        uint32_t m = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i == 3 && !mv_row3) break;
          const uint4 c4 = *reinterpret_cast<const uint4*>(cur + mv_toff + i * 8 * kMmTW);
          const uint4 p4 = *reinterpret_cast<const uint4*>(prv + mv_toff + i * 8 * kMmTW);
          m = __dp4a(bytes_ge(__vabsdiffu4(c4.x, p4.x), kM), 0x01010101u, m);
          m = __dp4a(bytes_ge(__vabsdiffu4(c4.y, p4.y), kM), 0x01010101u, m);
          m = __dp4a(bytes_ge(__vabsdiffu4(c4.z, p4.z), kM), 0x01010101u, m);
          m = __dp4a(bytes_ge(__vabsdiffu4(c4.w, p4.w), kM), 0x01010101u, m);
        }
        return (m >> 7) << mv_shift;                           // <= 64 hits per lane, in the half of its field
      };
      auto move_step = [&](int j, uint32_t (&pend)[NS]) {
        const uint32_t gl = gload + (uint32_t)j;
        const int st = (int)(gl & (kMmStages - 1)), sp = (int)((gl - 1) & (kMmStages - 1));
        mm_wait(&full_bar[st], (gl / kMmStages) & 1u, dbg, 4, j);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          pend[s] = move_tile(slots + st * kStageBytes + s * kMmSlot, slots + sp * kStageBytes + s * kMmSlot, kMS[s]);
      };
      // thresholds of one accumulator (this thread's pixel column, 64 columns = 32 row pairs each of pos and neg).
      // HSET2.BF writes 1.0 per hit; the hits are counted with HADD2 (FMA-heavy pipe: the ALU pipe, which carries the
      // compares, is the one that saturates) -- exact, <= 32 per half.  One multiply by 2^-24 turns the fp16 counts into
      // their integer bit patterns (k * 2^-24 is the subnormal with bits k): low half = even rows, high half = odd rows.
      auto comb_tile = [&](const uint32_t (&pos)[32], const uint32_t (&neg)[32], uint32_t tS, uint32_t tL, uint32_t& cntS, uint32_t& cntL) {
        const __half2 thS = *reinterpret_cast<const __half2*>(&tS);
        const __half2 thL = *reinterpret_cast<const __half2*>(&tL);
        __half2 aS[4], aL[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { aS[c] = __float2half2_rn(0.0f); aL[c] = __float2half2_rn(0.0f); }
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const __half2 r = __habs2(__hsub2(*reinterpret_cast<const __half2*>(&pos[q]), *reinterpret_cast<const __half2*>(&neg[q])));
          aS[q & 3] = __hadd2(aS[q & 3], __hge2(r, thS));
          aL[q & 3] = __hadd2(aL[q & 3], __hge2(r, thL));
        }
        const uint32_t one_ulp = 0x00010001u;                  // 2^-24 in both halves
        const __half2 ulp = *reinterpret_cast<const __half2*>(&one_ulp);
        const __half2 sS = __hmul2(__hadd2(__hadd2(aS[0], aS[1]), __hadd2(aS[2], aS[3])), ulp);
        const __half2 sL = __hmul2(__hadd2(__hadd2(aL[0], aL[1]), __hadd2(aL[2], aL[3])), ulp);
        cntS = *reinterpret_cast<const uint32_t*>(&sS);
        cntL = *reinterpret_cast<const uint32_t*>(&sL);
      };

      {                                                      // L_0: the frame before the first one of this item
        const int st = (int)(gload & (kMmStages - 1));
        mm_wait(&full_bar[st], (gload / kMmStages) & 1u, dbg, 5, 0);
      }
      uint32_t pendM[NS];
      move_step(1, pendM);
      for (int k = 1; k <= nf; ++k) {
        mm_wait(&mma_bar, nmma & 1u, dbg, 6, k); ++nmma;       // both accumulators of step k are complete
        tc_fence_after();
        uint32_t pos[NS][32], neg[NS][32];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          tc_ld_pack16(tmem_lane + s * 128u, *reinterpret_cast<uint32_t(*)[16]>(&pos[s][0]));
          tc_ld_pack16(tmem_lane + s * 128u + 32u, *reinterpret_cast<uint32_t(*)[16]>(&pos[s][16]));
          tc_ld_pack16(tmem_lane + s * 128u + 64u, *reinterpret_cast<uint32_t(*)[16]>(&neg[s][0]));
          tc_ld_pack16(tmem_lane + s * 128u + 96u, *reinterpret_cast<uint32_t(*)[16]>(&neg[s][16]));
        }
        uint32_t fixv[NS] = {};                       // edge tiles: hits of the excluded rows, packed [S even | S odd<<8 | L even<<16 | L odd<<24]
        if (any_fix) {
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            const int thS_i = (int)(tSS[s] & 0xFFFFu), thL_i = (int)(tLS[s] & 0xFFFFu);
            for (unsigned long long m = fixS[s]; m; m &= m - 1ull) {
              const int n = __ffsll((long long)m) - 1;
              const int pv = (int)tc_ld1(tmem_lane + s * 128u + (uint32_t)n), nv = (int)tc_ld1(tmem_lane + s * 128u + 64u + (uint32_t)n);
              tc_wait_ld();
              const int r = abs(pv - nv);
              fixv[s] += ((r >= thS_i) ? 1u : 0u) << ((n & 1) * 8);
              fixv[s] += ((r >= thL_i) ? 1u : 0u) << (16 + (n & 1) * 8);
            }
          }
        }
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&free_bar);               // accumulators and the stage of load k-1 may be overwritten
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          uint32_t cntS, cntL;                               // low half = hits in even rows (top field), high half = odd rows
          comb_tile(pos[s], neg[s], tSS[s], tLS[s], cntS, cntL);
          const uint32_t rS = __reduce_add_sync(0xFFFFFFFFu, cntS), rL = __reduce_add_sync(0xFFFFFFFFu, cntL);   // <= 32 x 32 per half
          const uint32_t rM = __reduce_add_sync(0xFFFFFFFFu, pendM[s]);
          uint32_t v = (((rM & pub_selM) | (rS & pub_selS) | (rL & pub_selL)) >> pub_shift) & 0xFFFFu;
          if (any_fix) {
            // <= 4 excluded rows per lane -> <= 4 hits per 8-bit field and lane, x 32 lanes = 128 < 256
            const uint32_t fS = __reduce_add_sync(0xFFFFFFFFu, fixv[s] & 0xFFFFu), fL = __reduce_add_sync(0xFFFFFFFFu, fixv[s] >> 16);
            if (lane == 1) v -= fS & 0xFFu;
            if (lane == 4) v -= (fS >> 8) & 0xFFu;
            if (lane == 2) v -= fL & 0xFFu;
            if (lane == 5) v -= (fL >> 8) & 0xFFu;
          }
          if (lane < 6 && v) atomicAdd(pub_red + ((k & 3) * 16 + s * 8), v);
        }
        if (k < nf) move_step(k + 1, pendM);
      }
    }
    gload += (uint32_t)nloads;
  }
  tc_fence_before();
  __syncthreads();
  if (producer) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
}

}  // namespace amtk
