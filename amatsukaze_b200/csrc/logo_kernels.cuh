// logo_kernels.cuh -- logo-template correlation on the GPU.
//
// Replaces, per frame: DeintY/CopyY (LogoScan.hpp:763-790), LogoDataParam::EvaluateLogo (:231-255) and
// LogoDataParam::CorrelationScore (:288-318) incl. CalcCorrelation5x5_AVX (ComputeKernel.cpp:77-121).
//
// Structure (two launches per evaluation job):
//   logo_scores_kernel : grid (pixel slices, frame lanes).  Each thread OWNS up to PXT feature pixels of the logo
//                        and keeps their 25 zero-mean taps in registers for the whole launch, so the 129 KB tap
//                        table is read once per CTA instead of once per frame (it would otherwise need ~400 TB/s
//                        of L2 bandwidth at the frame rates the streaming pass reaches).  Per frame the CTA stages
//                        the ROI as float (deinterlaced or raw) in shared memory, and per fade level builds the
//                        logo-removed image `work` in shared memory and lets every thread emit the score of its
//                        pixels, using the exact AVX expression tree (exact_math.h).
//   logo_sum_kernel    : the reference adds the ~1.3k pixel scores of an evaluation SEQUENTIALLY in float
//                        (LogoScan.hpp:310).  That order is kept (bit-exact results), but one thread per
//                        (frame, fade) runs its chain, so all 32 lanes of a warp carry independent chains.
#pragma once
#include "amtk_internal.h"
#include "exact_math.h"

namespace amtk {

constexpr int kEvalThreads = 512;
constexpr int kMaxFades = 24;

struct EvalJob {
  const void* ybase;         // Y plane of frame 0 of the (device-resident) clip window
  long long frame_stride;    // bytes
  int pitch;                 // ELEMENTS
  int frame0, nframes;       // frames [frame0, frame0+nframes) of the window
  int imgx, imgy;            // ROI origin in the frame (full-frame coordinates)
  int roi_w, roi_h;          // staged ROI size (always the FULL logo rectangle, also for field logos)
  int src_mode;              // 0: DeintY, 1: CopyY
  int src_off, src_stride;   // view of the staged ROI this logo reads (field logos: off = 0|w, stride = 2w)
  LogoDev logo;              // logo.w x logo.h = evaluated size (h/2 for field logos)
  float maxv;
  int nfades;
  float fades[kMaxFades];
  float* scores;             // [nframes][nfades][countPad]
};

template <typename pixel_t, int PXT>
__global__ void __launch_bounds__(kEvalThreads, 1) logo_scores_kernel(const EvalJob job) {
  extern __shared__ float smem_f[];
  float* src = smem_f;                                           // roi_w*roi_h
  float* work = smem_f + ((job.roi_w * job.roi_h + 3) & ~3);     // logo.w*logo.h (+8 pad)
  const int tid = threadIdx.x;
  const LogoDev& lg = job.logo;
  const int w = lg.w, npx = lg.w * lg.h;

  // ---- one-time: adopt feature pixels, pull their taps into registers ----
  float taps[PXT][25];
  int pxy[PXT];
  int cidx[PXT];
#pragma unroll
  for (int p = 0; p < PXT; ++p) {
    const int c = (blockIdx.x * PXT + p) * kEvalThreads + tid;
    cidx[p] = c;
    if (c < lg.count) {
      const uint32_t v = lg.pix[c];
      pxy[p] = (int)((v & 0xFFFFu) - 2) + (int)((v >> 16) - 2) * w;     // top-left of the 5x5 window
#pragma unroll
      for (int t = 0; t < 25; ++t) taps[p][t] = lg.tapsT[(size_t)t * lg.countPad + c];
    } else {
      pxy[p] = 0;
#pragma unroll
      for (int t = 0; t < 25; ++t) taps[p][t] = 0.0f;
    }
  }

  const int roi_n = job.roi_w * job.roi_h;
  for (int f = blockIdx.y; f < job.nframes; f += gridDim.y) {
    // ---- stage the ROI as float: DeintY (:763-780) or CopyY (:782-790) ----
    const pixel_t* fr = reinterpret_cast<const pixel_t*>(
        reinterpret_cast<const uint8_t*>(job.ybase) + (long long)(job.frame0 + f) * job.frame_stride);
    const pixel_t* roi = fr + job.imgx + (long long)job.imgy * job.pitch;
    for (int i = tid; i < roi_n; i += kEvalThreads) {
      const int y = i / job.roi_w, x = i - y * job.roi_w;
      const pixel_t* p = roi + x + (long long)y * job.pitch;
      float v;
      if (job.src_mode == 0 && y > 0 && y < job.roi_h - 1) {
        const int a = p[-job.pitch], b = p[0], c = p[job.pitch];
        v = (float)(a + 2 * b + c + 2) / 4.0f;       // exact: integer < 2^24, division by 4
      } else {
        v = (float)p[0];
      }
      src[i] = v;
    }
    __syncthreads();

    for (int fi = 0; fi < job.nfades; ++fi) {
      const float fade = job.fades[fi];
      const float omf = AMTK_FSUB(1.0f, fade);
      // ---- logo removal at this fade level (LogoScan.hpp:241-251) ----
      for (int i = tid; i < npx; i += kEvalThreads) {
        const int y = i / w, x = i - y * w;
        const float srcv = src[job.src_off + x + y * job.src_stride];
        work[i] = remove_logo(srcv, __ldg(lg.A + i), __ldg(lg.B + i), job.maxv, fade, omf);
      }
      __syncthreads();
      // ---- per-feature score (LogoScan.hpp:298-308) ----
      float* out = job.scores + ((size_t)f * job.nfades + fi) * lg.countPad;
#pragma unroll
      for (int p = 0; p < PXT; ++p) {
        if (cidx[p] < lg.count) {
          const float* wp = work + pxy[p];
          float avg;
          const float sum = corr5x5_tree(taps[p], [&](int dy, int dx) { return wp[dy * w + dx]; }, &avg);
          const float2 sc = __ldg(lg.scales + (size_t)cidx[p] * 32 + scale_bin(avg));
          out[cidx[p]] = pixel_score(sum, sc.x, sc.y);
        }
      }
      __syncthreads();     // `work` is rewritten by the next fade / `src` by the next frame
    }
  }
}

// One thread per (frame, fade): ordered float sum of the pixel scores, divided by blackScore (:252-254,310).
// out index = frame*out_frame_stride + out_off + fade*out_fade_stride; take_abs for AMTAnalyzeLogo (:1152-1154).
__global__ void __launch_bounds__(128) logo_sum_kernel(const float* __restrict__ scores, int count, int countPad,
                                                       int nframes, int nfades, float blackScore, int take_abs,
                                                       float* __restrict__ out, int out_frame_stride, int out_off,
                                                       int out_fade_stride) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nframes * nfades) return;
  const float4* row = reinterpret_cast<const float4*>(scores + (size_t)t * countPad);
  float r = 0.0f;
  const int n4 = count >> 2;
  int c = 0;
#pragma unroll 4
  for (int i = 0; i < n4; ++i) {
    const float4 v = row[i];
    r = AMTK_FADD(r, v.x); r = AMTK_FADD(r, v.y); r = AMTK_FADD(r, v.z); r = AMTK_FADD(r, v.w);
  }
  c = n4 << 2;
  const float* tail = scores + (size_t)t * countPad;
  for (; c < count; ++c) r = AMTK_FADD(r, tail[c]);
  float v = AMTK_FDIV(r, blackScore);
  if (take_abs) v = fabsf(v);
  const int f = t / nfades, fi = t - f * nfades;
  out[(size_t)f * out_frame_stride + out_off + (size_t)fi * out_fade_stride] = v;
}

__global__ void fill_pairs_kernel(float* out, int nframes, int stride, int off, float v0, float v1) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < nframes) { out[(size_t)f * stride + off] = v0; out[(size_t)f * stride + off + 1] = v1; }
}

}  // namespace amtk
