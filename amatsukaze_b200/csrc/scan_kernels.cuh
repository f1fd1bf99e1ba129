// scan_kernels.cuh -- LogoScan accumulation and logo erase on the GPU.
//
// LogoScan::AddFrame (LogoScan.hpp:594-659): per frame, the ROI border pixels decide whether the background is
// flat (max-min <= thy on Y, U and V) and give the background level (mean of the middle half of the sorted border
// values, :414-428); valid frames add f, bg, f^2, bg^2, f*bg to per-pixel accumulators (LogoColor::Add, :357-364).
// The reference accumulates ints in doubles; every partial sum is an exact integer < 2^53, so exact u64 integer
// accumulation is bit-identical after conversion.  A 256-bin histogram replaces the sort exactly.
#pragma once
#include "amtk_internal.h"
#include "exact_math.h"

namespace amtk {

struct ScanClip {
  const uint8_t* base; long long frame_stride; long long offU, offV;
  int pitchY, pitchUV;
  int scanx, scany, scanw, scanh, logUVx, logUVy, thy;
  int frame0, nframes;
};

// One CTA (256 threads) per frame: border histogram per plane -> {valid, bgY, bgU, bgV}.
__global__ void __launch_bounds__(256) scan_border_kernel(const ScanClip c, const uint8_t* __restrict__ select,
                                                          int4* __restrict__ frame_bg) {
  __shared__ unsigned int hist[3][256];
  __shared__ int res[3][2];
  const int f = blockIdx.x, tid = threadIdx.x;
  if (select && !select[f]) { if (tid == 0) frame_bg[f] = make_int4(0, 0, 0, 0); return; }
  for (int i = tid; i < 3 * 256; i += 256) (&hist[0][0])[i] = 0u;
  __syncthreads();
  const uint8_t* fr = c.base + (long long)(c.frame0 + f) * c.frame_stride;
  for (int pl = 0; pl < 3; ++pl) {
    const int w = pl ? (c.scanw >> c.logUVx) : c.scanw, h = pl ? (c.scanh >> c.logUVy) : c.scanh;
    const int pitch = pl ? c.pitchUV : c.pitchY;
    const uint8_t* p = fr + (pl == 0 ? 0 : (pl == 1 ? c.offU : c.offV)) +
                       (pl ? ((c.scanx >> c.logUVx) + (long long)(c.scany >> c.logUVy) * pitch)
                           : (c.scanx + (long long)c.scany * pitch));
    // border = rows 0 and h-1 (all x) + columns 0 and w-1 for y in [1,h-1)  (:616-635)
    const int nb = 2 * w + 2 * (h - 2);
    for (int i = tid; i < nb; i += 256) {
      int x, y;
      if (i < w) { x = i; y = 0; }
      else if (i < 2 * w) { x = i - w; y = h - 1; }
      else { const int k = i - 2 * w; y = 1 + (k >> 1); x = (k & 1) ? (w - 1) : 0; }
      atomicAdd(&hist[pl][p[x + (long long)y * pitch]], 1u);
    }
  }
  __syncthreads();
  if (tid < 3) {
    const int pl = tid;
    const int w = pl ? (c.scanw >> c.logUVx) : c.scanw, h = pl ? (c.scanh >> c.logUVy) : c.scanh;
    const int n = 2 * w + 2 * (h - 2);
    const int lo = n / 4, hi = n - n / 4;            // sorted ranks [lo, hi) are averaged (:421-423)
    int vmin = -1, vmax = 0, rank = 0;
    long long sum = 0;
    for (int v = 0; v < 256; ++v) {
      const int cnt = (int)hist[pl][v];
      if (cnt) {
        if (vmin < 0) vmin = v;
        vmax = v;
        const int a = max(rank, lo), b = min(rank + cnt, hi);
        if (b > a) sum += (long long)(b - a) * v;
        rank += cnt;
      }
    }
    const int nn = hi - lo;
    res[pl][0] = (vmax - vmin > c.thy) ? 0 : 1;      // abs(front-back) > thy rejects (:639-649)
    res[pl][1] = (int)((sum + nn / 2) / nn);         // (int)((t + nn/2)/nn) on exact integers (:425-427)
  }
  __syncthreads();
  if (tid == 0) frame_bg[f] = make_int4(res[0][0] & res[1][0] & res[2][0], res[0][1], res[1][1], res[2][1]);
}

// grid (pixel blocks, frame splits): thread per ROI pixel (Y then U then V), loops over its share of the frames.
// sums: [npix][3] u64 = sumF, sumF2, sumFB.  plane scalars bgsum[pl*2+{0,1}] = sumB, sumB2; bgsum[6] = nvalid.
__global__ void __launch_bounds__(256) scan_accumulate_kernel(const ScanClip c, const int4* __restrict__ frame_bg,
                                                              unsigned long long* __restrict__ sums,
                                                              unsigned long long* __restrict__ bgsum,
                                                              uint8_t* __restrict__ valid_out) {
  const int ny = c.scanw * c.scanh, wc = c.scanw >> c.logUVx, hc = c.scanh >> c.logUVy, nc = wc * hc;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = (c.nframes + gridDim.y - 1) / gridDim.y;
  const int f_lo = blockIdx.y * per, f_hi = min(c.nframes, f_lo + per);
  if (i < ny + 2 * nc) {
    int pl, x, y, pitch; long long off;
    if (i < ny) { pl = 0; y = i / c.scanw; x = i - y * c.scanw; pitch = c.pitchY; off = c.scanx + (long long)c.scany * pitch; }
    else {
      const int k = (i - ny) % nc; pl = 1 + (i - ny) / nc; y = k / wc; x = k - y * wc; pitch = c.pitchUV;
      off = (pl == 1 ? c.offU : c.offV) + (c.scanx >> c.logUVx) + (long long)(c.scany >> c.logUVy) * pitch;
    }
    const uint8_t* p = c.base + (long long)c.frame0 * c.frame_stride + off + x + (long long)y * pitch;
    unsigned long long sF = 0, sF2 = 0, sFB = 0;
    for (int f = f_lo; f < f_hi; ++f) {
      const int4 bg = frame_bg[f];
      if (bg.x) {
        const unsigned v = p[(long long)f * c.frame_stride];
        const unsigned b = (unsigned)(pl == 0 ? bg.y : (pl == 1 ? bg.z : bg.w));
        sF += v; sF2 += v * v; sFB += v * b;
      }
    }
    if (sF | sF2 | sFB) {
      atomicAdd(&sums[(size_t)i * 3 + 0], sF); atomicAdd(&sums[(size_t)i * 3 + 1], sF2); atomicAdd(&sums[(size_t)i * 3 + 2], sFB);
    }
  }
  // per-plane background sums + valid count: one thread per frame split
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long sb[3] = { 0, 0, 0 }, sb2[3] = { 0, 0, 0 }, nv = 0;
    for (int f = f_lo; f < f_hi; ++f) {
      const int4 bg = frame_bg[f];
      if (valid_out) valid_out[f] = (uint8_t)bg.x;
      if (bg.x) {
        ++nv;
        sb[0] += bg.y; sb2[0] += (unsigned long long)bg.y * bg.y;
        sb[1] += bg.z; sb2[1] += (unsigned long long)bg.z * bg.z;
        sb[2] += bg.w; sb2[2] += (unsigned long long)bg.w * bg.w;
      }
    }
    for (int pl = 0; pl < 3; ++pl) { atomicAdd(&bgsum[pl * 2], sb[pl]); atomicAdd(&bgsum[pl * 2 + 1], sb2[pl]); }
    atomicAdd(&bgsum[6], nv);
  }
}

// ---- AMTEraseLogo::Delogo (LogoScan.hpp:1248-1261) on the Y,U,V ROIs of each frame, in place -------------------
struct EraseJob {
  uint8_t* base; long long frame_stride; long long offU, offV;
  int pitchY, pitchUV;           // ELEMENTS
  int frame0, nframes;
  int w, h, logUVx, logUVy, imgx, imgy;
  int uvparity;                  // ((imgy / 2) % 2) of the logo's REAL frame position (the clip may be an ROI-only staging copy)
  const float *aY, *bY, *aU, *bU, *aV, *bV;
  const float* fades;            // [nframes][2] fadeT, fadeB (device)
  float maxv;
};

template <typename pixel_t>
__global__ void __launch_bounds__(256) erase_logo_kernel(const EraseJob j) {
  const int f = blockIdx.x;
  const float fadeT = j.fades[f * 2], fadeB = j.fades[f * 2 + 1];
  pixel_t* fr = reinterpret_cast<pixel_t*>(j.base + (long long)(j.frame0 + f) * j.frame_stride);
  const int wc = j.w >> j.logUVx, hc = j.h >> j.logUVy, ny = j.w * j.h, nc = wc * hc;
  const bool frame_mode = (fadeT == fadeB);          // :1374
  const int uvparity = j.uvparity;                   // :1385
  for (int i = threadIdx.x; i < ny + 2 * nc; i += blockDim.x) {
    pixel_t* p; float a, b, fade;
    if (i < ny) {
      const int y = i / j.w, x = i - y * j.w;
      p = fr + j.imgx + x + (long long)(j.imgy + y) * j.pitchY;
      if (!frame_mode && y >= 2 * (j.h / 2)) continue;           // field passes cover h/2 rows each (:1380-1381)
      a = j.aY[i]; b = j.bY[i];
      fade = frame_mode ? fadeT : ((y & 1) ? fadeB : fadeT);      // rows of the top field take fadeT (:1380-1381)
    } else {
      const int k = (i - ny) % nc, pl = (i - ny) / nc;
      const int y = k / wc, x = k - y * wc;
      p = reinterpret_cast<pixel_t*>(reinterpret_cast<uint8_t*>(fr) + (pl == 0 ? j.offU : j.offV)) +
          (j.imgx >> j.logUVx) + x + (long long)((j.imgy >> j.logUVy) + y) * j.pitchUV;
      if (!frame_mode && y >= 2 * (hc / 2)) continue;            // hUV/2 rows per field pass (:1391-1395)
      a = (pl == 0 ? j.aU : j.aV)[k]; b = (pl == 0 ? j.bU : j.bV)[k];
      // chroma row y belongs to the top-field group when (y & 1) == uvparity (:1385-1396)
      fade = frame_mode ? fadeT : (((y & 1) == uvparity) ? fadeT : fadeB);
    }
    const float srcv = (float)*p;
    const float tmp = remove_logo(srcv, a, b, j.maxv, fade, AMTK_FSUB(1.0f, fade));
    const float t = AMTK_FADD(tmp, 0.5f);
    const float m = (t > 0.0f) ? t : 0.0f;           // std::max(tmp + 0.5f, 0.0f)
    const float cl = (j.maxv < m) ? j.maxv : m;      // std::min(.., maxv)
    *p = (pixel_t)cl;
  }
}

// ---- AMTSource::MergeField (AMTSource.hpp:291-355): weave two decoded frames, optional NV12 chroma split ------------
struct WeaveJob {
  const uint8_t* src; uint8_t* dst;
  long long sstride, dstride, s_offu, s_offv, d_offu, d_offv;
  int s_pitchY, s_pitchUV, d_pitchY, d_pitchUV;   // BYTES
  int row_bytes_y, row_bytes_c;                   // payload bytes per luma / chroma row (planar)
  int H, HC, bps, nv12;
  const int* top_idx; const int* bot_idx;         // device
  int dst_frame0;
};

// grid (row blocks, 3 planes, frames); each thread moves 16 bytes of a row (tail bytes one by one)
__global__ void __launch_bounds__(256) weave_kernel(const WeaveJob j) {
  const int k = blockIdx.z, pl = blockIdx.y;
  const int rows = pl ? j.HC : j.H;
  const int rb = pl ? j.row_bytes_c : j.row_bytes_y;
  const uint8_t* ft = j.src + (long long)j.top_idx[k] * j.sstride;
  const uint8_t* fb = j.src + (long long)j.bot_idx[k] * j.sstride;
  uint8_t* fd = j.dst + (long long)(j.dst_frame0 + k) * j.dstride;
  const int vec_per_row = (rb + 15) / 16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * vec_per_row;
       i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / vec_per_row), v = (int)(i - (long long)y * vec_per_row);
    const uint8_t* fs = (y & 1) ? fb : ft;             // even rows from `top`, odd rows from `bottom` (Copy1 :292-302)
    const int nbytes = min(16, rb - v * 16);
    if (pl == 0 || !j.nv12) {
      const uint8_t* s = fs + (pl == 0 ? 0 : (pl == 1 ? j.s_offu : j.s_offv)) + (long long)y * (pl ? j.s_pitchUV : j.s_pitchY) + v * 16;
      uint8_t* d = fd + (pl == 0 ? 0 : (pl == 1 ? j.d_offu : j.d_offv)) + (long long)y * (pl ? j.d_pitchUV : j.d_pitchY) + v * 16;
      if (nbytes == 16 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0)
        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
      else
        for (int b = 0; b < nbytes; ++b) d[b] = s[b];
    } else {
      // NV12: interleaved UV row -> U (pl 1) or V (pl 2) samples (Copy2 :304-321)
      const uint8_t* s = fs + j.s_offu + (long long)y * j.s_pitchUV;
      uint8_t* d = fd + (pl == 1 ? j.d_offu : j.d_offv) + (long long)y * j.d_pitchUV + v * 16;
      const int comp = pl - 1;
      for (int b = 0; b < nbytes; b += j.bps) {
        const int xs = (v * 16 + b) / j.bps;           // sample index in the row
        for (int q = 0; q < j.bps; ++q) d[b + q] = s[(xs * 2 + comp) * j.bps + q];
      }
    }
  }
}

// ---- read-bandwidth probe: what a do-nothing streaming read achieves on this GPU ---------------------------------
__global__ void __launch_bounds__(256) read_probe_kernel(const uint4* __restrict__ p, size_t n16, unsigned* __restrict__ sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {          // 4 independent 16-byte loads in flight per thread
    const uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
    acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y; acc.z ^= a.z ^ b.z ^ c.z ^ d.z; acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
  }
  for (; i < n16; i += stride) { const uint4 a = __ldcs(p + i); acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w; }
  const unsigned v = acc.x ^ acc.y ^ acc.z ^ acc.w;
  if (v == 0x9E3779B9u) atomicAdd(sink, 1u);               // keeps the loads alive; practically never taken
}

}  // namespace amtk
