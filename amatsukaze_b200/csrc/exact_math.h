// exact_math.h -- IEEE binary32 primitives that must never be contracted into FMAs.
//
// The reference evaluates its logo arithmetic with separate multiply/add instructions (MSVC /arch:AVX has no
// FMA; ComputeKernel.cpp:88-114 spells the vector ops out).  To reproduce its scores bit for bit the same
// expression trees are written once here and used by BOTH the host table builder (logo_host.cpp) and the
// device kernels (logo_kernels.cu).  Device: round-to-nearest intrinsics (never fused).  Host: plain operators,
// compiled with -ffp-contract=off.
#pragma once
#if defined(__CUDA_ARCH__)
#define AMTK_HD __host__ __device__ __forceinline__
#define AMTK_FADD(a, b) __fadd_rn((a), (b))
#define AMTK_FSUB(a, b) __fsub_rn((a), (b))
#define AMTK_FMUL(a, b) __fmul_rn((a), (b))
#define AMTK_FDIV(a, b) __fdiv_rn((a), (b))
#elif defined(__CUDACC__)
#define AMTK_HD __host__ __device__ __forceinline__
#define AMTK_FADD(a, b) ((a) + (b))
#define AMTK_FSUB(a, b) ((a) - (b))
#define AMTK_FMUL(a, b) ((a) * (b))
#define AMTK_FDIV(a, b) ((a) / (b))
#else
#define AMTK_HD inline
#define AMTK_FADD(a, b) ((a) + (b))
#define AMTK_FSUB(a, b) ((a) - (b))
#define AMTK_FMUL(a, b) ((a) * (b))
#define AMTK_FDIV(a, b) ((a) / (b))
#endif

namespace amtk {

// Horizontal sum of 5 column values exactly as hsum256_ps sees lanes (c0..c4,0,0,0) (ComputeKernel.cpp:54-74):
// sumQuad = (c0+c4, c1+0, c2+0, c3+0); sumDual = (q0+q2, q1+q3); sum = d0+d1.  The "+0" lanes are kept so that
// even the sign of a zero result matches.
AMTK_HD float hsum5_tree(float c0, float c1, float c2, float c3, float c4) {
  float q0 = AMTK_FADD(c0, c4), q1 = AMTK_FADD(c1, 0.0f), q2 = AMTK_FADD(c2, 0.0f), q3 = AMTK_FADD(c3, 0.0f);
  float d0 = AMTK_FADD(q0, q2), d1 = AMTK_FADD(q1, q3);
  return AMTK_FADD(d0, d1);
}

// Zero-mean 5x5 correlation with the AVX summation tree of CalcCorrelation5x5_AVX (ComputeKernel.cpp:77-121).
// RowFn r(dy, dx) returns the image sample at (x-2+dx, y-2+dy); k = 25 kernel taps, row-major.
template <typename RowFn>
AMTK_HD float corr5x5_tree(const float* k, RowFn r, float* pavg) {
  float y[5][5];
#pragma unroll
  for (int dy = 0; dy < 5; ++dy)
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) y[dy][dx] = r(dy, dx);
  float c[5];
#pragma unroll
  for (int j = 0; j < 5; ++j)
    c[j] = AMTK_FADD(AMTK_FADD(AMTK_FADD(y[0][j], y[1][j]), AMTK_FADD(y[2][j], y[3][j])), y[4][j]);
  float avg = AMTK_FDIV(hsum5_tree(c[0], c[1], c[2], c[3], c[4]), 25.0f);
  float p[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    float t0 = AMTK_FMUL(k[j], AMTK_FSUB(y[0][j], avg));
    float t1 = AMTK_FMUL(k[5 + j], AMTK_FSUB(y[1][j], avg));
    float t2 = AMTK_FMUL(k[10 + j], AMTK_FSUB(y[2][j], avg));
    float t3 = AMTK_FMUL(k[15 + j], AMTK_FSUB(y[3][j], avg));
    float t4 = AMTK_FMUL(k[20 + j], AMTK_FSUB(y[4][j], avg));
    p[j] = AMTK_FADD(AMTK_FADD(AMTK_FADD(t0, t1), AMTK_FADD(t2, t3)), t4);
  }
  *pavg = avg;
  return hsum5_tree(p[0], p[1], p[2], p[3], p[4]);
}

// Per-mask-pixel score of LogoDataParam::CorrelationScore (LogoScan.hpp:301-308).
// scale/scale2 come from scales[count*32 + (clamp((int)avg,0,255)>>3)].
AMTK_HD int scale_bin(float avg) {
  int ai = (int)avg;           // truncation toward zero, like the reference's (int)avg
  ai = ai < 0 ? 0 : (ai > 255 ? 255 : ai);
  return ai >> 3;
}
AMTK_HD float pixel_score(float sum, float scale, float scale2) {
  float v = AMTK_FMUL(sum, scale);
  float m = (v < 1.0f) ? v : 1.0f;          // std::min(1.0f, v)
  float n = (-1.0f < m) ? m : -1.0f;        // std::max(-1.0f, m)
  return AMTK_FMUL(n, scale2);
}

// Logo removal at a fade level (LogoScan.hpp:245-250; also Delogo :1254-1257).
AMTK_HD float remove_logo(float srcv, float a, float b, float maxv, float fade, float one_minus_fade) {
  float bg = AMTK_FADD(AMTK_FMUL(a, srcv), AMTK_FMUL(b, maxv));
  return AMTK_FADD(AMTK_FMUL(fade, bg), AMTK_FMUL(one_minus_fade, srcv));
}

}  // namespace amtk
