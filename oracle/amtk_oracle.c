/* oracle/amtk_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product (see amtk_oracle.h).
 *
 * Plain-C restatement of the reference algorithm; every function cites the reference lines it follows
 * (paths relative to /root/reference/Amatsukaze/).  Must be compiled with -ffp-contract=off and without
 * -ffast-math so that float expressions evaluate exactly like the reference's MSVC /fp:precise build
 * (no FMA contraction: Amatsukaze.vcxproj:224-228 enables /arch:AVX only, which has no FMA).
 */
#include "amtk_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------
 * a1. 5x5 zero-mean correlation, AVX summation order (ComputeKernel.cpp:54-121)
 * ---------------------------------------------------------------------------------------------- */
static float hsum5(const float c[5]) {
  /* hsum256_ps (ComputeKernel.cpp:54-74) on lanes (c0..c4,0,0,0):
   * sumQuad = (c0+c4, c1+0, c2+0, c3+0); sumDual = (sumQuad0+sumQuad2, sumQuad1+sumQuad3); sum = dual0+dual1 */
  float q0 = c[0] + c[4], q1 = c[1] + 0.0f, q2 = c[2] + 0.0f, q3 = c[3] + 0.0f;
  float d0 = q0 + q2, d1 = q1 + q3;
  return d0 + d1;
}

float amtk_or_corr5x5(const float* k, const float* Y, int x, int y, int w, float* pavg) {
  const float* r0 = Y + (x - 2) + w * (y - 2);   /* ComputeKernel.cpp:82-86 */
  const float* r1 = r0 + w; const float* r2 = r1 + w; const float* r3 = r2 + w; const float* r4 = r3 + w;
  float c[5], p[5];
  for (int j = 0; j < 5; ++j) c[j] = ((r0[j] + r1[j]) + (r2[j] + r3[j])) + r4[j];   /* :88-94 */
  float avg = hsum5(c);
  avg /= 25;                                                                       /* :96-98 */
  for (int j = 0; j < 5; ++j) {                                                    /* :108-114 */
    float t0 = k[j] * (r0[j] - avg), t1 = k[5 + j] * (r1[j] - avg);
    float t2 = k[10 + j] * (r2[j] - avg), t3 = k[15 + j] * (r3[j] - avg);
    float t4 = k[20 + j] * (r4[j] - avg);
    p[j] = ((t0 + t1) + (t2 + t3)) + t4;
  }
  if (pavg) *pavg = avg;
  return hsum5(p);
}

float amtk_or_corr5x5_scalar_order(const float* k, const float* Y, int x, int y, int w, float* pavg) {
  float avg = 0.0f;                                                                /* LogoScan.hpp:26-32 */
  for (int ky = -2; ky <= 2; ++ky) for (int kx = -2; kx <= 2; ++kx) avg += Y[(x + kx) + (y + ky) * w];
  avg /= 25;
  float sum = 0.0f;                                                                /* :33-38 */
  for (int ky = -2; ky <= 2; ++ky) for (int kx = -2; kx <= 2; ++kx)
    sum += k[(kx + 2) + (ky + 2) * 5] * (Y[(x + kx) + (y + ky) * w] - avg);
  if (pavg) *pavg = avg;
  return sum;
}

/* ------------------------------------------------------------------------------------------------
 * a4. DeintY / CopyY (LogoScan.hpp:763-790)
 * ---------------------------------------------------------------------------------------------- */
#define DEF_DEINT(NAME, T)                                                                     \
  void NAME(float* dst, const T* src, int pitch, int w, int h) {                               \
    for (int x = 0; x < w; ++x) {                                                              \
      dst[x] = src[x];                                                                         \
      dst[x + (h - 1) * w] = src[x + (h - 1) * pitch];                                         \
    }                                                                                          \
    for (int y = 1; y < h - 1; ++y)                                                            \
      for (int x = 0; x < w; ++x) {                                                            \
        int a = src[x + (y - 1) * pitch], b = src[x + y * pitch], c = src[x + (y + 1) * pitch];\
        dst[x + y * w] = (a + 2 * b + c + 2) / 4.0f;                                           \
      }                                                                                        \
  }
DEF_DEINT(amtk_or_deint_y_u8, uint8_t)
DEF_DEINT(amtk_or_deint_y_u16, uint16_t)
#define DEF_COPY(NAME, T)                                                                      \
  void NAME(float* dst, const T* src, int pitch, int w, int h) {                               \
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) dst[x + y * w] = src[x + y * pitch]; \
  }
DEF_COPY(amtk_or_copy_y_u8, uint8_t)
DEF_COPY(amtk_or_copy_y_u16, uint16_t)

/* ------------------------------------------------------------------------------------------------
 * LogoData / LogoDataParam
 * ---------------------------------------------------------------------------------------------- */
amtk_or_logo* amtk_or_logo_new(int w, int h, int logUVx, int logUVy, int imgw, int imgh, int imgx, int imgy,
                               const float* data) {
  amtk_or_logo* l = (amtk_or_logo*)calloc(1, sizeof(*l));
  int wUV = w >> logUVx, hUV = h >> logUVy;
  size_t n = (size_t)(w * h + wUV * hUV * 2) * 2;          /* AMTLogo.hpp:203-212 */
  l->w = w; l->h = h; l->logUVx = logUVx; l->logUVy = logUVy;
  l->imgw = imgw; l->imgh = imgh; l->imgx = imgx; l->imgy = imgy;
  l->data = (float*)malloc(n * sizeof(float));
  if (data) memcpy(l->data, data, n * sizeof(float));
  l->aY = l->data; l->bY = l->aY + w * h; l->aU = l->bY + w * h;
  l->bU = l->aU + wUV * hUV; l->aV = l->bU + wUV * hUV; l->bV = l->aV + wUV * hUV;
  return l;
}
void amtk_or_logo_free(amtk_or_logo* l) {
  if (!l) return;
  free(l->data); free(l->mask); free(l->kernels); free(l->scales); free(l);
}

amtk_or_logo* amtk_or_logo_deint(const amtk_or_logo* s) {            /* LogoScan.hpp:734-761 */
  int w = s->w, h = s->h;
  amtk_or_logo* d = amtk_or_logo_new(w, h, s->logUVx, s->logUVy, s->imgw, s->imgh, s->imgx, s->imgy, NULL);
  /* NOTE: like the reference, only the Y planes are written; chroma of the deint logo stays uninitialised
   * (never read by the evaluation path).  We zero it for determinism. */
  memset(d->aU, 0, (size_t)((w >> s->logUVx) * (h >> s->logUVy)) * 4 * sizeof(float));
  for (int x = 0; x < w; ++x) {
    d->aY[x] = s->aY[x]; d->bY[x] = s->bY[x];
    d->aY[x + (h - 1) * w] = s->aY[x + (h - 1) * w]; d->bY[x + (h - 1) * w] = s->bY[x + (h - 1) * w];
  }
  for (int y = 1; y < h - 1; ++y)
    for (int x = 0; x < w; ++x) {
      d->aY[x + y * w] = (s->aY[x + (y - 1) * w] + 2 * s->aY[x + y * w] + s->aY[x + (y + 1) * w]) / 4.0f;
      d->bY[x + y * w] = (s->bY[x + (y - 1) * w] + 2 * s->bY[x + y * w] + s->bY[x + (y + 1) * w]) / 4.0f;
    }
  return d;
}

amtk_or_logo* amtk_or_logo_field(const amtk_or_logo* s, int bottom) {  /* LogoScan.hpp:257-283 */
  int w = s->w;
  amtk_or_logo* f = amtk_or_logo_new(w, s->h / 2, s->logUVx, s->logUVy, s->imgw, s->imgh / 2, s->imgx, s->imgy / 2, NULL);
  bottom = bottom ? 1 : 0;
  for (int y = 0; y < f->h; ++y)
    for (int x = 0; x < f->w; ++x) {
      f->aY[x + y * w] = s->aY[x + (bottom + y * 2) * w];
      f->bY[x + y * w] = s->bY[x + (bottom + y * 2) * w];
    }
  int UVoffset = bottom ^ (f->imgy % 2);
  int wUV = f->w >> s->logUVx, hUV = f->h >> s->logUVy;
  for (int y = 0; y < hUV; ++y)
    for (int x = 0; x < wUV; ++x) {
      f->aU[x + y * wUV] = s->aU[x + (UVoffset + y * 2) * wUV];
      f->bU[x + y * wUV] = s->bU[x + (UVoffset + y * 2) * wUV];
      f->aV[x + y * wUV] = s->aV[x + (UVoffset + y * 2) * wUV];
      f->bV[x + y * wUV] = s->bV[x + (UVoffset + y * 2) * wUV];
    }
  return f;
}

static void add_logo(const amtk_or_logo* l, float* Y, int maxv) {       /* LogoScan.hpp:320-333 */
  int n = l->w * l->h;
  for (int i = 0; i < n; ++i) {
    float a = l->aY[i], b = l->bY[i];
    if (a > 0) Y[i] = (Y[i] - b * maxv) / a;
  }
}

static void make_kernel(float* k, const float* Y, int x, int y, int w) { /* LogoScan.hpp:135-147 */
  for (int ky = -2; ky <= 2; ++ky) for (int kx = -2; kx <= 2; ++kx)
    k[(kx + 2) + (ky + 2) * 5] = Y[(x + kx) + (y + ky) * w];
  float acc = 0.0f;
  for (int i = 0; i < 25; ++i) acc = acc + k[i];          /* std::accumulate, left to right */
  float avg = acc / 25;
  for (int i = 0; i < 25; ++i) k[i] = k[i] - avg;
}

typedef struct { float var; int idx; } var_pair;
static int cmp_var_desc(const void* pa, const void* pb) {  /* std::greater<pair<float,int>> LogoScan.hpp:169 */
  const var_pair* a = (const var_pair*)pa; const var_pair* b = (const var_pair*)pb;
  if (a->var > b->var) return -1;
  if (a->var < b->var) return 1;
  if (a->idx > b->idx) return -1;
  if (a->idx < b->idx) return 1;
  return 0;
}

float amtk_or_logo_corr_score(const amtk_or_logo* l, const float* work, float maxv) {  /* LogoScan.hpp:288-318 */
  (void)maxv;
  int w = l->w, h = l->h, count = 0;
  float result = 0;
  for (int y = 2; y < h - 2; ++y)
    for (int x = 2; x < w - 2; ++x)
      if (l->mask[x + y * w]) {
        const float* k = &l->kernels[count * 25];
        float avg;
        float sum = amtk_or_corr5x5(k, work, x, y, w, &avg);
        int ai = (int)avg; if (ai > 255) ai = 255; if (ai < 0) ai = 0;                 /* :304 */
        const float* s = &l->scales[(count * 32 + (ai >> 3)) * 2];
        float v = sum * s[0];
        float m = (v < 1.0f) ? v : 1.0f;                /* std::min(1.0f, v)  */
        float normalized = (-1.0f < m) ? m : -1.0f;     /* std::max(-1.0f, m) */
        float score = normalized * s[1];
        result += score;                                /* sequential float sum, :310 */
        ++count;
      }
  return result;
}

void amtk_or_logo_create_mask(amtk_or_logo* l, float maskratio) {          /* LogoScan.hpp:112-229 */
  const float corrLowerLimit = 0.2f;
  int w = l->w, h = l->h, YSize = w * h;
  float* memWork = (float*)malloc(((size_t)YSize * 32 + 8) * sizeof(float));
  for (int c = 0; c < 32; ++c) {                                           /* :128-133 */
    float* slice = memWork + (size_t)c * YSize;
    for (int i = 0; i < YSize; ++i) slice[i] = (float)(c << 3);
    add_logo(l, slice, 255);
  }
  var_pair* variance = (var_pair*)calloc((size_t)YSize, sizeof(var_pair));
  for (int y = 2; y < h - 2; ++y)                                          /* :154-163 */
    for (int x = 2; x < w - 2; ++x) {
      float k[25];
      make_kernel(k, memWork + (size_t)16 * YSize, x, y, w);
      float acc = 0.0f;
      for (int i = 0; i < 25; ++i) acc = acc + k[i] * k[i];
      variance[x + y * w].var = acc;
    }
  for (int i = 0; i < YSize; ++i) variance[i].idx = i;
  qsort(variance, (size_t)YSize, sizeof(var_pair), cmp_var_desc);          /* :169 */
  free(l->mask); free(l->kernels); free(l->scales);
  l->mask = (uint8_t*)calloc((size_t)YSize, 1);
  int maskpixels = (int)(YSize * maskratio);                               /* :172 */
  if (maskpixels > YSize) maskpixels = YSize;
  l->maskpixels = maskpixels;
  for (int i = 0; i < maskpixels; ++i) l->mask[variance[i].idx] = 1;
  free(variance);

  l->kernels = (float*)calloc((size_t)maskpixels * 25 + 8, sizeof(float));
  l->scales = (float*)calloc((size_t)maskpixels * 32 * 2 + 2, sizeof(float));
  int count = 0;
  float avgCorr = 0.0f;
  for (int y = 2; y < h - 2; ++y)                                          /* :188-200 */
    for (int x = 2; x < w - 2; ++x)
      if (l->mask[x + y * w]) {
        float* k = &l->kernels[count * 25];
        make_kernel(k, memWork, x, y, w);
        for (int i = 0; i < 32; ++i) {
          float v = fabsf(amtk_or_corr5x5(k, memWork + (size_t)i * YSize, x, y, w, NULL));
          l->scales[(count * 32 + i) * 2] = v;
          avgCorr += v;
        }
        ++count;
      }
  l->count = count;
  avgCorr /= maskpixels * 32;                                              /* :202 (divides by maskpixels, not count) */
  float limitCorr = avgCorr * corrLowerLimit;
  /* :205-209 -- the reference also walks the never-visited tail [count,maskpixels) of an uninitialised array;
   * those entries are never read afterwards, so only the visited ones are restated. */
  for (int i = 0; i < count * 32; ++i) {
    float corr = l->scales[i * 2];
    l->scales[i * 2] = (corr > 0) ? (1.0f / corr) : 0.0f;
    float q = corr / limitCorr;
    l->scales[i * 2 + 1] = (q < 1.0f) ? q : 1.0f;                          /* std::min(1.0f, q) */
  }
  l->blackScore = amtk_or_logo_corr_score(l, memWork + (size_t)2 * YSize, 255);   /* :227-228 */
  free(memWork);
}

float amtk_or_logo_evaluate(const amtk_or_logo* l, const float* src, float maxv, float fade,
                            float* work, int stride) {                     /* LogoScan.hpp:231-255 */
  int w = l->w, h = l->h;
  if (stride == -1) stride = w;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float srcv = src[x + y * stride];
      float a = l->aY[x + y * w], b = l->bY[x + y * w];
      float bg = a * srcv + b * maxv;
      float dstv = fade * bg + (1 - fade) * srcv;
      work[x + y * w] = dstv;
    }
  return amtk_or_logo_corr_score(l, work, maxv) / l->blackScore;
}

/* ------------------------------------------------------------------------------------------------
 * a6. LogoFrame::ScanFrame (LogoScan.hpp:1543-1568), one logo
 * ---------------------------------------------------------------------------------------------- */
#define DEF_SCAN_FRAME(NAME, T, DEINT)                                                          \
  void NAME(const amtk_or_logo* lg, const T* planeY, int pitch, float maxv, float* out2) {      \
    int n = lg->w * lg->h;                                                                      \
    float* deint = (float*)malloc(((size_t)n + 8) * sizeof(float));                             \
    float* work = (float*)malloc(((size_t)n + 8) * sizeof(float));                              \
    memset(deint + n, 0, 8 * sizeof(float)); memset(work + n, 0, 8 * sizeof(float));            \
    int off = lg->imgx + lg->imgy * pitch;                                                      \
    DEINT(deint, planeY + off, pitch, lg->w, lg->h);                                            \
    out2[0] = amtk_or_logo_evaluate(lg, deint, maxv, 0, work, -1);                              \
    out2[1] = amtk_or_logo_evaluate(lg, deint, maxv, 1, work, -1);                              \
    free(deint); free(work);                                                                    \
  }
DEF_SCAN_FRAME(amtk_or_scan_frame_u8, uint8_t, amtk_or_deint_y_u8)
DEF_SCAN_FRAME(amtk_or_scan_frame_u16, uint16_t, amtk_or_deint_y_u16)

/* ------------------------------------------------------------------------------------------------
 * a7. AMTAnalyzeLogo::GetFrameT, one source frame (LogoScan.hpp:1136-1158)
 * ---------------------------------------------------------------------------------------------- */
#define DEF_ANALYZE(NAME, T, DEINT, COPY)                                                       \
  void NAME(const amtk_or_logo* dl, const amtk_or_logo* ft, const amtk_or_logo* fb,             \
            const T* planeY, int pitch, float maxv, float* out33) {                             \
    int w = dl->w, h = dl->h, n = w * h;                                                        \
    float* copy = (float*)calloc((size_t)n + 8, sizeof(float));                                 \
    float* deint = (float*)calloc((size_t)n + 8, sizeof(float));                                \
    float* work = (float*)calloc((size_t)n + 8, sizeof(float));                                 \
    int off = dl->imgx + dl->imgy * pitch;                                                      \
    COPY(copy, planeY + off, pitch, w, h);                                                      \
    DEINT(deint, planeY + off, pitch, w, h);                                                    \
    for (int f = 0; f <= 10; ++f) {                                                             \
      float fade = (float)f / 10.0f;                                                            \
      out33[f] = fabsf(amtk_or_logo_evaluate(dl, deint, maxv, fade, work, -1));                 \
      out33[11 + f] = fabsf(amtk_or_logo_evaluate(ft, copy, maxv, fade, work, w * 2));          \
      out33[22 + f] = fabsf(amtk_or_logo_evaluate(fb, copy + w, maxv, fade, work, w * 2));      \
    }                                                                                           \
    free(copy); free(deint); free(work);                                                        \
  }
DEF_ANALYZE(amtk_or_analyze_frame_u8, uint8_t, amtk_or_deint_y_u8, amtk_or_copy_y_u8)
DEF_ANALYZE(amtk_or_analyze_frame_u16, uint16_t, amtk_or_deint_y_u16, amtk_or_copy_y_u16)

/* ------------------------------------------------------------------------------------------------
 * a8. AMTEraseLogo (LogoScan.hpp:1248-1419)
 * ---------------------------------------------------------------------------------------------- */
#define DEF_DELOGO(NAME, T)                                                                     \
  void NAME(T* dst, int w, int h, int logopitch, int imgpitch, float maxv,                      \
            const float* A, const float* B, float fade) {                                       \
    for (int y = 0; y < h; ++y)                                                                 \
      for (int x = 0; x < w; ++x) {                                                             \
        float srcv = dst[x + y * imgpitch];                                                     \
        float a = A[x + y * logopitch], b = B[x + y * logopitch];                               \
        float bg = a * srcv + b * maxv;                                                         \
        float tmp = fade * bg + (1 - fade) * srcv;                                              \
        float t = tmp + 0.5f;                                                                   \
        float m = (t > 0.0f) ? t : 0.0f;              /* std::max(tmp+0.5f, 0.0f) */            \
        float c = (maxv < m) ? maxv : m;              /* std::min(.., maxv)       */            \
        dst[x + y * imgpitch] = (T)c;                                                           \
      }                                                                                         \
  }
DEF_DELOGO(amtk_or_delogo_u8, uint8_t)
DEF_DELOGO(amtk_or_delogo_u16, uint16_t)

void amtk_or_erase_frame_u8(const amtk_or_logo* lg, uint8_t* Y, uint8_t* U, uint8_t* V, int pitchY, int pitchUV,
                            float maxv, float fadeT, float fadeB) {        /* LogoScan.hpp:1354-1397 */
  int off = lg->imgx + lg->imgy * pitchY;
  int offUV = (lg->imgx >> lg->logUVx) + (lg->imgy >> lg->logUVy) * pitchUV;
  int w = lg->w, h = lg->h, wUV = w >> lg->logUVx, hUV = h >> lg->logUVy;
  if (fadeT == fadeB) {
    amtk_or_delogo_u8(Y + off, w, h, w, pitchY, maxv, lg->aY, lg->bY, fadeT);
    amtk_or_delogo_u8(U + offUV, wUV, hUV, wUV, pitchUV, maxv, lg->aU, lg->bU, fadeT);
    amtk_or_delogo_u8(V + offUV, wUV, hUV, wUV, pitchUV, maxv, lg->aV, lg->bV, fadeT);
  } else {
    amtk_or_delogo_u8(Y + off, w, h / 2, w * 2, pitchY * 2, maxv, lg->aY, lg->bY, fadeT);
    amtk_or_delogo_u8(Y + off + pitchY, w, h / 2, w * 2, pitchY * 2, maxv, lg->aY + w, lg->bY + w, fadeB);
    int uvparity = ((lg->imgy / 2) % 2);
    int tuvoff = uvparity * pitchUV, buvoff = !uvparity * pitchUV;
    int tuvoffl = uvparity * wUV, buvoffl = !uvparity * wUV;
    amtk_or_delogo_u8(U + offUV + tuvoff, wUV, hUV / 2, wUV * 2, pitchUV * 2, maxv, lg->aU + tuvoffl, lg->bU + tuvoffl, fadeT);
    amtk_or_delogo_u8(V + offUV + tuvoff, wUV, hUV / 2, wUV * 2, pitchUV * 2, maxv, lg->aV + tuvoffl, lg->bV + tuvoffl, fadeT);
    amtk_or_delogo_u8(U + offUV + buvoff, wUV, hUV / 2, wUV * 2, pitchUV * 2, maxv, lg->aU + buvoffl, lg->bU + buvoffl, fadeB);
    amtk_or_delogo_u8(V + offUV + buvoff, wUV, hUV / 2, wUV * 2, pitchUV * 2, maxv, lg->aV + buvoffl, lg->bV + buvoffl, fadeB);
  }
}

static int argmin11(const float* v) {   /* std::min_element: first minimum */
  int best = 0;
  for (int i = 1; i < 11; ++i) if (v[i] < v[best]) best = i;
  return best;
}

void amtk_or_calc_fade2(const float* records, int num_records, int num_frames, int n, float* fadeT, float* fadeB) {
  /* LogoScan.hpp:1263-1315.  The analyze clip's frame k holds records 8k..8k+7 (clamped to the last source
   * frame, :1133); record (nsrc+i) is therefore min(nsrc+i, num_records-1) with num_records = source frames.
   * Quirk kept: the index is nsrc+i with nsrc=clamp(n+i), i.e. effectively n+2i (:1273-1275). */
  enum { DIST = 4 };
  const float* frames[DIST * 2 + 1];
  for (int i = -DIST; i <= DIST; ++i) {
    int nsrc = n + i; if (nsrc > num_frames - 1) nsrc = num_frames - 1; if (nsrc < 0) nsrc = 0;
    int rec = nsrc + i;
    /* analyze_n=(rec)>>3 may address a frame outside the analyze clip; AviSynth clamps GetFrame to the
     * clip range, and AMTAnalyzeLogo clamps its source index (:1133).  rec<0 only arises as (-1..-4)>>3=-1
     * -> clamped to analyze frame 0 with idx=(rec&7). */
    int analyze_n = rec >> 3, idx = rec & 7;
    int nblk = (num_records + 7) / 8;
    if (analyze_n < 0) analyze_n = 0; if (analyze_n > nblk - 1) analyze_n = nblk - 1;
    int src = analyze_n * 8 + idx; if (src > num_records - 1) src = num_records - 1; if (src < 0) src = 0;
    frames[i + DIST] = records + (size_t)src * 33;
  }
  int minfades[DIST * 2 + 1];
  for (int i = 0; i < DIST * 2 + 1; ++i) minfades[i] = argmin11(frames[i]);
  int minT = argmin11(frames[DIST] + 11), minB = argmin11(frames[DIST] + 22);
  float before_fades = 0, after_fades = 0;
  for (int i = 1; i <= 4; ++i) { before_fades += minfades[DIST - i]; after_fades += minfades[DIST + i]; }
  before_fades /= 4 * 10; after_fades /= 4 * 10;
  if ((before_fades < 0.3 && after_fades > 0.7) || (before_fades > 0.7 && after_fades < 0.3)) {
    *fadeT = minT / 10.0f; *fadeB = minB / 10.0f;
  } else {
    *fadeT = *fadeB = (minfades[DIST] / 10.0f);
  }
}

/* ------------------------------------------------------------------------------------------------
 * a9/a10. LogoScan (LogoScan.hpp:336-660)
 * ---------------------------------------------------------------------------------------------- */
amtk_or_scan* amtk_or_scan_new(int scanw, int scanh, int logUVx, int logUVy, int thy) {
  amtk_or_scan* s = (amtk_or_scan*)calloc(1, sizeof(*s));
  s->scanw = scanw; s->scanh = scanh; s->logUVx = logUVx; s->logUVy = logUVy; s->thy = thy;
  int ny = scanw * scanh, nc = ny >> (logUVx + logUVy);
  s->sums = (double*)calloc((size_t)(ny + 2 * nc) * 5, sizeof(double));
  return s;
}
void amtk_or_scan_free(amtk_or_scan* s) { if (s) { free(s->sums); free(s); } }

static int cmp_short(const void* a, const void* b) { return (int)*(const short*)a - (int)*(const short*)b; }
static int med_average(const short* s, int n) {            /* LogoScan.hpp:414-428 */
  double t = 0; int nn = 0;
  for (int i = n / 4; i < n - (n / 4); i++, nn++) t += s[i];
  t = (t + nn / 2) / nn;
  return (int)t;
}
static void color_add(double* c, int f, int b) {            /* LogoColor::Add :357-364 */
  c[0] += f; c[1] += b; c[2] += f * f; c[3] += b * b; c[4] += f * b;
}

int amtk_or_scan_add_frame_u8(amtk_or_scan* s, const uint8_t* Y, const uint8_t* U, const uint8_t* V, int pitchY, int pitchUV) {
  int scanw = s->scanw, scanh = s->scanh;                   /* LogoScan.hpp:594-659 */
  int uw = scanw >> s->logUVx, uh = scanh >> s->logUVy;
  short* tY = (short*)malloc(sizeof(short) * (size_t)(scanw + scanh) * 2);
  short* tU = (short*)malloc(sizeof(short) * (size_t)(uw + uh) * 2);
  short* tV = (short*)malloc(sizeof(short) * (size_t)(uw + uh) * 2);
  int ny = 0, nu = 0, nv = 0, ok = 1;
  for (int x = 0; x < scanw; ++x) { tY[ny++] = Y[x]; tY[ny++] = Y[x + (scanh - 1) * pitchY]; }
  for (int y = 1; y < scanh - 1; ++y) { tY[ny++] = Y[y * pitchY]; tY[ny++] = Y[scanw - 1 + y * pitchY]; }
  for (int x = 0; x < uw; ++x) {
    tU[nu++] = U[x]; tU[nu++] = U[x + (uh - 1) * pitchUV];
    tV[nv++] = V[x]; tV[nv++] = V[x + (uh - 1) * pitchUV];
  }
  for (int y = 1; y < uh - 1; ++y) {
    tU[nu++] = U[y * pitchUV]; tU[nu++] = U[uw - 1 + y * pitchUV];
    tV[nv++] = V[y * pitchUV]; tV[nv++] = V[uw - 1 + y * pitchUV];
  }
  qsort(tY, (size_t)ny, sizeof(short), cmp_short);
  if (abs(tY[0] - tY[ny - 1]) > s->thy) ok = 0;
  if (ok) { qsort(tU, (size_t)nu, sizeof(short), cmp_short); if (abs(tU[0] - tU[nu - 1]) > s->thy) ok = 0; }
  if (ok) { qsort(tV, (size_t)nv, sizeof(short), cmp_short); if (abs(tV[0] - tV[nv - 1]) > s->thy) ok = 0; }
  if (ok) {
    int bgY = med_average(tY, ny), bgU = med_average(tU, nu), bgV = med_average(tV, nv);
    double* sy = s->sums; double* su = sy + (size_t)scanw * scanh * 5; double* sv = su + (size_t)uw * uh * 5;
    for (int y = 0; y < scanh; ++y) for (int x = 0; x < scanw; ++x) color_add(sy + (size_t)(x + y * scanw) * 5, Y[x + y * pitchY], bgY);
    for (int y = 0; y < uh; ++y) for (int x = 0; x < uw; ++x) {
      color_add(su + (size_t)(x + y * uw) * 5, U[x + y * pitchUV], bgU);
      color_add(sv + (size_t)(x + y * uw) * 5, V[x + y * pitchUV], bgV);
    }
    s->nframes++;
  }
  free(tY); free(tU); free(tV);
  return ok;
}

static void approxim_line(int n, double sum_x, double sum_y, double sum_x2, double sum_xy, double* a, double* b) {
  double temp = (double)n * sum_x2 - sum_x * sum_x;        /* LogoScan.hpp:336-342 */
  *a = ((double)n * sum_xy - sum_x * sum_y) / temp;
  *b = (sum_x2 * sum_y - sum_x * sum_xy) / temp;
}
static int get_ab(const double* c5, int maxv, int n, float* A, float* B) {
  /* LogoColor::Normalize :367-374 then GetAB :380-395 */
  double sumF = c5[0] / (double)maxv, sumB = c5[1] / (double)maxv;
  double sumF2 = c5[2] / ((double)maxv * maxv), sumB2 = c5[3] / ((double)maxv * maxv), sumFB = c5[4] / ((double)maxv * maxv);
  double A1, A2, B1, B2;
  approxim_line(n, sumF, sumB, sumF2, sumFB, &A1, &B1);
  approxim_line(n, sumB, sumF, sumB2, sumFB, &A2, &B2);
  *A = (float)((A1 + (1 / A2)) / 2);
  *B = (float)((B1 + (-B2 / A2)) / 2);
  if (isnan(*A) || isnan(*B) || isinf(*A) || isinf(*B) || *A == 0) return 0;
  return 1;
}
static float calc_dist(float a, float b) {                  /* LogoScan.hpp:430-432 */
  return (1.0f / 3.0f) * (a - 1) * (a - 1) + (a - 1) * b + b * b;
}

int amtk_or_scan_get_logo(const amtk_or_scan* s, int maxv, int clean, float* out) {   /* LogoScan.hpp:490-566 */
  int scanw = s->scanw, scanh = s->scanh, uw = scanw >> s->logUVx, uh = scanh >> s->logUVy;
  int ny = scanw * scanh, nc = uw * uh;
  float *aY = out, *bY = aY + ny, *aU = bY + ny, *bU = aU + nc, *aV = bU + nc, *bV = aV + nc;
  const double* sy = s->sums; const double* su = sy + (size_t)ny * 5; const double* sv = su + (size_t)nc * 5;
  for (int i = 0; i < ny; ++i) if (!get_ab(sy + (size_t)i * 5, maxv, s->nframes, &aY[i], &bY[i])) return 0;
  for (int i = 0; i < nc; ++i) {
    if (!get_ab(su + (size_t)i * 5, maxv, s->nframes, &aU[i], &bU[i])) return 0;
    if (!get_ab(sv + (size_t)i * 5, maxv, s->nframes, &aV[i], &bV[i])) return 0;
  }
  if (clean) {
    /* :536-561.  The three maxfilter() calls (:544-546) only write their scratch buffer `work` and never
     * copy it back into `dist` (:434-454) -- a no-op in the reference, so it is a no-op here. */
    float* dist = (float*)malloc(sizeof(float) * (size_t)ny);
    for (int y = 0; y < scanh; ++y) for (int x = 0; x < scanw; ++x) {
      int off = x + y * scanw, offUV = (x >> s->logUVx) + (y >> s->logUVy) * uw;
      float d = calc_dist(aY[off], bY[off]) + calc_dist(aU[offUV], bU[offUV]) + calc_dist(aV[offUV], bV[offUV]);
      d *= 1000;
      dist[off] = d;
    }
    for (int y = 0; y < scanh; ++y) for (int x = 0; x < scanw; ++x) {
      int off = x + y * scanw, offUV = (x >> s->logUVx) + (y >> s->logUVy) * uw;
      if (dist[off] < 0.3f) { aY[off] = 1; bY[off] = 0; aU[offUV] = 1; bU[offUV] = 0; aV[offUV] = 1; bV[offUV] = 0; }
    }
    free(dist);
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------------
 * a15. combing / field-difference metric -- THIS REPO's normative spec. PARITY UNPINNED.
 *   comb(y,x) = | p[y-2] + 4 p[y] + p[y+2] - 3 (p[y-1] + p[y+1]) |        for 2 <= y < H-2
 *   shima[f]  += comb >= thS ; lshima[f] += comb >= thL                    f = y & 1 (0 top field, 1 bottom)
 *   move[f]   += | p_n[y][x] - p_{n-1}[y][x] | >= thM                      all rows; caller passes prev=cur for n=0
 * ---------------------------------------------------------------------------------------------- */
#define DEF_COMB_PLANE(NAME, T)                                                                  \
  static void NAME(const T* cur, const T* prev, int w, int h, int pitch,                         \
                   int thM, int thS, int thL, int32_t* c6) {                                     \
    for (int y = 0; y < h; ++y) {                                                                \
      int f = y & 1;                                                                             \
      const T* r = cur + (size_t)y * pitch; const T* q = prev + (size_t)y * pitch;               \
      for (int x = 0; x < w; ++x) {                                                              \
        int d = (int)r[x] - (int)q[x]; if (d < 0) d = -d;                                        \
        if (d >= thM) c6[f * 3 + 0]++;                                                           \
      }                                                                                          \
      if (y >= 2 && y < h - 2) {                                                                 \
        const T* a = r - 2 * (size_t)pitch; const T* b = r - (size_t)pitch;                      \
        const T* d1 = r + (size_t)pitch; const T* e = r + 2 * (size_t)pitch;                     \
        for (int x = 0; x < w; ++x) {                                                            \
          int v = (int)a[x] + 4 * (int)r[x] + (int)e[x] - 3 * ((int)b[x] + (int)d1[x]);          \
          if (v < 0) v = -v;                                                                     \
          if (v >= thS) c6[f * 3 + 1]++;                                                         \
          if (v >= thL) c6[f * 3 + 2]++;                                                         \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
  }
DEF_COMB_PLANE(comb_plane_u8, uint8_t)
DEF_COMB_PLANE(comb_plane_u16, uint16_t)

void amtk_or_comb_frame_u8(const uint8_t* curY, const uint8_t* curU, const uint8_t* curV,
                           const uint8_t* prevY, const uint8_t* prevU, const uint8_t* prevV,
                           int w, int h, int pitchY, int pitchUV, int logUVx, int logUVy,
                           const int* th, int32_t* c) {
  memset(c, 0, 12 * sizeof(int32_t));
  comb_plane_u8(curY, prevY, w, h, pitchY, th[0], th[1], th[2], c);
  comb_plane_u8(curU, prevU, w >> logUVx, h >> logUVy, pitchUV, th[3], th[4], th[5], c + 6);
  comb_plane_u8(curV, prevV, w >> logUVx, h >> logUVy, pitchUV, th[3], th[4], th[5], c + 6);
}
void amtk_or_comb_frame_u16(const uint16_t* curY, const uint16_t* curU, const uint16_t* curV,
                            const uint16_t* prevY, const uint16_t* prevU, const uint16_t* prevV,
                            int w, int h, int pitchY, int pitchUV, int logUVx, int logUVy,
                            const int* th, int32_t* c) {
  memset(c, 0, 12 * sizeof(int32_t));
  comb_plane_u16(curY, prevY, w, h, pitchY, th[0], th[1], th[2], c);
  comb_plane_u16(curU, prevU, w >> logUVx, h >> logUVy, pitchUV, th[3], th[4], th[5], c + 6);
  comb_plane_u16(curV, prevV, w >> logUVx, h >> logUVy, pitchUV, th[3], th[4], th[5], c + 6);
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline loop for bench.py ("port" leg): ScanFrame (1 logo) + comb metric per frame, frames are
 * tightly packed YV12 (Y w*h, U, V (w/2)*(h/2)).  Threads split the frame range (OpenMP).
 * ---------------------------------------------------------------------------------------------- */
static double now_sec(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

void amtk_or_comb_frame_u8_avx2(const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*,
                                int, int, int, int, int, int, const int*, int32_t*);   /* oracle/amtk_comb_avx2.c */

/* mode: bit 0 = logo ScanFrame, bit 1 = combing counters; comb_impl: 0 scalar spec, 1 AVX2 spec.  The thread team is
 * spun up before the clock starts. */
double amtk_or_bench_run(const amtk_or_logo* lg, const uint8_t* frames, int nframes, int w, int h, const int* th6,
                         int nthreads, int mode, int comb_impl, float* out_scores, int32_t* out_counts) {
  size_t ysz = (size_t)w * h, csz = (size_t)(w / 2) * (h / 2), fsz = ysz + 2 * csz;
  volatile int sink = 0;
#pragma omp parallel num_threads(nthreads)
  { sink += 1; }
  double t0 = now_sec();
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int n = 0; n < nframes; ++n) {
    const uint8_t* cur = frames + (size_t)n * fsz;
    const uint8_t* prev = frames + (size_t)(n > 0 ? n - 1 : 0) * fsz;
    if (mode & 1) amtk_or_scan_frame_u8(lg, cur, w, 255.0f, out_scores + (size_t)n * 2);
    if (mode & 2) {
      if (comb_impl == 1)
        amtk_or_comb_frame_u8_avx2(cur, cur + ysz, cur + ysz + csz, prev, prev + ysz, prev + ysz + csz,
                                   w, h, w, w / 2, 1, 1, th6, out_counts + (size_t)n * 12);
      else
        amtk_or_comb_frame_u8(cur, cur + ysz, cur + ysz + csz, prev, prev + ysz, prev + ysz + csz,
                              w, h, w, w / 2, 1, 1, th6, out_counts + (size_t)n * 12);
    }
  }
  return now_sec() - t0;
}

double amtk_or_bench_scan_comb_u8(const amtk_or_logo* lg, const uint8_t* frames, int nframes,
                                  int w, int h, const int* th6, int nthreads, float* out_scores, int32_t* out_counts) {
  return amtk_or_bench_run(lg, frames, nframes, w, h, th6, nthreads, 3, 0, out_scores, out_counts);
}
