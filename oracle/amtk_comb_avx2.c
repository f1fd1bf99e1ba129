/* oracle/amtk_comb_avx2.c -- TEST INFRASTRUCTURE ONLY (CPU baseline + checker; never linked into the product).
 *
 * AVX2 implementation of THIS REPO's combing / field-difference spec (the scalar normative
 * form is comb_plane_u8 in oracle/amtk_oracle.c).  NOT Amatsukaze code: the reference delegates this arithmetic to
 * an external plugin whose source is absent from /root/reference, so "parity unpinned" applies here
 * exactly as to the scalar spec.  It exists so that bench.py's CPU arm compares the B200 kernel against a vectorised
 * CPU loop rather than against unvectorised C ("What's weak" 2(d)); tests/test_comb_spec.py checks it
 * against the scalar spec bit for bit.
 *
 *   comb(y,x) = | p[y-2] + 4 p[y] + p[y+2] - 3 (p[y-1] + p[y+1]) |        2 <= y < H-2
 *   shima[f] += comb >= thS ; lshima[f] += comb >= thL                     f = y & 1
 *   move[f]  += | cur[y][x] - prev[y][x] | >= thM                          every row
 *
 * Compiled with -mavx2 as its own object; amtk_or_comb_have_avx2() says whether the CPU can run it, and the entry
 * point falls back to the scalar spec when it cannot.
 */
#include <immintrin.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

void amtk_or_comb_frame_u8(const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*,
                           int, int, int, int, int, int, const int*, int32_t*);

int amtk_or_comb_have_avx2(void) { return __builtin_cpu_supports("avx2") ? 1 : 0; }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* 16 pixels: 5 rows -> |a + 4 r + e - 3 (b + d)| as 16 x s16 */
static inline __m256i comb16(const uint8_t* a, const uint8_t* b, const uint8_t* r, const uint8_t* d, const uint8_t* e) {
  const __m256i va = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)a));
  const __m256i vb = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)b));
  const __m256i vr = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)r));
  const __m256i vd = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)d));
  const __m256i ve = _mm256_cvtepu8_epi16(_mm_loadu_si128((const __m128i*)e));
  const __m256i pos = _mm256_add_epi16(_mm256_add_epi16(va, ve), _mm256_slli_epi16(vr, 2));
  const __m256i bd = _mm256_add_epi16(vb, vd);
  const __m256i neg = _mm256_add_epi16(bd, _mm256_slli_epi16(bd, 1));
  return _mm256_abs_epi16(_mm256_sub_epi16(pos, neg));
}

static void comb_plane_u8_avx2(const uint8_t* cur, const uint8_t* prev, int w, int h, int pitch,
                               int thM, int thS, int thL, int32_t* c6) {
  /* byte compare d >= thM as max(d, thM) == d; thM outside 1..255 handled without the vector compare */
  const int move_all = thM <= 0, move_none = thM > 255;
  const __m256i vM = _mm256_set1_epi8((char)clampi(thM, 0, 255));
  /* v >= th  <=>  v > th-1 (s16); v is in [0, 1530] */
  const __m256i vS = _mm256_set1_epi16((short)clampi(thS - 1, -1, 32767));
  const __m256i vL = _mm256_set1_epi16((short)clampi(thL - 1, -1, 32767));
  for (int y = 0; y < h; ++y) {
    const int f = y & 1;
    const uint8_t* r = cur + (size_t)y * pitch;
    const uint8_t* q = prev + (size_t)y * pitch;
    int x = 0;
    long long nm = 0;
    if (move_all) nm = w;
    else if (!move_none) {
      for (; x + 32 <= w; x += 32) {
        const __m256i c = _mm256_loadu_si256((const __m256i*)(r + x));
        const __m256i p = _mm256_loadu_si256((const __m256i*)(q + x));
        const __m256i d = _mm256_or_si256(_mm256_subs_epu8(c, p), _mm256_subs_epu8(p, c));
        const __m256i m = _mm256_cmpeq_epi8(_mm256_max_epu8(d, vM), d);
        nm += __builtin_popcount((unsigned)_mm256_movemask_epi8(m));
      }
      for (; x < w; ++x) { int d = (int)r[x] - (int)q[x]; if (d < 0) d = -d; nm += d >= thM; }
    }
    c6[f * 3 + 0] += (int32_t)nm;
    if (y >= 2 && y < h - 2) {
      const uint8_t* a = r - 2 * (size_t)pitch; const uint8_t* b = r - (size_t)pitch;
      const uint8_t* d1 = r + (size_t)pitch; const uint8_t* e = r + 2 * (size_t)pitch;
      __m256i accS = _mm256_setzero_si256(), accL = _mm256_setzero_si256();     /* 16 x s16 lanes, -1 per hit */
      int lanes_used = 0;
      long long ns = 0, nl = 0;
      for (x = 0; x + 16 <= w; x += 16) {
        const __m256i v = comb16(a + x, b + x, r + x, d1 + x, e + x);
        accS = _mm256_add_epi16(accS, _mm256_cmpgt_epi16(v, vS));
        accL = _mm256_add_epi16(accL, _mm256_cmpgt_epi16(v, vL));
        if (++lanes_used == 30000) {      /* s16 lanes hold up to 32767 hits: flush long before */
          short tS[16], tL[16];
          _mm256_storeu_si256((__m256i*)tS, accS); _mm256_storeu_si256((__m256i*)tL, accL);
          for (int k = 0; k < 16; ++k) { ns -= tS[k]; nl -= tL[k]; }
          accS = _mm256_setzero_si256(); accL = _mm256_setzero_si256(); lanes_used = 0;
        }
      }
      {
        short tS[16], tL[16];
        _mm256_storeu_si256((__m256i*)tS, accS); _mm256_storeu_si256((__m256i*)tL, accL);
        for (int k = 0; k < 16; ++k) { ns -= tS[k]; nl -= tL[k]; }
      }
      for (; x < w; ++x) {
        int v = (int)a[x] + 4 * (int)r[x] + (int)e[x] - 3 * ((int)b[x] + (int)d1[x]);
        if (v < 0) v = -v;
        ns += v >= thS; nl += v >= thL;
      }
      c6[f * 3 + 1] += (int32_t)ns;
      c6[f * 3 + 2] += (int32_t)nl;
    }
  }
}

/* same contract as amtk_or_comb_frame_u8 (oracle/amtk_oracle.c) */
void amtk_or_comb_frame_u8_avx2(const uint8_t* curY, const uint8_t* curU, const uint8_t* curV,
                                const uint8_t* prevY, const uint8_t* prevU, const uint8_t* prevV,
                                int w, int h, int pitchY, int pitchUV, int logUVx, int logUVy,
                                const int* th, int32_t* c) {
  if (!amtk_or_comb_have_avx2()) {
    amtk_or_comb_frame_u8(curY, curU, curV, prevY, prevU, prevV, w, h, pitchY, pitchUV, logUVx, logUVy, th, c);
    return;
  }
  memset(c, 0, 12 * sizeof(int32_t));
  comb_plane_u8_avx2(curY, prevY, w, h, pitchY, th[0], th[1], th[2], c);
  comb_plane_u8_avx2(curU, prevU, w >> logUVx, h >> logUVy, pitchUV, th[3], th[4], th[5], c + 6);
  comb_plane_u8_avx2(curV, prevV, w >> logUVx, h >> logUVy, pitchUV, th[3], th[4], th[5], c + 6);
}
