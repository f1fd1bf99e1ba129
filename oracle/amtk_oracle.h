/* oracle/amtk_oracle.h -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
 *
 * CPU restatement (plain C99) of the reference's per-frame pixel-analysis hot path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product library (amatsukaze_b200/lib/libamtk_b200.so) never links, loads or calls anything here.
 *
 * Parity status
 *   logo half  : PINNED -- every function below is checked bit-for-bit against the reference's own code
 *                compiled from /root/reference (oracle/_ref/libamtk_ref.so, see build_ref.sh) and against the
 *                golden vectors that code produced (tests/golden/, generator tests/golden/gen_golden.py).
 *   combing half: PARITY UNPINNED -- the KFM field-difference/combing arithmetic is not in /root/reference
 *                (external plugin nekopanda/AviSynthCUDAFilters, URL only at README.md:598, no version pin).
 *                amtk_or_comb_* implement THIS REPO's normative integer spec.
 *
 * All reference citations are file:line under /root/reference/Amatsukaze/.
 */
#ifndef AMTK_ORACLE_H
#define AMTK_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { AMTK_OR_KLEN = 25, AMTK_OR_CLEN = 32 };

/* ComputeKernel.cpp:77-121 (AVX summation tree, restated in scalar code) */
float amtk_or_corr5x5(const float* k, const float* Y, int x, int y, int w, float* pavg);
/* LogoScan.hpp:24-41 (scalar twin -- NOT what the reference runs on AVX hosts; kept for the a1 order test) */
float amtk_or_corr5x5_scalar_order(const float* k, const float* Y, int x, int y, int w, float* pavg);

/* LogoScan.hpp:763-790 */
void amtk_or_deint_y_u8(float* dst, const uint8_t* src, int pitch, int w, int h);
void amtk_or_deint_y_u16(float* dst, const uint16_t* src, int pitch, int w, int h);
void amtk_or_copy_y_u8(float* dst, const uint8_t* src, int pitch, int w, int h);
void amtk_or_copy_y_u16(float* dst, const uint16_t* src, int pitch, int w, int h);

/* One evaluation logo = LogoDataParam after CreateLogoMask (LogoScan.hpp:61-334). */
typedef struct amtk_or_logo {
  int w, h, logUVx, logUVy;
  int imgw, imgh, imgx, imgy;
  float* data;            /* aY,bY,aU,bU,aV,bV (AMTLogo.hpp:206-212) */
  float *aY, *bY, *aU, *bU, *aV, *bV;
  /* CreateLogoMask products */
  uint8_t* mask;          /* w*h */
  int maskpixels;         /* min(YSize,(int)(YSize*maskratio))  LogoScan.hpp:172 */
  int count;              /* mask pixels actually visited by the y,x in [2,dim-2) scan (LogoScan.hpp:188-200) */
  float* kernels;         /* count*25 */
  float* scales;          /* count*32*{scale,scale2} */
  float blackScore;
} amtk_or_logo;

amtk_or_logo* amtk_or_logo_new(int w, int h, int logUVx, int logUVy, int imgw, int imgh, int imgx, int imgy,
                               const float* data /* may be NULL: left uninitialised */);
void amtk_or_logo_free(amtk_or_logo* l);
amtk_or_logo* amtk_or_logo_deint(const amtk_or_logo* src);              /* LogoScan.hpp:734-761 */
amtk_or_logo* amtk_or_logo_field(const amtk_or_logo* src, int bottom);  /* LogoScan.hpp:257-283 */
void amtk_or_logo_create_mask(amtk_or_logo* l, float maskratio);        /* LogoScan.hpp:112-229 */
float amtk_or_logo_corr_score(const amtk_or_logo* l, const float* work, float maxv);            /* :288-318 */
float amtk_or_logo_evaluate(const amtk_or_logo* l, const float* src, float maxv, float fade,
                            float* work, int stride /* -1 => w */);                             /* :231-255 */

/* LogoFrame::ScanFrame for one logo (LogoScan.hpp:1543-1568): out[0]=corr0, out[1]=corr1.
 * pitch in ELEMENTS of the pixel type as the caller means it (the reference passes the BYTE pitch even for
 * 16-bit, :1547,1561 -- a caller that wants that quirk passes the byte pitch here too). */
void amtk_or_scan_frame_u8(const amtk_or_logo* deint_logo, const uint8_t* planeY, int pitch, float maxv, float* out2);
void amtk_or_scan_frame_u16(const amtk_or_logo* deint_logo, const uint16_t* planeY, int pitch, float maxv, float* out2);

/* AMTAnalyzeLogo::GetFrameT body for ONE source frame (LogoScan.hpp:1136-1158): out33 = p[11],t[11],b[11]. */
void amtk_or_analyze_frame_u8(const amtk_or_logo* deint_logo, const amtk_or_logo* field_t, const amtk_or_logo* field_b,
                              const uint8_t* planeY, int pitch, float maxv, float* out33);
void amtk_or_analyze_frame_u16(const amtk_or_logo* deint_logo, const amtk_or_logo* field_t, const amtk_or_logo* field_b,
                               const uint16_t* planeY, int pitch, float maxv, float* out33);

/* AMTEraseLogo::Delogo (LogoScan.hpp:1248-1261), in place. */
void amtk_or_delogo_u8(uint8_t* dst, int w, int h, int logopitch, int imgpitch, float maxv,
                       const float* A, const float* B, float fade);
void amtk_or_delogo_u16(uint16_t* dst, int w, int h, int logopitch, int imgpitch, float maxv,
                        const float* A, const float* B, float fade);
/* AMTEraseLogo::GetFrameT mode 0 on one frame given the two fades (LogoScan.hpp:1374-1397). */
void amtk_or_erase_frame_u8(const amtk_or_logo* logo, uint8_t* Y, uint8_t* U, uint8_t* V, int pitchY, int pitchUV,
                            float maxv, float fadeT, float fadeB);
/* AMTEraseLogo::CalcFade2 (LogoScan.hpp:1263-1315).  records = all LogoAnalyzeFrame of the clip, 33 floats each,
 * indexed by SOURCE frame number; num_frames = vi.num_frames of the erase clip. */
void amtk_or_calc_fade2(const float* records, int num_records, int num_frames, int n, float* fadeT, float* fadeB);

/* LogoScan accumulation (LogoScan.hpp:357-364,414-428,568-659). sums: plane-major Y,U,V; per pixel
 * {sumF,sumB,sumF2,sumB2,sumFB} as double exactly like LogoColor. */
typedef struct amtk_or_scan {
  int scanw, scanh, logUVx, logUVy, thy, nframes;
  double* sums;
} amtk_or_scan;
amtk_or_scan* amtk_or_scan_new(int scanw, int scanh, int logUVx, int logUVy, int thy);
void amtk_or_scan_free(amtk_or_scan* s);
int amtk_or_scan_add_frame_u8(amtk_or_scan* s, const uint8_t* y, const uint8_t* u, const uint8_t* v, int pitchY, int pitchUV);
/* LogoScan::Normalize + GetLogo (LogoScan.hpp:367-395,471-566): returns 1 and fills out (LogoData layout) or 0. */
int amtk_or_scan_get_logo(const amtk_or_scan* s, int maxv, int clean, float* out);

/* ---- combing / field-difference metric: THIS REPO's spec (parity unpinned, see header) ----
 * counts[12] = [plane class Y,C][field top,bottom][move, shima, lshima]; th[6] = {thMY,thSY,thLY,thMC,thSC,thLC}. */
void amtk_or_comb_frame_u8(const uint8_t* curY, const uint8_t* curU, const uint8_t* curV,
                           const uint8_t* prevY, const uint8_t* prevU, const uint8_t* prevV,
                           int w, int h, int pitchY, int pitchUV, int logUVx, int logUVy,
                           const int* th6, int32_t* counts12);
void amtk_or_comb_frame_u16(const uint16_t* curY, const uint16_t* curU, const uint16_t* curV,
                            const uint16_t* prevY, const uint16_t* prevU, const uint16_t* prevV,
                            int w, int h, int pitchY, int pitchUV, int logUVx, int logUVy,
                            const int* th6, int32_t* counts12);

/* Bounded CPU-baseline loops used by bench.py (port leg): returns seconds of wall time. */
double amtk_or_bench_scan_comb_u8(const amtk_or_logo* deint_logo, const uint8_t* frames, int nframes,
                                  int w, int h, const int* th6, int nthreads, float* out_scores, int32_t* out_counts);
/* mode: bit 0 = logo ScanFrame, bit 1 = combing counters; comb_impl: 0 = scalar spec, 1 = AVX2 spec (amtk_comb_avx2.c) */
double amtk_or_bench_run(const amtk_or_logo* deint_logo, const uint8_t* frames, int nframes, int w, int h, const int* th6,
                         int nthreads, int mode, int comb_impl, float* out_scores, int32_t* out_counts);
/* AVX2 form of the combing spec (oracle/amtk_comb_avx2.c; NOT Amatsukaze code); same contract as amtk_or_comb_frame_u8 */
void amtk_or_comb_frame_u8_avx2(const uint8_t* curY, const uint8_t* curU, const uint8_t* curV,
                                const uint8_t* prevY, const uint8_t* prevU, const uint8_t* prevV,
                                int w, int h, int pitchY, int pitchUV, int logUVx, int logUVy, const int* th6, int32_t* counts12);
int amtk_or_comb_have_avx2(void);

#ifdef __cplusplus
}
#endif
#endif
