/* oracle/ref_glue.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * extern "C" handles around the reference's own classes, which build_ref.sh extracts VERBATIM by line
 * range from /root/reference into oracle/_ref/ref_extract.inc (git-ignored, never committed):
 *   AMTLogo.hpp:17-282      LogoHeader, LogoData (+ .lgd Save/Load)
 *   LogoScan.hpp:24-45      scalar CalcCorrelation5x5 + decl of the AVX one (ComputeKernel.cpp:77-121)
 *   LogoScan.hpp:59-660     LogoDataParam, approxim_line, LogoColor, LogoScan
 *   LogoScan.hpp:734-790    DeintLogo, DeintY, CopyY
 *   LogoScan.hpp:1100-1103, 1248-1341, 1421-1461   LogoAnalyzeFrame, AMTEraseLogo::Delogo / CalcFade2 / CalcFade / ReadLogoFrameFile
 *                                                  (oracle/_ref/ref_erase.inc)
 *   AMTSource.hpp:291-355   AMTSource::Copy1 / Copy2 / MergeField (oracle/_ref/ref_mergefield.inc)
 *   LogoScan.hpp:1119-1161, 1543-1568   AMTAnalyzeLogo::GetFrameT, LogoFrame::ScanFrame (ref_analyze.inc, ref_scanframe.inc)
 *   FilteredSource.hpp:163-188,197-210,645-660,663-666   readTimecodeFile + base-fps estimate, AMTDecimate ctor body + GetFrame
 * This file adds no arithmetic of its own: every function forwards to the extracted code.
 * `private`/`protected` are opened (oracle build only) so the tests can read the tables the reference
 * keeps private (scales, blackScore, LogoColor sums) -- SURVEY.md Appendix A item 4. */
#include "ref_shim.h"
#define private public
#define protected public
#define class struct   /* members before the first access specifier are private in the reference's classes */
#include "ref_extract.inc"
#undef class
#undef private
#undef protected

using namespace logo;

/* ---- LogoFrame::selectLogo / writeResult (LogoScan.hpp:1645-1827), verbatim inside a shim class that only supplies
 * the members those two functions touch (numLogos, evalResults, numFrames, framesPerSec, bestLogo, logoRatio, ctx). */
#include <cstdarg>
struct RefCtxShim { void debugF(const char*, ...) {} };
class StringBuilder {
  std::string s_;
public:
  template <typename... Args> StringBuilder& append(const char* fmt, Args const&... args) {
    char buf[256]; snprintf(buf, sizeof(buf), fmt, args...); s_ += buf; return *this;
  }
  MemoryChunk getMC() { return MemoryChunk((uint8_t*)s_.data(), s_.size()); }
};
struct RefLogoFrame {
  RefCtxShim ctx;
  int numLogos = 0, numFrames = 0, framesPerSec = 30;
  struct EvalResult { float corr0, corr1; };
  std::unique_ptr<EvalResult[]> evalResults;
  const float THRESH = 0.2f;
  int bestLogo = -1;
  float logoRatio = 0;
#include "ref_logoframe.inc"
};

extern "C" {

/* ---- ComputeKernel.cpp / LogoScan.hpp:24-41 ---- */
float ref_corr5x5_avx(const float* k, const float* Y, int x, int y, int w, float* pavg) {
  return CalcCorrelation5x5_AVX(k, Y, x, y, w, pavg);
}
float ref_corr5x5_scalar(const float* k, const float* Y, int x, int y, int w, float* pavg) {
  return CalcCorrelation5x5(k, Y, x, y, w, pavg);
}
int ref_is_avx(void) { return IsAVXAvailable() ? 1 : 0; }

/* ---- LogoDataParam ---- */
/* data = aY,bY,aU,bU,aV,bV contiguous, exactly LogoData's own layout (AMTLogo.hpp:206-212) */
void* ref_logo_create(int w, int h, int logUVx, int logUVy, int imgw, int imgh, int imgx, int imgy, const float* data) {
  LogoData d(w, h, logUVx, logUVy);
  size_t n = (size_t)(w * h + (w >> logUVx) * (h >> logUVy) * 2) * 2;
  memcpy(d.data.get(), data, n * sizeof(float));
  return new LogoDataParam(std::move(d), imgw, imgh, imgx, imgy);
}
void ref_logo_free(void* p) { delete (LogoDataParam*)p; }
/* LogoFrame ctor / AMTAnalyzeLogo ctor: deint logo built from a loaded logo (LogoScan.hpp:1605-1607,1177-1179) */
void* ref_logo_deint(void* src) {
  LogoDataParam* s = (LogoDataParam*)src;
  LogoDataParam* d = new LogoDataParam(LogoData(s->w, s->h, s->logUVx, s->logUVy), s->imgw, s->imgh, s->imgx, s->imgy);
  DeintLogo(*d, *s, s->w, s->h);
  return d;
}
void* ref_logo_field(void* src, int bottom) { return ((LogoDataParam*)src)->MakeFieldLogo(bottom != 0).release(); }
void ref_logo_create_mask(void* p, float maskratio) { ((LogoDataParam*)p)->CreateLogoMask(maskratio); }
void ref_logo_dims(void* p, int* out10) {
  LogoDataParam* s = (LogoDataParam*)p;
  int v[10] = { s->w, s->h, s->logUVx, s->logUVy, s->imgw, s->imgh, s->imgx, s->imgy, s->maskpixels, 0 };
  memcpy(out10, v, sizeof(v));
}
void ref_logo_get_data(void* p, float* out) {
  LogoDataParam* s = (LogoDataParam*)p;
  size_t n = (size_t)(s->w * s->h + (s->w >> s->logUVx) * (s->h >> s->logUVy) * 2) * 2;
  memcpy(out, s->data.get(), n * sizeof(float));
}
float ref_logo_black_score(void* p) { return ((LogoDataParam*)p)->blackScore; }
void ref_logo_get_mask(void* p, uint8_t* out) { LogoDataParam* s = (LogoDataParam*)p; memcpy(out, s->mask.get(), (size_t)s->w * s->h); }
/* only the first `count` (visited) entries are initialised by the reference (LogoScan.hpp:186-200) */
void ref_logo_get_kernels(void* p, float* out, int count) { memcpy(out, ((LogoDataParam*)p)->kernels.get(), (size_t)count * 25 * sizeof(float)); }
void ref_logo_get_scales(void* p, float* out, int count) { memcpy(out, ((LogoDataParam*)p)->scales.get(), (size_t)count * 32 * 2 * sizeof(float)); }
float ref_logo_evaluate(void* p, const float* src, float maxv, float fade, float* work, int stride) {
  return ((LogoDataParam*)p)->EvaluateLogo(src, maxv, fade, work, stride);
}
float ref_logo_corr_score(void* p, const float* work, float maxv) { return ((LogoDataParam*)p)->CorrelationScore(work, maxv); }

/* ---- .lgd I/O (AMTLogo.hpp:239-279) ---- */
int ref_logo_save(void* p, const char* path, int imgw, int imgh, int imgx, int imgy, const char* name, int serviceId) {
  LogoDataParam* s = (LogoDataParam*)p;
  try {
    LogoHeader hd(s->w, s->h, s->logUVx, s->logUVy, imgw, imgh, imgx, imgy, name);
    hd.serviceId = serviceId;
    s->Save(path, &hd);
    return 1;
  } catch (const IOException&) { return 0; }
}
void* ref_logo_load(const char* path, void* header540) {
  try {
    LogoHeader hd;
    LogoData d = LogoData::Load(path, &hd);
    if (header540) memcpy(header540, &hd, sizeof(hd));
    return new LogoDataParam(std::move(d), &hd);
  } catch (const IOException&) { return nullptr; }
}
int ref_sizeof(int which) {
  switch (which) { case 0: return (int)sizeof(LogoHeader); case 1: return (int)sizeof(LOGO_FILE_HEADER);
                   case 2: return (int)sizeof(LOGO_HEADER); case 3: return (int)sizeof(LOGO_PIXEL); }
  return -1;
}

/* ---- DeintY / CopyY (LogoScan.hpp:763-790) ---- */
void ref_deint_y_u8(float* dst, const uint8_t* src, int pitch, int w, int h) { DeintY<uint8_t>(dst, src, pitch, w, h); }
void ref_deint_y_u16(float* dst, const uint16_t* src, int pitch, int w, int h) { DeintY<uint16_t>(dst, src, pitch, w, h); }
void ref_copy_y_u8(float* dst, const uint8_t* src, int pitch, int w, int h) { CopyY<uint8_t>(dst, src, pitch, w, h); }
void ref_copy_y_u16(float* dst, const uint16_t* src, int pitch, int w, int h) { CopyY<uint16_t>(dst, src, pitch, w, h); }

/* ---- LogoScan (LogoScan.hpp:398-660) ---- */
void* ref_scan_create(int scanw, int scanh, int logUVx, int logUVy, int thy) { return new LogoScan(scanw, scanh, logUVx, logUVy, thy); }
void ref_scan_free(void* p) { delete (LogoScan*)p; }
int ref_scan_add_frame_u8(void* p, const uint8_t* y, const uint8_t* u, const uint8_t* v, int pitchY, int pitchUV) {
  return ((LogoScan*)p)->AddFrame<uint8_t>(y, u, v, pitchY, pitchUV) ? 1 : 0;
}
int ref_scan_add_frame_u16(void* p, const uint16_t* y, const uint16_t* u, const uint16_t* v, int pitchY, int pitchUV) {
  return ((LogoScan*)p)->AddFrame<uint16_t>(y, u, v, pitchY, pitchUV) ? 1 : 0;
}
int ref_scan_nframes(void* p) { return ((LogoScan*)p)->nframes; }
/* raw accumulators, plane-major (Y then U then V), 5 doubles per pixel: sumF,sumB,sumF2,sumB2,sumFB */
void ref_scan_get_sums(void* p, double* out) {
  LogoScan* s = (LogoScan*)p;
  int ny = s->scanw * s->scanh, nc = ny >> (s->logUVx + s->logUVy);
  const LogoColor* planes[3] = { s->logoY.get(), s->logoU.get(), s->logoV.get() };
  int cnt[3] = { ny, nc, nc };
  for (int pl = 0; pl < 3; ++pl) for (int i = 0; i < cnt[pl]; ++i) {
    const LogoColor& c = planes[pl][i];
    *out++ = c.sumF; *out++ = c.sumB; *out++ = c.sumF2; *out++ = c.sumB2; *out++ = c.sumFB;
  }
}
/* overwrite accumulators (lets a test finalise GPU-produced integer sums through the reference's own GetLogo) */
void ref_scan_set_sums(void* p, const double* in, int nframes) {
  LogoScan* s = (LogoScan*)p;
  int ny = s->scanw * s->scanh, nc = ny >> (s->logUVx + s->logUVy);
  LogoColor* planes[3] = { s->logoY.get(), s->logoU.get(), s->logoV.get() };
  int cnt[3] = { ny, nc, nc };
  for (int pl = 0; pl < 3; ++pl) for (int i = 0; i < cnt[pl]; ++i) {
    LogoColor& c = planes[pl][i];
    c.sumF = *in++; c.sumB = *in++; c.sumF2 = *in++; c.sumB2 = *in++; c.sumFB = *in++;
  }
  s->nframes = nframes;
}
void ref_scan_normalize(void* p, int maxv) { ((LogoScan*)p)->Normalize(maxv); }
/* returns 1 and fills out (LogoData layout) or 0 when the reference returns nullptr (LogoScan.hpp:503,508-509) */
int ref_scan_get_logo(void* p, int clean, float* out) {
  LogoScan* s = (LogoScan*)p;
  std::unique_ptr<LogoData> d = s->GetLogo(clean != 0);
  if (!d) return 0;
  size_t n = (size_t)(s->scanw * s->scanh + (s->scanw >> s->logUVx) * (s->scanh >> s->logUVy) * 2) * 2;
  memcpy(out, d->data.get(), n * sizeof(float));
  return 1;
}

/* ---- LogoFrame post-processing ---- */
int ref_logoframe_write(const float* eval /* [numFrames][numLogos][2] */, int numFrames, int numLogos, int framesPerSec,
                        int numCandidates, const char* outpath, int* bestLogo, float* logoRatio) {
  try {
    RefLogoFrame lf;
    lf.numLogos = numLogos; lf.numFrames = numFrames; lf.framesPerSec = framesPerSec;
    lf.evalResults.reset(new RefLogoFrame::EvalResult[(size_t)numFrames * numLogos]);
    memcpy(lf.evalResults.get(), eval, sizeof(float) * 2 * (size_t)numFrames * numLogos);
    lf.selectLogo(numCandidates);
    if (bestLogo) *bestLogo = lf.bestLogo;
    if (logoRatio) *logoRatio = lf.logoRatio;
    if (outpath) lf.writeResult(outpath);
    return 1;
  } catch (const IOException&) { return 0; }
}

} /* extern "C" */

/* ---- AMTEraseLogo::Delogo / CalcFade2 (LogoScan.hpp:1248-1315), verbatim inside a shim class.  The shim supplies what the two
 * member functions touch: vi.num_frames, and an analyze clip whose GetFrame(n) hands out block n of a caller-provided array of
 * LogoAnalyzeFrame records (8 per block, as AMTAnalyzeLogo lays them out, LogoScan.hpp:1128-1157). */
#include <climits>
#include "ref_analyzeframe.inc"
struct RefShimFrameObj { const uint8_t* p; const uint8_t* GetReadPtr() const { return p; } };
struct RefShimPFrame {
  RefShimFrameObj o{ nullptr };
  RefShimPFrame() {}
  RefShimPFrame(std::nullptr_t) {}
  explicit RefShimPFrame(const uint8_t* p) { o.p = p; }
  const RefShimFrameObj* operator->() const { return &o; }
};
struct RefShimAnalyzeClip {
  const LogoAnalyzeFrame* records; int nblocks;
  RefShimPFrame GetFrame(int n, void*) const { return RefShimPFrame(reinterpret_cast<const uint8_t*>(records + (size_t)std::max(0, std::min(nblocks - 1, n)) * 8)); }
};
struct RefShimAvsError { std::string msg; };
struct RefShimEnv {
  void ThrowError(const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    throw RefShimAvsError{ buf };
  }
};
#include <regex>
struct RefEraseShim {
  struct { int num_frames; } vi;
  const RefShimAnalyzeClip* analyzeclip;
  std::vector<int> frameResult;            /* members of AMTEraseLogo the extracted functions use (LogoScan.hpp:1242-1246) */
  int maxFadeLength;
#define PVideoFrame RefShimPFrame
#define IScriptEnvironment2 void
#define IScriptEnvironment RefShimEnv
#include "ref_erase.inc"
#undef PVideoFrame
#undef IScriptEnvironment2
#undef IScriptEnvironment
};
extern "C" {
void ref_delogo_u8(uint8_t* dst, int w, int h, int logopitch, int imgpitch, float maxv, const float* A, const float* B, float fade) {
  RefEraseShim s; s.Delogo<uint8_t>(dst, w, h, logopitch, imgpitch, maxv, A, B, fade);
}
void ref_delogo_u16(uint16_t* dst, int w, int h, int logopitch, int imgpitch, float maxv, const float* A, const float* B, float fade) {
  RefEraseShim s; s.Delogo<uint16_t>(dst, w, h, logopitch, imgpitch, maxv, A, B, fade);
}
/* records: [nblocks][8] LogoAnalyzeFrame = float[nblocks*8][33] */
void ref_calc_fade2(const float* records, int nblocks, int num_frames, int n, float* fadeT, float* fadeB) {
  RefShimAnalyzeClip clip{ reinterpret_cast<const LogoAnalyzeFrame*>(records), nblocks };
  RefEraseShim s; s.vi.num_frames = num_frames; s.analyzeclip = &clip;
  s.CalcFade2(n, *fadeT, *fadeB, nullptr);
}
/* AMTEraseLogo's fade selection for every frame: ReadLogoFrameFile (when logof_path is given) then CalcFade(n) for n in [0, num_frames).
 * frame_result (optional, [num_frames]) receives the 0/1/2 state table.  Returns 1, or 0 with the ThrowError text in err. */
int ref_erase_fades(const float* records, int nblocks, int num_frames, const char* logof_path, int maxFadeLength,
                    float* out, int* frame_result, char* err, int errlen) {
  RefShimAnalyzeClip clip{ reinterpret_cast<const LogoAnalyzeFrame*>(records), nblocks };
  RefEraseShim s; s.vi.num_frames = num_frames; s.analyzeclip = &clip; s.maxFadeLength = maxFadeLength;
  try {
    RefShimEnv env;
    if (logof_path) s.ReadLogoFrameFile(logof_path, &env);
    for (int n = 0; n < num_frames; ++n) s.CalcFade(n, out[2 * n], out[2 * n + 1], nullptr);
    if (frame_result && logof_path) memcpy(frame_result, s.frameResult.data(), sizeof(int) * (size_t)num_frames);
    return 1;
  } catch (const RefShimAvsError& e) {
    if (err && errlen > 0) { strncpy(err, e.msg.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; }
    return 0;
  }
}

} /* extern "C" */

/* ---- AMTSource::Copy1 / Copy2 / MergeField (AMTSource.hpp:291-355), verbatim inside a shim class ---- */
enum AVPixelFormat { AV_PIX_FMT_YUV420P = 0, AV_PIX_FMT_NV12 = 23 };
struct AVPixFmtDescriptor { int log2_chroma_w, log2_chroma_h; };
static const AVPixFmtDescriptor* av_pix_fmt_desc_get(AVPixelFormat) { static const AVPixFmtDescriptor d = { 1, 1 }; return &d; }   /* 4:2:0 */
struct AVFrame { uint8_t* data[8]; int linesize[8]; int format; };
struct RefShimWFrameObj   /* PLANAR_Y/U/V come from ref_shim.h */ {
  uint8_t* p[3]; int pitch[3];
  uint8_t* GetWritePtr(int pl) { return p[pl == PLANAR_Y ? 0 : pl == PLANAR_U ? 1 : 2]; }
  int GetPitch(int pl) { return pitch[pl == PLANAR_Y ? 0 : pl == PLANAR_U ? 1 : 2]; }
};
struct RefShimWFrame { RefShimWFrameObj o; RefShimWFrameObj* operator->() { return &o; } };
struct RefSourceShim {
  struct { int width, height; } vi;
#define PVideoFrame RefShimWFrame
#include "ref_mergefield.inc"
#undef PVideoFrame
};
extern "C" {
/* 8-bit 4:2:0: planar source (nv12 = 0: tU/tV and bU/bV separate planes) or NV12 (nv12 = 1: tU / bU point at the interleaved plane) */
void ref_merge_field_u8(uint8_t* dY, uint8_t* dU, uint8_t* dV, int dpY, int dpUV,
                        const uint8_t* tY, const uint8_t* tU, const uint8_t* tV, int tpY, int tpUV,
                        const uint8_t* bY, const uint8_t* bU, const uint8_t* bV, int bpY, int bpUV, int w, int h, int nv12) {
  AVFrame top = {}, bottom = {};
  top.data[0] = (uint8_t*)tY; top.data[1] = (uint8_t*)tU; top.data[2] = (uint8_t*)tV; top.linesize[0] = tpY; top.linesize[1] = top.linesize[2] = tpUV;
  bottom.data[0] = (uint8_t*)bY; bottom.data[1] = (uint8_t*)bU; bottom.data[2] = (uint8_t*)bV; bottom.linesize[0] = bpY; bottom.linesize[1] = bottom.linesize[2] = bpUV;
  top.format = bottom.format = nv12 ? AV_PIX_FMT_NV12 : AV_PIX_FMT_YUV420P;
  RefShimWFrame dst; dst.o.p[0] = dY; dst.o.p[1] = dU; dst.o.p[2] = dV; dst.o.pitch[0] = dpY; dst.o.pitch[1] = dst.o.pitch[2] = dpUV;
  RefSourceShim s; s.vi.width = w; s.vi.height = h;
  s.MergeField<uint8_t>(dst, &top, &bottom);
}

} /* extern "C" */

/* ---- AMTAnalyzeLogo::GetFrameT (LogoScan.hpp:1119-1161) and LogoFrame::ScanFrame (:1543-1568), verbatim inside shim classes.
 * The stand-ins supply exactly the members the two functions touch: a frame with GetReadPtr/GetPitch/GetWritePtr, a child clip
 * handing out packed planar frames, NewVideoFrame returning the caller's output block, VideoInfo with the fields read. */
struct RefShimVideoInfo { int width, height, num_frames, bits; int BitsPerComponent() const { return bits; } };
struct RefShimRFrameObj {
  const uint8_t* p[3]; int pitch[3]; uint8_t* w;
  const uint8_t* GetReadPtr(int pl = PLANAR_Y) const { return p[pl == PLANAR_Y ? 0 : pl == PLANAR_U ? 1 : 2]; }
  int GetPitch(int pl = PLANAR_Y) const { return pitch[pl == PLANAR_Y ? 0 : pl == PLANAR_U ? 1 : 2]; }
  uint8_t* GetWritePtr() { return w; }
};
struct RefShimRFrame { RefShimRFrameObj o; RefShimRFrameObj* operator->() { return &o; } };
struct RefShimChild {                       /* packed planar 4:2:0 frames, tight pitches, bytes per sample bps */
  const uint8_t* base; long long frame_stride; int w, h, bps;
  RefShimRFrame GetFrame(int n, void*) const {
    RefShimRFrame f; const uint8_t* fr = base + (long long)n * frame_stride;
    f.o.p[0] = fr; f.o.p[1] = fr + (size_t)w * h * bps; f.o.p[2] = f.o.p[1] + (size_t)(w / 2) * (h / 2) * bps;
    f.o.pitch[0] = w * bps; f.o.pitch[1] = f.o.pitch[2] = (w / 2) * bps; f.o.w = nullptr;
    return f;
  }
};
struct RefShimEnv2 {
  uint8_t* out;
  RefShimRFrame NewVideoFrame(const RefShimVideoInfo&) { RefShimRFrame f; f.o.p[0] = f.o.p[1] = f.o.p[2] = nullptr; f.o.pitch[0] = f.o.pitch[1] = f.o.pitch[2] = 256; f.o.w = out; return f; }
};
struct RefAnalyzeShim {
  RefShimVideoInfo vi, srcvi;
  const RefShimChild* child;
  std::unique_ptr<LogoDataParam> deintLogo, fieldLogoT, fieldLogoB;       /* borrowed: released again before destruction */
  LogoHeader header;
#define PVideoFrame RefShimRFrame
#define IScriptEnvironment2 RefShimEnv2
#include "ref_analyze.inc"
#undef PVideoFrame
#undef IScriptEnvironment2
};
struct RefDeintArr { LogoDataParam** p; LogoDataParam& operator[](int i) { return *p[i]; } };
struct RefScanShim {
  int numLogos; RefShimVideoInfo vi; RefDeintArr deintArr;
  struct EvalResult { float corr0, corr1; };
#define PVideoFrame RefShimRFrame
#include "ref_scanframe.inc"
#undef PVideoFrame
};
extern "C" {
/* out: [8][33] floats = the 8 LogoAnalyzeFrame records of analyze frame n (LogoScan.hpp:1128-1157) */
void ref_analyze_getframe(void* deint, void* top, void* bottom, int imgx, int imgy, const void* frames, int num_frames, int w, int h,
                          int bits, int n, float* out) {
  RefShimChild child{ (const uint8_t*)frames, (long long)w * h * 3 / 2 * (bits > 8 ? 2 : 1), w, h, bits > 8 ? 2 : 1 };
  RefAnalyzeShim s;
  s.child = &child; s.srcvi = RefShimVideoInfo{ w, h, num_frames, bits }; s.vi = s.srcvi;
  s.deintLogo.reset((LogoDataParam*)deint); s.fieldLogoT.reset((LogoDataParam*)top); s.fieldLogoB.reset((LogoDataParam*)bottom);
  LogoDataParam* d = (LogoDataParam*)deint;
  s.header = LogoHeader(d->getWidth(), d->getHeight(), d->getLogUVx(), d->getLogUVy(), w, h, imgx, imgy, "");
  RefShimEnv2 env{ (uint8_t*)out };
  if (bits > 8) s.GetFrameT<uint16_t>(n, &env); else s.GetFrameT<uint8_t>(n, &env);
  s.deintLogo.release(); s.fieldLogoT.release(); s.fieldLogoB.release();
}
/* out: [numLogos][2]; logos[i] may be NULL (an invalid logo, LogoScan.hpp:1551-1558).  pitch_y as the reference passes it (BYTES, :1547) */
void ref_scan_frame(void** logos, int numLogos, const void* frame, int w, int h, int bits, float* memDeint, float* memWork, float* out) {
  static LogoDataParam invalid;
  std::vector<LogoDataParam*> arr(numLogos);
  for (int i = 0; i < numLogos; ++i) arr[i] = logos[i] ? (LogoDataParam*)logos[i] : &invalid;
  RefShimChild child{ (const uint8_t*)frame, 0, w, h, bits > 8 ? 2 : 1 };
  RefShimRFrame f = child.GetFrame(0, nullptr);
  RefScanShim s; s.numLogos = numLogos; s.vi = RefShimVideoInfo{ w, h, 1, bits }; s.deintArr = RefDeintArr{ arr.data() };
  float maxv = (float)((1 << bits) - 1);
  if (bits > 8) s.ScanFrame<uint16_t>(f, memDeint, memWork, maxv, reinterpret_cast<RefScanShim::EvalResult*>(out));
  else s.ScanFrame<uint8_t>(f, memDeint, memWork, maxv, reinterpret_cast<RefScanShim::EvalResult*>(out));
}

} /* extern "C" */

/* ---- FilteredSource.hpp: readTimecodeFile (:163-188) + the base-fps estimate (:197-210), AMTDecimate ctor body (:645-660) + GetFrame
 * (:663-666), verbatim inside shim classes ---- */
#include <numeric>
static inline const std::string& to_tstring(const std::string& s) { return s; }
struct RefTimecodeShim {
  std::vector<double> timeCodes_;
  int vfrTimingFps_ = 0;
#include "ref_timecode.inc"
  void estimateFps() {
#include "ref_timecode_fps.inc"
  }
};
struct RefDecimateChild { int GetFrame(int n, void*) { return n; } };
struct RefDecimateShim {
  std::vector<int> durations, framesMap;
  struct { int num_frames; } vi;
  RefDecimateChild childobj, *child = &childobj;
  void construct(const std::string& duration, RefShimEnv* env) {
#include "ref_decimate_ctor.inc"
  }
#define PVideoFrame int
#define __stdcall
#define IScriptEnvironment RefShimEnv
#include "ref_decimate_getframe.inc"
#undef PVideoFrame
#undef __stdcall
#undef IScriptEnvironment
};
extern "C" {
/* returns the number of time codes (incl. the total), -1 when the file cannot be opened; *fps = vfrTimingFps (0 = no grid fits better) */
int ref_read_timecode(const char* path, double* out, int cap, int* fps) {
  try {
    RefTimecodeShim s; s.readTimecodeFile(path);
    if (!s.timeCodes_.empty()) s.estimateFps();
    int n = (int)s.timeCodes_.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = s.timeCodes_[i];
    *fps = s.vfrTimingFps_;
    return n;
  } catch (const IOException&) { return -1; }
}
/* AMTDecimate on a source of num_source_frames: writes the source frame of every output frame; returns the output frame count,
 * -1 unreadable file, -2 with the ThrowError text in err */
int ref_decimate_map(const char* duration_path, int num_source_frames, int* map, int cap, char* err, int errlen) {
  try {
    RefDecimateShim s; s.vi.num_frames = num_source_frames;
    RefShimEnv env;
    s.construct(duration_path, &env);
    for (int i = 0; i < s.vi.num_frames && i < cap; ++i) map[i] = s.GetFrame(i, &env);
    return s.vi.num_frames;
  } catch (const IOException&) { return -1; }
  catch (const RefShimAvsError& e) { if (err && errlen > 0) { strncpy(err, e.msg.c_str(), (size_t)errlen - 1); err[errlen - 1] = 0; } return -2; }
}

/* ---- CPU baseline loop for bench.py (--impl reference and the cpu_baseline / parity legs) ---------------------
 * Per frame: LogoFrame::ScanFrame (LogoScan.hpp:1559-1566) with the reference's OWN DeintY + EvaluateLogo
 * (incl. CalcCorrelation5x5_AVX), followed by the combing metric, which the reference does not contain and is
 * therefore this repo's spec -- scalar (amtk_or_comb_frame_u8) or its AVX2 form (amtk_or_comb_frame_u8_avx2), both
 * linked in from oracle/.  frames: packed YV12.  OpenMP over independent frames.
 *
 * A bench handle owns the thread team and every per-thread scratch buffer, so the timed region (ref_bench_run)
 * contains no thread creation and no allocation.  EvaluateLogo only writes its `work` argument,
 * so all threads share the one LogoDataParam. */
void amtk_or_comb_frame_u8(const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*,
                           int, int, int, int, int, int, const int*, int32_t*);
void amtk_or_comb_frame_u8_avx2(const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*,
                                int, int, int, int, int, int, const int*, int32_t*);
}
#include <time.h>
#include <omp.h>
struct RefBench {
  LogoDataParam* lg; int w, h, nthreads;
  std::vector<std::vector<float>> deint, work;
};
extern "C" {
void* ref_bench_create(void* deint_logo, int w, int h, int nthreads) {
  RefBench* b = new RefBench();
  b->lg = (LogoDataParam*)deint_logo; b->w = w; b->h = h; b->nthreads = nthreads < 1 ? 1 : nthreads;
  b->deint.resize(b->nthreads); b->work.resize(b->nthreads);
  const size_t n = (size_t)b->lg->w * b->lg->h + 8;
  volatile int sink = 0;
#pragma omp parallel num_threads(b->nthreads)
  {   /* spins the team up and makes every thread touch its own scratch (first-touch placement) */
    const int t = omp_get_thread_num();
    b->deint[t].assign(n, 0.0f); b->work[t].assign(n, 0.0f);
    sink += (int)b->deint[t][0];
  }
  return b;
}
void ref_bench_free(void* p) { delete (RefBench*)p; }
int ref_bench_threads(void* p) { return ((RefBench*)p)->nthreads; }
/* mode: bit 0 = logo ScanFrame (reference code), bit 1 = combing counters; comb_impl: 0 scalar spec, 1 AVX2 spec.
 * Returns wall seconds of one pass over nframes frames. */
double ref_bench_run(void* p, const uint8_t* frames, int nframes, const int* th6, int mode, int comb_impl,
                     float* out_scores, int32_t* out_counts) {
  RefBench* b = (RefBench*)p;
  LogoDataParam* lg = b->lg;
  const int w = b->w, h = b->h;
  const size_t ysz = (size_t)w * h, csz = (size_t)(w / 2) * (h / 2), fsz = ysz + 2 * csz;
  const int off = lg->imgx + lg->imgy * w;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel num_threads(b->nthreads)
  {
    const int t = omp_get_thread_num();
    float* deint = b->deint[t].data(); float* work = b->work[t].data();
#pragma omp for schedule(static)
    for (int i = 0; i < nframes; ++i) {
      const uint8_t* cur = frames + (size_t)i * fsz;
      const uint8_t* prev = frames + (size_t)(i > 0 ? i - 1 : 0) * fsz;
      if (mode & 1) {
        DeintY<uint8_t>(deint, cur + off, w, lg->w, lg->h);
        out_scores[(size_t)i * 2 + 0] = lg->EvaluateLogo(deint, 255.0f, 0, work);
        out_scores[(size_t)i * 2 + 1] = lg->EvaluateLogo(deint, 255.0f, 1, work);
      }
      if (mode & 2) {
        if (comb_impl == 1)
          amtk_or_comb_frame_u8_avx2(cur, cur + ysz, cur + ysz + csz, prev, prev + ysz, prev + ysz + csz,
                                     w, h, w, w / 2, 1, 1, th6, out_counts + (size_t)i * 12);
        else
          amtk_or_comb_frame_u8(cur, cur + ysz, cur + ysz + csz, prev, prev + ysz, prev + ysz + csz,
                                w, h, w, w / 2, 1, 1, th6, out_counts + (size_t)i * 12);
      }
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* one-shot form kept for older callers: create + run + free (allocation and team start-up are NOT timed) */
double ref_bench_scan_comb_u8(void* deint_logo, const uint8_t* frames, int nframes, int w, int h, const int* th6,
                              int nthreads, float* out_scores, int32_t* out_counts) {
  void* b = ref_bench_create(deint_logo, w, h, nthreads);
  const double s = ref_bench_run(b, frames, nframes, th6, 3, 0, out_scores, out_counts);
  ref_bench_free(b);
  return s;
}

} /* extern "C" */
