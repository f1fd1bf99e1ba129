// filters.hpp -- host-side (C++) mirror of the reference's filter / in-process interface for the hot path, written
// against avs_compat.h and calling the CUDA library ONLY through the C ABI (include/amtk_b200.h).
//
//   av::AMTSource          frame provider behind IClip::GetFrame           (reference AMTSource.hpp:721-830,873-882)
//   logo::AMTAnalyzeLogo   8 x LogoAnalyzeFrame{p,t,b} per output frame    (reference LogoScan.hpp:1100-1236)
//   logo::AMTEraseLogo     CalcFade / CalcFade2 / Delogo                   (reference LogoScan.hpp:1238-1519)
//   logo::LogoFrame        scanFrames / selectLogo / writeResult           (reference LogoScan.hpp:1521-1836;
//                                                                            the CMAnalyze entry, CMAnalyze.hpp:291-311)
//   AMTCombAnalyze + ReadAllFrames   the telecine pre-pass pull loop       (reference FilteredSource.hpp:417-439,519-544;
//                                                                            the arithmetic lives in the external KFM plugin)
//   AvisynthPluginInit3    registration with the reference's names/arg specs (reference Amatsukaze.cpp:43-66)
//
// Same names, argument meaning and error behaviour as the reference; the bodies are new: frames live in HBM, every
// per-pixel loop is a CUDA kernel, and whole-clip passes are single batched calls instead of per-frame loops.
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <regex>
#include <stdexcept>
#include "avs_compat.h"
#include "../../include/amtk_b200.h"

// ---------------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------------
typedef std::string tstring;

struct AMTContext {                                      // logging subset of StreamUtils.hpp:343-511
  bool quiet = true;
  void infoF(const char* fmt, ...) const { if (quiet) return; va_list ap; va_start(ap, fmt); fprintf(stderr, "AMT [info] "); vfprintf(stderr, fmt, ap); fputc('\n', stderr); va_end(ap); }
  void debugF(const char* fmt, ...) const { if (quiet) return; va_list ap; va_start(ap, fmt); fprintf(stderr, "AMT [debug] "); vfprintf(stderr, fmt, ap); fputc('\n', stderr); va_end(ap); }
  void info(const char* s) const { if (!quiet) fprintf(stderr, "AMT [info] %s\n", s); }
};

struct IOException : std::runtime_error { using std::runtime_error::runtime_error; };

inline void amtk_check(int ok, IScriptEnvironment* env) {      // C-ABI failure -> AvisynthError, like env->ThrowError
  if (!ok) env->ThrowError("%s", amtk_last_error());
}
inline int nblocks(int n, int block) { return (n + block - 1) / block; }

// A clip whose frames are resident in HBM can hand filters a descriptor for batched processing.
class IDeviceClip {
public:
  virtual ~IDeviceClip() {}
  virtual bool GetDeviceClip(amtk_clip* out) = 0;
};

inline amtk_clip HostFrameClip(const PVideoFrame& f, const VideoInfo& vi) {     // one CPU frame as a 1-frame host clip
  amtk_clip c; memset(&c, 0, sizeof(c));
  c.base = f->Base(); c.frame_stride = (int64_t)((f->TotalBytes() + 15) & ~(size_t)15);
  c.off_u = (int64_t)f->GetOffset(PLANAR_U); c.off_v = (int64_t)f->GetOffset(PLANAR_V);
  c.width = vi.width; c.height = vi.height; c.pitch_y = f->GetPitch(PLANAR_Y); c.pitch_uv = f->GetPitch(PLANAR_U);
  c.log_uvx = c.log_uvy = 1; c.bytes_per_sample = vi.ComponentSize(); c.bits_per_sample = vi.BitsPerComponent();
  c.num_frames = 1; c.on_device = 0;
  return c;
}

namespace av {

// ---------------------------------------------------------------------------------------------------------------
// AMTSource: frame provider.  The reference decodes MPEG2/H.264 with FFmpeg into CPU frames on demand
// (AMTSource.hpp:585-780); decode is out of scope here, so the source is a raw planar clip file (the stand-in for
// the `amts%d.dat` artefact, AMTSource.hpp:835-871) that is uploaded once and stays resident in HBM.
// File: "AMTSRAW1" + int32 {width,height,bits,num_frames,fps_num,fps_den} + tightly packed planar 4:2:0 frames.
// ---------------------------------------------------------------------------------------------------------------
class AMTSource : public IClip, public IDeviceClip {
  VideoInfo vi;
  amtk_ctx* ctx;
  std::vector<uint8_t> host;          // CPU copy (serves GetFrame)
  void* dev = nullptr;                // HBM copy (serves batched filters)
  bool interlaced = true;
  size_t ysz() const { return (size_t)vi.width * vi.height * vi.ComponentSize(); }
  size_t csz() const { return (size_t)(vi.width / 2) * (vi.height / 2) * vi.ComponentSize(); }
  size_t fsz() const { return ysz() + 2 * csz(); }
public:
  AMTSource(const tstring& path, IScriptEnvironment* env) : ctx(env->GetAmtkContext()) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) env->ThrowError("AMTSource: failed to open %s", path.c_str());
    char magic[8]; int32_t h[6];
    if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "AMTSRAW1", 8) != 0 || fread(h, 4, 6, fp) != 6) { fclose(fp); env->ThrowError("AMTSource: bad header in %s", path.c_str()); }
    vi.width = h[0]; vi.height = h[1]; vi.num_frames = h[3]; vi.fps_numerator = (unsigned)h[4]; vi.fps_denominator = (unsigned)h[5];
    switch (h[2]) {                                       // AMTSource.hpp:428-442
      case 8: vi.pixel_type = VideoInfo::CS_YV12; break;
      case 10: vi.pixel_type = VideoInfo::CS_YUV420P10; break;
      case 12: vi.pixel_type = VideoInfo::CS_YUV420P12; break;
      default: fclose(fp); env->ThrowError("AMTSource: unsupported bit depth %d", h[2]);
    }
    host.resize(fsz() * vi.num_frames);
    const bool ok = fread(host.data(), 1, host.size(), fp) == host.size();
    fclose(fp);
    if (!ok) env->ThrowError("AMTSource: truncated file %s", path.c_str());
    if (!ctx) env->ThrowError("AMTSource: no device bound to the script environment");
    amtk_check(amtk_device_alloc(ctx, host.size(), &dev), env);
    amtk_check(amtk_memcpy_h2d(ctx, dev, host.data(), host.size()), env);
  }
  ~AMTSource() { if (dev) amtk_device_free(ctx, dev); }

  PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override {
    n = std::max(0, std::min(vi.num_frames - 1, n));
    PVideoFrame f = env->NewVideoFrame(vi);
    const uint8_t* src = host.data() + fsz() * n;
    const int planes[3] = { PLANAR_Y, PLANAR_U, PLANAR_V };
    size_t off = 0;
    for (int p = 0; p < 3; ++p) {
      const int rows = f->GetHeight(planes[p]), rb = f->GetRowSize(planes[p]);
      for (int y = 0; y < rows; ++y) memcpy(f->GetWritePtr(planes[p]) + (size_t)y * f->GetPitch(planes[p]), src + off + (size_t)y * rb, rb);
      off += (size_t)rows * rb;
    }
    return f;
  }
  bool GetDeviceClip(amtk_clip* c) override {
    memset(c, 0, sizeof(*c));
    c->base = dev; c->frame_stride = (int64_t)fsz(); c->off_u = (int64_t)ysz(); c->off_v = (int64_t)(ysz() + csz());
    c->width = vi.width; c->height = vi.height; c->pitch_y = vi.width * vi.ComponentSize(); c->pitch_uv = (vi.width / 2) * vi.ComponentSize();
    c->log_uvx = c->log_uvy = 1; c->bytes_per_sample = vi.ComponentSize(); c->bits_per_sample = vi.BitsPerComponent();
    c->num_frames = vi.num_frames; c->on_device = 1;
    return true;
  }
  // device frames edited in place by a batched filter become visible to GetFrame after this
  void SyncHostFromDevice(IScriptEnvironment* env) { amtk_check(amtk_memcpy_d2h(ctx, host.data(), dev, host.size()), env); }
  void __stdcall GetAudio(void*, int64_t, int64_t, IScriptEnvironment*) override {}
  const VideoInfo& __stdcall GetVideoInfo() override { return vi; }
  bool __stdcall GetParity(int) override { return interlaced; }                         // AMTSource.hpp:821-823
  int __stdcall SetCacheHints(int cachehints, int) override { return cachehints == CACHE_GET_MTMODE ? MT_NICE_FILTER : 0; }   // :825-830
};

inline AVSValue __cdecl CreateAMTSource(AVSValue args, void*, IScriptEnvironment* env) {     // AMTSource.hpp:873-882
  return AVSValue(PClip(new AMTSource(args[0].AsString(), env)));        // [filter]s [outqp]b are decode options: ignored
}

}  // namespace av

namespace logo {

struct LogoAnalyzeFrame { float p[11], t[11], b[11]; };              // LogoScan.hpp:1100-1103
static_assert(sizeof(LogoAnalyzeFrame) == 132, "LogoAnalyzeFrame layout");

struct LogoHandle {                                                     // RAII over amtk_logo
  amtk_logo* h = nullptr;
  LogoHandle() {}
  explicit LogoHandle(amtk_logo* p) : h(p) {}
  LogoHandle(LogoHandle&& o) : h(o.h) { o.h = nullptr; }
  LogoHandle& operator=(LogoHandle&& o) { if (this != &o) { reset(); h = o.h; o.h = nullptr; } return *this; }
  ~LogoHandle() { reset(); }
  void reset() { if (h) amtk_logo_destroy(h); h = nullptr; }
  bool valid() const { return h != nullptr; }
};

// ---------------------------------------------------------------------------------------------------------------
// AMTAnalyzeLogo (LogoScan.hpp:1106-1236)
// ---------------------------------------------------------------------------------------------------------------
class AMTAnalyzeLogo : public GenericVideoFilter {
  VideoInfo srcvi;
  LogoHandle logo, deintLogo, fieldLogoT, fieldLogoB;
  float maskratio;
public:
  AMTAnalyzeLogo(PClip clip, const tstring& logoPath, float maskratio, IScriptEnvironment* env)
      : GenericVideoFilter(clip), srcvi(vi), maskratio(maskratio) {
    amtk_logo* p = nullptr;
    if (!amtk_logo_load(env->GetAmtkContext(), logoPath.c_str(), &p, nullptr))
      env->ThrowError("Failed to read logo file (%s)", logoPath.c_str());                 // :1173-1175
    logo = LogoHandle(p);
    amtk_check(amtk_logo_deint(logo.h, &p), env); deintLogo = LogoHandle(p);             // :1177-1180
    amtk_check(amtk_logo_create_mask(deintLogo.h, maskratio), env);
    amtk_check(amtk_logo_field(logo.h, 0, &p), env); fieldLogoT = LogoHandle(p);         // :1182-1185
    amtk_check(amtk_logo_create_mask(fieldLogoT.h, maskratio), env);
    amtk_check(amtk_logo_field(logo.h, 1, &p), env); fieldLogoB = LogoHandle(p);
    amtk_check(amtk_logo_create_mask(fieldLogoB.h, maskratio), env);
    const int out_bytes = (int)sizeof(LogoAnalyzeFrame) * 8;                              // :1195-1200
    vi.pixel_type = VideoInfo::CS_BGR32;
    vi.width = 64;
    vi.height = nblocks(out_bytes, vi.width * 4);
    vi.num_frames = nblocks(vi.num_frames, 8);
  }

  // records of SOURCE frames [first, first+count) (count <= srcvi.num_frames - first), batched on the device
  void AnalyzeSourceFrames(int first, int count, LogoAnalyzeFrame* out, IScriptEnvironment* env) {
    amtk_ctx* ctx = env->GetAmtkContext();
    amtk_clip dc;
    IDeviceClip* d = dynamic_cast<IDeviceClip*>(child.get());
    if (d && d->GetDeviceClip(&dc)) {       // HBM-resident source: one batched call, results returned to the host
      amtk_check(amtk_logo_analyze_frames(ctx, &dc, deintLogo.h, fieldLogoT.h, fieldLogoB.h, first, count,
                                          reinterpret_cast<float*>(out), 0), env);
      return;
    }
    for (int i = 0; i < count; ++i) {                                                      // any other IClip: frame by frame
      PVideoFrame f = child->GetFrame(first + i, env);
      amtk_clip hc = HostFrameClip(f, srcvi);
      amtk_check(amtk_logo_analyze_frames(ctx, &hc, deintLogo.h, fieldLogoT.h, fieldLogoB.h, 0, 1, reinterpret_cast<float*>(out + i), 0), env);
    }
  }

  PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override {
    const int pixelSize = srcvi.ComponentSize();
    if (pixelSize != 1 && pixelSize != 2) env->ThrowError("[AMTAnalyzeLogo] Unsupported pixel format");   // :1213-1215
    PVideoFrame dst = env->NewVideoFrame(vi);
    LogoAnalyzeFrame* pDst = reinterpret_cast<LogoAnalyzeFrame*>(dst->GetWritePtr());
    const int first = std::min(srcvi.num_frames - 1, n * 8);
    const int count = std::max(1, std::min(8, srcvi.num_frames - n * 8));
    AnalyzeSourceFrames(first, count, pDst, env);
    for (int i = count; i < 8; ++i) pDst[i] = pDst[count - 1];        // nsrc clamps to the last source frame (:1133)
    return dst;
  }
  int __stdcall SetCacheHints(int cachehints, int) override { return cachehints == CACHE_GET_MTMODE ? MT_NICE_FILTER : 0; }   // :1220-1225

  static AVSValue __cdecl Create(AVSValue args, void*, IScriptEnvironment* env) {         // :1227-1235
    return AVSValue(PClip(new AMTAnalyzeLogo(args[0].AsClip(), args[1].AsString(), (float)args[2].AsFloat(35) / 100.0f, env)));
  }
};

// ---------------------------------------------------------------------------------------------------------------
// AMTEraseLogo (LogoScan.hpp:1238-1519)
// ---------------------------------------------------------------------------------------------------------------
class AMTEraseLogo : public GenericVideoFilter {
  PClip analyzeclip;
  std::vector<int> frameResult;
  LogoHandle logo;
  int mode, maxFadeLength;

  void CalcFade2(int n, float& fadeT, float& fadeB, IScriptEnvironment* env) {           // :1263-1315
    // the 9 records around n sit in at most 3 analyze frames; fetch them through the analyze clip's GetFrame
    const int nrec = vi.num_frames;
    std::vector<float> rec((size_t)nrec * 33, 0.0f);
    std::vector<char> have((size_t)nblocks(nrec, 8), 0);
    for (int i = -4; i <= 4; ++i) {
      const int nsrc = std::max(0, std::min(vi.num_frames - 1, n + i));
      const int blk = std::max(0, std::min((int)have.size() - 1, (nsrc + i) >> 3));
      if (have[blk]) continue;
      PVideoFrame f = analyzeclip->GetFrame(blk, env);
      const int cnt = std::min(8, nrec - blk * 8);
      memcpy(&rec[(size_t)blk * 8 * 33], f->GetReadPtr(), (size_t)cnt * sizeof(LogoAnalyzeFrame));
      have[blk] = 1;
    }
    amtk_calc_fade2(rec.data(), nrec, vi.num_frames, n, &fadeT, &fadeB);
  }
  void CalcFade(int n, float& fadeT, float& fadeB, IScriptEnvironment* env) {            // :1317-1341
    if (frameResult.empty()) { CalcFade2(n, fadeT, fadeB, env); return; }
    const int halfWidth = maxFadeLength >> 1;
    bool uniform = true; int first = 0;
    for (int i = -halfWidth; i <= halfWidth; ++i) {
      const int v = frameResult[std::max(0, std::min(vi.num_frames - 1, n + i))];
      if (i == -halfWidth) first = v; else if (v != first) uniform = false;
    }
    if (uniform) fadeT = fadeB = (frameResult[std::max(0, std::min(vi.num_frames - 1, n))] == 2) ? 1.0f : 0.0f;
    else CalcFade2(n, fadeT, fadeB, env);
  }
  void ReadLogoFrameFile(const tstring& path, IScriptEnvironment* env) {                  // :1421-1461
    struct Elem { bool isStart; int best, start, end; };
    std::vector<Elem> el;
    FILE* fp = fopen(path.c_str(), "r");
    if (!fp) env->ThrowError("Failed to read dat file (%s)", path.c_str());
    std::regex re("^\\s*(\\d+)\\s+(\\S)\\s+(\\d+)\\s+(\\S+)\\s+(\\d+)\\s+(\\d+)");
    char line[512];
    while (fgets(line, sizeof(line), fp)) {
      std::cmatch m;
      if (std::regex_search(line, m, re))
        el.push_back(Elem{ std::tolower(m[2].str()[0]) == 's', std::stoi(m[1].str()), std::stoi(m[5].str()), std::stoi(m[6].str()) });
    }
    fclose(fp);
    frameResult.assign(vi.num_frames, 0);
    auto fill = [&](int a, int b, int v) { std::fill(frameResult.begin() + std::min(vi.num_frames, a), frameResult.begin() + std::min(vi.num_frames, std::max(a, b)), v); };
    for (size_t i = 0; i + 1 < el.size() || i < el.size(); i += 2) {
      if (i + 1 >= el.size() || !el[i].isStart || el[i + 1].isStart)
        env->ThrowError("Invalid logoframe file. Start and End must be cyclic.");
      fill(el[i].start, el[i].end + 1, 1);
      fill(el[i].end, el[i + 1].start + 1, 2);
      fill(el[i + 1].start + 1, el[i + 1].end + 1, 1);
    }
  }
public:
  AMTEraseLogo(PClip clip, PClip analyzeclip, const tstring& logoPath, const tstring& logofPath, int mode, int maxFadeLength, IScriptEnvironment* env)
      : GenericVideoFilter(clip), analyzeclip(analyzeclip), mode(mode), maxFadeLength(maxFadeLength) {
    amtk_logo* p = nullptr;
    if (!amtk_logo_load(env->GetAmtkContext(), logoPath.c_str(), &p, nullptr))
      env->ThrowError("Failed to read logo file (%s)", logoPath.c_str());                 // :1471-1477
    logo = LogoHandle(p);
    if (logofPath.size() > 0) ReadLogoFrameFile(logofPath, env);
  }
  void GetFades(int n, float& fadeT, float& fadeB, IScriptEnvironment* env) { CalcFade(n, fadeT, fadeB, env); }

  PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override {               // :1343-1419
    const int pixelSize = vi.ComponentSize();
    if (pixelSize != 1 && pixelSize != 2) env->ThrowError("[AMTEraseLogo] Unsupported pixel format");
    PVideoFrame frame = child->GetFrame(n, env);
    env->MakeWritable(&frame);
    float fades[2];
    CalcFade(n, fades[0], fades[1], env);
    if (mode != 0) return frame;                          // debug overlay mode of the reference draws text only
    // one frame through HBM: upload, Delogo kernel in place, download
    amtk_ctx* ctx = env->GetAmtkContext();
    const size_t bytes = (frame->TotalBytes() + 15) & ~(size_t)15;
    void* d = nullptr;
    amtk_check(amtk_device_alloc(ctx, bytes, &d), env);
    amtk_clip c = HostFrameClip(frame, vi);
    int ok = amtk_memcpy_h2d(ctx, d, frame->Base(), frame->TotalBytes());
    c.base = d; c.on_device = 1;
    ok = ok && amtk_erase_logo_frames(ctx, &c, logo.h, 0, 1, fades);
    ok = ok && amtk_memcpy_d2h(ctx, frame->GetWritePtr(PLANAR_Y), d, frame->TotalBytes());
    amtk_device_free(ctx, d);
    amtk_check(ok, env);
    return frame;
  }
  // Batched form for an HBM-resident source: every frame of [first, first+count) erased in place with one launch.
  void EraseInPlace(int first, int count, IScriptEnvironment* env) {
    amtk_clip dc;
    IDeviceClip* d = dynamic_cast<IDeviceClip*>(child.get());
    if (!d || !d->GetDeviceClip(&dc)) env->ThrowError("[AMTEraseLogo] EraseInPlace needs a device-resident source");
    std::vector<float> fades((size_t)count * 2);
    for (int i = 0; i < count; ++i) CalcFade(first + i, fades[2 * i], fades[2 * i + 1], env);
    amtk_check(amtk_erase_logo_frames(env->GetAmtkContext(), &dc, logo.h, first, count, fades.data()), env);
  }
  int __stdcall SetCacheHints(int cachehints, int) override { return cachehints == CACHE_GET_MTMODE ? MT_NICE_FILTER : 0; }   // :1500-1505

  static AVSValue __cdecl Create(AVSValue args, void*, IScriptEnvironment* env) {         // :1507-1518
    return AVSValue(PClip(new AMTEraseLogo(args[0].AsClip(), args[1].AsClip(), args[2].AsString(), args[3].AsString(""),
                                           args[4].AsInt(0), args[5].AsInt(16), env)));
  }
};

// ---------------------------------------------------------------------------------------------------------------
// LogoFrame (LogoScan.hpp:1521-1836): whole-clip logo scan used by CMAnalyze::logoFrame (CMAnalyze.hpp:273-317)
// ---------------------------------------------------------------------------------------------------------------
class LogoFrame {
  AMTContext& ctx;
  int numLogos;
  std::vector<LogoHandle> logoArr, deintArr;
  int numFrames = 0, framesPerSec = 30;
  VideoInfo vi;
  struct EvalResult { float corr0, corr1; };
  std::vector<EvalResult> evalResults;
  const float THRESH = 0.2f;                                            // |score| below this is "unknown" (:1538)
  int bestLogo = -1;
  float logoRatio = 0.0f;
public:
  LogoFrame(AMTContext& ctx, const std::vector<tstring>& logofiles, float maskratio) : ctx(ctx) {   // :1592-1616
    numLogos = (int)logofiles.size();
    logoArr.resize(numLogos); deintArr.resize(numLogos);
    for (int i = 0; i < numLogos; ++i) {
      amtk_logo* p = nullptr;
      if (!amtk_logo_load(nullptr, logofiles[i].c_str(), &p, nullptr)) continue;          // load errors are ignored (:1612-1614)
      logoArr[i] = LogoHandle(p);
      if (amtk_logo_deint(logoArr[i].h, &p)) { deintArr[i] = LogoHandle(p); amtk_logo_create_mask(deintArr[i].h, maskratio); }
    }
  }

  void scanFrames(PClip clip, IScriptEnvironment2* env) {                                   // :1618-1630
    vi = clip->GetVideoInfo();
    const int pixelSize = vi.ComponentSize();
    if (pixelSize != 1 && pixelSize != 2) env->ThrowError("[LogoFrame] Unsupported pixel format");
    amtk_ctx* actx = env->GetAmtkContext();
    std::vector<amtk_logo*> hs(numLogos);
    for (int i = 0; i < numLogos; ++i) hs[i] = deintArr[i].h;          // invalid logos stay NULL -> (0,-1) (:1551-1558)
    evalResults.assign((size_t)vi.num_frames * numLogos, EvalResult{ 0, -1 });
    // the reference passes the BYTE pitch as element pitch also for 16-bit clips (:1547,1561); kept for parity
    amtk_clip dc;
    IDeviceClip* d = dynamic_cast<IDeviceClip*>(clip.get());
    if (d && d->GetDeviceClip(&dc)) {                                    // one batched call for the whole clip
      const int quirk = pixelSize == 2 ? dc.pitch_y : 0;
      amtk_check(amtk_logo_scan_frames(actx, &dc, hs.data(), numLogos, 0, vi.num_frames, quirk,
                                       reinterpret_cast<float*>(evalResults.data()), 0), env);
    } else {
      for (int n = 0; n < vi.num_frames; ++n) {                          // generic IClip: the reference's pull loop (:1577-1579)
        PVideoFrame f = clip->GetFrame(n, env);
        amtk_clip hc = HostFrameClip(f, vi);
        amtk_check(amtk_logo_scan_frames(actx, &hc, hs.data(), numLogos, 0, 1, pixelSize == 2 ? hc.pitch_y : 0,
                                         reinterpret_cast<float*>(&evalResults[(size_t)n * numLogos]), 0), env);
        if ((n % 5000) == 0) ctx.infoF("%6d/%d", n, vi.num_frames);
      }
    }
    numFrames = vi.num_frames;
    framesPerSec = (int)std::round((float)vi.fps_numerator / vi.fps_denominator);
    ctx.info("Finished");
  }

  const float* results() const { return reinterpret_cast<const float*>(evalResults.data()); }
  // Offline re-analysis of saved scores: float[nframes][numLogos][2] as scanFrames produces them (the reference can only
  // dump them, dumpResult :1632-1645).  Lets selectLogo / writeResult run without a device.
  void setResults(const float* corr, int nframes, unsigned fps_numerator, unsigned fps_denominator) {
    evalResults.resize((size_t)nframes * numLogos);
    memcpy(evalResults.data(), corr, evalResults.size() * sizeof(EvalResult));
    numFrames = nframes;
    framesPerSec = (int)std::round((float)fps_numerator / fps_denominator);
  }

  // choose the logo that is detected most often with the least residue after removal (:1647-1682)
  void selectLogo(int numCandidates = -1) {
    if (numCandidates < 0) numCandidates = numLogos;
    struct Summary { float cost = 0; int numFrames = 0; };
    std::vector<Summary> sum(numCandidates);
    for (int n = 0; n < numFrames; ++n)
      for (int i = 0; i < numCandidates; ++i) {
        const EvalResult& r = evalResults[(size_t)n * numLogos + i];
        if (r.corr0 > THRESH && std::abs(r.corr1) < THRESH) { sum[i].numFrames++; sum[i].cost += std::abs(r.corr1); }
      }
    std::vector<float> score(numCandidates);
    for (int i = 0; i < numCandidates; ++i) {
      const Summary& s = sum[i];
      score[i] = (s.numFrames == 0) ? INFINITY : (s.cost / s.numFrames) * (numFrames / (float)s.numFrames);
      ctx.debugF("logo%d: %f * %f = %f", i + 1, (s.cost / s.numFrames), (numFrames / (float)s.numFrames), score[i]);
    }
    bestLogo = (int)(std::min_element(score.begin(), score.end()) - score.begin());
    logoRatio = (float)sum[bestLogo].numFrames / numFrames;
  }

  // logoframe file for join_logo_scp / AMTEraseLogo::ReadLogoFrameFile (:1686-1827)
  void writeResult(const tstring& outpath, int logoIndex = -1) {
    if (logoIndex < 0) { if (bestLogo < 0) selectLogo(); logoIndex = bestLogo; }
    const float threshL = 0.5f, avgDur = 1.0f, medianDur = 0.5f;
    const int halfAvg = int(framesPerSec * avgDur / 2 + 0.5f), aveFrames = halfAvg * 2 + 1;
    const int halfMed = int(framesPerSec * medianDur / 2 + 0.5f), medFrames = halfMed * 2 + 1;
    const int win = std::max(aveFrames, medFrames), halfWin = win / 2;
    const int N = numFrames;
    // raw score per frame (negative corr0 and positive corr1 are noise), edge-padded by half a window
    std::vector<float> padded((size_t)N + win);
    float* raw = padded.data() + halfWin;
    for (int n = 0; n < N; ++n) {
      const EvalResult& r = evalResults[(size_t)n * numLogos + logoIndex];
      raw[n] = std::max(0.0f, r.corr0) + std::min(0.0f, r.corr1);
    }
    std::fill(padded.data(), raw, raw[0]);
    std::fill(raw + N, padded.data() + padded.size(), raw[N - 1]);

    struct FR { int result; float score; };
    std::vector<FR> fr(N);
    std::vector<float> med(medFrames);
    for (int i = 0; i < N; ++i) {
      // min of the maxima before and after: rescues frames where motion washes the logo out
      const float beforeMax = *std::max_element(raw + i - halfAvg, raw + i);
      const float afterMax = *std::max_element(raw + i + 1, raw + i + 1 + halfAvg);
      const float mm = std::min(beforeMax, afterMax);
      const int mmRes = (std::abs(mm) < threshL) ? 1 : (mm < 0.0f) ? 0 : 2;
      const float avg = std::accumulate(raw + i - halfAvg, raw + i + halfAvg + 1, 0.0f) / aveFrames;
      const int avgRes = (std::abs(avg) < THRESH) ? 1 : (avg < 0.0f) ? 0 : 2;
      fr[i].result = (mmRes != avgRes) ? 1 : mmRes;
      std::copy(raw + i - halfMed, raw + i + halfMed + 1, med.begin());
      std::sort(med.begin(), med.end());
      fr[i].score = med[halfMed];
    }
    // unknown runs bounded by equal states take that state
    for (int it = 0; it != N;) {
      int first1 = it; while (first1 < N && fr[first1].result != 1) ++first1;
      it = first1; while (it < N && fr[it].result == 1) ++it;
      const int prev = (first1 == 0) ? 0 : fr[first1 - 1].result;
      const int next = (it == N) ? 0 : fr[it].result;
      if (prev == next) for (int k = first1; k < it; ++k) fr[k].result = prev;
    }
    // emit logo sections, refining the boundaries on the median-filtered score
    std::string out;
    auto last_before = [&](int hi, int lo, auto pred, int none) {        // reverse find in [lo,hi): index+1 of the hit, else `none`
      for (int k = hi - 1; k >= lo; --k) if (pred(fr[k])) return k + 1;
      return none;
    };
    auto first_from = [&](int lo, int hi, auto pred) { for (int k = lo; k < hi; ++k) if (pred(fr[k])) return k; return hi; };
    for (int it = 0; it != N;) {
      const int sEnd0 = first_from(it, N, [](const FR& r) { return r.result == 2; });
      const int eEnd0 = first_from(sEnd0, N, [](const FR& r) { return r.result == 0; });
      int sEnd = sEnd0, eEnd = eEnd0;
      if (sEnd != N) {
        if (fr[sEnd].score >= THRESH) sEnd = last_before(sEnd, 0, [&](const FR& r) { return r.score < THRESH; }, 0);
        else sEnd = first_from(sEnd, N, [&](const FR& r) { return r.score >= THRESH; });
      }
      if (eEnd != N) {
        if (fr[eEnd].score <= -THRESH) eEnd = last_before(eEnd, sEnd, [&](const FR& r) { return r.score > -THRESH; }, sEnd);
        else eEnd = first_from(eEnd, N, [&](const FR& r) { return r.score <= -THRESH; });
      }
      const int sStart = last_before(sEnd, it, [&](const FR& r) { return r.score <= -THRESH; }, it);
      const int eStart = last_before(eEnd, sEnd, [&](const FR& r) { return r.score >= THRESH; }, sEnd);
      const int sBest = first_from(sStart, sEnd, [](const FR& r) { return r.score > 0; });
      const int eBest = last_before(eEnd, eStart, [](const FR& r) { return r.score > 0; }, eStart);
      if (sEnd != eEnd) {
        char buf[128];
        snprintf(buf, sizeof(buf), "%6d S 0 ALL %6d %6d\n", sBest, sStart, sEnd); out += buf;
        snprintf(buf, sizeof(buf), "%6d E 0 ALL %6d %6d\n", eBest - 1, eStart - 1, eEnd - 1); out += buf;
      }
      it = eEnd0;
    }
    FILE* fp = fopen(outpath.c_str(), "w");
    if (!fp) throw IOException("failed to open " + outpath);
    fwrite(out.data(), 1, out.size(), fp);
    fclose(fp);
  }
  int getBestLogo() const { return bestLogo; }
  float getLogoRatio() const { return logoRatio; }
};

}  // namespace logo

// ---------------------------------------------------------------------------------------------------------------
// Telecine pre-pass.  In the product the script calls KFMDeint(..., pass=..., filepath=AMT_TMP) from an external plugin
// and AMTFilterSource pulls every frame and discards it (FilteredSource.hpp:417-439,519-544).  AMTCombAnalyze is that
// pre-pass filter for the field-difference / combing counters: the whole clip is analysed by ONE streaming launch
// on first use; GetFrame returns the source frame untouched (pre-process semantics), results go to
// <AMT_TMP>.combstat.txt (one line per frame: 12 integers) when a path is given.
// ---------------------------------------------------------------------------------------------------------------
class AMTCombAnalyze : public GenericVideoFilter {
  std::vector<int32_t> counts;
  tstring outpath;
  amtk_comb_params prm;
  bool done = false;
  void Run(IScriptEnvironment* env) {
    if (done) return;
    amtk_ctx* ctx = env->GetAmtkContext();
    counts.assign((size_t)vi.num_frames * 12, 0);
    amtk_clip dc;
    IDeviceClip* d = dynamic_cast<IDeviceClip*>(child.get());
    if (d && d->GetDeviceClip(&dc)) {
      amtk_check(amtk_comb_frames(ctx, &dc, &prm, 0, vi.num_frames, counts.data(), 0), env);
    } else {                                  // generic source: frames are packed pairwise (prev, cur) on the host
      PVideoFrame prev = child->GetFrame(0, env);
      for (int n = 0; n < vi.num_frames; ++n) {
        PVideoFrame cur = n ? child->GetFrame(n, env) : prev;
        const size_t fb = (cur->TotalBytes() + 15) & ~(size_t)15;
        std::vector<uint8_t> two(2 * fb);
        memcpy(two.data(), prev->Base(), prev->TotalBytes()); memcpy(two.data() + fb, cur->Base(), cur->TotalBytes());
        amtk_clip hc = HostFrameClip(cur, vi);
        hc.base = two.data(); hc.frame_stride = (int64_t)fb; hc.num_frames = 2;
        amtk_check(amtk_comb_frames(ctx, &hc, &prm, 1, 1, &counts[(size_t)n * 12], 0), env);
        if (n == 0) for (int k : { 0, 3, 6, 9 }) counts[k] = 0;       // prev(0) = frame 0 itself
        prev = cur;
      }
    }
    if (!outpath.empty()) {
      FILE* fp = fopen(outpath.c_str(), "w");
      if (!fp) env->ThrowError("AMTCombAnalyze: failed to write %s", outpath.c_str());
      for (int n = 0; n < vi.num_frames; ++n) {
        for (int k = 0; k < 12; ++k) fprintf(fp, k ? " %d" : "%d", counts[(size_t)n * 12 + k]);
        fputc('\n', fp);
      }
      fclose(fp);
    }
    done = true;
  }
public:
  AMTCombAnalyze(PClip clip, const tstring& outpath, IScriptEnvironment*) : GenericVideoFilter(clip), outpath(outpath) { amtk_comb_default_params(&prm); }
  PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override { Run(env); return child->GetFrame(n, env); }
  const std::vector<int32_t>& Counts(IScriptEnvironment* env) { Run(env); return counts; }
  int __stdcall SetCacheHints(int cachehints, int) override { return cachehints == CACHE_GET_MTMODE ? MT_SERIALIZED : 0; }
  static AVSValue __cdecl Create(AVSValue args, void*, IScriptEnvironment* env) {
    return AVSValue(PClip(new AMTCombAnalyze(args[0].AsClip(), args[1].AsString(""), env)));
  }
};

// AMTFilterSource::ReadAllFrames (FilteredSource.hpp:417-439): pull every frame of a pre-process pass and discard it.
inline void ReadAllFrames(PClip clip, IScriptEnvironment* env) {
  const int n = clip->GetVideoInfo().num_frames;
  for (int i = 0; i < n; ++i) clip->GetFrame(i, env);
}

// ---------------------------------------------------------------------------------------------------------------
// Telecine side files and their consumers
// ---------------------------------------------------------------------------------------------------------------
// AMTDecimate (FilteredSource.hpp:637-676): <tmp>.duration.txt holds one integer per OUTPUT frame = how many source
// frames it lasts; output frame i shows source frame sum(durations[0..i)).
class AMTDecimate : public GenericVideoFilter {
  std::vector<int> durations, framesMap;
public:
  AMTDecimate(PClip source, const std::string& duration, IScriptEnvironment* env) : GenericVideoFilter(source) {
    FILE* fp = fopen(duration.c_str(), "r");
    if (!fp) env->ThrowError("[AMTDecimate] failed to open %s", duration.c_str());
    char line[256];
    while (fgets(line, sizeof(line), fp)) durations.push_back(std::atoi(line));
    fclose(fp);
    const int numSourceFrames = std::accumulate(durations.begin(), durations.end(), 0);
    if (vi.num_frames != numSourceFrames)
      env->ThrowError("[AMTDecimate] # of frames does not match. %d(%s) vs %d(source clip)", numSourceFrames, duration.c_str(), vi.num_frames);
    vi.num_frames = (int)durations.size();
    framesMap.assign(durations.size(), 0);
    for (size_t i = 0; i + 1 < durations.size(); ++i) framesMap[i + 1] = framesMap[i] + durations[i];
  }
  PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override {
    return child->GetFrame(framesMap[std::max(0, std::min(n, vi.num_frames - 1))], env);
  }
  int SourceFrame(int n) const { return framesMap[std::max(0, std::min(n, (int)framesMap.size() - 1))]; }
  static AVSValue __cdecl Create(AVSValue args, void*, IScriptEnvironment* env) {
    return AVSValue(PClip(new AMTDecimate(args[0].AsClip(), args[1].AsString(), env)));
  }
};

// AMTFilterSource::readTimecodeFile + readTimecode (FilteredSource.hpp:163-212): timestamps in ms, one per line,
// '#' comments, optional "# total: <seconds>"; the end time is extrapolated when absent; the VFR base rate is the one
// of 60/120/240 (x1000/1001) whose grid fits the timestamps best.
struct TimecodeFile {
  std::vector<double> timeCodes;
  int vfrTimingFps = 0;
  bool read(const std::string& path) {
    FILE* fp = fopen(path.c_str(), "r");
    if (!fp) return false;
    std::regex re("#\\s*total:\\s*([+-]?([0-9]*[.])?[0-9]+).*");
    char line[512];
    timeCodes.clear();
    bool total = false;
    while (!total && fgets(line, sizeof(line), fp)) {
      std::string str(line);
      while (!str.empty() && (str.back() == '\n' || str.back() == '\r')) str.pop_back();
      if (str.empty()) continue;
      std::smatch m;
      if (std::regex_search(str, m, re)) { timeCodes.push_back(std::atof(m[1].str().c_str()) * 1000); total = true; }
      else if (str[0] != '#') timeCodes.push_back(std::atoi(str.c_str()));
    }
    fclose(fp);
    if (!total) {
      const size_t n = timeCodes.size();
      if (n >= 2) timeCodes.push_back(timeCodes[n - 1] * 2 - timeCodes[n - 2]);
      else if (n == 1) timeCodes.push_back(timeCodes[0] + 1000.0 / 60.0);
    }
    if (timeCodes.empty()) return true;
    double minDiff = timeCodes.back();
    const double epsilon = timeCodes.size() * 10e-10;
    for (int fps : { 60, 120, 240 }) {
      const double mult = fps / 1001.0, inv = 1.0 / mult;
      double diff = 0;
      for (double ts : timeCodes) diff += std::abs(inv * std::round(ts * mult) - ts);
      if (diff < minDiff - epsilon) { vfrTimingFps = fps; minDiff = diff; }
    }
    return true;
  }
};

// Pulldown classification from the combing counters (this repo's heuristic; the reference delegates the decision to
// the external KFM plugin).  3:2 telecine shows up as two adjacent combed frames in every 5-frame cycle.  The cycle
// phase is taken from the whole clip (the frame pair position with the largest summed comb response); a cycle whose
// pair stands out by `ratio` against its other three frames becomes 4 film frames (durations 1,1,2,1: the film frame
// that straddles the combed pair lasts two video frames), any other cycle passes through as 5 x 1.
// Writes <base>.duration.txt (AMTDecimate) and <base>.timecode.txt (ms per output frame, "# total:" trailer).
// Returns the number of film cycles, -1 on I/O error.
inline int WriteTelecineFiles(const std::vector<int32_t>& counts, int num_frames, unsigned fps_num, unsigned fps_den,
                              const std::string& base, double ratio = 2.0) {
  // the large-threshold response ("lshima", Y top+bottom) separates real combing from vertical detail best
  auto shima = [&](int n) { return (long long)counts[(size_t)n * 12 + 2] + counts[(size_t)n * 12 + 5]; };
  int phase = 0; long long best = -1;
  for (int p = 0; p < 5; ++p) {
    long long acc = 0;
    for (int n = p; n + 1 < num_frames; n += 5) acc += std::min(shima(n), shima(n + 1));
    if (acc > best) { best = acc; phase = p; }
  }
  const int start = (phase + 3) % 5;                      // cycles begin two frames before the combed pair
  std::vector<int> durations;
  int film_cycles = 0, n = 0;
  for (; n < start && n < num_frames; ++n) durations.push_back(1);
  for (; n + 5 <= num_frames; n += 5) {
    const long long pair = std::min(shima(n + 2), shima(n + 3));
    const long long rest = std::max(std::max(shima(n), shima(n + 1)), shima(n + 4));
    if ((double)pair > ratio * (double)std::max<long long>(rest, 1)) {
      ++film_cycles;
      const int d[4] = { 1, 1, 2, 1 };
      durations.insert(durations.end(), d, d + 4);
    } else {
      durations.insert(durations.end(), 5, 1);
    }
  }
  for (; n < num_frames; ++n) durations.push_back(1);
  FILE* fd = fopen((base + ".duration.txt").c_str(), "w");
  FILE* ft = fopen((base + ".timecode.txt").c_str(), "w");
  if (!fd || !ft) { if (fd) fclose(fd); if (ft) fclose(ft); return -1; }
  fprintf(ft, "# timecode format v2\n");
  const double frame_ms = 1000.0 * fps_den / fps_num;
  int src = 0;
  for (int d : durations) { fprintf(fd, "%d\n", d); fprintf(ft, "%d\n", (int)std::round(src * frame_ms)); src += d; }
  fprintf(ft, "# total: %.6f\n", src * frame_ms / 1000.0);
  fclose(fd); fclose(ft);
  return film_cycles;
}

// ---------------------------------------------------------------------------------------------------------------
// CMAnalyze (CMAnalyze.hpp:22-317) -- the logo-analysis half only: the constructor runs logoFrame() when logos are
// configured and exposes getLogoPath().  chapter_exe / join_logo_scp subprocesses, Trim/zone parsing (:319-679) are
// out of scope.  ConfigWrapper is reduced to the accessors logoFrame() reads.
// ---------------------------------------------------------------------------------------------------------------
struct ConfigWrapper {
  std::vector<tstring> logoPath, eraseLogoPath;          // --logo / --erase-logo (AmatsukazeCLI.hpp:358-366)
  bool looseLogoDetection = false;                       // --loose-logo-detection (:370)
  tstring tmpDir = ".";
  const std::vector<tstring>& getLogoPath() const { return logoPath; }
  const std::vector<tstring>& getEraseLogoPath() const { return eraseLogoPath; }
  bool isLooseLogoDetection() const { return looseLogoDetection; }
  tstring getTmpAMTSourcePath(int v) const { return tmpDir + "/amts" + std::to_string(v) + ".dat"; }            // TranscodeSetting.hpp:926-928
  tstring getTmpLogoFramePath(int v, int logoIndex = -1) const {                                               // :934-939
    return tmpDir + "/logof" + std::to_string(v) + (logoIndex == -1 ? std::string() : "-" + std::to_string(logoIndex)) + ".txt";
  }
};

struct AviSynthException : std::runtime_error { using std::runtime_error::runtime_error; };

class CMAnalyze {
public:
  // env must have the plugin registered (AvisynthPluginInit3) and an amtk context bound; the reference creates its
  // own script environment and loads itself as a plugin (:275-280) -- on the B200 build the caller owns the device.
  CMAnalyze(AMTContext& ctx, const ConfigWrapper& setting, int videoFileIndex, int /*numFrames*/, IScriptEnvironment2* env)
      : ctx(ctx), setting_(setting) {
    if (setting_.getLogoPath().size() > 0 || setting_.getEraseLogoPath().size() > 0) {                        // :34
      ctx.info("[logo analysis]");
      logoFrame(videoFileIndex, env);
      if (logopath.size() > 0) ctx.infoF("matched logo: %s", logopath.c_str());
    }
  }
  const tstring& getLogoPath() const { return logopath; }
  float getLogoRatio() const { return logoRatio; }
private:
  AMTContext& ctx;
  const ConfigWrapper& setting_;
  tstring logopath;
  float logoRatio = 0.0f;

  void logoFrame(int videoFileIndex, IScriptEnvironment2* env) {                                              // :273-317
    try {
      PClip clip = env->Invoke("AMTSource", AVSValue(std::vector<AVSValue>{ AVSValue(setting_.getTmpAMTSourcePath(videoFileIndex)) })).AsClip();
      const VideoInfo vi = clip->GetVideoInfo();
      const int duration = (int)((int64_t)vi.num_frames * vi.fps_denominator / vi.fps_numerator);
      const auto& logoPath = setting_.getLogoPath();
      const auto& eraseLogoPath = setting_.getEraseLogoPath();
      std::vector<tstring> allLogoPath = logoPath;
      allLogoPath.insert(allLogoPath.end(), eraseLogoPath.begin(), eraseLogoPath.end());
      logo::LogoFrame logof(ctx, allLogoPath, 0.35f);
      logof.scanFrames(clip, env);                       // ONE batched device pass for all logos
      if (logoPath.size() > 0) {
        logof.selectLogo((int)logoPath.size());
        logof.writeResult(setting_.getTmpLogoFramePath(videoFileIndex));
        logoRatio = logof.getLogoRatio();
        const float threshold = setting_.isLooseLogoDetection() ? 0.03f : (duration <= 60 * 7) ? 0.03f : 0.1f;
        if (logoRatio < threshold) ctx.info("no logo matched in this section");
        else logopath = setting_.getLogoPath()[logof.getBestLogo()];
      }
      for (int i = 0; i < (int)eraseLogoPath.size(); ++i)
        logof.writeResult(setting_.getTmpLogoFramePath(videoFileIndex, i), (int)logoPath.size() + i);
    } catch (const AvisynthError& avserror) {
      throw AviSynthException(avserror.msg);                                                                  // :314-316
    }
  }
};

// Registration with the reference's names and argument specs (Amatsukaze.cpp:43-66).
extern "C" inline const char* __stdcall AvisynthPluginInit3(IScriptEnvironment* env, const AVS_Linkage* const) {
  env->AddFunction("AMTSource", "s[filter]s[outqp]b", av::CreateAMTSource, 0);
  env->AddFunction("AMTAnalyzeLogo", "cs[maskratio]i", logo::AMTAnalyzeLogo::Create, 0);
  env->AddFunction("AMTEraseLogo", "ccs[logof]s[mode]i[maxfade]i", logo::AMTEraseLogo::Create, 0);
  env->AddFunction("AMTDecimate", "c[duration]s", AMTDecimate::Create, 0);
  env->AddFunction("AMTCombAnalyze", "c[filepath]s", AMTCombAnalyze::Create, 0);
  return "Amatsukaze plugin (B200 hot path)";
}
