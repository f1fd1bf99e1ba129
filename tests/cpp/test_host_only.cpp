// tests/cpp/test_host_only.cpp -- the parts of the host-side filter mirror that need NO device: LogoFrame::selectLogo /
// writeResult on saved scores (LogoScan.hpp:1645-1827), AMTEraseLogo's logoframe-file state machine and fade selection
// (:1263-1341,1421-1461), AMTDecimate, the timecode reader (FilteredSource.hpp:163-212) and the telecine side files.
// Driven by tests/test_host_only.py (runs in the CPU suite).  usage: test_host_only <mode> args...
#include "../../amatsukaze_b200/host/filters.hpp"
#include <fstream>
#include <string>

static std::vector<char> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// a clip that only knows its VideoInfo; frames are blank (AMTDecimate / AMTEraseLogo never look at the pixels here)
class BlankClip : public IClip {
  VideoInfo vi_;
public:
  explicit BlankClip(const VideoInfo& vi) : vi_(vi) {}
  PVideoFrame __stdcall GetFrame(int, IScriptEnvironment* env) override { return env->NewVideoFrame(vi_); }
  bool __stdcall GetParity(int) override { return true; }
  void __stdcall GetAudio(void*, int64_t, int64_t, IScriptEnvironment*) override {}
  int __stdcall SetCacheHints(int, int) override { return 0; }
  const VideoInfo& __stdcall GetVideoInfo() override { return vi_; }
};

// the analyze clip of AMTEraseLogo: 8 LogoAnalyzeFrame records per frame, served from a float[nrec][33] file
class RecordClip : public IClip {
  VideoInfo vi_;
  std::vector<logo::LogoAnalyzeFrame> rec_;
public:
  RecordClip(const std::vector<char>& raw) {
    rec_.resize(raw.size() / sizeof(logo::LogoAnalyzeFrame));
    memcpy(rec_.data(), raw.data(), rec_.size() * sizeof(logo::LogoAnalyzeFrame));
    vi_.width = 8 * (int)sizeof(logo::LogoAnalyzeFrame); vi_.height = 1; vi_.pixel_type = VideoInfo::CS_BGR32;
    vi_.width = (vi_.width + 3) / 4;
    vi_.num_frames = ((int)rec_.size() + 7) / 8;
  }
  PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override {
    PVideoFrame f = env->NewVideoFrame(vi_);
    const int cnt = std::min(8, (int)rec_.size() - n * 8);
    memcpy(f->GetWritePtr(), &rec_[(size_t)n * 8], (size_t)cnt * sizeof(logo::LogoAnalyzeFrame));
    return f;
  }
  bool __stdcall GetParity(int) override { return true; }
  void __stdcall GetAudio(void*, int64_t, int64_t, IScriptEnvironment*) override {}
  int __stdcall SetCacheHints(int, int) override { return 0; }
  const VideoInfo& __stdcall GetVideoInfo() override { return vi_; }
};

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: test_host_only <mode> ...\n"); return 2; }
  const std::string mode = argv[1];
  IScriptEnvironment envObj; IScriptEnvironment* env = &envObj;      // no amtk context: nothing here may touch the device
  try {
    if (mode == "logoframe" && argc == 8) {              // scores.bin nframes nlogos fpsnum fpsden out.txt
      const std::vector<char> raw = slurp(argv[2]);
      const int nframes = atoi(argv[3]), nlogos = atoi(argv[4]);
      if (raw.size() != (size_t)nframes * nlogos * 2 * sizeof(float)) { fprintf(stderr, "bad score file\n"); return 2; }
      AMTContext ctx;
      std::vector<tstring> none((size_t)nlogos, "/nonexistent.lgd");      // scores are injected; logo files are not needed
      logo::LogoFrame logof(ctx, none, 0.35f);
      logof.setResults(reinterpret_cast<const float*>(raw.data()), nframes, (unsigned)atoi(argv[5]), (unsigned)atoi(argv[6]));
      logof.selectLogo();
      logof.writeResult(argv[7]);
      printf("bestLogo=%d logoRatio=%.9g\n", logof.getBestLogo(), logof.getLogoRatio());
    } else if (mode == "timecode" && argc == 3) {
      TimecodeFile tc;
      const bool ok = tc.read(argv[2]);
      printf("ok=%d n=%zu fps=%d", (int)ok, tc.timeCodes.size(), tc.vfrTimingFps);
      for (double t : tc.timeCodes) printf(" %.6f", t);
      printf("\n");
    } else if (mode == "decimate" && argc == 4) {        // duration.txt nframes
      VideoInfo vi; vi.width = 64; vi.height = 32; vi.pixel_type = VideoInfo::CS_YV12; vi.num_frames = atoi(argv[3]);
      PClip src(new BlankClip(vi));
      AMTDecimate dec(src, argv[2], env);
      printf("frames=%d map:", dec.GetVideoInfo().num_frames);
      for (int i = 0; i < dec.GetVideoInfo().num_frames; ++i) printf(" %d", dec.SourceFrame(i));
      printf("\n");
    } else if (mode == "telecine" && argc == 7) {        // counts.bin nframes fpsnum fpsden base
      const std::vector<char> raw = slurp(argv[2]);
      const int nframes = atoi(argv[3]);
      std::vector<int32_t> counts((size_t)nframes * 12);
      if (raw.size() != counts.size() * 4) { fprintf(stderr, "bad counts file\n"); return 2; }
      memcpy(counts.data(), raw.data(), raw.size());
      printf("film_cycles=%d\n", WriteTelecineFiles(counts, nframes, (unsigned)atoi(argv[4]), (unsigned)atoi(argv[5]), argv[6]));
    } else if (mode == "fades" && argc == 8) {           // logo.lgd logof.txt|- records.bin nframes maxfade out.bin
      const int nframes = atoi(argv[5]);
      VideoInfo vi; vi.width = 64; vi.height = 32; vi.pixel_type = VideoInfo::CS_YV12; vi.num_frames = nframes;
      PClip src(new BlankClip(vi));
      PClip ana(new RecordClip(slurp(argv[4])));
      const std::string logof = std::string(argv[3]) == "-" ? "" : argv[3];
      logo::AMTEraseLogo er(src, ana, argv[2], logof, 0, atoi(argv[6]), env);
      std::vector<float> fades((size_t)nframes * 2);
      for (int n = 0; n < nframes; ++n) er.GetFades(n, fades[2 * n], fades[2 * n + 1], env);
      FILE* fp = fopen(argv[7], "wb"); fwrite(fades.data(), sizeof(float), fades.size(), fp); fclose(fp);
      printf("fades=%d\n", nframes);
    } else if (mode == "fieldplan" && argc == 3) {       // picture structures as digits, e.g. 0343663
      std::vector<uint8_t> pics;
      for (const char* p = argv[2]; *p; ++p) pics.push_back((uint8_t)(*p - '0'));
      const std::vector<av::FilterSourceFrame> fr = av::MakeFilterSourceFrames(pics);
      std::vector<int32_t> top, bottom;
      av::MakeFieldPlan(fr, top, bottom);
      printf("frames=%zu:", fr.size());
      for (size_t k = 0; k < fr.size(); ++k) printf(" %d%s/%d,%d@%.1f", fr[k].decoded, fr[k].halfDelay ? "h" : "", top[k], bottom[k], fr[k].pts);
      printf("\n");
    } else if (mode == "filterpass" && argc == 5) {      // tmpdir nframes script(vfr|cfr|none|four|throw)
      // AMTFilterSource's pass loop on a CPU-only source: which passes run, what the script sees, what gets appended.
      struct CountingClip : BlankClip {
        int* pulls;
        CountingClip(const VideoInfo& vi, int* p) : BlankClip(vi), pulls(p) {}
        PVideoFrame __stdcall GetFrame(int n, IScriptEnvironment* env) override { ++*pulls; return BlankClip::GetFrame(n, env); }
      };
      static int g_pulls = 0, g_opens = 0;
      static VideoInfo g_vi;
      g_vi.width = 64; g_vi.height = 32; g_vi.pixel_type = VideoInfo::CS_YV12; g_vi.num_frames = atoi(argv[3]);
      ConfigWrapper setting; setting.tmpDir = argv[2];
      const std::string which = argv[4];
      std::string log;
      auto hook = [](IScriptEnvironment* env) {           // AMTSource without a device: a counting blank clip
        env->AddFunction("AMTSource", "s[filter]s[outqp]b", [](AVSValue, void*, IScriptEnvironment*) -> AVSValue {
          ++g_opens; return AVSValue(PClip(new CountingClip(g_vi, &g_pulls))); }, nullptr);
      };
      FilterScript script;
      if (which == "vfr" || which == "cfr") {
        // the decision passes need the device; here the pre-process passes just leave the side files behind
        script = [&](IScriptEnvironment* env) {
          const int pass = env->GetVar("AMT_PASS").AsInt();
          const int npre = which == "vfr" ? 2 : 1;
          log += "pass" + std::to_string(pass) + "(dev=" + std::to_string(env->GetVar("AMT_DEV").AsInt()) + ",tmp=" + env->GetVar("AMT_TMP").AsString() + ") ";
          env->SetVar("AMT_PRE_PROC", AVSValue(pass < npre));
          if (pass == npre - 1) {                         // last pre-process: duration (every 5th frame lasts 2) + timecode
            const std::string base = env->GetVar("AMT_TMP").AsString();
            FILE* fd = fopen((base + ".duration.txt").c_str(), "w"); FILE* ft = fopen((base + ".timecode.txt").c_str(), "w");
            int src = 0, nout = 0;
            while (src < g_vi.num_frames) { const int d = (nout % 4 == 3 && src + 2 <= g_vi.num_frames) ? 2 : 1; fprintf(fd, "%d\n", d); fprintf(ft, "%d\n", (int)(src * 1001.0 / 30 + 0.5)); src += d; ++nout; }
            fprintf(ft, "# total: %.6f\n", src * 1.001 / 30);
            fclose(fd); fclose(ft);
          }
        };
      } else if (which == "four") {                       // a script that is ALWAYS a pre-process: 4 passes, then the output build
        script = [&](IScriptEnvironment* env) { log += "pass" + std::to_string(env->GetVar("AMT_PASS").AsInt()) + " "; env->SetVar("AMT_PRE_PROC", AVSValue(true)); };
      } else if (which == "throw") {
        script = [&](IScriptEnvironment* env) { env->ThrowError("script failed in pass %d", env->GetVar("AMT_PASS").AsInt()); };
      }
      AMTContext ctx;
      try {
        AMTFilterSource fs(ctx, setting, nullptr, 3, EncodeFileKey{ 0 }, "", script, nullptr, DEV_TYPE_CPU, hook);
        printf("script: %s\n", log.c_str());
        printf("preproc_passes=%zu opens=%d pulls=%d out_frames=%d timecodes=%zu vfrfps=%d\n", fs.getPasses().size(), g_opens, g_pulls,
               fs.getVideoInfo().num_frames, fs.getTimeCodes().size(), fs.getVfrTimingFps());
        printf("is_decimate=%d\n", dynamic_cast<AMTDecimate*>(fs.getClip().get()) != nullptr);
      } catch (const AviSynthException& e) {
        printf("AviSynthException: %s\n", e.what());
      }
    } else {
      fprintf(stderr, "unknown mode / wrong arguments\n");
      return 2;
    }
  } catch (const AvisynthError& e) {
    printf("AvisynthError: %s\n", e.msg.c_str());
    return 4;
  } catch (const std::exception& e) {
    printf("exception: %s\n", e.what());
    return 5;
  }
  return 0;
}
