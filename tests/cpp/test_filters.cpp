// tests/cpp/test_filters.cpp -- drives the host-side mirror of the reference's filter interface
// (amatsukaze_b200/host/filters.hpp) the way the reference's callers do: CMAnalyze::logoFrame (CMAnalyze.hpp:273-317)
// and AMTFilterSource's MakeSource + pre-pass pull loop (FilteredSource.hpp:417-475).  Writes raw results into <outdir>;
// tests/test_host_filters.py compares them with the oracle.  usage: test_filters <clip.amtsraw> <logo.lgd> <outdir>
#include "../../amatsukaze_b200/host/filters.hpp"
#include <string>

static void dump(const std::string& path, const void* p, size_t n) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
  fwrite(p, 1, n, fp); fclose(fp);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: test_filters clip logo outdir\n"); return 2; }
  const std::string clipPath = argv[1], logoPath = argv[2], out = argv[3];
  amtk_ctx* actx = nullptr;
  if (!amtk_ctx_create(0, nullptr, &actx)) { fprintf(stderr, "ctx: %s\n", amtk_last_error()); return 3; }
  IScriptEnvironment envObj; IScriptEnvironment* env = &envObj;
  env->SetAmtkContext(actx);
  int rc = 0;
  try {
    printf("plugin: %s\n", AvisynthPluginInit3(env, nullptr));
    printf("params: %s | %s | %s\n", env->FunctionParams("AMTSource"), env->FunctionParams("AMTAnalyzeLogo"), env->FunctionParams("AMTEraseLogo"));
    // ---- CMAnalyze::logoFrame ----
    PClip clip = env->Invoke("AMTSource", AVSValue(std::vector<AVSValue>{ AVSValue(clipPath) })).AsClip();
    const VideoInfo vi = clip->GetVideoInfo();
    printf("clip: %dx%d %d frames %d bits\n", vi.width, vi.height, vi.num_frames, vi.BitsPerComponent());
    AMTContext ctx;
    logo::LogoFrame logof(ctx, { logoPath, out + "/does-not-exist.lgd" }, 0.35f);
    logof.scanFrames(clip, env);
    logof.selectLogo();
    logof.writeResult(out + "/logof.txt");
    printf("bestLogo=%d logoRatio=%.6f\n", logof.getBestLogo(), logof.getLogoRatio());
    dump(out + "/eval.bin", logof.results(), sizeof(float) * 2 * 2 * vi.num_frames);
    // ---- MakeSource: AMTEraseLogo(AMTAnalyzeLogo(src, logo), logo, logof, maxfade) ----
    PClip analyze = env->Invoke("AMTAnalyzeLogo", AVSValue(std::vector<AVSValue>{ AVSValue(clip), AVSValue(logoPath), AVSValue(35) })).AsClip();
    const VideoInfo avi = analyze->GetVideoInfo();
    printf("analyze vi: %dx%d %d frames type %d\n", avi.width, avi.height, avi.num_frames, avi.pixel_type);
    std::vector<logo::LogoAnalyzeFrame> recs((size_t)avi.num_frames * 8);
    for (int n = 0; n < avi.num_frames; ++n) {
      PVideoFrame f = analyze->GetFrame(n, env);
      memcpy(&recs[(size_t)n * 8], f->GetReadPtr(), sizeof(logo::LogoAnalyzeFrame) * 8);
    }
    dump(out + "/analyze.bin", recs.data(), recs.size() * sizeof(logo::LogoAnalyzeFrame));
    for (int pass = 0; pass < 2; ++pass) {          // without / with the logoframe file
      PClip erase = env->Invoke("AMTEraseLogo", AVSValue(std::vector<AVSValue>{ AVSValue(clip), AVSValue(analyze), AVSValue(logoPath),
                                 pass ? AVSValue(out + "/logof.txt") : AVSValue(), AVSValue(0), AVSValue(16) })).AsClip();
      logo::AMTEraseLogo* er = dynamic_cast<logo::AMTEraseLogo*>(erase.get());
      std::vector<float> fades((size_t)vi.num_frames * 2);
      for (int n = 0; n < vi.num_frames; ++n) er->GetFades(n, fades[2 * n], fades[2 * n + 1], env);
      dump(out + (pass ? "/fades_logof.bin" : "/fades.bin"), fades.data(), fades.size() * sizeof(float));
      if (pass == 0) {
        std::vector<uint8_t> packed;
        for (int n = 0; n < vi.num_frames; n += 7) {                 // a sample of frames through IClip::GetFrame
          PVideoFrame f = erase->GetFrame(n, env);
          const int pl[3] = { PLANAR_Y, PLANAR_U, PLANAR_V };
          for (int p = 0; p < 3; ++p)
            for (int y = 0; y < f->GetHeight(pl[p]); ++y)
              packed.insert(packed.end(), f->GetReadPtr(pl[p]) + (size_t)y * f->GetPitch(pl[p]), f->GetReadPtr(pl[p]) + (size_t)y * f->GetPitch(pl[p]) + f->GetRowSize(pl[p]));
        }
        dump(out + "/erased.bin", packed.data(), packed.size());
      }
    }
    // ---- telecine pre-pass: every frame pulled and discarded (FilteredSource.hpp:417-439) ----
    PClip pre = env->Invoke("AMTCombAnalyze", AVSValue(std::vector<AVSValue>{ AVSValue(clip), AVSValue(out + "/combstat.txt") })).AsClip();
    ReadAllFrames(pre, env);
    // ---- side files of the telecine pass and their consumers (FilteredSource.hpp:163-212,265-271,637-676) ----
    {
      AMTCombAnalyze* ca = dynamic_cast<AMTCombAnalyze*>(pre.get());
      const int film = WriteTelecineFiles(ca->Counts(env), vi.num_frames, vi.fps_numerator, vi.fps_denominator, out + "/tc");
      PClip dec = env->Invoke("AMTDecimate", AVSValue(std::vector<AVSValue>{ AVSValue(clip), AVSValue(out + "/tc.duration.txt") })).AsClip();
      TimecodeFile tc;
      const bool ok = tc.read(out + "/tc.timecode.txt");
      printf("telecine: film_cycles=%d decimated=%d timecodes=%zu total_ms=%.3f vfrfps=%d ok=%d\n", film,
             dec->GetVideoInfo().num_frames, tc.timeCodes.size(), tc.timeCodes.empty() ? 0.0 : tc.timeCodes.back(), tc.vfrTimingFps, (int)ok);
      AMTDecimate* d = dynamic_cast<AMTDecimate*>(dec.get());
      printf("decimate map:");
      for (int i = 0; i < std::min(10, dec->GetVideoInfo().num_frames); ++i) printf(" %d", d->SourceFrame(i));
      printf("\n");
      try {
        FILE* fp = fopen((out + "/bad.duration.txt").c_str(), "w"); fprintf(fp, "1\n2\n"); fclose(fp);
        env->Invoke("AMTDecimate", AVSValue(std::vector<AVSValue>{ AVSValue(clip), AVSValue(out + "/bad.duration.txt") }));
        rc = 1;
      } catch (const AvisynthError& e) { printf("expected error: %s\n", e.msg.c_str()); }
    }
    // ---- CMAnalyze ctor -> logoFrame (CMAnalyze.hpp:25-47,273-317): one match logo + the same logo as an erase logo ----
    {
      ConfigWrapper setting;
      setting.tmpDir = out;
      setting.logoPath = { out + "/does-not-exist.lgd", logoPath };
      setting.eraseLogoPath = { logoPath };
      { FILE* a = fopen(clipPath.c_str(), "rb"); FILE* b = fopen(setting.getTmpAMTSourcePath(0).c_str(), "wb");
        std::vector<char> buf(1 << 20); size_t n;
        while ((n = fread(buf.data(), 1, buf.size(), a)) > 0) fwrite(buf.data(), 1, n, b);
        fclose(a); fclose(b); }
      CMAnalyze cma(ctx, setting, 0, vi.num_frames, env);
      printf("cmanalyze: logopath=%s ratio=%.6f\n", cma.getLogoPath().c_str(), cma.getLogoRatio());
      ConfigWrapper none; none.tmpDir = out;
      CMAnalyze idle(ctx, none, 0, vi.num_frames, env);          // no logos configured: nothing runs, empty path
      printf("cmanalyze idle: '%s'\n", idle.getLogoPath().c_str());
    }
    // ---- error behaviour ----
    try {
      env->Invoke("AMTAnalyzeLogo", AVSValue(std::vector<AVSValue>{ AVSValue(clip), AVSValue(out + "/nope.lgd"), AVSValue(35) }));
      printf("ERROR: missing logo did not throw\n"); rc = 1;
    } catch (const AvisynthError& e) { printf("expected error: %s\n", e.msg.c_str()); }
    try { env->Invoke("NoSuchFilter", AVSValue()); rc = 1; } catch (const AvisynthError& e) { printf("expected error: %s\n", e.msg.c_str()); }
    printf("launches=%lld\n", (long long)amtk_ctx_launch_count(actx));
  } catch (const AvisynthError& e) {
    fprintf(stderr, "AvisynthError: %s\n", e.msg.c_str()); rc = 4;
  }
  amtk_ctx_destroy(actx);
  printf(rc == 0 ? "OK\n" : "FAILED\n");
  return rc;
}
