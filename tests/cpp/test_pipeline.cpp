// tests/cpp/test_pipeline.cpp -- round-2 host-side pieces on a real device: AMTFilterSource's multi-pass telecine driver on
// ONE HBM-resident clip (FilteredSource.hpp:232-275,417-544), AMTSource's ingest semantics (picture structure -> frame
// list, half-delay field weave, NV12 split, FrameType property; AMTSource.hpp:291-408,524-551; StreamReform.hpp:874-904)
// and the AviSynthNeo-style device-frame path of the AMTEraseLogo(AMTAnalyzeLogo(...)) chain.
// usage: test_pipeline <mode> ...   (driven by tests/test_host_pipeline.py)
#include "../../amatsukaze_b200/host/filters.hpp"
#include <string>

static void dump(const std::string& path, const void* p, size_t n) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
  fwrite(p, 1, n, fp); fclose(fp);
}
static double now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

// packs a CPU frame (planes with pitch) into tight planar bytes
static void pack(const PVideoFrame& f, std::vector<uint8_t>& out) {
  const int pl[3] = { PLANAR_Y, PLANAR_U, PLANAR_V };
  for (int p = 0; p < 3; ++p)
    for (int y = 0; y < f->GetHeight(pl[p]); ++y)
      out.insert(out.end(), f->GetReadPtr(pl[p]) + (size_t)y * f->GetPitch(pl[p]), f->GetReadPtr(pl[p]) + (size_t)y * f->GetPitch(pl[p]) + f->GetRowSize(pl[p]));
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: test_pipeline <mode> ...\n"); return 2; }
  const std::string mode = argv[1];
  amtk_ctx* actx = nullptr;
  if (!amtk_ctx_create(0, nullptr, &actx)) { fprintf(stderr, "ctx: %s\n", amtk_last_error()); return 3; }
  int rc = 0;
  try {
    if (mode == "passes" && argc == 5) {                 // tmpdir logo.lgd(or -) script(vfr|cfr)   (clip at <tmpdir>/amts0.dat)
      ConfigWrapper setting; setting.tmpDir = argv[2];
      const std::string logo = std::string(argv[3]) == "-" ? "" : argv[3];
      AMTContext ctx;
      const double t0 = now();
      AMTFilterSource fs(ctx, setting, actx, 0, EncodeFileKey{ 0 }, logo, std::string(argv[4]) == "cfr" ? KFMCfrScript : KFMVfrScript);
      const double t1 = now();
      printf("passes: preproc=%zu uploads=%d out_frames=%d timecodes=%zu vfrfps=%d seconds=%.4f\n", fs.getPasses().size(), fs.numSourceUploads(),
             fs.getVideoInfo().num_frames, fs.getTimeCodes().size(), fs.getVfrTimingFps(), t1 - t0);
      for (auto& p : fs.getPasses()) printf("pass %d: preproc=%d frames=%d seconds=%.4f\n", p.pass, (int)p.preproc, p.frames, p.seconds);
      AMTDecimate* d = dynamic_cast<AMTDecimate*>(fs.getClip().get());
      printf("decimate=%d map:", d != nullptr);
      for (int i = 0; d && i < std::min(12, fs.getVideoInfo().num_frames); ++i) printf(" %d", d->SourceFrame(i));
      printf("\n");
      // the output clip, pulled like the encoder would: device frames, downloaded through OnCPU
      PClip cpu(new av::OnCPU(fs.getClip()));
      std::vector<uint8_t> packed;
      for (int n = 0; n < fs.getVideoInfo().num_frames; n += 5) pack(cpu->GetFrame(n, fs.getEnv()), packed);
      dump(std::string(argv[2]) + "/out_frames.bin", packed.data(), packed.size());
      printf("launches=%lld\n", (long long)amtk_ctx_launch_count(actx));
    } else if (mode == "ingest" && argc == 4) {          // clip.amtsraw2 out.bin : every output frame through GetFrame (CPU consumer)
      IScriptEnvironment envObj; IScriptEnvironment* env = &envObj;
      BindDevice(env, actx, DEV_TYPE_CPU);
      AvisynthPluginInit3(env, nullptr);
      PClip clip = env->Invoke("AMTSource", AVSValue(std::vector<AVSValue>{ AVSValue(std::string(argv[2])) })).AsClip();
      const VideoInfo vi = clip->GetVideoInfo();
      std::vector<uint8_t> packed;
      printf("ingest: frames=%d types:", vi.num_frames);
      for (int n = 0; n < vi.num_frames; ++n) { PVideoFrame f = clip->GetFrame(n, env); pack(f, packed); printf(" %d", f->GetProperty("FrameType", -1)); }
      printf("\n");
      printf("mt=%d parity=%d devtypes=%d\n", clip->SetCacheHints(CACHE_GET_MTMODE, 0), (int)clip->GetParity(0), clip->SetCacheHints(CACHE_GET_DEV_TYPE, 0));
      dump(argv[3], packed.data(), packed.size());
    } else if (mode == "devframes" && argc == 5) {       // clip.amtsraw logo.lgd outdir : the erase chain with CPU frames and with device frames
      std::vector<uint8_t> res[2];
      double secs[2] = { 0, 0 };
      for (int devmode = 0; devmode < 2; ++devmode) {
        IScriptEnvironment envObj; IScriptEnvironment* env = &envObj;
        BindDevice(env, actx, devmode ? DEV_TYPE_CUDA : DEV_TYPE_CPU);
        AvisynthPluginInit3(env, nullptr);
        PClip src = env->Invoke("AMTSource", AVSValue(std::vector<AVSValue>{ AVSValue(std::string(argv[2])) })).AsClip();
        PClip ana = env->Invoke("AMTAnalyzeLogo", AVSValue(std::vector<AVSValue>{ AVSValue(src), AVSValue(std::string(argv[3])), AVSValue(35) })).AsClip();
        PClip er = env->Invoke("AMTEraseLogo", AVSValue(std::vector<AVSValue>{ AVSValue(src), AVSValue(ana), AVSValue(std::string(argv[3])), AVSValue(), AVSValue(0), AVSValue(16) })).AsClip();
        PClip out(new av::OnCPU(er));
        const int n = src->GetVideoInfo().num_frames;
        const double t0 = now();
        int ndev = 0;
        for (int i = 0; i < n; ++i) { PVideoFrame raw = er->GetFrame(i, env); ndev += raw->IsDevice(); }
        secs[devmode] = now() - t0;
        for (int i = 0; i < n; ++i) pack(out->GetFrame(i, env), res[devmode]);
        printf("devframes mode=%d: device_frames=%d of %d, %.4f s for the chain\n", devmode, ndev, n, secs[devmode]);
      }
      printf("identical=%d\n", (int)(res[0] == res[1]));
      dump(std::string(argv[4]) + "/erased_chain.bin", res[1].data(), res[1].size());
    } else {
      fprintf(stderr, "unknown mode\n"); rc = 2;
    }
  } catch (const AvisynthError& e) {
    fprintf(stderr, "AvisynthError: %s\n", e.msg.c_str()); rc = 4;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what()); rc = 5;
  }
  amtk_ctx_destroy(actx);
  printf(rc == 0 ? "OK\n" : "FAILED\n");
  return rc;
}
