// amtk_b200.cu -- the C ABI of libamtk_b200.so (include/amtk_b200.h) and all kernel launches.
// Host side of the drop-in boundary: validates arguments, moves host-resident clips through HBM staging buffers,
// builds TMA descriptors and the static work partition, launches the sm_100a kernels.  No CPU compute fallback.
#include "amtk_internal.h"
#include "logo_kernels.cuh"
#include "comb_kernels.cuh"
#include "comb_stream.cuh"
#include "comb_mma.cuh"
#include "scan_kernels.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

extern "C" { static int logo_ensure_device(const amtk_logo* cl, amtk_ctx* ctx, bool need_tables); }

namespace amtk {

static thread_local std::string g_error;
static constexpr size_t kEvalSmemLimit = 226 * 1024;     // dynamic shared memory we ask for at most (227 KB per CTA on sm_100)
void set_error(const std::string& msg) { g_error = msg; }
bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
  return false;
}

LogoDev logo_dev(const amtk_logo* l) {
  LogoDev d;
  d.w = l->host.w; d.h = l->host.h; d.count = l->host.count(); d.countPad = l->countPad;
  d.blackScore = l->host.blackScore; d.A = l->dA; d.B = l->dB; d.pix = l->dPix; d.tapsT = l->dTapsT; d.scales = l->dScales;
  return d;
}

static bool ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return true;
  if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
  size_t sz = std::max(need, (size_t)1 << 20);
  if (!cuda_ok(cudaMalloc(p, sz), "cudaMalloc(scratch)")) return false;
  *cap = sz;
  return true;
}

struct DevSelect {   // RAII: make a device (the context's, or an ordinal) current for the duration of a call
  int prev = -1; bool ok = true;
  std::unique_lock<std::recursive_mutex> lock;     // held for the whole entry point when constructed from a context
  explicit DevSelect(const amtk_ctx* c) : DevSelect(c->device) { lock = std::unique_lock<std::recursive_mutex>(c->mu); }
  explicit DevSelect(int device) { ok = cuda_ok(cudaGetDevice(&prev), "cudaGetDevice") && cuda_ok(cudaSetDevice(device), "cudaSetDevice"); }
  ~DevSelect() { if (prev >= 0) cudaSetDevice(prev); }
};

static bool validate_clip(const amtk_clip* c, bool need_chroma) {
  if (!c || !c->base) { set_error("clip: null"); return false; }
  if (c->bytes_per_sample != 1 && c->bytes_per_sample != 2) { set_error("Unsupported pixel format"); return false; }
  if (c->width <= 0 || c->height <= 0 || c->num_frames <= 0) { set_error("clip: bad geometry"); return false; }
  if (c->pitch_y < c->width * c->bytes_per_sample) { set_error("clip: pitch_y smaller than a row"); return false; }
  if (need_chroma && c->pitch_uv < (c->width >> c->log_uvx) * c->bytes_per_sample) { set_error("clip: pitch_uv smaller than a row"); return false; }
  return true;
}

// A device-resident window of a clip: frames [first, first+count) of the clip are at dev_base + i*frame_stride
// (i = frame - first).
struct Window { const uint8_t* dev_base; int first; int count; };

// Runs fn(window, lo, hi) so that frames [lo,hi) (clip numbering) are resident; with need_prev the frame lo-1 is
// resident too when lo > 0.  Device clips: one call, zero copies.  Host clips: double-buffered H2D staging on the
// copy stream, overlapped with the kernels of the previous chunk.
template <typename Fn>
static int for_each_window(amtk_ctx* ctx, const amtk_clip* clip, int frame0, int nframes, bool need_prev, Fn fn) {
  if (frame0 < 0 || nframes < 0 || frame0 + nframes > clip->num_frames) AMTK_FAIL("frame range outside the clip");
  if (nframes == 0) return 1;
  if (clip->on_device) {
    Window w{ reinterpret_cast<const uint8_t*>(clip->base), 0, clip->num_frames };
    return fn(w, frame0, frame0 + nframes);
  }
  const size_t fs = (size_t)clip->frame_stride;
  size_t budget = (size_t)256 << 20;          // HBM staging per buffer (two buffers); AMTK_STAGE_MB overrides (tests)
  if (const char* e = getenv("AMTK_STAGE_MB")) budget = (size_t)std::max(1, atoi(e)) << 20;
  int per = (int)std::max<size_t>(1, std::min<size_t>((size_t)nframes, budget / fs));
  if (need_prev && per > 1) per -= 1;
  const size_t need = (size_t)(per + (need_prev ? 1 : 0)) * fs;
  if (ctx->stage_bytes < need) {
    for (int b = 0; b < 2; ++b) { if (ctx->stage[b]) cudaFree(ctx->stage[b]); ctx->stage[b] = nullptr; }
    ctx->stage_bytes = 0;
    for (int b = 0; b < 2; ++b) AMTK_CUDA(cudaMalloc(&ctx->stage[b], need));
    ctx->stage_bytes = need;
  }
  int chunk = 0;
  long long h2d = 0;
  for (int lo = frame0; lo < frame0 + nframes; lo += per, ++chunk) {
    const int hi = std::min(frame0 + nframes, lo + per);
    const int b = chunk & 1;
    const int first = (need_prev && lo > 0) ? lo - 1 : lo;
    AMTK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));      // previous user of this buffer
    AMTK_CUDA(cudaMemcpyAsync(ctx->stage[b], reinterpret_cast<const uint8_t*>(clip->base) + (size_t)first * fs,
                              (size_t)(hi - first) * fs, cudaMemcpyHostToDevice, ctx->copy_stream));
    AMTK_CUDA(cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
    AMTK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
    Window w{ reinterpret_cast<const uint8_t*>(ctx->stage[b]), first, hi - first };
    if (!fn(w, lo, hi)) return 0;
    AMTK_CUDA(cudaEventRecord(ctx->ev_done[b], ctx->stream));
    h2d += (long long)(hi - first) * (long long)fs;
  }
  ctx->h2d_bytes_last = h2d;
  return 1;
}

// ROI-only staging of HOST clips for the entry points that read nothing but a rectangle of every frame
// (LogoFrame::ScanFrame, AMTAnalyzeLogo, ReMakeLogo's fade sweep, LogoScan::AddFrame, AMTEraseLogo -- the reference
// itself touches only the ROI there, LogoScan.hpp:1559-1566,1146-1155,606-635,1374-1397).  Instead of moving whole frames
// over PCIe (3.1 MB each for a 6 KB rectangle) the library copies the rectangle rows of Y (and U, V) into a compact
// clip in HBM with strided 3-D copies and runs the same kernels on that clip with shifted coordinates.
// fn(vclip, window, lo, hi, dx, dy): vclip describes the resident data; luma sample (x, y) of the real frame is at
// (x - dx, y - dy) in vclip (chroma: shifted by dx >> log_uvx, dy >> log_uvy).  write_back copies the rectangles back
// to the host frames after fn (in-place erase).
template <typename Fn>
static int for_each_roi_window(amtk_ctx* ctx, const amtk_clip* clip, int frame0, int nframes, int rx, int ry, int rw, int rh,
                               bool with_chroma, bool write_back, Fn fn) {
  if (frame0 < 0 || nframes < 0 || frame0 + nframes > clip->num_frames) AMTK_FAIL("frame range outside the clip");
  if (nframes == 0) return 1;
  if (clip->on_device) {
    Window w{ reinterpret_cast<const uint8_t*>(clip->base), 0, clip->num_frames };
    return fn(*clip, w, frame0, frame0 + nframes, 0, 0);
  }
  const int bps = clip->bytes_per_sample, lx = clip->log_uvx, ly = clip->log_uvy;
  const int A = 16 << lx;                                              // luma byte alignment that keeps chroma 16-byte aligned
  const int xb0 = ((rx * bps) / A) * A;
  const int xb1 = std::min(clip->pitch_y, (((rx + rw) * bps + A - 1) / A) * A);
  const int cp = ((xb1 - xb0 + A - 1) / A) * A;                        // compact luma pitch (bytes); chroma pitch = cp >> lx
  const int dy = ry & ~((1 << ly) - 1);
  const int y1 = std::min(clip->height, (ry + rh + (1 << ly) - 1) & ~((1 << ly) - 1));
  const int rowsY = y1 - dy, rowsC = with_chroma ? (rowsY >> ly) : 0;
  const int cpc = cp >> lx, spanY = xb1 - xb0;
  const int xc0 = xb0 >> lx, spanC = std::min(clip->pitch_uv - xc0, spanY >> lx);
  // plane offsets are whole rows of the luma pitch, so the frame stride is a multiple of both pitches (3-D copies)
  const int rows_u = rowsY, rows_c_as_y = (int)(((long long)cpc * rowsC + cp - 1) / cp);
  const long long offU = (long long)cp * rows_u, offV = offU + (long long)cp * rows_c_as_y;
  const long long fs = with_chroma ? offV + (long long)cp * rows_c_as_y : (long long)cp * rowsY;
  if (cp <= 0 || rowsY <= 0 || (with_chroma && spanC <= 0)) AMTK_FAIL("ROI staging: empty rectangle");
  size_t budget = (size_t)256 << 20;
  if (const char* e = getenv("AMTK_STAGE_MB")) budget = (size_t)std::max(1, atoi(e)) << 20;
  const int per = (int)std::max<size_t>(1, std::min<size_t>((size_t)nframes, budget / (size_t)fs));
  const size_t need = (size_t)per * (size_t)fs;
  if (ctx->stage_bytes < need) {
    for (int b = 0; b < 2; ++b) { if (ctx->stage[b]) cudaFree(ctx->stage[b]); ctx->stage[b] = nullptr; }
    ctx->stage_bytes = 0;
    for (int b = 0; b < 2; ++b) AMTK_CUDA(cudaMalloc(&ctx->stage[b], std::max(need, (size_t)1 << 20)));
    ctx->stage_bytes = std::max(need, (size_t)1 << 20);
  }
  amtk_clip v = *clip;
  v.frame_stride = fs; v.off_u = offU; v.off_v = offV;
  v.width = cp / bps; v.height = rowsY; v.pitch_y = cp; v.pitch_uv = cpc; v.on_device = 1;
  const uint8_t* hbase = reinterpret_cast<const uint8_t*>(clip->base);
  // one strided copy per plane and chunk (cudaMemcpy3D: x = bytes of a rectangle row, y = rows, z = frames); falls back
  // to one 2-D copy per frame when the frame stride is not a whole number of rows
  auto copy_plane = [&](int pl, uint8_t* dev, int first, int count, bool to_host, cudaStream_t st) -> bool {
    const long long hoff = pl == 0 ? 0 : (pl == 1 ? clip->off_u : clip->off_v);
    const int hp = pl ? clip->pitch_uv : clip->pitch_y, dp = pl ? cpc : cp;
    const int span = pl ? spanC : spanY, rows = pl ? rowsC : rowsY;
    const long long doff = pl == 0 ? 0 : (pl == 1 ? offU : offV);
    uint8_t* h = const_cast<uint8_t*>(hbase) + (long long)first * clip->frame_stride + hoff + (long long)(pl ? (dy >> ly) : dy) * hp + (pl ? xc0 : xb0);
    uint8_t* d = dev + doff;
    if (clip->frame_stride % hp == 0 && fs % dp == 0) {
      cudaMemcpy3DParms p3; memset(&p3, 0, sizeof(p3));
      const cudaPitchedPtr hptr = make_cudaPitchedPtr(h, (size_t)hp, (size_t)hp, (size_t)(clip->frame_stride / hp));
      const cudaPitchedPtr dptr = make_cudaPitchedPtr(d, (size_t)dp, (size_t)dp, (size_t)(fs / dp));
      p3.srcPtr = to_host ? dptr : hptr; p3.dstPtr = to_host ? hptr : dptr;
      p3.extent = make_cudaExtent((size_t)span, (size_t)rows, (size_t)count);
      p3.kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
      return cuda_ok(cudaMemcpy3DAsync(&p3, st), "cudaMemcpy3DAsync(roi)");
    }
    for (int f = 0; f < count; ++f) {
      uint8_t* hf = h + (long long)f * clip->frame_stride; uint8_t* df = d + (long long)f * fs;
      if (!cuda_ok(to_host ? cudaMemcpy2DAsync(hf, hp, df, dp, span, rows, cudaMemcpyDeviceToHost, st)
                           : cudaMemcpy2DAsync(df, dp, hf, hp, span, rows, cudaMemcpyHostToDevice, st), "cudaMemcpy2DAsync(roi)")) return false;
    }
    return true;
  };
  int chunk = 0;
  for (int lo = frame0; lo < frame0 + nframes; lo += per, ++chunk) {
    const int hi = std::min(frame0 + nframes, lo + per);
    const int b = chunk & 1;
    uint8_t* dev = reinterpret_cast<uint8_t*>(ctx->stage[b]);
    AMTK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0));      // previous user of this buffer
    for (int pl = 0; pl < (with_chroma ? 3 : 1); ++pl) if (!copy_plane(pl, dev, lo, hi - lo, false, ctx->copy_stream)) return 0;
    AMTK_CUDA(cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
    AMTK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
    v.base = dev; v.num_frames = hi - lo;
    Window w{ dev, lo, hi - lo };
    if (!fn(v, w, lo, hi, xb0 / bps, dy)) return 0;
    if (write_back)
      for (int pl = 0; pl < (with_chroma ? 3 : 1); ++pl) if (!copy_plane(pl, dev, lo, hi - lo, true, ctx->stream)) return 0;
    AMTK_CUDA(cudaEventRecord(ctx->ev_done[b], ctx->stream));
  }
  ctx->h2d_bytes_last = (long long)nframes * ((long long)spanY * rowsY + (with_chroma ? 2LL * spanC * rowsC : 0));
  return 1;
}

// ---------------------------------------------------------------------------------------------------------
// logo evaluation launches
// ---------------------------------------------------------------------------------------------------------
struct EvalSpec {
  const amtk_logo* logo;      // evaluated logo (mask built)
  int roi_x, roi_y;           // staged rectangle (the FULL logo rectangle, frame coordinates)
  int roi_w, roi_h;
  int src_mode, src_off, src_stride;
  int nfades; const float* fades;
  int take_abs;
  int out_off, out_fade_stride;     // where in an output row the values go
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is sticky per (device, kernel): set it once, raise it when a larger logo comes along
static bool want_smem(amtk_ctx* ctx, const void* fn, int bytes) {
  for (auto& e : ctx->smem_attr) if (e.first == fn) {
    if (e.second >= bytes) return true;
    if (!cuda_ok(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute")) return false;
    e.second = bytes; return true;
  }
  if (!cuda_ok(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), "cudaFuncSetAttribute")) return false;
  ctx->smem_attr.emplace_back(fn, bytes);
  return true;
}

static int launch_eval(amtk_ctx* ctx, const amtk_clip* clip, const Window& win, int lo, int hi, int pitch_elems,
                       const EvalSpec& sp, float* dout, int out_frame_stride, int out_row0) {
  const amtk::HostLogo& hl = sp.logo->host;
  if (!logo_ensure_device(sp.logo, ctx, true)) return 0;
  if (sp.nfades < 1 || sp.nfades > kMaxFades) AMTK_FAIL("too many fade levels");
  const int count = hl.count();
  const int bits = clip->bits_per_sample;
  const float maxv = (float)((1 << bits) - 1);
  if (count == 0) {   // degenerate logo: CorrelationScore is 0 -> 0/blackScore
    AMTK_FAIL("logo has no feature pixels");
  }
  const int countPad = sp.logo->countPad;
  // ROI staging: TMA box of (roi_w rounded up to 16 bytes) x roi_h samples per frame, when the layout allows it
  const int bps = clip->bytes_per_sample;
  // the box starts at imgx rounded down to 16 bytes: the TMA unit raises "illegal instruction" when the innermost
  // start address is not 16-byte aligned (measured on B200; tools/tma_align_probe.py)
  const int box_x = ((sp.roi_x * bps) & ~15) / bps;
  const int box_w = ((((sp.roi_x - box_x) + sp.roi_w) * bps + 15) & ~15) / bps;
  const long long pitch_bytes = (long long)pitch_elems * bps;
  const long long plane_rows = ((long long)clip->pitch_y * clip->height) / pitch_bytes;     // rows as addressed with pitch_elems
  const bool tma_ok = ctx->encode_tiled && box_w <= 256 && sp.roi_h <= 256 && (pitch_bytes & 15) == 0 &&
                      (clip->frame_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(win.dev_base) & 15) == 0;
  // shared-memory plan: everything for logos up to ~100x100, A/B through L1 up to ~16k px, one fade per pass beyond
  int ab_smem = 1, pair_fades = 1;
  size_t smem = logo_scores_smem_bytes(sp.roi_w * sp.roi_h, hl.w * hl.h, box_w * sp.roi_h * bps, ab_smem, pair_fades);
  if (smem > kEvalSmemLimit) { ab_smem = 0; smem = logo_scores_smem_bytes(sp.roi_w * sp.roi_h, hl.w * hl.h, box_w * sp.roi_h * bps, ab_smem, pair_fades); }
  if (smem > kEvalSmemLimit) { pair_fades = 0; smem = logo_scores_smem_bytes(sp.roi_w * sp.roi_h, hl.w * hl.h, box_w * sp.roi_h * bps, ab_smem, pair_fades); }
  if (smem > kEvalSmemLimit) AMTK_FAIL("logo too large for the shared-memory evaluation path (more than ~24k pixels)");
  CUtensorMap roi_map;
  memset(&roi_map, 0, sizeof(roi_map));
  if (tma_ok) {
    cuuint64_t gdim[3] = { (cuuint64_t)pitch_elems, (cuuint64_t)plane_rows, (cuuint64_t)win.count };
    cuuint64_t gstr[2] = { (cuuint64_t)pitch_bytes, (cuuint64_t)clip->frame_stride };
    cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)sp.roi_h, 1u };
    cuuint32_t estr[3] = { 1u, 1u, 1u };
    CUresult r = ctx->encode_tiled(&roi_map, bps == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3,
                                   const_cast<uint8_t*>(win.dev_base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) AMTK_FAIL("cuTensorMapEncodeTiled(roi) failed (" + std::to_string((int)r) + ")");
  }
  // frames per launch bounded by the score scratch (<= 96 MB)
  const size_t per_frame = (size_t)sp.nfades * countPad * sizeof(float);
  const int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)(hi - lo), ((size_t)96 << 20) / per_frame));
  if (!ensure(&ctx->scratch, &ctx->scratch_bytes, ctx->scratch_off + per_frame * batch)) return 0;
  for (int f0 = lo; f0 < hi; f0 += batch) {
    const int n = std::min(batch, hi - f0);
    EvalJob job;
    job.ybase = win.dev_base; job.frame_stride = clip->frame_stride; job.pitch = pitch_elems;
    job.frame0 = f0 - win.first; job.nframes = n;
    job.imgx = sp.roi_x; job.imgy = sp.roi_y;
    job.roi_w = sp.roi_w; job.roi_h = sp.roi_h;
    job.src_mode = sp.src_mode; job.src_off = sp.src_off; job.src_stride = sp.src_stride;
    job.logo = logo_dev(sp.logo); job.maxv = maxv; job.nfades = sp.nfades;
    for (int i = 0; i < sp.nfades; ++i) job.fades[i] = sp.fades[i];
    job.scores = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->scratch) + ctx->scratch_off);
    job.use_tma = tma_ok ? 1 : 0; job.roi_box_w = box_w; job.roi_box_x = box_x; job.roi_map = roi_map;
    job.ab_smem = ab_smem; job.pair_fades = pair_fades;
    const int slices3 = (count + kEvalThreads * 3 - 1) / (kEvalThreads * 3);
    int pxt = 3, slices = slices3;
    if (count <= kEvalThreads) { pxt = 1; slices = 1; }
    else if (count <= kEvalThreads * 2) { pxt = 2; slices = 1; }
    // one CTA per SM (the kernel needs the whole register file): a single wave, so the tap tables are loaded once per SM
    const int lanes = std::max(1, std::min(n, (ctx->sm_count * ctx->knobs.eval_waves) / slices));
    dim3 grid(slices, lanes);
    const bool u16 = clip->bytes_per_sample == 2;
    const bool w64 = hl.w == 64 && sp.roi_w == 64 && ctx->knobs.eval_cw;     // compile-time width variant
    const bool wh64 = w64 && hl.h == 64 && sp.roi_h == 64;                   // ... and height (LogoFrame::ScanFrame on 64x64 logos)
#define AMTK_LAUNCH_SCORES(T, P)                                                                              \
  do {                                                                                                        \
    void (*kfn)(const EvalJob) = wh64 ? logo_scores_kernel<T, P, 64, 64> : w64 ? logo_scores_kernel<T, P, 64, 0> : logo_scores_kernel<T, P, 0, 0>; \
    if (!want_smem(ctx, (const void*)kfn, (int)smem)) return 0;                                                \
    kfn<<<grid, kEvalThreads, smem, ctx->stream>>>(job);                                                      \
  } while (0)
    if (!u16) { if (pxt == 1) AMTK_LAUNCH_SCORES(uint8_t, 1); else if (pxt == 2) AMTK_LAUNCH_SCORES(uint8_t, 2); else AMTK_LAUNCH_SCORES(uint8_t, 3); }
    else      { if (pxt == 1) AMTK_LAUNCH_SCORES(uint16_t, 1); else if (pxt == 2) AMTK_LAUNCH_SCORES(uint16_t, 2); else AMTK_LAUNCH_SCORES(uint16_t, 3); }
#undef AMTK_LAUNCH_SCORES
    AMTK_CUDA(cudaGetLastError());
    const int total = n * sp.nfades;
    float* sum_out = dout + (size_t)(f0 - out_row0) * out_frame_stride;
    const size_t sum_smem = (size_t)32 * (countPad + 4) * sizeof(float);
    if (sum_smem <= 200 * 1024) {
      if (!want_smem(ctx, (const void*)logo_sum_bulk_kernel, (int)sum_smem)) return 0;
      logo_sum_bulk_kernel<<<(total + 31) / 32, 32, sum_smem, ctx->stream>>>(
          job.scores, count, countPad, n, sp.nfades, hl.blackScore, sp.take_abs, sum_out, out_frame_stride, sp.out_off, sp.out_fade_stride);
    } else {
      logo_sum_kernel<<<(total + kSumThreads - 1) / kSumThreads, kSumThreads, 0, ctx->stream>>>(
          job.scores, count, countPad, n, sp.nfades, hl.blackScore, sp.take_abs, sum_out, out_frame_stride, sp.out_off, sp.out_fade_stride);
    }
    AMTK_CUDA(cudaGetLastError());
    ctx->launches += 2;
  }
  return 1;
}

static bool roi_inside(const amtk::HostLogo& full, const amtk_clip* clip, int pitch_elems) {
  // the ROI rows must lie inside the frame allocation as addressed with pitch_elems
  if (full.imgx < 0 || full.imgy < 0) return false;
  const long long last = (long long)full.imgx + full.w - 1 + (long long)(full.imgy + full.h - 1) * pitch_elems;
  const long long plane_elems = (long long)clip->pitch_y / clip->bytes_per_sample * clip->height;
  return full.imgx + full.w <= pitch_elems && last < plane_elems;
}

// ---------------------------------------------------------------------------------------------------------
// comb launch
// ---------------------------------------------------------------------------------------------------------
static int comb_thresholds_ok(const amtk_comb_params* p, int bytes_per_sample) {
  if (bytes_per_sample == 2) {
    const int all[6] = { p->th_move_y, p->th_shima_y, p->th_lshima_y, p->th_move_c, p->th_shima_c, p->th_lshima_c };
    for (int v : all) if (v < 1) { set_error("comb: thresholds must be >= 1"); return 0; }
    if (p->th_move_y > 32768 || p->th_move_c > 32768) { set_error("comb: th_move must be in [1,32768] for 16-bit samples"); return 0; }
    return 1;
  }
  const int m[2] = { p->th_move_y, p->th_move_c };
  const int s[4] = { p->th_shima_y, p->th_lshima_y, p->th_shima_c, p->th_lshima_c };
  for (int v : m) if (v < 1 || v > 128) { set_error("comb: th_move must be in [1,128]"); return 0; }
  for (int v : s) if (v < 1 || v > 2047) { set_error("comb: th_shima/th_lshima must be in [1,2047]"); return 0; }
  return 1;
}

// Everything the host needs to know about one compiled comb-kernel variant.
struct CombVariant {
  int R, strip, stages, sync, TH, boxH, threads, smem;
  void (*kernel)(const CombArgs);        // 8-bit samples
  void (*kernel16)(const CombArgs);      // 16-bit samples (only for the default variants; else NULL)
};
template <typename Cfg> static CombVariant make_variant() {
  return CombVariant{ Cfg::R, Cfg::STRIP, Cfg::STAGES, Cfg::SYNC, Cfg::TH, Cfg::BOXH, Cfg::THREADS, Cfg::SMEM, comb_tma_kernel<Cfg, 1>, nullptr };
}
template <typename Cfg> static CombVariant make_variant16() {     // 8-byte strips: 8 px of u8 or 4 px of u16
  return CombVariant{ Cfg::R, Cfg::STRIP, Cfg::STAGES, Cfg::SYNC, Cfg::TH, Cfg::BOXH, Cfg::THREADS, Cfg::SMEM, comb_tma_kernel<Cfg, 1>, comb_tma_kernel<Cfg, 2> };
}
static const CombVariant* comb_variants(int* n) {
  static const CombVariant v[] = {
    // production: rows-per-run chosen per clip (pick_comb_R), 8-byte strips, 3-stage ring, block barrier per tile-frame
    make_variant16<CombCfg<15, 8, 3, 0>>(), make_variant16<CombCfg<16, 8, 3, 0>>(), make_variant16<CombCfg<17, 8, 3, 0>>(),
    // kept for tools/tune_comb.py: 2- and 4-stage rings, mbarrier 'release' sync (all measured below the production
    // variant)
    make_variant<CombCfg<17, 8, 2, 0>>(), make_variant<CombCfg<17, 8, 4, 0>>(), make_variant<CombCfg<17, 8, 3, 1>>(),
  };
  *n = (int)(sizeof(v) / sizeof(v[0]));
  return v;
}

// rows per run: the R in {15,16,17} that wastes the fewest rows over luma + chroma (1080/540 -> 17, 720/360 -> 15)
static int pick_comb_R(int hY, int hC) {
  int best = 16; long long best_waste = -1;
  for (int R = 17; R >= 15; --R) {
    const int th = 8 * R;
    const long long waste = (long long)((hY + th - 1) / th) * th - hY + 2LL * (((hC + th - 1) / th) * th - hC) / 2;
    if (best_waste < 0 || waste < best_waste) { best_waste = waste; best = R; }
  }
  return best;
}

// ---- round-2 streaming kernel (comb_stream.cuh): independent warp streams, 8-bit samples ----------------------
struct WsVariant { int R, stages, warps, bps, TH, boxH, smem; void (*kernel)(const WsArgs); };
template <typename Cfg> static WsVariant make_ws() { return WsVariant{ Cfg::R, Cfg::STAGES, Cfg::WARPS, Cfg::BPS, Cfg::TH, Cfg::BOXH, Cfg::SMEM, comb_ws_kernel<Cfg> }; }
static const WsVariant* ws_variants(int* n) {
  static const WsVariant v[] = { make_ws<WsCfg<17, 2>>(), make_ws<WsCfg<15, 2>>(), make_ws<WsCfg<16, 2>>(), make_ws<WsCfg<9, 2>>(), make_ws<WsCfg<15, 3>>(), make_ws<WsCfg<12, 2>>(), make_ws<WsCfg<10, 2>>(),
                                 make_ws<WsCfg<15, 2, 7>>(), make_ws<WsCfg<13, 2, 5>>(), make_ws<WsCfg<15, 2, 2>>(), make_ws<WsCfg<15, 3, 3>>(),
                                 // 16-bit containers with <= 10 significant bits (YUV420P10): integer-lane stencil, no conversion
                                 make_ws<WsCfg<15, 2, 4, 2>>(), make_ws<WsCfg<16, 2, 4, 2>>(), make_ws<WsCfg<17, 2, 4, 2>>() };
  *n = (int)(sizeof(v) / sizeof(v[0]));
  return v;
}
// rows per run for 4-run tiles: fewest wasted rows over luma + chroma, ties to the larger R (less halo per row)
static int pick_ws_R(int hY, int hC) {
  int best = 17; long long best_waste = -1;
  for (int R : { 17, 16, 15 }) {
    const int th = 4 * R;
    const long long waste = 2LL * ((long long)((hY + th - 1) / th) * th - hY) + 2LL * ((long long)((hC + th - 1) / th) * th - hC);
    if (best_waste < 0 || waste < best_waste) { best_waste = waste; best = R; }
  }
  return best;
}

static int launch_comb_ws(amtk_ctx* ctx, const amtk_clip* clip, const Window& win, int lo, int hi,
                          const amtk_comb_params* prm, int* dcounts, int out_row0) {
  const int hY = clip->height, hC = clip->height >> clip->log_uvy;
  const int wY = clip->width, wC = clip->width >> clip->log_uvx;
  const int R = ctx->knobs.comb_R ? ctx->knobs.comb_R : pick_ws_R(hY, hC);
  int nvar = 0; const WsVariant* vars = ws_variants(&nvar); const WsVariant* V = nullptr;
  const int bps = clip->bytes_per_sample;
  for (int i = 0; i < nvar; ++i) if (vars[i].R == R && vars[i].stages == ctx->knobs.comb_ws_stages && vars[i].warps == ctx->knobs.comb_ws_warps && vars[i].bps == bps) V = &vars[i];
  if (!V) AMTK_FAIL("comb: no warp-stream kernel variant for the requested AMTK_COMB_* settings");
  const int WW = V->warps;
  WsArgs args;
  memset(&args, 0, sizeof(args));
  const CUtensorMapL2promotion promo = ctx->knobs.comb_l2 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : ctx->knobs.comb_l2 == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                       ctx->knobs.comb_l2 == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  for (int pl = 0; pl < 3; ++pl) {
    const long long off = pl == 0 ? 0 : (pl == 1 ? clip->off_u : clip->off_v);
    cuuint64_t gdim[3] = { (cuuint64_t)(pl ? wC : wY) * bps, (cuuint64_t)(pl ? hC : hY), (cuuint64_t)win.count };     // x in BYTES (u8 element type also for 16-bit containers)
    cuuint64_t gstr[2] = { (cuuint64_t)(pl ? clip->pitch_uv : clip->pitch_y), (cuuint64_t)clip->frame_stride };
    cuuint32_t box[3] = { (cuuint32_t)kWsTW, (cuuint32_t)V->boxH, 1u };
    cuuint32_t estr[3] = { 1u, 1u, 1u };
    if (ctx->encode_tiled(&args.map[pl], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(win.dev_base) + off, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      AMTK_FAIL("cuTensorMapEncodeTiled failed");
  }
  // chroma remainder columns of at most 64 bytes: U and V side by side in one tile through a 4-D map (x, plane, y, frame)
  const int remC = (wC * bps) % kWsTW;
  const long long uv_dist = clip->off_v - clip->off_u;
  const bool pair_uv = ctx->knobs.comb_merge_uv && remC > 0 && remC <= kWsTW / 2 && uv_dist > 0 && (uv_dist & 15) == 0;
  if (pair_uv) {
    cuuint64_t gdim[4] = { (cuuint64_t)wC * bps, 2u, (cuuint64_t)hC, (cuuint64_t)win.count };
    cuuint64_t gstr[3] = { (cuuint64_t)uv_dist, (cuuint64_t)clip->pitch_uv, (cuuint64_t)clip->frame_stride };
    cuuint32_t box[4] = { (cuuint32_t)(kWsTW / 2), 2u, (cuuint32_t)V->boxH, 1u };
    cuuint32_t estr[4] = { 1u, 1u, 1u, 1u };
    if (ctx->encode_tiled(&args.map_uv, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(win.dev_base) + clip->off_u, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      AMTK_FAIL("cuTensorMapEncodeTiled(uv pair) failed");
  }
  const int tyY = (hY + V->TH - 1) / V->TH, tyC = (hC + V->TH - 1) / V->TH;
  int tile0 = 0, nc = 0;
  auto thresholds = [&](WsClass& C, bool chroma) {
    C.cls = chroma ? 1 : 0; C.H = chroma ? hC : hY;
    const int tM = chroma ? prm->th_move_c : prm->th_move_y, tS = chroma ? prm->th_shima_c : prm->th_shima_y, tL = chroma ? prm->th_lshima_c : prm->th_lshima_y;
    if (bps == 1) {
      C.thM = (unsigned)(0x80 - tM) * 0x01010101u;
      C.thS = (unsigned)tS * 0x00010001u;     // integer k in [1,2047] IS the fp16 bit pattern of k*2^-24
      C.thL = (unsigned)tL * 0x00010001u;
    } else {                                   // <= 10-bit samples: |d| <= 1023, |r| <= 6138, r is compared as 8192 + |r| (bit patterns)
      C.thM = (unsigned)std::min(tM, 2047) * 0x00010001u;
      C.thS = (unsigned)(8192 + std::min(tS, 8191)) * 0x00010001u;
      C.thL = (unsigned)(8192 + std::min(tL, 8191)) * 0x00010001u;
    }
  };
  for (int pl = 0; pl < 3; ++pl) {                         // 128-byte tiles of Y, U, V
    WsClass& C = args.cl[nc];
    const int w = (pl ? wC : wY) * bps;                    // bytes
    C.kind = 0; C.map = pl; thresholds(C, pl != 0);
    C.tilesX = (pl && pair_uv) ? w / kWsTW : (w + kWsTW - 1) / kWsTW;
    C.tile0 = tile0; C.ntiles = C.tilesX * (pl ? tyC : tyY);
    if (C.ntiles == 0) continue;
    tile0 += C.ntiles; ++nc;
  }
  if (pair_uv) {
    WsClass& C = args.cl[nc];
    C.kind = 1; C.x0 = wC * bps - remC; C.tilesX = 1; thresholds(C, true);
    C.tile0 = tile0; C.ntiles = tyC; tile0 += C.ntiles; ++nc;
  }
  args.nclasses = nc;
  const int ntiles = tile0, nf = hi - lo;
  amtk_ctx::CombPlan& plan = ctx->plan;
  if (plan.occ_kernel != (const void*)V->kernel) {          // once per kernel variant, not per launch
    int occ = 0;
    AMTK_CUDA(cudaFuncSetAttribute(V->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V->smem));
    AMTK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, V->kernel, 32 * WW, V->smem));
    if (occ < 1) AMTK_FAIL("comb kernel does not fit on an SM");
    plan.occ = occ; plan.occ_kernel = (const void*)V->kernel; plan.valid = false;
  }
  int occ = plan.occ;
  if (ctx->knobs.comb_ctas > 0) occ = std::min(occ, ctx->knobs.comb_ctas);
  // Work queue: every tile's frame range is cut into items; the warps pull items from a global counter.  Long items
  // (little halo overhead: one extra tile load per item) make up the first ~85 % of the work, short ones the rest, so
  // that all warps run dry within about one short item of each other.  The item list depends only on the geometry and
  // the frame range, so it stays on the device between calls (a 1-frame GetFrame call re-uses it without any copy).
  const long long total = (long long)ntiles * nf;
  const int nwarps = ctx->sm_count * occ * WW;
  const int grid = (int)std::min<long long>((long long)ctx->sm_count * occ, (total + WW - 1) / WW);
  const int f0 = lo - win.first;
  if (!(plan.valid && plan.wY == wY && plan.hY == hY && plan.wC == wC && plan.hC == hC && plan.nf == nf && plan.f0 == f0 &&
        plan.R == V->R + 100 * bps && plan.item == ctx->knobs.comb_item + 1000 * ctx->knobs.comb_tail && plan.ctas == occ * WW)) {
    int big = ctx->knobs.comb_item > 0 ? ctx->knobs.comb_item : 64, small = std::max(4, big / 4);
    // each warp should see at least ~6 big items; shrink for short clips
    while (big > 8 && (long long)ntiles * (nf / big) < 6LL * nwarps) { big /= 2; small = std::max(4, big / 4); }
    const int tail_frames = std::min(nf, std::max(small, (int)(nf * 0.15)));
    const int head_frames = nf - tail_frames;
    // third tier (AMTK_COMB_TAIL = frames per item, 0 = off): the last ~4 % of the frames in very short items, so that the warps
    // run dry within one such item of each other (a 16-frame item is ~48 us of a warp stream's time; measured -1.2 %)
    const int tiny = ctx->knobs.comb_tail;
    const int end_frames = (tiny > 0 && tiny < small) ? std::min(tail_frames, std::max(tiny, (int)(nf * 0.04))) : 0;
    const int mid_end = nf - end_frames;
    std::vector<CombSegment> segs;
    segs.reserve((size_t)ntiles * (head_frames / big + tail_frames / small + (tiny > 0 ? end_frames / tiny : 0) + 3));
    for (int t = 0; t < ntiles; ++t)
      for (int f = 0; f < head_frames; f += big) segs.push_back(CombSegment{ t, f0 + f, f0 + std::min(head_frames, f + big) });
    for (int t = 0; t < ntiles; ++t)
      for (int f = head_frames; f < mid_end; f += small) segs.push_back(CombSegment{ t, f0 + f, f0 + std::min(mid_end, f + small) });
    for (int t = 0; t < ntiles; ++t)
      for (int f = mid_end; f < nf; f += tiny) segs.push_back(CombSegment{ t, f0 + f, f0 + std::min(nf, f + tiny) });
    const size_t seg_bytes = segs.size() * sizeof(CombSegment);
    plan.q_off = (seg_bytes + 255) & ~(size_t)255;
    plan.valid = false;
    if (!ensure(&plan.dev, &plan.cap, plan.q_off + 256)) return 0;
    AMTK_CUDA(cudaMemcpyAsync(plan.dev, segs.data(), seg_bytes, cudaMemcpyHostToDevice, ctx->stream));
    AMTK_CUDA(cudaStreamSynchronize(ctx->stream));           // pageable source vector dies at the end of this scope
    plan.nitems = (int)segs.size();
    plan.wY = wY; plan.hY = hY; plan.wC = wC; plan.hC = hC; plan.nf = nf; plan.f0 = f0; plan.R = V->R + 100 * bps;
    plan.item = ctx->knobs.comb_item + 1000 * ctx->knobs.comb_tail; plan.ctas = occ * WW; plan.valid = true;
  }
  AMTK_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(plan.dev) + plan.q_off, 0, 256, ctx->stream));
  args.segs = reinterpret_cast<const CombSegment*>(plan.dev);
  args.nitems = plan.nitems;
  args.queue = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(plan.dev) + plan.q_off);
  args.counts = dcounts;
  args.out_frame0 = out_row0 - win.first;
  AMTK_CUDA(cudaMemsetAsync(dcounts + (size_t)(lo - out_row0) * 12, 0, (size_t)nf * 12 * sizeof(int), ctx->stream));
  std::pair<cudaEvent_t, cudaEvent_t> ev{ nullptr, nullptr };
  if (ctx->timing) {
    if (!ctx->timing_pool.empty()) { ev = ctx->timing_pool.back(); ctx->timing_pool.pop_back(); }
    else { AMTK_CUDA(cudaEventCreate(&ev.first)); AMTK_CUDA(cudaEventCreate(&ev.second)); }
    AMTK_CUDA(cudaEventRecord(ev.first, ctx->stream));
  }
  args.prefetch = ctx->knobs.comb_ws_prefetch;
  if (ctx->want_side_mark) {                                // fused step: the logo kernels may be queued on the side stream from here on
    AMTK_CUDA(cudaEventRecord(ctx->ev_side, ctx->stream));
    ctx->want_side_mark = false;
  }
  V->kernel<<<grid, 32 * WW, V->smem, ctx->stream>>>(args);
  AMTK_CUDA(cudaGetLastError());
  if (ctx->timing) { AMTK_CUDA(cudaEventRecord(ev.second, ctx->stream)); ctx->timing_events.push_back(ev); }
  ctx->launches += 1;
  return 1;
}

// ---- tensor-core streaming kernel (comb_mma.cuh): one CTA = one tile stream, stencil as two tcgen05.mma per tile-frame ----
static int launch_comb_mma(amtk_ctx* ctx, const amtk_clip* clip, const Window& win, int lo, int hi,
                           const amtk_comb_params* prm, int* dcounts, int out_row0) {
  const int hY = clip->height, hC = clip->height >> clip->log_uvy;
  const int wY = clip->width, wC = clip->width >> clip->log_uvx;
  WsArgs args;
  memset(&args, 0, sizeof(args));
  const CUtensorMapL2promotion promo = ctx->knobs.comb_l2 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : ctx->knobs.comb_l2 == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                       ctx->knobs.comb_l2 == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  for (int pl = 0; pl < 3; ++pl) {
    const long long off = pl == 0 ? 0 : (pl == 1 ? clip->off_u : clip->off_v);
    cuuint64_t gdim[3] = { (cuuint64_t)(pl ? wC : wY), (cuuint64_t)(pl ? hC : hY), (cuuint64_t)win.count };
    cuuint64_t gstr[2] = { (cuuint64_t)(pl ? clip->pitch_uv : clip->pitch_y), (cuuint64_t)clip->frame_stride };
    cuuint32_t box[3] = { (cuuint32_t)kMmTW, (cuuint32_t)kMmBoxH, 1u };
    cuuint32_t estr[3] = { 1u, 1u, 1u };
    // 128-byte swizzle: the staged tile is the MN-major A operand of the MMA as it lands
    if (ctx->encode_tiled(&args.map[pl], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(win.dev_base) + off, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      AMTK_FAIL("cuTensorMapEncodeTiled failed");
  }
  const int tyY = (hY + kMmTH - 1) / kMmTH, tyC = (hC + kMmTH - 1) / kMmTH;
  int tile0 = 0, nc = 0;
  for (int pl = 0; pl < 3; ++pl) {
    WsClass& C = args.cl[nc];
    const bool chroma = pl != 0;
    C.kind = 0; C.map = pl; C.cls = chroma ? 1 : 0; C.H = chroma ? hC : hY;
    C.thM = (unsigned)(0x80 - (chroma ? prm->th_move_c : prm->th_move_y)) * 0x01010101u;
    C.thS = (unsigned)(chroma ? prm->th_shima_c : prm->th_shima_y) * 0x00010001u;
    C.thL = (unsigned)(chroma ? prm->th_lshima_c : prm->th_lshima_y) * 0x00010001u;
    C.tilesX = ((chroma ? wC : wY) + kMmTW - 1) / kMmTW;
    C.tile0 = tile0; C.ntiles = C.tilesX * (chroma ? tyC : tyY);
    if (C.ntiles == 0) continue;
    tile0 += C.ntiles; ++nc;
  }
  args.nclasses = nc;
  const int ntiles = tile0, nf = hi - lo;
  amtk_ctx::CombPlan& plan = ctx->plan;
  // NS tiles per CTA step (AMTK_COMB_MMA = 1 or 2).  Dynamic shared memory is padded so that exactly 4 / NS CTAs share an
  // SM: their 128 * NS tensor-memory columns each add up to all 512.
  const int NS = ctx->knobs.comb_mma == 2 ? 2 : 1;
  void (*kern)(const WsArgs) = NS == 2 ? comb_mma_kernel<2> : comb_mma_kernel<1>;
  const int smem = NS == 2 ? 110 * 1024 : 55 * 1024;
  static_assert(2 * kMmSmemPerStream + kMmBandBytes + 1024 <= 110 * 1024 && kMmSmemPerStream + kMmBandBytes + 1024 <= 55 * 1024, "comb_mma: shared-memory budget");
  if (plan.occ_kernel != (const void*)kern) {
    AMTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // Residency is set by construction, not queried (the occupancy API answers 1 for a kernel that allocates tensor memory)
    plan.occ = 4 / NS; plan.occ_kernel = (const void*)kern; plan.valid = false;
  }
  int occ = plan.occ;
  if (ctx->knobs.comb_ctas > 0) occ = std::min(occ, ctx->knobs.comb_ctas);
  const int npairs_t = (ntiles + NS - 1) / NS;               // a CTA streams NS tiles at a time
  const long long total = (long long)npairs_t * nf;
  const int nstreams = ctx->sm_count * occ;
  const int grid = (int)std::min<long long>(nstreams, total);
  const int f0 = lo - win.first;
  if (!(plan.valid && plan.wY == wY && plan.hY == hY && plan.wC == wC && plan.hC == hC && plan.nf == nf && plan.f0 == f0 &&
        plan.R == 1000 * NS + kMmTH && plan.item == ctx->knobs.comb_item && plan.ctas == occ)) {
    int big = ctx->knobs.comb_item > 0 ? ctx->knobs.comb_item : 64, small = std::max(4, big / 4);
    while (big > 8 && (long long)npairs_t * (nf / big) < 6LL * nstreams) { big /= 2; small = std::max(4, big / 4); }
    const int tail_frames = std::min(nf, std::max(small, (int)(nf * 0.15)));
    const int head_frames = nf - tail_frames;
    // Items come in PAIRS (2i, 2i+1) with the same frame range; an odd tile count is padded with a filler (tile = ~t: the
    // last tile once more, results dropped).  Frame-block-major order: CTAs running at the same time work on the same
    // frames of neighbouring tiles, so halo rows and straddled lines are shared through L2.
    std::vector<CombSegment> segs;
    segs.reserve((size_t)NS * npairs_t * (head_frames / big + tail_frames / small + 2));
    auto push_block = [&](int fa, int fz) {
      for (int t = 0; t < ntiles; t += NS) {
        segs.push_back(CombSegment{ t, fa, fz });
        if (NS == 2) segs.push_back(CombSegment{ t + 1 < ntiles ? t + 1 : ~t, fa, fz });
      }
    };
    for (int f = 0; f < head_frames; f += big) push_block(f0 + f, f0 + std::min(head_frames, f + big));
    for (int f = head_frames; f < nf; f += small) push_block(f0 + f, f0 + std::min(nf, f + small));
    const size_t seg_bytes = segs.size() * sizeof(CombSegment);
    plan.q_off = (seg_bytes + 255) & ~(size_t)255;
    plan.valid = false;
    if (!ensure(&plan.dev, &plan.cap, plan.q_off + 256)) return 0;
    AMTK_CUDA(cudaMemcpyAsync(plan.dev, segs.data(), seg_bytes, cudaMemcpyHostToDevice, ctx->stream));
    AMTK_CUDA(cudaStreamSynchronize(ctx->stream));
    plan.nitems = (int)segs.size();
    plan.wY = wY; plan.hY = hY; plan.wC = wC; plan.hC = hC; plan.nf = nf; plan.f0 = f0; plan.R = 1000 * NS + kMmTH;
    plan.item = ctx->knobs.comb_item; plan.ctas = occ; plan.valid = true;
  }
  AMTK_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(plan.dev) + plan.q_off, 0, 256, ctx->stream));
  args.segs = reinterpret_cast<const CombSegment*>(plan.dev);
  args.nitems = plan.nitems;
  args.queue = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(plan.dev) + plan.q_off);
  args.counts = dcounts;
  args.out_frame0 = out_row0 - win.first;
  AMTK_CUDA(cudaMemsetAsync(dcounts + (size_t)(lo - out_row0) * 12, 0, (size_t)nf * 12 * sizeof(int), ctx->stream));
  std::pair<cudaEvent_t, cudaEvent_t> ev{ nullptr, nullptr };
  if (ctx->timing) {
    if (!ctx->timing_pool.empty()) { ev = ctx->timing_pool.back(); ctx->timing_pool.pop_back(); }
    else { AMTK_CUDA(cudaEventCreate(&ev.first)); AMTK_CUDA(cudaEventCreate(&ev.second)); }
    AMTK_CUDA(cudaEventRecord(ev.first, ctx->stream));
  }
  kern<<<grid, kMmThreads, smem, ctx->stream>>>(args);
  AMTK_CUDA(cudaGetLastError());
  if (ctx->timing) { AMTK_CUDA(cudaEventRecord(ev.second, ctx->stream)); ctx->timing_events.push_back(ev); }
  ctx->launches += 1;
  // experimental kernel: read its watchdog record back (costs a stream synchronisation per launch)
  int dbg[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
  AMTK_CUDA(cudaMemcpyAsync(dbg, args.queue + 16, sizeof(dbg), cudaMemcpyDeviceToHost, ctx->stream));
  AMTK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (dbg[0]) {
    char msg[256];
    snprintf(msg, sizeof(msg), "comb_mma<%d>: a device-side wait timed out (wait %d, step %d, CTA %d, thread %d, parity %d); results discarded",
             NS, dbg[1], dbg[2], dbg[3], dbg[4], dbg[5]);
    AMTK_FAIL(msg);
  }
  return 1;
}

static int launch_comb(amtk_ctx* ctx, const amtk_clip* clip, const Window& win, int lo, int hi,
                       const amtk_comb_params* prm, int* dcounts, int out_row0) {
  const bool tma_layout = !((clip->frame_stride & 15) || (clip->pitch_y & 15) || (clip->pitch_uv & 15) || (clip->off_u & 15) ||
                            (clip->off_v & 15) || (reinterpret_cast<uintptr_t>(win.dev_base) & 15));
  if (!tma_layout || !ctx->encode_tiled || ctx->knobs.comb_generic) {
    // generic kernel: any sample size / pitch (DESIGN.md 3.1 "fallback")
    CombGenericArgs g;
    g.base = win.dev_base; g.frame_stride = clip->frame_stride;
    g.off[0] = 0; g.off[1] = clip->off_u; g.off[2] = clip->off_v;
    const int bps = clip->bytes_per_sample;
    for (int pl = 0; pl < 3; ++pl) {
      g.pitch[pl] = (pl ? clip->pitch_uv : clip->pitch_y) / bps;
      g.W[pl] = pl ? (clip->width >> clip->log_uvx) : clip->width;
      g.H[pl] = pl ? (clip->height >> clip->log_uvy) : clip->height;
      g.thM[pl] = pl ? prm->th_move_c : prm->th_move_y;
      g.thS[pl] = pl ? prm->th_shima_c : prm->th_shima_y;
      g.thL[pl] = pl ? prm->th_lshima_c : prm->th_lshima_y;
    }
    g.first_frame = lo - win.first; g.prev_of_first = lo > 0 ? lo - 1 - win.first : lo - win.first;
    g.nframes = hi - lo; g.counts = dcounts + (size_t)(lo - out_row0) * 12;
    AMTK_CUDA(cudaMemsetAsync(g.counts, 0, (size_t)(hi - lo) * 12 * sizeof(int), ctx->stream));
    for (int f0 = 0; f0 < hi - lo; f0 += 16384) {            // gridDim.z limit
      CombGenericArgs gg = g;
      gg.first_frame = g.first_frame + f0; gg.prev_of_first = f0 ? gg.first_frame - 1 : g.prev_of_first;
      gg.nframes = std::min(16384, hi - lo - f0); gg.counts = g.counts + (size_t)f0 * 12;
      dim3 grid((g.W[0] + kGenTW - 1) / kGenTW, (g.H[0] + kGenTH - 1) / kGenTH, gg.nframes * 3);
      if (bps == 1) comb_generic_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(gg);
      else comb_generic_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(gg);
      AMTK_CUDA(cudaGetLastError());
      ctx->launches += 1;
    }
    return 1;
  }
  if (clip->bytes_per_sample == 1 && ctx->knobs.comb_mma) return launch_comb_mma(ctx, clip, win, lo, hi, prm, dcounts, out_row0);
  if (ctx->knobs.comb_ws && (clip->bytes_per_sample == 1 || (clip->bits_per_sample <= 10 && ctx->knobs.comb_ws10)))
    return launch_comb_ws(ctx, clip, win, lo, hi, prm, dcounts, out_row0);
  const int hY = clip->height, hC = clip->height >> clip->log_uvy;
  const int R = ctx->knobs.comb_R ? ctx->knobs.comb_R : pick_comb_R(hY, hC);
  int nvar = 0; const CombVariant* vars = comb_variants(&nvar); const CombVariant* V = nullptr;
  for (int i = 0; i < nvar; ++i) if (vars[i].R == R && vars[i].strip == ctx->knobs.comb_strip && vars[i].stages == ctx->knobs.comb_stages && vars[i].sync == ctx->knobs.comb_sync) V = &vars[i];
  if (!V) AMTK_FAIL("comb: no kernel variant for the requested AMTK_COMB_* settings");
  const int bps = clip->bytes_per_sample;
  if (bps == 2 && !V->kernel16) AMTK_FAIL("comb: the selected AMTK_COMB_* variant has no 16-bit kernel");
  const int twe = kCombTW / bps;                 // samples per tile row
  CombArgs args;
  memset(&args, 0, sizeof(args));
  // Chroma width 960 = 7.5 tiles: the 64-pixel remainders of U and V share ONE tile (two half-width TMA boxes)
  // instead of two half-empty ones.
  const int wC = clip->width >> clip->log_uvx;
  const int rem = wC % twe;
  const bool merge_uv = ctx->knobs.comb_merge_uv && rem > 0 && rem <= twe / 2 && (V->boxH * (kCombTW / 2)) % 128 == 0;
  int tile0 = 0;
  for (int pl = 0; pl < 4; ++pl) {
    CombPlane& P = args.plane[pl];
    const bool chroma = pl != 0;
    P.W = chroma ? wC : clip->width;
    P.H = chroma ? hC : hY;
    P.tilesX = (P.W + twe - 1) / twe; P.tilesY = (P.H + V->TH - 1) / V->TH;
    if (merge_uv && (pl == 1 || pl == 2)) P.tilesX -= 1;           // remainder column handled by the pseudo plane
    if (pl == 3) { P.tilesX = merge_uv ? 1 : 0; if (!merge_uv) P.tilesY = 0; }
    P.tile0 = tile0; tile0 += P.tilesX * P.tilesY;
    P.cls = chroma ? 1 : 0;
    const int thM = chroma ? prm->th_move_c : prm->th_move_y;
    const int thS = chroma ? prm->th_shima_c : prm->th_shima_y, thL = chroma ? prm->th_lshima_c : prm->th_lshima_y;
    if (bps == 1) {
      P.thM = (unsigned)(0x80 - thM) * 0x01010101u;
      P.thS = (unsigned)thS * 0x00010001u;       // integer k in [1,2047] IS the fp16 bit pattern of k*2^-24
      P.thL = (unsigned)thL * 0x00010001u;
    } else {
      P.thM = (unsigned)(0x8000 - thM) * 0x00010001u;
      const float fs = (float)thS, fl = (float)thL;
      memcpy(&P.thS, &fs, 4); memcpy(&P.thL, &fl, 4);
    }
    if (pl == 3) break;
    const long long off = pl == 0 ? 0 : (pl == 1 ? clip->off_u : clip->off_v);
    const int pitch = pl ? clip->pitch_uv : clip->pitch_y;
    cuuint64_t gdim[3] = { (cuuint64_t)P.W, (cuuint64_t)P.H, (cuuint64_t)win.count };
    cuuint64_t gstr[2] = { (cuuint64_t)pitch, (cuuint64_t)clip->frame_stride };
    cuuint32_t estr[3] = { 1u, 1u, 1u };
    const CUtensorMapL2promotion promo = ctx->knobs.comb_l2 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : ctx->knobs.comb_l2 == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                                         ctx->knobs.comb_l2 == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    for (int half = 0; half < (pl && merge_uv ? 2 : 1); ++half) {
      cuuint32_t box[3] = { (cuuint32_t)(half ? twe / 2 : twe), (cuuint32_t)V->boxH, 1u };
      CUtensorMap* m = half ? &args.map_half[pl - 1] : &args.map[pl];
      CUresult r = ctx->encode_tiled(m, bps == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint8_t*>(win.dev_base) + off, gdim, gstr, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) AMTK_FAIL("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    }
  }
  args.half_x = wC - rem;
  const int ntiles = tile0;
  const int nf = hi - lo;
  // ---- static partition of (tile, frame) pairs over the resident CTAs: tile-major, frame-minor, equal shares.
  // A tile-frame costs the same wherever it lies (the kernel is issue/latency bound per warp, and tile shapes are
  // chosen so that bands are full), so equal counts = equal time; each CTA touches at most ~2 tiles.
  int occ = 0;
  void (*kern)(const CombArgs) = bps == 1 ? V->kernel : V->kernel16;
  AMTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, V->smem));
  AMTK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, V->threads, V->smem));
  if (occ < 1) AMTK_FAIL("comb kernel does not fit on an SM");
  if (ctx->knobs.comb_ctas > 0) occ = std::min(occ, ctx->knobs.comb_ctas);
  const long long total = (long long)ntiles * nf;
  int grid = (int)std::min<long long>((long long)ctx->sm_count * occ, total);
  std::vector<CombSegment> segs;
  // "lock-step" partition: every tile is cut into the same C frame ranges, so neighbouring tiles stream the same
  // frames at the same time and share halo rows / straddled 128-byte lines through L2.  Used when the plane pitch
  // makes tile rows straddle lines (pitch % 128 != 0 on luma); costs a few idle CTA slots.
  const int chunks = std::max(1, std::min(nf, grid / std::max(1, ntiles)));
  // Since the tensor maps use 64-byte L2 promotion (straddled lines are no longer fetched whole) the equal-share
  // partition wins on every layout measured (tools/part_probe.py); lock-step stays available through the knob.
  const bool lockstep = ctx->knobs.comb_part == 1;
  if (lockstep) grid = chunks * ntiles;
  std::vector<int> seg_start((size_t)grid + 1, 0);
  if (lockstep) {
    for (int b = 0; b < grid; ++b) {                 // CTA b = (chunk c, tile t), tile-minor so that co-resident CTAs are neighbours
      const int c = b / ntiles, t = b % ntiles;
      const int f0 = (int)((long long)nf * c / chunks), f1 = (int)((long long)nf * (c + 1) / chunks);
      seg_start[b] = (int)segs.size();
      if (f1 > f0) segs.push_back(CombSegment{ t, lo - win.first + f0, lo - win.first + f1 });
    }
  } else
  for (int b = 0; b < grid; ++b) {
    const long long lo_i = total * b / grid, hi_i = total * (b + 1) / grid;     // [lo_i, hi_i) of the tile-major order
    seg_start[b] = (int)segs.size();
    long long i = lo_i;
    while (i < hi_i) {
      const int t = (int)(i / nf), f = (int)(i % nf);
      const int take = (int)std::min<long long>(nf - f, hi_i - i);
      segs.push_back(CombSegment{ t, lo - win.first + f, lo - win.first + f + take });
      i += take;
    }
  }
  seg_start[grid] = (int)segs.size();
  const size_t seg_bytes = segs.size() * sizeof(CombSegment), st_bytes = seg_start.size() * sizeof(int);
  const size_t st_off = (seg_bytes + 255) & ~(size_t)255;
  if (!ensure(&ctx->small, &ctx->small_bytes, st_off + st_bytes)) return 0;
  AMTK_CUDA(cudaMemcpyAsync(ctx->small, segs.data(), seg_bytes, cudaMemcpyHostToDevice, ctx->stream));
  AMTK_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ctx->small) + st_off, seg_start.data(), st_bytes, cudaMemcpyHostToDevice, ctx->stream));
  // pageable host vectors: the async copies above have completed their host reads on return
  args.segs = reinterpret_cast<const CombSegment*>(ctx->small);
  args.seg_start = reinterpret_cast<const int*>(reinterpret_cast<uint8_t*>(ctx->small) + st_off);
  args.counts = dcounts;
  args.out_frame0 = out_row0 - win.first;
  AMTK_CUDA(cudaMemsetAsync(dcounts + (size_t)(lo - out_row0) * 12, 0, (size_t)nf * 12 * sizeof(int), ctx->stream));
  std::pair<cudaEvent_t, cudaEvent_t> ev{ nullptr, nullptr };
  if (ctx->timing) {
    if (!ctx->timing_pool.empty()) { ev = ctx->timing_pool.back(); ctx->timing_pool.pop_back(); }
    else { AMTK_CUDA(cudaEventCreate(&ev.first)); AMTK_CUDA(cudaEventCreate(&ev.second)); }
    AMTK_CUDA(cudaEventRecord(ev.first, ctx->stream));
  }
  kern<<<grid, V->threads, V->smem, ctx->stream>>>(args);
  AMTK_CUDA(cudaGetLastError());
  if (ctx->timing) { AMTK_CUDA(cudaEventRecord(ev.second, ctx->stream)); ctx->timing_events.push_back(ev); }
  ctx->launches += 1;
  return 1;
}

static std::once_flag g_driver_once;
static amtk_encode_tiled_fn g_encode = nullptr;

}  // namespace amtk

using namespace amtk;

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" {

const char* amtk_last_error(void) { return g_error.c_str(); }
int amtk_version(void) { return 100; }

int amtk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int amtk_ctx_create(int device, void* cuda_stream, amtk_ctx** out) {
  if (!out) AMTK_FAIL("amtk_ctx_create: out is null");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); AMTK_FAIL("no CUDA device: this library has no CPU fallback"); }
  if (device < 0 || device >= n) AMTK_FAIL("amtk_ctx_create: bad device ordinal");
  int prev = 0; AMTK_CUDA(cudaGetDevice(&prev));
  AMTK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop; AMTK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { cudaSetDevice(prev); AMTK_FAIL("this build targets sm_100a (Blackwell B200) only"); }
  amtk_ctx* c = new amtk_ctx();
  c->device = device; c->sm_count = prop.multiProcessorCount;
  // NULL selects the legacy default stream (which orders with every blocking stream, e.g. torch's default one)
  c->stream = reinterpret_cast<cudaStream_t>(cuda_stream); c->own_stream = false;
  bool ok = cuda_ok(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate(copy)") &&
            cuda_ok(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking), "cudaStreamCreate(side)") &&
            cuda_ok(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming), "cudaEventCreate") &&
            cuda_ok(cudaEventCreateWithFlags(&c->ev_side_done, cudaEventDisableTiming), "cudaEventCreate") &&
            cuda_ok(cudaStreamCreateWithFlags(&c->side_stream2, cudaStreamNonBlocking), "cudaStreamCreate(side2)") &&
            cuda_ok(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming), "cudaEventCreate") &&
            cuda_ok(cudaEventCreateWithFlags(&c->ev_join1, cudaEventDisableTiming), "cudaEventCreate") &&
            cuda_ok(cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming), "cudaEventCreate");
  for (int b = 0; b < 2 && ok; ++b) {
    ok = cuda_ok(cudaEventCreateWithFlags(&c->ev_copy[b], cudaEventDisableTiming), "cudaEventCreate") &&
         cuda_ok(cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming), "cudaEventCreate");
  }
  std::call_once(g_driver_once, [] {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<amtk_encode_tiled_fn>(fn);
    else cudaGetLastError();
  });
  c->encode_tiled = g_encode;
  // per-context tuning knobs (defaults are the measured best; env AMTK_* overrides them for tools/tune_comb.py)
  if (const char* e = getenv("AMTK_COMB_STRIP")) c->knobs.comb_strip = atoi(e);
  if (const char* e = getenv("AMTK_COMB_STAGES")) c->knobs.comb_stages = atoi(e);
  if (const char* e = getenv("AMTK_COMB_R")) c->knobs.comb_R = atoi(e);
  if (const char* e = getenv("AMTK_COMB_CTAS")) c->knobs.comb_ctas = atoi(e);
  if (const char* e = getenv("AMTK_COMB_SYNC")) c->knobs.comb_sync = atoi(e);
  if (const char* e = getenv("AMTK_COMB_GENERIC")) c->knobs.comb_generic = atoi(e);
  if (const char* e = getenv("AMTK_COMB_MERGE_UV")) c->knobs.comb_merge_uv = atoi(e);
  if (const char* e = getenv("AMTK_COMB_PART")) c->knobs.comb_part = atoi(e);
  if (const char* e = getenv("AMTK_EVAL_WAVES")) c->knobs.eval_waves = std::max(1, atoi(e));
  if (const char* e = getenv("AMTK_COMB_L2")) c->knobs.comb_l2 = atoi(e);
  if (const char* e = getenv("AMTK_COMB_WS")) c->knobs.comb_ws = atoi(e);
  if (const char* e = getenv("AMTK_SCAN_LITE")) c->knobs.scan_lite = atoi(e);
  if (const char* e = getenv("AMTK_LITE_CTAS")) c->knobs.lite_ctas = std::max(1, atoi(e));
  if (const char* e = getenv("AMTK_COMB_WS_STAGES")) c->knobs.comb_ws_stages = atoi(e);
  if (const char* e = getenv("AMTK_COMB_ITEM")) c->knobs.comb_item = atoi(e);
  if (const char* e = getenv("AMTK_COMB_TAIL")) c->knobs.comb_tail = atoi(e);
  if (const char* e = getenv("AMTK_COMB_MMA")) c->knobs.comb_mma = atoi(e);
  if (const char* e = getenv("AMTK_COMB_WS10")) c->knobs.comb_ws10 = atoi(e);
  if (const char* e = getenv("AMTK_EVAL_CW")) c->knobs.eval_cw = atoi(e);
  if (const char* e = getenv("AMTK_EVAL_PAR")) c->knobs.eval_par = atoi(e);
  if (const char* e = getenv("AMTK_SCAN_OVERLAP")) c->knobs.scan_overlap = atoi(e);
  if (const char* e = getenv("AMTK_COMB_WS_WARPS")) c->knobs.comb_ws_warps = atoi(e);
  if (const char* e = getenv("AMTK_COMB_WS_PF")) c->knobs.comb_ws_prefetch = atoi(e);
  cudaSetDevice(prev);
  if (!ok) { amtk_ctx_destroy(c); return 0; }
  *out = c;
  return 1;
}

void amtk_ctx_destroy(amtk_ctx* c) {
  if (!c) return;
  int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->side_stream) { cudaStreamSynchronize(c->side_stream); cudaStreamDestroy(c->side_stream); }
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  if (c->ev_side_done) cudaEventDestroy(c->ev_side_done);
  if (c->side_stream2) { cudaStreamSynchronize(c->side_stream2); cudaStreamDestroy(c->side_stream2); }
  for (cudaEvent_t e : { c->ev_fork, c->ev_join1, c->ev_join2 }) if (e) cudaEventDestroy(e);
  for (int b = 0; b < 2; ++b) { if (c->ev_copy[b]) cudaEventDestroy(c->ev_copy[b]); if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]); if (c->stage[b]) cudaFree(c->stage[b]); }
  if (c->scratch) cudaFree(c->scratch);
  if (c->small) cudaFree(c->small);
  if (c->dout) cudaFree(c->dout);
  if (c->dout2) cudaFree(c->dout2);
  if (c->hout) cudaFreeHost(c->hout);
  if (c->plan.dev) cudaFree(c->plan.dev);
  for (auto* v : { &c->timing_events, &c->timing_pool }) for (auto& ev : *v) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  cudaSetDevice(prev);
  delete c;
}

int amtk_ctx_synchronize(amtk_ctx* c) {
  if (!c) AMTK_FAIL("null context");
  DevSelect ds(c); if (!ds.ok) return 0;
  AMTK_CUDA(cudaStreamSynchronize(c->stream));
  return 1;
}
int64_t amtk_ctx_launch_count(const amtk_ctx* c) { return c ? c->launches : 0; }
int64_t amtk_ctx_last_h2d_bytes(const amtk_ctx* c) { return c ? c->h2d_bytes_last : 0; }

int amtk_ctx_set_kernel_timing(amtk_ctx* c, int enable) {
  if (!c) AMTK_FAIL("null context");
  c->timing = enable != 0;
  return 1;
}

int amtk_ctx_get_kernel_timing(amtk_ctx* c, double* ms_total, int64_t* launches, int reset) {
  if (!c) AMTK_FAIL("null context");
  DevSelect ds(c); if (!ds.ok) return 0;
  AMTK_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& ev : c->timing_events) {
    float ms = 0.0f;
    AMTK_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
    c->timing_ms += ms; c->timing_count += 1;
    c->timing_pool.push_back(ev);
  }
  c->timing_events.clear();
  if (ms_total) *ms_total = c->timing_ms;
  if (launches) *launches = c->timing_count;
  if (reset) { c->timing_ms = 0.0; c->timing_count = 0; }
  return 1;
}

int amtk_probe_read_ms(amtk_ctx* c, const void* ptr, size_t bytes, int reps, double* ms_out) {
  if (!c || !ptr || !ms_out || reps < 1) AMTK_FAIL("amtk_probe_read_ms: bad argument");
  if (reinterpret_cast<uintptr_t>(ptr) & 15) AMTK_FAIL("amtk_probe_read_ms: pointer must be 16-byte aligned");
  DevSelect ds(c); if (!ds.ok) return 0;
  if (!ensure(&c->small, &c->small_bytes, 256)) return 0;
  cudaEvent_t e0, e1;
  AMTK_CUDA(cudaEventCreate(&e0)); AMTK_CUDA(cudaEventCreate(&e1));
  const int grid = c->sm_count * 8;
  unsigned* sink = reinterpret_cast<unsigned*>(c->small);
  read_probe_kernel<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(ptr), bytes / 16, sink);
  AMTK_CUDA(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < reps; ++i) read_probe_kernel<<<grid, 256, 0, c->stream>>>(reinterpret_cast<const uint4*>(ptr), bytes / 16, sink);
  AMTK_CUDA(cudaEventRecord(e1, c->stream));
  AMTK_CUDA(cudaEventSynchronize(e1));
  float ms = 0; AMTK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  AMTK_CUDA(cudaGetLastError());
  *ms_out = ms / reps;
  return 1;
}

int amtk_host_alloc(size_t bytes, void** out) {
  if (!out) AMTK_FAIL("amtk_host_alloc: out is null");
  AMTK_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return 1;
}
void amtk_host_free(void* p) { if (p) cudaFreeHost(p); }

int amtk_device_alloc(amtk_ctx* c, size_t bytes, void** out) {
  if (!c || !out) AMTK_FAIL("amtk_device_alloc: null argument");
  DevSelect ds(c); if (!ds.ok) return 0;
  AMTK_CUDA(cudaMalloc(out, bytes));
  return 1;
}
void amtk_device_free(amtk_ctx* c, void* p) {
  if (!c || !p) return;
  DevSelect ds(c);
  cudaStreamSynchronize(c->stream);
  cudaFree(p);
}
int amtk_memcpy_h2d(amtk_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c || !dst || !src) AMTK_FAIL("amtk_memcpy_h2d: null argument");
  DevSelect ds(c); if (!ds.ok) return 0;
  AMTK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  AMTK_CUDA(cudaStreamSynchronize(c->stream));
  return 1;
}
int amtk_memcpy_d2d(amtk_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c || !dst || !src) AMTK_FAIL("amtk_memcpy_d2d: null argument");
  DevSelect ds(c); if (!ds.ok) return 0;
  AMTK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  AMTK_CUDA(cudaStreamSynchronize(c->stream));
  return 1;
}
int amtk_memcpy_d2h(amtk_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c || !dst || !src) AMTK_FAIL("amtk_memcpy_d2h: null argument");
  DevSelect ds(c); if (!ds.ok) return 0;
  AMTK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  AMTK_CUDA(cudaStreamSynchronize(c->stream));
  return 1;
}

// ---------------------------------------------------------------------------------------------------------
// logos
// ---------------------------------------------------------------------------------------------------------
static void logo_free_device(amtk_logo* l) {
  float* fp[] = { l->dA, l->dB, l->dAU, l->dBU, l->dAV, l->dBV, l->dTapsT };
  for (float* p : fp) if (p) cudaFree(p);
  if (l->dPix) cudaFree(l->dPix);
  if (l->dScales) cudaFree(l->dScales);
  l->dA = l->dB = l->dAU = l->dBU = l->dAV = l->dBV = l->dTapsT = nullptr; l->dPix = nullptr; l->dScales = nullptr;
}

static int logo_upload_planes(amtk_logo* l) {
  amtk::HostLogo& h = l->host;
  const size_t ny = h.ySize() * sizeof(float), nc = h.cSize() * sizeof(float);
  float** dst[6] = { &l->dA, &l->dB, &l->dAU, &l->dBU, &l->dAV, &l->dBV };
  const float* src[6] = { h.aY(), h.bY(), h.aU(), h.bU(), h.aV(), h.bV() };
  for (int i = 0; i < 6; ++i) {
    const size_t n = i < 2 ? ny : nc;
    AMTK_CUDA(cudaMalloc(dst[i], std::max<size_t>(n, 16)));
    AMTK_CUDA(cudaMemcpy(*dst[i], src[i], n, cudaMemcpyHostToDevice));
  }
  return 1;
}

static int logo_upload_tables(amtk_logo* l) {
  amtk::HostLogo& h = l->host;
  if (l->dPix) { cudaFree(l->dPix); l->dPix = nullptr; }
  if (l->dTapsT) { cudaFree(l->dTapsT); l->dTapsT = nullptr; }
  if (l->dScales) { cudaFree(l->dScales); l->dScales = nullptr; }
  const int count = h.count();
  if (count > 0) {
    std::vector<float> tapsT((size_t)25 * l->countPad, 0.0f);
    for (int c = 0; c < count; ++c) for (int t = 0; t < 25; ++t) tapsT[(size_t)t * l->countPad + c] = h.kernels[(size_t)c * 25 + t];
    AMTK_CUDA(cudaMalloc(&l->dPix, (size_t)count * sizeof(uint32_t)));
    AMTK_CUDA(cudaMalloc(&l->dTapsT, tapsT.size() * sizeof(float)));
    AMTK_CUDA(cudaMalloc(&l->dScales, (size_t)count * 32 * sizeof(float2)));
    AMTK_CUDA(cudaMemcpy(l->dPix, h.pix.data(), (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice));
    AMTK_CUDA(cudaMemcpy(l->dTapsT, tapsT.data(), tapsT.size() * sizeof(float), cudaMemcpyHostToDevice));
    AMTK_CUDA(cudaMemcpy(l->dScales, h.scales.data(), (size_t)count * 32 * sizeof(float2), cudaMemcpyHostToDevice));
  }
  l->tables_uploaded = true;
  return 1;
}

// Logos are host objects; their HBM copies are made on first use by a context (and stay on that device).
static int logo_ensure_device(const amtk_logo* cl, amtk_ctx* ctx, bool need_tables) {
  amtk_logo* l = const_cast<amtk_logo*>(cl);
  std::lock_guard<std::mutex> lock(l->mu);
  if (l->device >= 0 && l->dA && l->device != ctx->device) AMTK_FAIL("logo already resident on another device");
  l->device = ctx->device;
  if (!l->dA && !logo_upload_planes(l)) return 0;
  if (need_tables) {
    if (!l->has_mask) AMTK_FAIL("logo has no mask: call amtk_logo_create_mask first");
    if (!l->tables_uploaded && !logo_upload_tables(l)) return 0;
  }
  return 1;
}

static int logo_adopt(amtk_ctx* /*ctx: logos bind to a device on first use*/, amtk::HostLogo&& h, amtk_logo** out) {
  amtk_logo* l = new amtk_logo();
  l->host = std::move(h);
  *out = l;
  return 1;
}

int amtk_logo_create(amtk_ctx* ctx, const float* data, int w, int h, int lx, int ly, int imgw, int imgh, int imgx, int imgy, amtk_logo** out) {
  if (!data || !out) AMTK_FAIL("amtk_logo_create: null argument");   // ctx may be NULL: bound on first use
  if (w < 5 || h < 5 || w > 4096 || h > 4096 || lx < 0 || lx > 2 || ly < 0 || ly > 2) AMTK_FAIL("amtk_logo_create: bad logo geometry");
  amtk::HostLogo hl; hl.init(w, h, lx, ly, imgw, imgh, imgx, imgy);
  memcpy(hl.data.data(), data, hl.dataSize() * sizeof(float));
  return logo_adopt(ctx, std::move(hl), out);
}

int amtk_logo_load(amtk_ctx* ctx, const char* path, amtk_logo** out, void* header540) {
  if (!path || !out) AMTK_FAIL("amtk_logo_load: null argument");
  amtk::HostLogo hl; amtk::LgdHeader hdr; std::string err;
  if (!amtk::lgd_load(path, hl, &hdr, err)) AMTK_FAIL(err);
  if (header540) memcpy(header540, &hdr, sizeof(hdr));
  return logo_adopt(ctx, std::move(hl), out);
}

int amtk_logo_save(const amtk_logo* l, const char* path, const char* name, int service_id) {
  if (!l || !path) AMTK_FAIL("amtk_logo_save: null argument");
  std::string err;
  if (!amtk::lgd_save(l->host, path, name ? name : "No Name", service_id, err)) AMTK_FAIL(err);
  return 1;
}

void amtk_logo_destroy(amtk_logo* l) {
  if (!l) return;
  if (l->device >= 0 && l->dA) { DevSelect ds(l->device); logo_free_device(l); }
  delete l;
}

int amtk_logo_deint(const amtk_logo* src, amtk_logo** out) {
  if (!src || !out) AMTK_FAIL("amtk_logo_deint: null argument");
  amtk::HostLogo d; amtk::logo_deint(src->host, d);
  return logo_adopt(nullptr, std::move(d), out);
}

int amtk_logo_field(const amtk_logo* src, int bottom, amtk_logo** out) {
  if (!src || !out) AMTK_FAIL("amtk_logo_field: null argument");
  if (src->host.h / 2 < 5) AMTK_FAIL("amtk_logo_field: logo too small");
  amtk::HostLogo f; amtk::logo_field(src->host, bottom != 0, f);
  return logo_adopt(nullptr, std::move(f), out);
}

int amtk_logo_create_mask(amtk_logo* l, float maskratio) {
  if (!l) AMTK_FAIL("amtk_logo_create_mask: null logo");
  if (!(maskratio > 0.0f) || maskratio > 1.0f) AMTK_FAIL("amtk_logo_create_mask: maskratio must be in (0,1]");
  std::lock_guard<std::mutex> lock(l->mu);
  amtk::logo_create_mask(l->host, maskratio);
  l->countPad = std::max(32, (l->host.count() + 31) & ~31);
  l->has_mask = true;
  l->tables_uploaded = false;       // (re)uploaded by the next evaluation call
  return 1;
}

int amtk_logo_get_info(const amtk_logo* l, amtk_logo_info* o) {
  if (!l || !o) AMTK_FAIL("amtk_logo_get_info: null argument");
  const amtk::HostLogo& h = l->host;
  o->w = h.w; o->h = h.h; o->log_uvx = h.logUVx; o->log_uvy = h.logUVy;
  o->imgw = h.imgw; o->imgh = h.imgh; o->imgx = h.imgx; o->imgy = h.imgy;
  o->maskpixels = h.maskpixels; o->count = h.count(); o->black_score = h.blackScore;
  return 1;
}

int amtk_logo_get_tables(const amtk_logo* l, float* data, uint8_t* mask, float* kernels, float* scales) {
  if (!l) AMTK_FAIL("amtk_logo_get_tables: null logo");
  const amtk::HostLogo& h = l->host;
  if (data) memcpy(data, h.data.data(), h.dataSize() * sizeof(float));
  if ((mask || kernels || scales) && !l->has_mask) AMTK_FAIL("logo has no mask");
  if (mask) memcpy(mask, h.mask.data(), h.mask.size());
  if (kernels) memcpy(kernels, h.kernels.data(), h.kernels.size() * sizeof(float));
  if (scales) memcpy(scales, h.scales.data(), h.scales.size() * sizeof(float));
  return 1;
}

// ---------------------------------------------------------------------------------------------------------
// evaluation entry points
// ---------------------------------------------------------------------------------------------------------
// Small host outputs (a GetFrame-sized call returns 8 .. 1056 bytes): the result kernels write straight into a pinned, device-mapped
// buffer of the context, so the call ends with a stream synchronise and a CPU copy instead of a D2H copy operation and ITS
// completion (a third of a one-frame call's wall time).  Only for outputs that kernels write once (logo scores), never for the
// atomically accumulated combing counters.
constexpr size_t kHostOutBytes = 64 << 10;
static float* host_out_alias(amtk_ctx* ctx, size_t bytes) {
  if (bytes > kHostOutBytes) return nullptr;
  if (!ctx->hout) {
    if (cudaHostAlloc(&ctx->hout, kHostOutBytes, cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); ctx->hout = nullptr; return nullptr; }
    if (cudaHostGetDevicePointer(&ctx->hout_dev, ctx->hout, 0) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(ctx->hout); ctx->hout = nullptr; return nullptr; }
  }
  return reinterpret_cast<float*>(ctx->hout_dev);
}
// device buffer the result kernels of a host-output call write to: the mapped alias when the output is small, else ctx->dout
static float* host_out_buffer(amtk_ctx* ctx, size_t bytes) {
  if (float* a = host_out_alias(ctx, bytes)) return a;
  if (!ensure(&ctx->dout, &ctx->dout_bytes, bytes)) return nullptr;
  return reinterpret_cast<float*>(ctx->dout);
}

static int finish_output(amtk_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes, int out_on_device) {
  if (out_on_device) return 1;
  if (ctx->hout && dev_src == ctx->hout_dev) {
    AMTK_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(host_dst, ctx->hout, bytes);
    return 1;
  }
  AMTK_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  AMTK_CUDA(cudaStreamSynchronize(ctx->stream));
  return 1;
}

// `real` is the caller's clip (frame size checks), `clip` the resident data (the same, or an ROI-only staging copy whose
// luma origin sits at (dx, dy) of the real frame).
static int scan_frames_impl(amtk_ctx* ctx, const amtk_clip* real, const amtk_clip* clip, int dx, int dy, amtk_logo* const* logos, int nlogos,
                            const Window& win, int lo, int hi, int pitch_override, float* dscores, int row0) {
  static const float kFades01[2] = { 0.0f, 1.0f };
  const int pitch = pitch_override > 0 ? pitch_override : clip->pitch_y / clip->bytes_per_sample;
  const int real_pitch = pitch_override > 0 ? pitch_override : real->pitch_y / real->bytes_per_sample;
  for (int i = 0; i < nlogos; ++i) {
    const amtk_logo* lg = logos[i];
    float* o = dscores + (size_t)(lo - row0) * nlogos * 2;
    if (!lg || lg->host.imgw != real->width || lg->host.imgh != real->height) {      // LogoScan.hpp:1551-1558
      fill_pairs_kernel<<<(hi - lo + 127) / 128, 128, 0, ctx->stream>>>(o, hi - lo, nlogos * 2, i * 2, 0.0f, -1.0f);
      AMTK_CUDA(cudaGetLastError()); ctx->launches += 1;
      continue;
    }
    if (!roi_inside(lg->host, real, real_pitch)) AMTK_FAIL("logo rectangle lies outside the frame");
    EvalSpec sp{ lg, lg->host.imgx - dx, lg->host.imgy - dy, lg->host.w, lg->host.h, 0, 0, lg->host.w, 2, kFades01, 0, i * 2, 1 };
    if (!launch_eval(ctx, clip, win, lo, hi, pitch, sp, dscores, nlogos * 2, row0)) return 0;
  }
  return 1;
}

// bounding box of the rectangles of all logos that will be evaluated on `clip` (false when there is none)
static bool logos_bbox(const amtk_clip* clip, amtk_logo* const* logos, int nlogos, int* rx, int* ry, int* rw, int* rh) {
  int x0 = 1 << 30, y0 = 1 << 30, x1 = -1, y1 = -1;
  for (int i = 0; i < nlogos; ++i) {
    const amtk_logo* lg = logos[i];
    if (!lg || lg->host.imgw != clip->width || lg->host.imgh != clip->height) continue;
    x0 = std::min(x0, lg->host.imgx); y0 = std::min(y0, lg->host.imgy);
    x1 = std::max(x1, lg->host.imgx + lg->host.w); y1 = std::max(y1, lg->host.imgy + lg->host.h);
  }
  if (x1 < 0 || x0 < 0 || y0 < 0 || x1 > clip->width || y1 > clip->height) return false;
  *rx = x0; *ry = y0; *rw = x1 - x0; *rh = y1 - y0;
  return true;
}

int amtk_logo_scan_frames(amtk_ctx* ctx, const amtk_clip* clip, amtk_logo* const* logos, int nlogos,
                          int frame0, int nframes, int pitch_override, float* out, int out_on_device) {
  if (ctx && nframes == 0) return 1;                                 // empty range: nothing to do
  if (!ctx || !logos || !out || nlogos < 1) AMTK_FAIL("amtk_logo_scan_frames: bad argument");
  if (!validate_clip(clip, false)) return 0;
  DevSelect ds(ctx); if (!ds.ok) return 0;
  const size_t bytes = (size_t)nframes * nlogos * 2 * sizeof(float);
  float* d = out;
  if (!out_on_device) { d = host_out_buffer(ctx, bytes); if (!d) return 0; }
  int rx, ry, rw, rh;
  if (!clip->on_device && pitch_override <= 0 && logos_bbox(clip, logos, nlogos, &rx, &ry, &rw, &rh)) {
    // host frames: only the logo rectangles cross PCIe (the reference reads nothing else, LogoScan.hpp:1559-1566)
    if (!for_each_roi_window(ctx, clip, frame0, nframes, rx, ry, rw, rh, false, false,
                             [&](const amtk_clip& v, const Window& w, int lo, int hi, int dx, int dy) {
          return scan_frames_impl(ctx, clip, &v, dx, dy, logos, nlogos, w, lo, hi, 0, d, frame0); }))
      return 0;
  } else if (!for_each_window(ctx, clip, frame0, nframes, false, [&](const Window& w, int lo, int hi) {
        return scan_frames_impl(ctx, clip, clip, 0, 0, logos, nlogos, w, lo, hi, pitch_override, d, frame0); }))
    return 0;
  return finish_output(ctx, out, d, bytes, out_on_device);
}

static int analyze_impl(amtk_ctx* ctx, const amtk_clip* clip, int dx, int dy, const amtk_logo* dl, const amtk_logo* ft, const amtk_logo* fb,
                        const Window& win, int lo, int hi, float* dout, int row0) {
  float fades[11];
  for (int f = 0; f <= 10; ++f) fades[f] = (float)f / 10.0f;             // LogoScan.hpp:1152
  const int pitch = clip->pitch_y / clip->bytes_per_sample;
  const int w = dl->host.w, h = dl->host.h;
  const int rx = dl->host.imgx - dx, ry = dl->host.imgy - dy;
  EvalSpec sp{ dl, rx, ry, w, h, 0, 0, w, 11, fades, 1, 0, 1 };                // p[f]: deint logo on DeintY
  EvalSpec st{ ft, rx, ry, w, h, 1, 0, 2 * w, 11, fades, 1, 11, 1 };           // t[f]: top field logo on CopyY, stride 2w
  EvalSpec sb{ fb, rx, ry, w, h, 1, w, 2 * w, 11, fades, 1, 22, 1 };           // b[f]: bottom field logo on CopyY + w
  const int n = hi - lo;
  if (!(ctx->knobs.eval_par && n <= 16 && ctx->side_stream && ctx->side_stream2))
    return launch_eval(ctx, clip, win, lo, hi, pitch, sp, dout, 33, row0) &&
           launch_eval(ctx, clip, win, lo, hi, pitch, st, dout, 33, row0) &&
           launch_eval(ctx, clip, win, lo, hi, pitch, sb, dout, 33, row0);
  // GetFrame-sized call (AMTAnalyzeLogo::GetFrame = 8 source frames): each evaluation launches only n CTAs, so the three of them
  // run side by side on three streams, each with its own slice of the score scratch (115 -> ~70 us per call).  The context is
  // locked for the whole entry point (DevSelect), so swapping ctx->stream around a launch is invisible to other threads.
  const EvalSpec* specs[3] = { &sp, &st, &sb };
  size_t off[3], total = 0;
  for (int i = 0; i < 3; ++i) {
    if (!logo_ensure_device(specs[i]->logo, ctx, true)) return 0;                      // table uploads happen before the fork
    off[i] = total;
    total += (((size_t)n * 11 * specs[i]->logo->countPad * sizeof(float)) + 255) & ~(size_t)255;
  }
  if (!ensure(&ctx->scratch, &ctx->scratch_bytes, total)) return 0;                    // no reallocation once work is in flight
  cudaStream_t main_stream = ctx->stream, streams[3] = { ctx->stream, ctx->side_stream, ctx->side_stream2 };
  cudaEvent_t joins[3] = { nullptr, ctx->ev_join1, ctx->ev_join2 };
  AMTK_CUDA(cudaEventRecord(ctx->ev_fork, main_stream));
  int ok = 1;
  for (int i = 0; i < 3 && ok; ++i) {
    if (i) ok = cuda_ok(cudaStreamWaitEvent(streams[i], ctx->ev_fork, 0), "cudaStreamWaitEvent");
    ctx->stream = streams[i]; ctx->scratch_off = off[i];
    ok = ok && launch_eval(ctx, clip, win, lo, hi, pitch, *specs[i], dout, 33, row0);
    ctx->stream = main_stream; ctx->scratch_off = 0;
    if (i && ok) ok = cuda_ok(cudaEventRecord(joins[i], streams[i]), "cudaEventRecord") &&
                      cuda_ok(cudaStreamWaitEvent(main_stream, joins[i], 0), "cudaStreamWaitEvent");
  }
  if (!ok) { cudaStreamSynchronize(ctx->side_stream); cudaStreamSynchronize(ctx->side_stream2); }   // nothing of this call stays in flight
  return ok;
}

int amtk_logo_analyze_frames(amtk_ctx* ctx, const amtk_clip* clip, const amtk_logo* dl, const amtk_logo* ft, const amtk_logo* fb,
                             int frame0, int nframes, float* out, int out_on_device) {
  if (ctx && nframes == 0) return 1;
  if (!ctx || !dl || !ft || !fb || !out) AMTK_FAIL("amtk_logo_analyze_frames: bad argument");
  if (!validate_clip(clip, false)) return 0;
  if (ft->host.w != dl->host.w || fb->host.w != dl->host.w || ft->host.h != dl->host.h / 2 || fb->host.h != dl->host.h / 2)
    AMTK_FAIL("field logos do not match the deint logo");
  DevSelect ds(ctx); if (!ds.ok) return 0;
  const size_t bytes = (size_t)nframes * 33 * sizeof(float);
  float* d = out;
  if (!out_on_device) { d = host_out_buffer(ctx, bytes); if (!d) return 0; }
  if (!roi_inside(dl->host, clip, clip->pitch_y / clip->bytes_per_sample)) AMTK_FAIL("logo rectangle lies outside the frame");
  if (!for_each_roi_window(ctx, clip, frame0, nframes, dl->host.imgx, dl->host.imgy, dl->host.w, dl->host.h, false, false,
                           [&](const amtk_clip& v, const Window& w, int lo, int hi, int dx, int dy) {
        return analyze_impl(ctx, &v, dx, dy, dl, ft, fb, w, lo, hi, d, frame0); }))
    return 0;
  return finish_output(ctx, out, d, bytes, out_on_device);
}

int amtk_logo_eval_fades(amtk_ctx* ctx, const amtk_clip* clip, const amtk_logo* dl, const float* fades, int nfades,
                         int frame0, int nframes, float* out, int out_on_device) {
  if (ctx && nframes == 0) return 1;
  if (!ctx || !dl || !fades || !out) AMTK_FAIL("amtk_logo_eval_fades: bad argument");
  if (nfades < 1 || nfades > kMaxFades) AMTK_FAIL("amtk_logo_eval_fades: 1..24 fade levels");
  if (!validate_clip(clip, false)) return 0;
  DevSelect ds(ctx); if (!ds.ok) return 0;
  const size_t bytes = (size_t)nframes * nfades * sizeof(float);
  float* d = out;
  if (!out_on_device) { d = host_out_buffer(ctx, bytes); if (!d) return 0; }
  const int pitch = clip->pitch_y / clip->bytes_per_sample;
  if (!roi_inside(dl->host, clip, pitch)) AMTK_FAIL("logo rectangle lies outside the frame");
  if (!for_each_roi_window(ctx, clip, frame0, nframes, dl->host.imgx, dl->host.imgy, dl->host.w, dl->host.h, false, false,
                           [&](const amtk_clip& v, const Window& w, int lo, int hi, int dx, int dy) {
        EvalSpec sp{ dl, dl->host.imgx - dx, dl->host.imgy - dy, dl->host.w, dl->host.h, 0, 0, dl->host.w, nfades, fades, 0, 0, 1 };
        return launch_eval(ctx, &v, w, lo, hi, v.pitch_y / v.bytes_per_sample, sp, d, nfades, frame0); }))
    return 0;
  return finish_output(ctx, out, d, bytes, out_on_device);
}

// ---------------------------------------------------------------------------------------------------------
// combing metric + fused step
// ---------------------------------------------------------------------------------------------------------
void amtk_comb_default_params(amtk_comb_params* p) {
  if (!p) return;
  p->th_move_y = 20; p->th_shima_y = 12; p->th_lshima_y = 36;
  p->th_move_c = 24; p->th_shima_c = 16; p->th_lshima_c = 48;
}

int amtk_comb_frames(amtk_ctx* ctx, const amtk_clip* clip, const amtk_comb_params* prm, int frame0, int nframes,
                     int32_t* counts, int out_on_device) {
  if (ctx && nframes == 0) return 1;
  if (!ctx || !prm || !counts) AMTK_FAIL("amtk_comb_frames: bad argument");
  if (!validate_clip(clip, true) || !comb_thresholds_ok(prm, clip->bytes_per_sample)) return 0;
  DevSelect ds(ctx); if (!ds.ok) return 0;
  const size_t bytes = (size_t)nframes * 12 * sizeof(int32_t);
  int* d = counts;
  if (!out_on_device) { if (!ensure(&ctx->dout2, &ctx->dout2_bytes, bytes)) return 0; d = reinterpret_cast<int*>(ctx->dout2); }
  if (!for_each_window(ctx, clip, frame0, nframes, true, [&](const Window& w, int lo, int hi) {
        return launch_comb(ctx, clip, w, lo, hi, prm, d, frame0); }))
    return 0;
  return finish_output(ctx, counts, d, bytes, out_on_device);
}

// ScanFrame scores of ONE logo through the co-resident kernel, on the context's side stream: enqueued BEFORE the comb
// kernel so that each SM takes one of its CTAs and fills up with comb CTAs; joins the main stream at the end.
static int launch_scan_lite(amtk_ctx* ctx, const amtk_clip* clip, const Window& win, int lo, int hi, const amtk_logo* lg,
                            float* dscores, int nlogos, int logo_index, int row0, bool side) {
  cudaStream_t st = side ? ctx->side_stream : ctx->stream;
  if (!logo_ensure_device(lg, ctx, true)) return 0;
  const amtk::HostLogo& hl = lg->host;
  const int count = hl.count(), countPad = lg->countPad, n = hi - lo;
  if (!ensure(&ctx->scratch, &ctx->scratch_bytes, (size_t)n * 2 * countPad * sizeof(float))) return 0;
  LiteJob job;
  job.ybase = win.dev_base; job.frame_stride = clip->frame_stride; job.pitch = clip->pitch_y / clip->bytes_per_sample;
  job.frame0 = lo - win.first; job.nframes = n; job.imgx = hl.imgx; job.imgy = hl.imgy;
  job.logo = logo_dev(lg); job.maxv = (float)((1 << clip->bits_per_sample) - 1);
  job.nfades = 2; job.fades[0] = 0.0f; job.fades[1] = 1.0f;
  job.scores = reinterpret_cast<float*>(ctx->scratch);
  const size_t smem = logo_lite_smem_bytes(hl.w, hl.h, clip->bytes_per_sample);
  if (side) AMTK_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side, 0));      // ev_side: recorded by the caller on the main stream
  // under the comb kernel: one CTA per SM (all that fits); on its own: as many as the SMs hold, a few frames each
  const int grid = side ? std::min(n, ctx->sm_count) : std::min(n, ctx->sm_count * ctx->knobs.lite_ctas);
  // same shared-memory carveout as the comb kernel (which needs the maximum): an SM only hosts CTAs of both kernels at
  // once when they agree on the L1 / shared split -- with the default (small) carveout of this kernel the comb CTAs had
  // to wait until its CTA left the SM (measured: step 1.53 ms instead of 1.34)
  static bool carveout_set = false;
  if (!carveout_set) {
    AMTK_CUDA(cudaFuncSetAttribute(logo_lite_kernel<uint8_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    AMTK_CUDA(cudaFuncSetAttribute(logo_lite_kernel<uint16_t>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    AMTK_CUDA(cudaFuncSetAttribute(logo_sum_bulk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carveout_set = true;
  }
  if (clip->bytes_per_sample == 1) logo_lite_kernel<uint8_t><<<grid, kLiteThreads, smem, st>>>(job);
  else logo_lite_kernel<uint16_t><<<grid, kLiteThreads, smem, st>>>(job);
  AMTK_CUDA(cudaGetLastError());
  const int total = n * 2;
  float* sum_out = dscores + (size_t)(lo - row0) * nlogos * 2;
  const size_t sum_smem = (size_t)32 * (countPad + 4) * sizeof(float);
  if (sum_smem <= 200 * 1024) {
    AMTK_CUDA(cudaFuncSetAttribute(logo_sum_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sum_smem));
    logo_sum_bulk_kernel<<<(total + 31) / 32, 32, sum_smem, st>>>(job.scores, count, countPad, n, 2, hl.blackScore, 0, sum_out, nlogos * 2, logo_index * 2, 1);
  } else {
    logo_sum_kernel<<<(total + kSumThreads - 1) / kSumThreads, kSumThreads, 0, st>>>(job.scores, count, countPad, n, 2, hl.blackScore, 0, sum_out, nlogos * 2, logo_index * 2, 1);
  }
  AMTK_CUDA(cudaGetLastError());
  if (side) AMTK_CUDA(cudaEventRecord(ctx->ev_side_done, ctx->side_stream));
  ctx->launches += 2;
  return 1;
}

int amtk_scan_comb_frames(amtk_ctx* ctx, const amtk_clip* clip, amtk_logo* const* logos, int nlogos,
                          const amtk_comb_params* prm, int frame0, int nframes, float* scores, int32_t* counts, int out_on_device) {
  if (ctx && nframes == 0) return 1;
  if (!ctx || !prm || !counts || !scores || !logos || nlogos < 1) AMTK_FAIL("amtk_scan_comb_frames: bad argument");
  if (!validate_clip(clip, true) || !comb_thresholds_ok(prm, clip->bytes_per_sample)) return 0;
  DevSelect ds(ctx); if (!ds.ok) return 0;
  const size_t sbytes = (size_t)nframes * nlogos * 2 * sizeof(float), cbytes = (size_t)nframes * 12 * sizeof(int32_t);
  float* ds_ = scores; int* dc = counts;
  if (!out_on_device) {
    if (!ensure(&ctx->dout, &ctx->dout_bytes, sbytes) || !ensure(&ctx->dout2, &ctx->dout2_bytes, cbytes)) return 0;
    ds_ = reinterpret_cast<float*>(ctx->dout); dc = reinterpret_cast<int*>(ctx->dout2);
  }
  // One logo that fits the small-footprint kernel (the headline case): its evaluation runs UNDER the streaming pass on the
  // side stream.  Anything else (several logos, large logos) takes the serial path after the comb kernel.
  const amtk_logo* lg0 = logos[0];
  const bool lite = ctx->knobs.scan_lite && nlogos == 1 && lg0 && lg0->has_mask && lg0->host.count() > 0 &&
                    lg0->host.imgw == clip->width && lg0->host.imgh == clip->height &&
                    roi_inside(lg0->host, clip, clip->pitch_y / clip->bytes_per_sample) &&
                    logo_lite_smem_bytes(lg0->host.w, lg0->host.h, clip->bytes_per_sample) <= 29 * 1024;
  if (!for_each_window(ctx, clip, frame0, nframes, true, [&](const Window& w, int lo, int hi) -> int {
        if (lite && ctx->knobs.scan_lite == 2) {          // the small-footprint kernel on its own, after the comb kernel
          return launch_comb(ctx, clip, w, lo, hi, prm, dc, frame0) && launch_scan_lite(ctx, clip, w, lo, hi, lg0, ds_, nlogos, 0, frame0, false);
        }
        if (lite) {
          // comb first: its CTAs take three slots on every SM, and the only place left for the logo kernel's CTAs is the
          // one remaining slot per SM (launched the other way round the scheduler may stack several logo CTAs on one SM)
          AMTK_CUDA(cudaEventRecord(ctx->ev_side, ctx->stream));               // side stream starts after what is queued so far
          if (!launch_comb(ctx, clip, w, lo, hi, prm, dc, frame0)) return 0;
          if (!launch_scan_lite(ctx, clip, w, lo, hi, lg0, ds_, nlogos, 0, frame0, true)) return 0;
          return cuda_ok(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_done, 0), "cudaStreamWaitEvent") ? 1 : 0;
        }
        if (ctx->knobs.scan_overlap && clip->on_device) {
          // Opt-in (AMTK_SCAN_OVERLAP=1): the logo kernels go to the side stream right behind the comb launch.  They need a
          // whole SM each (512 threads x 128 registers), so they never share an SM with a comb CTA; the block scheduler packs
          // them into the gaps at both ends of the comb kernel.  Measured: step 1.307 ms instead of 1.328 ms, but some logo
          // CTAs take their SM BEFORE the comb kernel's CTAs arrive, which stretches the comb kernel's own duration by 60 us and
          // would misstate its roofline fraction; gating the side stream on an "all comb CTAs resident" word (stream memory
          // operation) kept the comb duration but cost the kernel as much as the overlap gained.  Default: serial.
          ctx->want_side_mark = true;
          if (!launch_comb(ctx, clip, w, lo, hi, prm, dc, frame0)) { ctx->want_side_mark = false; return 0; }
          const bool marked = !ctx->want_side_mark;            // the warp-stream launch recorded ev_side right before its kernel
          ctx->want_side_mark = false;
          if (marked) {
            AMTK_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side, 0));
            cudaStream_t main_stream = ctx->stream;
            ctx->stream = ctx->side_stream;                  // the context is locked (DevSelect): nobody else sees the swap
            const int ok = scan_frames_impl(ctx, clip, clip, 0, 0, logos, nlogos, w, lo, hi, 0, ds_, frame0);
            ctx->stream = main_stream;
            if (!ok) return 0;
            AMTK_CUDA(cudaEventRecord(ctx->ev_side_done, ctx->side_stream));
            return cuda_ok(cudaStreamWaitEvent(ctx->stream, ctx->ev_side_done, 0), "cudaStreamWaitEvent") ? 1 : 0;
          }
          return scan_frames_impl(ctx, clip, clip, 0, 0, logos, nlogos, w, lo, hi, 0, ds_, frame0);
        }
        return launch_comb(ctx, clip, w, lo, hi, prm, dc, frame0) &&
               scan_frames_impl(ctx, clip, clip, 0, 0, logos, nlogos, w, lo, hi, 0, ds_, frame0); }))
    return 0;
  if (out_on_device) return 1;
  AMTK_CUDA(cudaMemcpyAsync(scores, ds_, sbytes, cudaMemcpyDeviceToHost, ctx->stream));
  AMTK_CUDA(cudaMemcpyAsync(counts, dc, cbytes, cudaMemcpyDeviceToHost, ctx->stream));
  AMTK_CUDA(cudaStreamSynchronize(ctx->stream));
  return 1;
}

// ---------------------------------------------------------------------------------------------------------
// LogoScan accumulation
// ---------------------------------------------------------------------------------------------------------
int amtk_scan_create(amtk_ctx* ctx, int scanw, int scanh, int lx, int ly, int thy, amtk_scan** out) {
  if (!ctx || !out) AMTK_FAIL("amtk_scan_create: null argument");
  if (scanw < 4 || scanh < 4 || scanw > 4096 || scanh > 4096 || lx < 0 || lx > 2 || ly < 0 || ly > 2) AMTK_FAIL("amtk_scan_create: bad geometry");
  DevSelect ds(ctx); if (!ds.ok) return 0;
  amtk_scan* s = new amtk_scan();
  s->ctx = ctx; s->device = ctx->device; s->scanw = scanw; s->scanh = scanh; s->logUVx = lx; s->logUVy = ly; s->thy = thy;
  s->npix = (size_t)scanw * scanh + 2 * (size_t)(scanw >> lx) * (scanh >> ly);
  if (!cuda_ok(cudaMalloc(&s->dSums, s->npix * 3 * sizeof(unsigned long long)), "cudaMalloc") ||
      !cuda_ok(cudaMalloc(&s->dBg, 8 * sizeof(unsigned long long)), "cudaMalloc")) { amtk_scan_destroy(s); return 0; }
  cudaMemset(s->dSums, 0, s->npix * 3 * sizeof(unsigned long long));
  cudaMemset(s->dBg, 0, 8 * sizeof(unsigned long long));
  *out = s;
  return 1;
}

void amtk_scan_destroy(amtk_scan* s) {
  if (!s) return;
  { DevSelect ds(s->device); if (s->dSums) cudaFree(s->dSums); if (s->dBg) cudaFree(s->dBg); }
  delete s;
}

int amtk_scan_add_frames(amtk_scan* s, const amtk_clip* clip, int scanx, int scany, int frame0, int nframes,
                         const uint8_t* frame_select, uint8_t* valid_out) {
  if (!s) AMTK_FAIL("amtk_scan_add_frames: null scan");
  amtk_ctx* ctx = s->ctx;
  if (!validate_clip(clip, true)) return 0;
  if (clip->bytes_per_sample != 1) AMTK_FAIL("LogoScan supports 8-bit clips only (as the reference, LogoScan.hpp:812)");
  if (clip->log_uvx != s->logUVx || clip->log_uvy != s->logUVy) AMTK_FAIL("chroma subsampling mismatch");
  if (scanx < 0 || scany < 0 || scanx + s->scanw > clip->width || scany + s->scanh > clip->height) AMTK_FAIL("scan rectangle outside the frame");
  DevSelect ds(ctx); if (!ds.ok) return 0;
  // small per-frame buffers: int4 bg[n], u8 select[n], u8 valid[n]
  const size_t bg_bytes = (size_t)nframes * sizeof(int4), off_sel = (bg_bytes + 255) & ~(size_t)255;
  const size_t off_val = off_sel + (((size_t)nframes + 255) & ~(size_t)255);
  if (!ensure(&ctx->dout, &ctx->dout_bytes, off_val + (size_t)nframes + 256)) return 0;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dout);
  int4* dbg = reinterpret_cast<int4*>(base); uint8_t* dsel = base + off_sel; uint8_t* dval = base + off_val;
  if (frame_select) AMTK_CUDA(cudaMemcpyAsync(dsel, frame_select, (size_t)nframes, cudaMemcpyHostToDevice, ctx->stream));
  // host clips: only the scan rectangle (Y, U, V) is uploaded -- LogoScan::AddFrame reads nothing else (LogoScan.hpp:606-635)
  const int ok = for_each_roi_window(ctx, clip, frame0, nframes, scanx, scany, s->scanw, s->scanh, true, false,
                                     [&](const amtk_clip& v, const Window& w, int lo, int hi, int dx, int dy) {
    ScanClip c;
    c.base = w.dev_base; c.frame_stride = v.frame_stride; c.offU = v.off_u; c.offV = v.off_v;
    c.pitchY = v.pitch_y; c.pitchUV = v.pitch_uv;
    c.scanx = scanx - dx; c.scany = scany - dy; c.scanw = s->scanw; c.scanh = s->scanh; c.logUVx = s->logUVx; c.logUVy = s->logUVy; c.thy = s->thy;
    c.frame0 = lo - w.first; c.nframes = hi - lo;
    const int rel = lo - frame0;
    scan_border_kernel<<<hi - lo, 256, 0, ctx->stream>>>(c, frame_select ? dsel + rel : nullptr, dbg + rel);
    AMTK_CUDA(cudaGetLastError());
    const int pixblocks = (int)((s->npix + 255) / 256);
    const int splits = std::max(1, std::min(hi - lo, (ctx->sm_count * 8) / pixblocks));
    scan_accumulate_kernel<<<dim3(pixblocks, splits), 256, 0, ctx->stream>>>(c, dbg + rel, s->dSums, s->dBg, dval + rel);
    AMTK_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return 1;
  });
  if (!ok) return 0;
  std::vector<uint8_t> tmp;
  unsigned long long nv = 0;
  if (valid_out) AMTK_CUDA(cudaMemcpyAsync(valid_out, dval, (size_t)nframes, cudaMemcpyDeviceToHost, ctx->stream));
  AMTK_CUDA(cudaMemcpyAsync(&nv, s->dBg + 6, sizeof(nv), cudaMemcpyDeviceToHost, ctx->stream));
  AMTK_CUDA(cudaStreamSynchronize(ctx->stream));
  s->nvalid = (int)nv;
  return 1;
}

int amtk_scan_num_valid(const amtk_scan* s) { return s ? s->nvalid : 0; }

int amtk_scan_get_sums(amtk_scan* s, double* out) {
  if (!s || !out) AMTK_FAIL("amtk_scan_get_sums: null argument");
  DevSelect ds(s->ctx); if (!ds.ok) return 0;
  std::vector<unsigned long long> h(s->npix * 3), bg(8);
  AMTK_CUDA(cudaStreamSynchronize(s->ctx->stream));
  AMTK_CUDA(cudaMemcpy(h.data(), s->dSums, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  AMTK_CUDA(cudaMemcpy(bg.data(), s->dBg, bg.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  const size_t ny = (size_t)s->scanw * s->scanh, nc = (size_t)(s->scanw >> s->logUVx) * (s->scanh >> s->logUVy);
  for (size_t i = 0; i < s->npix; ++i) {
    const int pl = i < ny ? 0 : (i < ny + nc ? 1 : 2);
    out[i * 5 + 0] = (double)h[i * 3 + 0];        // sumF   (exact: < 2^53)
    out[i * 5 + 1] = (double)bg[pl * 2 + 0];      // sumB
    out[i * 5 + 2] = (double)h[i * 3 + 1];        // sumF2
    out[i * 5 + 3] = (double)bg[pl * 2 + 1];      // sumB2
    out[i * 5 + 4] = (double)h[i * 3 + 2];        // sumFB
  }
  s->nvalid = (int)bg[6];
  return 1;
}

int amtk_scan_get_logo(amtk_scan* s, int maxv, int clean, float* data) {
  if (!s || !data) AMTK_FAIL("amtk_scan_get_logo: null argument");
  std::vector<double> sums(s->npix * 5);
  if (!amtk_scan_get_sums(s, sums.data())) return 0;
  if (!amtk::scan_finalize(sums.data(), s->nvalid, s->scanw, s->scanh, s->logUVx, s->logUVy, maxv, clean != 0, data))
    AMTK_FAIL("Insufficient logo frames");
  return 1;
}

// ---------------------------------------------------------------------------------------------------------
// ScanLogo pipeline (LogoScan.hpp:794-1098)
// ---------------------------------------------------------------------------------------------------------
namespace {
struct ScanGuard { amtk_scan* s = nullptr; ~ScanGuard() { if (s) amtk_scan_destroy(s); } };
struct LogoGuard { amtk_logo* l = nullptr; ~LogoGuard() { if (l) amtk_logo_destroy(l); } };
}

int amtk_scan_logo(amtk_ctx* ctx, const amtk_clip* clip, int service_id, const char* dstpath,
                   int imgx, int imgy, int w, int h, int thy, int max_frames, amtk_logo_analyze_cb cb) {
  if (!ctx || !dstpath) AMTK_FAIL("amtk_scan_logo: null argument");
  if (!validate_clip(clip, true)) return 0;
  if (clip->bytes_per_sample != 1) AMTK_FAIL("LogoScan supports 8-bit clips only (as the reference, LogoScan.hpp:812)");
  const int n = clip->num_frames;
  // ---- MakeInitialLogo (:917-921): frames are offered in reading order until max_frames valid ones were gathered (:884);
  //      every 200 frames read the callback gets (position/size * 50, readCount, 0, numFrames) and may cancel (:905-910).
  //      The clip replaces the decoder, so "position / file size" is frames read / frames in the clip. ----
  std::vector<uint8_t> valid((size_t)n), select((size_t)n, 0);
  int numFrames = 0, nread = 0;
  {
    ScanGuard probe;                       // validity only; the accumulation proper runs once the cut-off frame is known
    if (!amtk_scan_create(ctx, w, h, clip->log_uvx, clip->log_uvy, thy, &probe.s)) return 0;
    while (nread < n && numFrames < max_frames) {
      const int blk = std::min(200, n - nread);
      if (!amtk_scan_add_frames(probe.s, clip, imgx, imgy, nread, blk, nullptr, valid.data() + nread)) return 0;
      int i = nread;
      for (; i < nread + blk && numFrames < max_frames; ++i) if (valid[i]) { select[i] = 1; ++numFrames; }
      // the frame on which the limit is reached is the last one processed (onFrame returns false on the NEXT call, :884)
      nread = (numFrames >= max_frames) ? i : nread + blk;
      if ((nread % 200) == 0 && cb && !cb(50.0f * (float)nread / (float)std::max(1, n), nread, 0, numFrames)) AMTK_FAIL("Cancel requested");
    }
  }
  const size_t ndata = ((size_t)w * h + 2 * (size_t)(w >> clip->log_uvx) * (h >> clip->log_uvy)) * 2;
  std::vector<float> logodata(ndata);
  {
    ScanGuard init;
    if (!amtk_scan_create(ctx, w, h, clip->log_uvx, clip->log_uvy, thy, &init.s)) return 0;
    if (!amtk_scan_add_frames(init.s, clip, imgx, imgy, 0, nread, select.data(), nullptr)) return 0;
    if (!amtk_scan_get_logo(init.s, 255, 0, logodata.data())) return 0;      // "Insufficient logo frames"
  }
  // ---- ReMakeLogo x2 (:923-1036): 20-fade sweep over the STORED frames; every 100 of them the callback gets
  //      (i / numFrames * 25 + progressbase, i, numFrames, numFrames) (:977-982); frames whose best fade index is > 8 are
  //      accumulated again (:1018-1021).  Only frames [0, nread) are touched. ----
  float fades[20];
  for (int fi = 0; fi < 20; ++fi) fades[fi] = 0.1f * fi;                      // :967
  const int kBlock = 128;                                                      // clip frames per sweep call
  std::vector<float> sweep((size_t)kBlock * 20);
  for (int round = 0; round < 2; ++round) {
    const float progressbase = 50.0f + 25.0f * round;                          // :1064-1068
    LogoGuard raw, deint;
    if (!amtk_logo_create(ctx, logodata.data(), w, h, clip->log_uvx, clip->log_uvy, w, h, imgx, imgy, &raw.l)) return 0;
    if (!amtk_logo_deint(raw.l, &deint.l) || !amtk_logo_create_mask(deint.l, 0.1f)) return 0;      // :929-931
    std::vector<uint8_t> sel2((size_t)n, 0);
    int stored = 0;                                                            // the reference's i: index among the stored frames
    for (int f0 = 0; f0 < nread; f0 += kBlock) {
      const int blk = std::min(kBlock, nread - f0);
      bool any = false;
      for (int i = f0; i < f0 + blk; ++i) any = any || select[i];
      if (!any) continue;
      if (!amtk_logo_eval_fades(ctx, clip, deint.l, fades, 20, f0, blk, sweep.data(), 0)) return 0;
      for (int i = f0; i < f0 + blk; ++i) {
        if (!select[i]) continue;
        float best = FLT_MAX; int bi = 0;                                      // :964-975, first strict minimum of |score|
        for (int fi = 0; fi < 20; ++fi) { const float r = std::fabs(sweep[(size_t)(i - f0) * 20 + fi]); if (r < best) { best = r; bi = fi; } }
        sel2[i] = bi > 8;                                                      // :1018-1021
        if ((stored % 100) == 0 && cb && !cb((float)stored / (float)numFrames * 25.0f + progressbase, stored, numFrames, numFrames))
          AMTK_FAIL("Cancel requested");
        ++stored;
      }
    }
    ScanGuard acc;
    if (!amtk_scan_create(ctx, w, h, clip->log_uvx, clip->log_uvy, thy, &acc.s)) return 0;
    if (!amtk_scan_add_frames(acc.s, clip, imgx, imgy, 0, nread, sel2.data(), nullptr)) return 0;
    if (!amtk_scan_get_logo(acc.s, 255, 1, logodata.data())) return 0;        // :1030-1035
  }
  if (cb && !cb(1.0f, numFrames, numFrames, numFrames)) AMTK_FAIL("Cancel requested");     // :1071-1073
  LogoGuard fin;                                                                // :1075-1078
  if (!amtk_logo_create(nullptr, logodata.data(), w, h, clip->log_uvx, clip->log_uvy, clip->width, clip->height, imgx, imgy, &fin.l)) return 0;
  return amtk_logo_save(fin.l, dstpath, "No Name", service_id);
}

// ---------------------------------------------------------------------------------------------------------
// erase
// ---------------------------------------------------------------------------------------------------------
int amtk_erase_logo_frames(amtk_ctx* ctx, const amtk_clip* clip, const amtk_logo* logo, int frame0, int nframes, const float* fades) {
  if (!ctx || !logo || !fades) AMTK_FAIL("amtk_erase_logo_frames: bad argument");
  if (!validate_clip(clip, true)) return 0;
  const amtk::HostLogo& h = logo->host;
  if (h.imgx < 0 || h.imgy < 0 || h.imgx + h.w > clip->width || h.imgy + h.h > clip->height) AMTK_FAIL("logo rectangle lies outside the frame");
  if (frame0 < 0 || nframes < 0 || frame0 + nframes > clip->num_frames) AMTK_FAIL("frame range outside the clip");
  if (nframes == 0) return 1;
  DevSelect ds(ctx); if (!ds.ok) return 0;
  if (!logo_ensure_device(logo, ctx, false)) return 0;
  if (!ensure(&ctx->dout, &ctx->dout_bytes, (size_t)nframes * 2 * sizeof(float))) return 0;
  AMTK_CUDA(cudaMemcpyAsync(ctx->dout, fades, (size_t)nframes * 2 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  // Device clips are edited in place in HBM.  Host clips (the IClip::GetFrame surface: one MakeWritable'd CPU frame) move
  // only the three logo rectangles: up, Delogo kernel, back down -- not the reference's full-frame copy (LogoScan.hpp:1347).
  const int ok = for_each_roi_window(ctx, clip, frame0, nframes, h.imgx, h.imgy, h.w, h.h, true, true,
                                     [&](const amtk_clip& v, const Window& w, int lo, int hi, int dx, int dy) {
    EraseJob j;
    j.base = const_cast<uint8_t*>(w.dev_base); j.frame_stride = v.frame_stride;
    j.offU = v.off_u; j.offV = v.off_v;
    j.pitchY = v.pitch_y / v.bytes_per_sample; j.pitchUV = v.pitch_uv / v.bytes_per_sample;
    j.frame0 = lo - w.first; j.nframes = hi - lo;
    j.w = h.w; j.h = h.h; j.logUVx = h.logUVx; j.logUVy = h.logUVy; j.imgx = h.imgx - dx; j.imgy = h.imgy - dy;
    j.uvparity = (h.imgy / 2) % 2;                                             // LogoScan.hpp:1385, real frame position
    j.aY = logo->dA; j.bY = logo->dB; j.aU = logo->dAU; j.bU = logo->dBU; j.aV = logo->dAV; j.bV = logo->dBV;
    j.fades = reinterpret_cast<const float*>(ctx->dout) + (size_t)(lo - frame0) * 2;
    j.maxv = (float)((1 << v.bits_per_sample) - 1);
    if (v.bytes_per_sample == 1) erase_logo_kernel<uint8_t><<<hi - lo, 256, 0, ctx->stream>>>(j);
    else erase_logo_kernel<uint16_t><<<hi - lo, 256, 0, ctx->stream>>>(j);
    AMTK_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return 1;
  });
  if (!ok) return 0;
  AMTK_CUDA(cudaStreamSynchronize(ctx->stream));    // `fades` staging buffer is reused by later calls; host frames are complete
  return 1;
}

int amtk_weave_frames(amtk_ctx* ctx, const amtk_clip* src, const amtk_clip* dst, int dst_frame0,
                      const int32_t* top_idx, const int32_t* bottom_idx, int n, int src_is_nv12) {
  if (!ctx || !top_idx || !bottom_idx) AMTK_FAIL("amtk_weave_frames: null argument");
  if (!validate_clip(src, true) || !validate_clip(dst, true)) return 0;
  if (!src->on_device || !dst->on_device) AMTK_FAIL("amtk_weave_frames: clips must be device resident");
  if (src->width != dst->width || src->height != dst->height || src->bytes_per_sample != dst->bytes_per_sample ||
      src->log_uvx != dst->log_uvx || src->log_uvy != dst->log_uvy)
    AMTK_FAIL("amtk_weave_frames: source and destination formats differ");
  if (n < 0 || dst_frame0 < 0 || dst_frame0 + n > dst->num_frames) AMTK_FAIL("frame range outside the clip");
  for (int k = 0; k < n; ++k)
    if (top_idx[k] < 0 || top_idx[k] >= src->num_frames || bottom_idx[k] < 0 || bottom_idx[k] >= src->num_frames)
      AMTK_FAIL("amtk_weave_frames: source frame index outside the clip");
  if (n == 0) return 1;
  DevSelect ds(ctx); if (!ds.ok) return 0;
  if (!ensure(&ctx->dout, &ctx->dout_bytes, (size_t)n * 2 * sizeof(int))) return 0;
  int* didx = reinterpret_cast<int*>(ctx->dout);
  AMTK_CUDA(cudaMemcpyAsync(didx, top_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  AMTK_CUDA(cudaMemcpyAsync(didx + n, bottom_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  WeaveJob j;
  j.src = reinterpret_cast<const uint8_t*>(src->base); j.dst = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(dst->base));
  j.sstride = src->frame_stride; j.dstride = dst->frame_stride;
  j.s_offu = src->off_u; j.s_offv = src->off_v; j.d_offu = dst->off_u; j.d_offv = dst->off_v;
  j.s_pitchY = src->pitch_y; j.s_pitchUV = src->pitch_uv; j.d_pitchY = dst->pitch_y; j.d_pitchUV = dst->pitch_uv;
  j.bps = src->bytes_per_sample; j.nv12 = src_is_nv12 ? 1 : 0;
  j.H = src->height; j.HC = src->height >> src->log_uvy;
  j.row_bytes_y = src->width * j.bps; j.row_bytes_c = (src->width >> src->log_uvx) * j.bps;
  j.top_idx = didx; j.bot_idx = didx + n; j.dst_frame0 = dst_frame0;
  for (int k0 = 0; k0 < n; k0 += 32768) {
    WeaveJob jj = j; jj.top_idx += k0; jj.bot_idx += k0; jj.dst_frame0 += k0;
    const int nn = std::min(32768, n - k0);
    const long long work = (long long)j.H * ((j.row_bytes_y + 15) / 16);
    dim3 grid((unsigned)std::min<long long>((work + 255) / 256, 4096), 3, nn);
    weave_kernel<<<grid, 256, 0, ctx->stream>>>(jj);
    AMTK_CUDA(cudaGetLastError());
    ctx->launches += 1;
  }
  AMTK_CUDA(cudaStreamSynchronize(ctx->stream));      // index staging buffer is reused by later calls
  return 1;
}

void amtk_calc_fade2(const float* records, int num_records, int num_frames, int n, float* ft, float* fb) {
  amtk::calc_fade2(records, num_records, num_frames, n, ft, fb);
}
int amtk_calc_fade2_index(int num_records, int num_frames, int n, int i) { return amtk::calc_fade2_index(num_records, num_frames, n, i); }
void amtk_calc_fade2_records(const float* rec9, float* ft, float* fb) { amtk::calc_fade2_records(rec9, ft, fb); }

}  // extern "C"

#include "group.cuh"
