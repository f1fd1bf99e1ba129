/* include/amtk_b200.h -- C ABI of libamtk_b200.so: the B200-native (sm_100a) implementation of Amatsukaze's
 * per-frame pixel-analysis hot path (logo-template correlation, LogoScan accumulation, logo erase, and the
 * field-difference / combing metric).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference's Amatsukaze/ directory).  A reference-side
 * binding (what a maintainer would add to LogoScan.hpp / FilteredSource.hpp) is shown in INTEGRATION.md.
 *
 * Conventions (mirroring the reference's C exports, StreamUtils.hpp:1037-1039 + LogoScan.hpp:1083-1098):
 *   - every function returns 1 on success, 0 on failure; after a failure amtk_last_error() returns the
 *     message for the calling thread (the reference: `return false` after ctx->setError(); text fetched with
 *     AMTContext_GetError()).
 *   - handles are opaque, owned by the caller, released with the matching *_destroy().
 *   - a context is bound to ONE CUDA device and ONE stream; calls on distinct contexts run concurrently, calls on the
 *     SAME context from several host threads are safe and serialise on the context (the reference filters answer
 *     CACHE_GET_MTMODE with MT_NICE_FILTER, LogoScan.hpp:1220-1225,1500-1505: AviSynth may call GetFrame from several
 *     Prefetch threads at once).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef AMTK_B200_H
#define AMTK_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AMTK_API __attribute__((visibility("default")))

typedef struct amtk_ctx amtk_ctx;
typedef struct amtk_logo amtk_logo;
typedef struct amtk_scan amtk_scan;

/* ---------------------------------------------------------------------------------------------
 * Context / errors   (replaces AMTContext_Create / ATMContext_Delete / AMTContext_GetError,
 *                     StreamUtils.hpp:1037-1039, for this path)
 * ------------------------------------------------------------------------------------------- */
AMTK_API const char* amtk_last_error(void);
AMTK_API int amtk_version(void);
/* number of CUDA devices visible (0 when there is no driver/GPU; never fails) */
AMTK_API int amtk_device_count(void);
/* device: CUDA ordinal.  stream: the cudaStream_t every kernel of this context is launched on (e.g. torch's
 * current stream); NULL = the legacy default stream. */
AMTK_API int amtk_ctx_create(int device, void* cuda_stream, amtk_ctx** out);
AMTK_API void amtk_ctx_destroy(amtk_ctx* ctx);
AMTK_API int amtk_ctx_synchronize(amtk_ctx* ctx);
/* kernels launched by this context since creation (bench.py reports it as gpu_launches) */
AMTK_API int64_t amtk_ctx_launch_count(const amtk_ctx* ctx);
/* payload bytes the last call on a HOST clip copied host->device (whole frames for the combing pass, only the logo /
 * scan rectangle rows for the logo entry points -- what the reference reads there, LogoScan.hpp:1559-1566) */
AMTK_API int64_t amtk_ctx_last_h2d_bytes(const amtk_ctx* ctx);

/* Per-launch timing of the dominant streaming kernel (comb) with CUDA events recorded on the context's stream
 * around each launch; get() synchronizes the stream, returns the accumulated milliseconds and launch count. */
AMTK_API int amtk_ctx_set_kernel_timing(amtk_ctx* ctx, int enable);
AMTK_API int amtk_ctx_get_kernel_timing(amtk_ctx* ctx, double* ms_total, int64_t* launches, int reset);

/* Measurement helper: a trivial streaming read (uint4 loads, XOR-reduced) over [ptr, ptr+bytes) on the context's
 * stream, timed with CUDA events; returns the average milliseconds of `reps` passes after one warm-up pass.
 * bench.py reports bytes/ms as the read-only HBM ceiling next to the roofline numbers. */
AMTK_API int amtk_probe_read_ms(amtk_ctx* ctx, const void* device_ptr, size_t bytes, int reps, double* ms_out);

/* Pinned host memory for the host-buffer entry points (optional; pageable memory works, only slower). */
AMTK_API int amtk_host_alloc(size_t bytes, void** out);
AMTK_API void amtk_host_free(void* p);
/* HBM buffers for device-resident clips (what AMTSource keeps its decoded frames in, replacing the reference's CPU
 * frame cache, AMTSource.hpp:419-425) and synchronous copies on the context's stream. */
AMTK_API int amtk_device_alloc(amtk_ctx* ctx, size_t bytes, void** out);
AMTK_API void amtk_device_free(amtk_ctx* ctx, void* p);
AMTK_API int amtk_memcpy_h2d(amtk_ctx* ctx, void* dst_device, const void* src_host, size_t bytes);
AMTK_API int amtk_memcpy_d2h(amtk_ctx* ctx, void* dst_host, const void* src_device, size_t bytes);
AMTK_API int amtk_memcpy_d2d(amtk_ctx* ctx, void* dst_device, const void* src_device, size_t bytes);   /* MakeWritable of a device frame */

/* ---------------------------------------------------------------------------------------------
 * Clip descriptor: a run of planar YUV frames, either resident in HBM or in host memory.
 * Describes what the reference's filters read through PVideoFrame::GetReadPtr/GetPitch(PLANAR_Y|U|V)
 * (LogoScan.hpp:1138-1145, AMTSource.hpp:357-408).
 * ------------------------------------------------------------------------------------------- */
typedef struct amtk_clip {
  const void* base;        /* first byte of frame 0 (its Y plane)                                  */
  int64_t frame_stride;    /* bytes from one frame to the next                                     */
  int64_t off_u, off_v;    /* byte offsets of the U and V planes inside a frame                    */
  int32_t width, height;   /* luma size in pixels                                                  */
  int32_t pitch_y, pitch_uv; /* bytes per row.  When base, frame_stride, off_u/off_v and the pitches are  */
                             /* all multiples of 16 bytes the TMA streaming kernels run; any other layout */
                             /* takes slower plain-load kernels with identical results                    */
  int32_t log_uvx, log_uvy;  /* chroma subsampling shifts (1,1 for YV12 / YUV420P10)               */
  int32_t bytes_per_sample;  /* 1 (YV12) or 2 (YUV420P10/P12/P16, little endian)                   */
  int32_t bits_per_sample;   /* 8, 10, 12 or 16: maxv = (1<<bits)-1 (LogoScan.hpp:1130,1575); every sample must be
                              * <= maxv (the 10-bit combing path relies on it, as the reference's 10-bit formats do) */
  int32_t num_frames;
  int32_t on_device;         /* 1: base is a device pointer on the context's device; 0: host pointer */
} amtk_clip;

/* ---------------------------------------------------------------------------------------------
 * Logos   (replaces logo::LogoData / logo::LogoDataParam, AMTLogo.hpp:49-280, LogoScan.hpp:61-334)
 * ------------------------------------------------------------------------------------------- */
typedef struct amtk_logo_info {
  int32_t w, h, log_uvx, log_uvy;
  int32_t imgw, imgh, imgx, imgy;
  int32_t maskpixels;      /* min(w*h,(int)(w*h*maskratio)), LogoScan.hpp:172; 0 before create_mask */
  int32_t count;           /* mask pixels the y,x in [2,dim-2) scan visits (= kernels/scales rows)   */
  float black_score;       /* LogoScan.hpp:227-228                                                   */
} amtk_logo_info;

/* data: aY,bY,aU,bU,aV,bV contiguous floats -- LogoData's own layout (AMTLogo.hpp:203-212).
 * Logos are host objects: ctx may be NULL; the HBM copy of a logo's tables is made by the first context that
 * evaluates it (and the logo then belongs to that device). */
AMTK_API int amtk_logo_create(amtk_ctx* ctx, const float* data, int w, int h, int log_uvx, int log_uvy,
                              int imgw, int imgh, int imgx, int imgy, amtk_logo** out);
/* LogoData::Load (AMTLogo.hpp:257-279): reads a .lgd, skipping the AviUtl base part; header540 (may be NULL)
 * receives the raw 540-byte LogoHeader (AMTLogo.hpp:19-47). */
AMTK_API int amtk_logo_load(amtk_ctx* ctx, const char* path, amtk_logo** out, void* header540);
/* LogoData::Save (AMTLogo.hpp:239-255) incl. the AviUtl-compatible base part (ToOutLGP, :96-167). */
AMTK_API int amtk_logo_save(const amtk_logo* logo, const char* path, const char* name, int service_id);
AMTK_API void amtk_logo_destroy(amtk_logo* logo);
/* DeintLogo (LogoScan.hpp:734-761): vertical [1 2 1]/4 of the Y planes, same image placement. */
AMTK_API int amtk_logo_deint(const amtk_logo* src, amtk_logo** out);
/* LogoDataParam::MakeFieldLogo (LogoScan.hpp:257-283). */
AMTK_API int amtk_logo_field(const amtk_logo* src, int bottom, amtk_logo** out);
/* LogoDataParam::CreateLogoMask (LogoScan.hpp:112-229): mask, kernels, scale tables, blackScore; uploads the
 * evaluation tables to HBM on next use.  Must be called before the logo is used by scan/analyze entry points.
 * (Setup-time host code, once per logo -- exactly where the reference runs it; the per-frame path is CUDA.) */
AMTK_API int amtk_logo_create_mask(amtk_logo* logo, float maskratio);
AMTK_API int amtk_logo_get_info(const amtk_logo* logo, amtk_logo_info* out);
/* Copies out host tables (any pointer may be NULL): data (LogoData layout), mask w*h, kernels count*25,
 * scales count*32*2 {scale,scale2}. */
AMTK_API int amtk_logo_get_tables(const amtk_logo* logo, float* data, uint8_t* mask, float* kernels, float* scales);

/* ---------------------------------------------------------------------------------------------
 * Logo evaluation
 * ------------------------------------------------------------------------------------------- */
/* LogoFrame::ScanFrame over frames [frame0, frame0+nframes) (LogoScan.hpp:1543-1589; the CMAnalyze entry,
 * CMAnalyze.hpp:291-292).  logos[i] are DEINT logos with masks.  out: float[nframes][nlogos][2] = corr0,corr1
 * (EvalResult, :1532-1535); a logo whose imgw/imgh differ from the clip yields (0,-1) (:1551-1558).
 * pitch_elems_override: 0 = use clip.pitch_y/bytes_per_sample; >0 = element pitch to use for addressing the Y
 * plane (the reference passes the BYTE pitch even for 16-bit samples, :1547,1561 -- see INTEGRATION.md).
 * out_on_device: 1 = out is a device pointer (stays in HBM), 0 = host pointer. */
AMTK_API int amtk_logo_scan_frames(amtk_ctx* ctx, const amtk_clip* clip, amtk_logo* const* logos, int nlogos,
                                   int frame0, int nframes, int pitch_elems_override, float* out, int out_on_device);
/* AMTAnalyzeLogo::GetFrameT body per SOURCE frame (LogoScan.hpp:1119-1161): out float[nframes][33] =
 * LogoAnalyzeFrame{p[11],t[11],b[11]} (:1100-1103) for source frames frame0.. (the caller groups 8 per output
 * frame and clamps, :1133). */
AMTK_API int amtk_logo_analyze_frames(amtk_ctx* ctx, const amtk_clip* clip, const amtk_logo* deint_logo,
                                      const amtk_logo* field_top, const amtk_logo* field_bottom,
                                      int frame0, int nframes, float* out, int out_on_device);
/* General form used by LogoAnalyzer::ReMakeLogo's 20-fade sweep (LogoScan.hpp:955-975): evaluates
 * EvaluateLogo(DeintY(roi), maxv, fades[i]) for every frame; out float[nframes][nfades] (signed scores). */
AMTK_API int amtk_logo_eval_fades(amtk_ctx* ctx, const amtk_clip* clip, const amtk_logo* deint_logo,
                                  const float* fades, int nfades, int frame0, int nframes, float* out, int out_on_device);

/* ---------------------------------------------------------------------------------------------
 * Field-difference / combing metric (the telecine pre-pass the reference drives through
 * AMTFilterSource::FilterPass/ReadAllFrames, FilteredSource.hpp:232-238,417-439,519-544, and computes in the
 * external KFM plugin).  Integer spec: DESIGN.md section 4; 8-bit and 16-bit (YUV420P10/12/16) samples.
 * counts int32[nframes][12] = [plane class Y,C][field top,bottom][move, shima, lshima].
 * Thresholds: 8-bit th_move in [1,128], th_shima/th_lshima in [1,2047]; 16-bit th_move in [1,32768], others >= 1.
 * ------------------------------------------------------------------------------------------- */
typedef struct amtk_comb_params {
  int32_t th_move_y, th_shima_y, th_lshima_y;   /* defaults 20, 12, 36 */
  int32_t th_move_c, th_shima_c, th_lshima_c;   /* defaults 24, 16, 48 */
} amtk_comb_params;
AMTK_API void amtk_comb_default_params(amtk_comb_params* p);
/* prev(frame0) is frame0-1 when frame0 > 0 (halo frame for range-sharded clips), else frame0 itself. */
AMTK_API int amtk_comb_frames(amtk_ctx* ctx, const amtk_clip* clip, const amtk_comb_params* params,
                              int frame0, int nframes, int32_t* counts, int out_on_device);

/* The fused hot-path step of BASELINE.json's headline metric: one pass over frames [frame0, frame0+nframes)
 * producing BOTH the ScanFrame scores (as amtk_logo_scan_frames) and the combing counters (as amtk_comb_frames). */
AMTK_API int amtk_scan_comb_frames(amtk_ctx* ctx, const amtk_clip* clip, amtk_logo* const* logos, int nlogos,
                                   const amtk_comb_params* params, int frame0, int nframes,
                                   float* scores, int32_t* counts, int out_on_device);

/* ---------------------------------------------------------------------------------------------
 * LogoScan accumulation (replaces logo::LogoScan, LogoScan.hpp:398-660; the ScanLogo C export's inner loop,
 * :1083-1098 -> :881-914)
 * ------------------------------------------------------------------------------------------- */
AMTK_API int amtk_scan_create(amtk_ctx* ctx, int scanw, int scanh, int log_uvx, int log_uvy, int thy, amtk_scan** out);
AMTK_API void amtk_scan_destroy(amtk_scan* s);
/* LogoScan::AddFrame for every frame in [frame0, frame0+nframes) with the ROI at (scanx, scany) (:594-659);
 * valid_out (may be NULL; host pointer) receives 1/0 per frame = AddFrame's return value.
 * frame_select (may be NULL; host pointer, nframes bytes): only frames with a non-zero byte are offered
 * (ReMakeLogo's `minFades[i] > 8` filter, :1018-1021). */
AMTK_API int amtk_scan_add_frames(amtk_scan* s, const amtk_clip* clip, int scanx, int scany, int frame0, int nframes,
                                  const uint8_t* frame_select, uint8_t* valid_out);
AMTK_API int amtk_scan_num_valid(const amtk_scan* s);
/* raw accumulators as doubles, plane-major Y,U,V, 5 per pixel {sumF,sumB,sumF2,sumB2,sumFB} (LogoColor, :346) */
AMTK_API int amtk_scan_get_sums(amtk_scan* s, double* out);
/* LogoScan::Normalize(maxv) + GetLogo(clean) (:471-566): fills data (LogoData layout); returns 0 with error
 * "Insufficient logo frames" when the reference would return nullptr (:847-849). */
AMTK_API int amtk_scan_get_logo(amtk_scan* s, int maxv, int clean, float* data);

/* The whole logo-generation pipeline of the reference's ScanLogo C export (LogoScan.hpp:1083-1098 ->
 * LogoAnalyzer::ScanLogo :1058-1079): MakeInitialLogo (:917-921, AddFrame on every frame until max_frames valid ones),
 * ReMakeLogo twice (:923-1036: deint logo, mask 0.1, 20-fade sweep per valid frame, re-accumulate frames whose best
 * fade index is > 8, GetLogo(clean)), then LogoData::Save.  The clip stands where the reference decodes `srcpath`;
 * the valid ROI frames stay in HBM instead of the UtVideo work file.  cb (may be NULL) has the reference's
 * LOGO_ANALYZE_CB signature (:792): bool(float progress, int nread, int total, int ngather); returning 0 cancels
 * ("Cancel requested", :908-910).  Only 8-bit clips, like the reference (:812). */
typedef int (*amtk_logo_analyze_cb)(float progress, int nread, int total, int ngather);
AMTK_API int amtk_scan_logo(amtk_ctx* ctx, const amtk_clip* clip, int service_id, const char* dstpath,
                            int imgx, int imgy, int w, int h, int thy, int max_frames, amtk_logo_analyze_cb cb);

/* ---------------------------------------------------------------------------------------------
 * Frame ingest: field weave + NV12 split on the device (replaces AMTSource::MergeField / Copy1 / Copy2,
 * AMTSource.hpp:291-355, used by MakeFrame :357-366 for half-delay (BFF / repeat-field) sources,
 * StreamReform.hpp:890-903).  dst frame k (k = 0..n-1, starting at dst_frame0) takes its EVEN rows (luma and chroma)
 * from src frame top_idx[k] and its ODD rows from src frame bottom_idx[k]; with src_is_nv12 the source chroma is one
 * interleaved UV plane at off_u (pitch_uv bytes per row) that is split into dst's U and V planes.
 * Both clips must be device resident, same size and sample format.  top_idx/bottom_idx are host pointers.
 * ------------------------------------------------------------------------------------------- */
AMTK_API int amtk_weave_frames(amtk_ctx* ctx, const amtk_clip* src, const amtk_clip* dst, int dst_frame0,
                               const int32_t* top_idx, const int32_t* bottom_idx, int n, int src_is_nv12);

/* ---------------------------------------------------------------------------------------------
 * Logo erase (replaces AMTEraseLogo::Delogo on Y,U,V, LogoScan.hpp:1248-1261,1374-1397), in place on a
 * device-resident or host clip.  fades float[nframes][2] = fadeT,fadeB per frame (host pointer).
 * ------------------------------------------------------------------------------------------- */
AMTK_API int amtk_erase_logo_frames(amtk_ctx* ctx, const amtk_clip* clip, const amtk_logo* logo,
                                    int frame0, int nframes, const float* fades);
/* AMTEraseLogo::CalcFade2 (LogoScan.hpp:1263-1315) on host records (float[num_records][33]). */
AMTK_API void amtk_calc_fade2(const float* records, int num_records, int num_frames, int n, float* fade_t, float* fade_b);
/* The same decision without materialising every record of the clip: CalcFade2 reads nine records around frame n
 * (offsets i = -4..4, with the reference's double offset, :1273-1275).  _index gives the record each offset reads,
 * _records decides from those nine (float[9][33], offset order). */
AMTK_API int amtk_calc_fade2_index(int num_records, int num_frames, int n, int i);
AMTK_API void amtk_calc_fade2_records(const float* rec9, float* fade_t, float* fade_b);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: ONE process drives several devices -- a context, a stream and a host thread per device
 * (each thread pinned to the CPUs next to its GPU), NCCL over NVLink only for the final gather of the small per-frame
 * result blocks and for the exact integer all-reduce of a frame-sharded LogoScan.  This is what the reference's
 * job-per-GPU scheduler (AmatsukazeServer/Server/ResourceManager.cs:81-85, Amatsukaze/InterProcessComm.hpp:87-95) would
 * drive from C++/C#.  NCCL is loaded with dlopen on first use; one device needs no NCCL at all.
 * ------------------------------------------------------------------------------------------- */
typedef struct amtk_group amtk_group;
/* devices: CUDA ordinals (NULL = 0..ndev-1) */
AMTK_API int amtk_group_create(int ndev, const int* devices, amtk_group** out);
AMTK_API void amtk_group_destroy(amtk_group* g);
AMTK_API int amtk_group_size(const amtk_group* g);
/* the context of member i: create logos / scans / device buffers for that device through it */
AMTK_API amtk_ctx* amtk_group_ctx(amtk_group* g, int i);
/* CPUs member i's host thread was bound to (0 = not bound), NCCL version in use (0 = none) */
AMTK_API int amtk_group_numa_cpus(const amtk_group* g, int i);
AMTK_API int amtk_group_nccl_version(const amtk_group* g);
/* pinned host memory allocated and first-touched by member i's thread (NUMA-local to its GPU) */
AMTK_API int amtk_group_host_alloc(amtk_group* g, int i, size_t bytes, void** out);
/* BASELINE configs[4]: clips[i] (host or device resident, all with >= nframes frames) is analysed on member i with logos[i]
 * (amtk_scan_comb_frames semantics, frames [0, nframes)), then ONE ncclAllGather of the per-member result blocks.
 * Device clips: returns when the work is enqueued (asynchronous); host clips: returns when the staging is done. */
AMTK_API int amtk_group_scan_comb_streams(amtk_group* g, const amtk_clip* clips, amtk_logo* const* logos,
                                          const amtk_comb_params* params, int nframes);
/* gathered results of the last pass as held by member `from`: scores float[ndev][nframes][2], counts int32[ndev][nframes][12] */
AMTK_API int amtk_group_fetch_results(amtk_group* g, int from, int nframes, float* scores, int32_t* counts);
AMTK_API int amtk_group_synchronize(amtk_group* g);
/* device-side timing: mark = one CUDA event per member after everything enqueued so far (compute and collective);
 * elapsed = milliseconds between two marks per member (take the maximum) */
AMTK_API int amtk_group_mark(amtk_group* g, int slot);
AMTK_API int amtk_group_elapsed_ms(amtk_group* g, int slot_a, int slot_b, double* ms_per_member);
/* frame-sharded LogoScan: member i adds frames [frame0[i], frame0[i]+nframes[i]) of clips[i] to scans[i] (created on
 * amtk_group_ctx(g, i)), then ONE ncclAllReduce(ncclSum, ncclUint64) leaves the whole-clip sums in every scans[i]. */
AMTK_API int amtk_group_scan_add_frames(amtk_group* g, amtk_scan* const* scans, const amtk_clip* clips, int scanx, int scany,
                                        const int* frame0, const int* nframes);

#ifdef __cplusplus
}
#endif
#endif /* AMTK_B200_H */
