// amtk_internal.h -- shared declarations of the CUDA translation unit (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/amtk_b200.h"
#include "logo_host.h"

// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no link-time dependency on libcuda.so).
typedef CUresult (*amtk_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

namespace amtk {

void set_error(const std::string& msg);
bool cuda_ok(cudaError_t e, const char* what);

#define AMTK_CUDA(call)                                          \
  do {                                                           \
    if (!::amtk::cuda_ok((call), #call)) return 0;               \
  } while (0)
#define AMTK_FAIL(msg)                                           \
  do {                                                           \
    ::amtk::set_error(msg);                                      \
    return 0;                                                    \
  } while (0)

// Evaluation tables of one logo as the kernels see them (all device pointers).
struct LogoDev {
  int w, h, count, countPad;      // countPad: kernel-tap row pitch (multiple of 32)
  float blackScore;
  const float* A;                 // w*h
  const float* B;                 // w*h
  const uint32_t* pix;            // count: x | y<<16, reference scan order
  const float* tapsT;             // 25 x countPad, tap-major (coalesced one-time load into registers)
  const float2* scales;           // count x 32 {scale, scale2}
};

}  // namespace amtk

struct amtk_ctx {
  // One context = one device + one stream + one set of scratch buffers.  Entry points serialise on this mutex (taken by
  // DevSelect), so concurrent calls on ONE context from several host threads (AviSynth Prefetch threads calling GetFrame
  // on MT_NICE_FILTER filters) are safe; calls on distinct contexts run concurrently.  Recursive: amtk_scan_logo calls
  // other entry points.
  mutable std::recursive_mutex mu;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;       // H2D staging for host-resident clips
  cudaStream_t side_stream = nullptr;       // the logo evaluation that runs UNDER the streaming comb kernel (fused step)
  cudaEvent_t ev_side = nullptr, ev_side_done = nullptr;
  cudaStream_t side_stream2 = nullptr;      // GetFrame-sized AMTAnalyzeLogo calls: the three logo evaluations run side by side (main + two side streams)
  cudaEvent_t ev_fork = nullptr, ev_join1 = nullptr, ev_join2 = nullptr;
  size_t scratch_off = 0;                   // byte offset into `scratch` the next launch_eval writes its per-pixel scores at
  cudaEvent_t ev_copy[2] = { nullptr, nullptr };
  cudaEvent_t ev_done[2] = { nullptr, nullptr };
  int sm_count = 0;
  int64_t launches = 0;
  long long h2d_bytes_last = 0;             // payload bytes the last host-clip call copied host->device
  // scratch (grown on demand, reused across calls)
  void* scratch = nullptr; size_t scratch_bytes = 0;       // per-pixel scores
  void* stage[2] = { nullptr, nullptr }; size_t stage_bytes = 0;   // device staging of host clips
  void* small = nullptr; size_t small_bytes = 0;           // misc small device buffers (counters, segments)
  void* dout = nullptr; size_t dout_bytes = 0;             // device-side outputs when the caller's are on the host
  void* dout2 = nullptr; size_t dout2_bytes = 0;
  std::vector<std::pair<const void*, int>> smem_attr;       // (kernel, dynamic shared memory limit already set on this device): per-call cudaFuncSetAttribute avoided
  void* hout = nullptr; void* hout_dev = nullptr;            // small host outputs: pinned, device-mapped; the kernels write it directly (no D2H copy operation)
  amtk_encode_tiled_fn encode_tiled = nullptr;
  bool want_side_mark = false;                // fused step (opt-in overlap): the next warp-stream comb launch records ev_side right before its kernel
  struct Knobs {            // kernel-variant selection; read from AMTK_* environment variables at context creation
    int eval_waves = 1;     // logo_scores_kernel CTAs per SM
    int eval_cw = 1;        // 1: 64-pixel-wide logos use the compile-time-width kernel variant
    int eval_par = 1;       // 1: AMTAnalyzeLogo calls of <= 16 frames run their three evaluations concurrently on three streams
    int scan_overlap = 0;   // 1: fused step on resident clips: logo kernels on the side stream (faster step, but stretches the comb kernel's own duration)
    int comb_generic = 0;   // 1: force the plain-load comb kernel
    int comb_merge_uv = 1;  // U|V remainder columns share one tile
    int comb_part = -1;     // partition: -1 auto, 0 equal-share, 1 lock-step
    int comb_strip = 8, comb_stages = 3, comb_R = 0, comb_ctas = 0, comb_sync = 0, comb_l2 = 64;
    int comb_item = 0;        // frames per long work item of the warp-stream kernel (0 = auto)
    int comb_tail = 4;        // third tier of the work queue: items of this many frames over the last ~4 % of the range (0 = two tiers; measured -1.2 %)
    int comb_ws_stages = 2;  // ring slots per warp stream
    int comb_mma = 0;        // 1|2: tensor-core streaming kernel (comb_mma.cuh) for 8-bit clips, NS tiles per CTA step
    int comb_ws10 = 1;       // 16-bit containers with <= 10 significant bits run the warp-stream kernel's integer-lane form
    int comb_ws_warps = 4;   // warp streams per CTA
    int comb_ws_prefetch = 0; // L2 prefetch distance of the warp streams' tile loads (steps ahead of the slot refill)
    int lite_ctas = 5;      // CTAs per SM of the small-footprint logo kernel when it runs on its own
    int scan_lite = 0;      // fused step: 0 = logo_scores after the comb kernel (default); 1 = logo_lite UNDER the comb kernel on the side
                            // stream (step 1.324 vs 1.339 ms, but the comb kernel itself stretches 1.225 -> 1.309 ms); 2 = logo_lite alone (1.370 ms)
    int comb_ws = 1;        // 1: round-2 warp-stream kernel for 8-bit clips (comb_stream.cuh); 0: round-1 CTA-ring kernel
  } knobs;
  // cached launch plan of the streaming comb kernel: work items on the device + occupancy, keyed by geometry and range
  struct CombPlan {
    bool valid = false;
    int wY = 0, hY = 0, wC = 0, hC = 0, nf = 0, f0 = 0, R = 0, item = 0, ctas = 0;
    void* dev = nullptr; size_t cap = 0;      // [items][CombSegment] + queue counter
    int nitems = 0; size_t q_off = 0;
    int occ = 0; const void* occ_kernel = nullptr;
  } plan;
  // optional per-launch timing of the dominant (comb) kernel with CUDA events on the launching stream
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;   // recorded, not yet resolved
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_pool;     // free pairs
  double timing_ms = 0.0; int64_t timing_count = 0;
};

struct amtk_logo {
  // Device the HBM copies live on (-1 = none yet).  Deliberately NOT a pointer to the context that first evaluated the
  // logo: logos are host objects that may outlive any context (use-after-free in destroy/ensure_device).
  int device = -1;
  amtk::HostLogo host;
  // device copies (valid after create_mask; A/B valid from creation)
  float* dA = nullptr; float* dB = nullptr;       // Y planes
  float* dAU = nullptr; float* dBU = nullptr; float* dAV = nullptr; float* dBV = nullptr;
  uint32_t* dPix = nullptr; float* dTapsT = nullptr; float2* dScales = nullptr;
  int countPad = 0;
  bool has_mask = false;
  bool tables_uploaded = false;
  std::mutex mu;
};

struct amtk_scan {
  amtk_ctx* ctx = nullptr;                 // used by add_frames/get_* only (the context must be alive for those calls)
  int device = 0;                          // amtk_scan_destroy needs nothing but the ordinal
  int scanw = 0, scanh = 0, logUVx = 1, logUVy = 1, thy = 0;
  int nvalid = 0;
  unsigned long long* dSums = nullptr;     // [npix][3] u64: sumF, sumF2, sumFB  (exact integers)
  unsigned long long* dBg = nullptr;       // [3 planes][2]: sumB, sumB2 (per plane scalars) + [6] = nvalid
  size_t npix = 0;
};
