// pipe_probe.cu -- which sm_100a issue pipes do the comb kernel's instructions share?  Times warp-instruction
// throughput (instr/clk/SM) of single opcodes and 1:1 mixes at the comb kernel's occupancy (16 warps/SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define OP_HFMA2(r, x, y)   asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y))
#define OP_HFMA2S(r, x, y)  asm volatile("fma.rn.sat.f16x2 %0, %0, %0, %1;" : "+r"(r) : "r"(x))
#define OP_HADD2(r, x, y)   asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r) : "r"(x))
#define OP_HSET2(r, x, y)   asm volatile("set.ge.u32.f16x2 %0, %0, %1;" : "+r"(r) : "r"(x))
#define OP_IDP(r, x, y)     asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(x), "r"(y))
#define OP_IMAD(r, x, y)    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y))
#define OP_IADD(r, x, y)    asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(x))
#define OP_LOP3(r, x, y)    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(x), "r"(y))
#define OP_PRMT(r, x, y)    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y))
#define OP_VABS(r, x, y)    asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(x), "r"(y))
#define OP_FFMA(r, x, y)    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&r) : "f"(*(float*)&x), "f"(*(float*)&y))
#define OP_NONE(r, x, y)

#define KERNEL(name, A, B, C)                                                                             \
  __global__ void __launch_bounds__(128) name(uint32_t* out, uint32_t x, uint32_t y, int iters, long long* clk) { \
    uint32_t a[8], b[8], c[8];                                                                            \
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; b[i] = threadIdx.x * 3 + i; c[i] = threadIdx.x * 7 + i; } \
    __syncthreads();                                                                                      \
    long long t0 = clock64();                                                                             \
    for (int it = 0; it < iters; ++it) {                                                                  \
      _Pragma("unroll") for (int i = 0; i < 8; ++i) { A(a[i], x, y); B(b[i], x, y); C(c[i], x, y); }      \
    }                                                                                                     \
    long long t1 = clock64();                                                                             \
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s += a[i] ^ b[i] ^ c[i];                                  \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                                       \
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;                                                      \
  }

KERNEL(k_hfma2, OP_HFMA2, OP_NONE, OP_NONE)
KERNEL(k_hfma2s, OP_HFMA2S, OP_NONE, OP_NONE)
KERNEL(k_hadd2, OP_HADD2, OP_NONE, OP_NONE)
KERNEL(k_hset2, OP_HSET2, OP_NONE, OP_NONE)
KERNEL(k_idp, OP_IDP, OP_NONE, OP_NONE)
KERNEL(k_imad, OP_IMAD, OP_NONE, OP_NONE)
KERNEL(k_iadd, OP_IADD, OP_NONE, OP_NONE)
KERNEL(k_lop3, OP_LOP3, OP_NONE, OP_NONE)
KERNEL(k_prmt, OP_PRMT, OP_NONE, OP_NONE)
KERNEL(k_vabs, OP_VABS, OP_NONE, OP_NONE)
KERNEL(k_ffma, OP_FFMA, OP_NONE, OP_NONE)
KERNEL(k_hfma2_idp, OP_HFMA2, OP_IDP, OP_NONE)
KERNEL(k_hfma2_imad, OP_HFMA2, OP_IMAD, OP_NONE)
KERNEL(k_hfma2_ffma, OP_HFMA2, OP_FFMA, OP_NONE)
KERNEL(k_hfma2_hset2, OP_HFMA2, OP_HSET2, OP_NONE)
KERNEL(k_hfma2_lop3, OP_HFMA2, OP_LOP3, OP_NONE)
KERNEL(k_hfma2_hfma2s, OP_HFMA2, OP_HFMA2S, OP_NONE)
KERNEL(k_idp_imad, OP_IDP, OP_IMAD, OP_NONE)
KERNEL(k_idp_lop3, OP_IDP, OP_LOP3, OP_NONE)
KERNEL(k_idp_ffma, OP_IDP, OP_FFMA, OP_NONE)
KERNEL(k_hset2_lop3, OP_HSET2, OP_LOP3, OP_NONE)
KERNEL(k_prmt_lop3, OP_PRMT, OP_LOP3, OP_NONE)
KERNEL(k_vabs_lop3, OP_VABS, OP_LOP3, OP_NONE)
KERNEL(k_hfma2_idp_lop3, OP_HFMA2, OP_IDP, OP_LOP3)
KERNEL(k_hfma2_ffma_lop3, OP_HFMA2, OP_FFMA, OP_LOP3)
KERNEL(k_hfma2_imad_lop3, OP_HFMA2, OP_IMAD, OP_LOP3)
KERNEL(k_hfma2_hfma2s_idp, OP_HFMA2, OP_HFMA2S, OP_IDP)

typedef void (*kern_t)(uint32_t*, uint32_t, uint32_t, int, long long*);
struct Entry { const char* name; kern_t k; int nops; };

int main() {
  Entry tab[] = {
    {"HFMA2", k_hfma2, 1}, {"HFMA2.SAT", k_hfma2s, 1}, {"HADD2", k_hadd2, 1}, {"HSET2", k_hset2, 1}, {"IDP.4A", k_idp, 1},
    {"IMAD", k_imad, 1}, {"IADD", k_iadd, 1}, {"LOP3", k_lop3, 1}, {"PRMT", k_prmt, 1}, {"VABSDIFF4", k_vabs, 1}, {"FFMA", k_ffma, 1},
    {"HFMA2+IDP", k_hfma2_idp, 2}, {"HFMA2+IMAD", k_hfma2_imad, 2}, {"HFMA2+FFMA", k_hfma2_ffma, 2},
    {"HFMA2+HSET2", k_hfma2_hset2, 2}, {"HFMA2+LOP3", k_hfma2_lop3, 2}, {"HFMA2+HFMA2.SAT", k_hfma2_hfma2s, 2},
    {"IDP+IMAD", k_idp_imad, 2}, {"IDP+LOP3", k_idp_lop3, 2}, {"IDP+FFMA", k_idp_ffma, 2}, {"HSET2+LOP3", k_hset2_lop3, 2},
    {"PRMT+LOP3", k_prmt_lop3, 2}, {"VABSDIFF4+LOP3", k_vabs_lop3, 2},
    {"HFMA2+IDP+LOP3", k_hfma2_idp_lop3, 3}, {"HFMA2+FFMA+LOP3", k_hfma2_ffma_lop3, 3}, {"HFMA2+IMAD+LOP3", k_hfma2_imad_lop3, 3},
    {"HFMA2+HFMA2.SAT+IDP", k_hfma2_hfma2s_idp, 3},
  };
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  const int ctas_per_sm = 4, threads = 128, iters = 4000;
  const int grid = prop.multiProcessorCount * ctas_per_sm;
  uint32_t* out; long long* clk; cudaMalloc(&out, grid * threads * 4); cudaMalloc(&clk, grid * 8);
  long long* h = (long long*)malloc(grid * 8);
  printf("%-24s %10s  (16 warps/SM; 4 = issue limit)\n", "mix", "instr/clk/SM");
  for (auto& e : tab) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    e.k<<<grid, threads>>>(out, 0x3C003C00u, 0x00010001u, iters, clk);
    cudaEventRecord(e0);
    e.k<<<grid, threads>>>(out, 0x3C003C00u, 0x00010001u, iters * 4, clk);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", e.name, cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    // per SM: ctas_per_sm CTAs x 4 warps x iters x 8 x nops warp-instructions, all CTAs run concurrently
    const double instr = (double)ctas_per_sm * 4 * iters * 8 * e.nops;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    printf("%-24s clock64: %7.3f   events@%d MHz: %7.3f\n", e.name, 4.0 * instr / avg, khz / 1000, 4.0 * instr / (ms * 1e-3 * khz * 1e3));
  }
  return 0;
}
