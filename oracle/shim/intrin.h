/* oracle/shim/intrin.h -- TEST INFRASTRUCTURE ONLY.
 * Stand-in for MSVC <intrin.h> so that /root/reference/Amatsukaze/ComputeKernel.cpp
 * (which does `#include <intrin.h>` and calls __cpuid(int[4], int), ComputeKernel.cpp:10,24,35)
 * compiles unmodified with g++.  _xgetbv comes from <immintrin.h> under -mxsave. */
#pragma once
#include <cpuid.h>
#include <immintrin.h>
#ifdef __cpuid
#undef __cpuid
#endif
static inline void __cpuid(int regs[4], int leaf) {
  unsigned a = 0, b = 0, c = 0, d = 0;
  __cpuid_count((unsigned)leaf, 0u, a, b, c, d);
  regs[0] = (int)a; regs[1] = (int)b; regs[2] = (int)c; regs[3] = (int)d;
}
