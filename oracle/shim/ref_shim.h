/* oracle/shim/ref_shim.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal environment that lets verbatim line ranges of the reference
 * (AMTLogo.hpp:17-282, LogoScan.hpp:24-45,59-660,734-790) compile under g++ on Linux.
 * Nothing here is algorithmic: it only supplies the Windows/CoreUtils names those ranges use
 * (CoreUtils.hpp:91-95 MemoryChunk, CoreUtils.hpp:257-337 File, tstring/_T, strncpy_s, stdext).
 * The extracted ranges are never committed: oracle/build_ref.sh regenerates them from
 * /root/reference into oracle/_ref/ (git-ignored). */
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <climits>
#include <string>
#include <vector>
#include <memory>
#include <algorithm>
#include <numeric>
#include <functional>
#include <iterator>

typedef std::string tstring;
typedef char tchar;
#ifndef _T
#define _T(x) x
#endif

/* AviSynth plane ids (include/avisynth.h PLANAR_Y/U/V values) */
enum { PLANAR_Y = 1 << 0, PLANAR_U = 1 << 1, PLANAR_V = 1 << 2 };

struct IOException { };

struct MemoryChunk {
  uint8_t* data; size_t length;
  MemoryChunk() : data(nullptr), length(0) {}
  MemoryChunk(uint8_t* d, size_t l) : data(d), length(l) {}
};

class File {
  FILE* fp_;
public:
  File(const tstring& path, const tchar* mode) { fp_ = fopen(path.c_str(), mode); if (!fp_) throw IOException(); }
  ~File() { if (fp_) fclose(fp_); }
  File(const File&) = delete;
  void write(MemoryChunk mc) const { if (mc.length && fwrite(mc.data, mc.length, 1, fp_) != 1) throw IOException(); }
  template <typename T> void writeValue(T v) const { write(MemoryChunk((uint8_t*)&v, sizeof(T))); }
  size_t read(MemoryChunk mc) const { if (!mc.length) return 0; return fread(mc.data, 1, mc.length, fp_); }
  template <typename T> T readValue() const { T v; if (read(MemoryChunk((uint8_t*)&v, sizeof(T))) != sizeof(T)) throw IOException(); return v; }
  void seek(int64_t off, int origin) const { if (fseeko(fp_, (off_t)off, origin) != 0) throw IOException(); }
  /* one text line without its line ending; false at end of file (contract of CoreUtils.hpp:352-375) */
  bool getline(std::string& line) {
    line.clear();
    int c, got = 0;
    while ((c = fgetc(fp_)) != EOF) { got = 1; if (c == '\n') break; line.push_back((char)c); }
    if (!line.empty() && line.back() == '\r') line.pop_back();
    return got != 0;
  }
};

template <size_t N> static inline void strncpy_s(char (&dst)[N], const char* src, size_t count) {
  size_t n = std::min(count, N - 1); strncpy(dst, src, n); dst[n] = 0;
}
template <size_t N> static inline void strcpy_s(char (&dst)[N], const char* src) {
  /* reference copies LOGO_FILE_HEADER_STR (28 bytes incl. embedded NULs) with strcpy_s: copy up to first NUL */
  strncpy(dst, src, N - 1); dst[N - 1] = 0;
}
namespace stdext { template <class P> static inline P checked_array_iterator(P p, size_t) { return p; } }
