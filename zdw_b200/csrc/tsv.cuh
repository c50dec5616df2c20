// tsv.cuh -- escape-aware TSV delimiter classification shared by the row indexer and both encode passes.
//
// Rules restated from the reference:
//  * '\n' ends a logical row iff it is preceded by an EVEN number of consecutive backslashes
//    (GetNextRow, getnextrow.cpp:44-53,72-77);
//  * '\t' separates fields iff it is preceded by an EVEN number of consecutive backslashes
//    (get_next_column, ConvertToZDW.cpp:1048-1067).  The backward scans of the reference stop at the
//    row / previous field start, which a maximal-run count reproduces because a delimiter byte itself
//    is never a backslash;
//  * a physical line of length < 2, i.e. a bare "\n" at the start of a logical row, is skipped
//    (getnextrow.cpp:39-43): an unescaped newline directly after another unescaped newline (or at
//    offset 0) ends no row.
#pragma once

#include "common.cuh"

namespace zdwb {

struct ChunkMasks {
  uint32_t tab;   // field separators
  uint32_t term;  // row terminators
  uint32_t skip;  // blank-line newlines (neither delimiter nor data)
};

// parity of the backslash run that ends right before position p (p >= 0)
__device__ __forceinline__ bool odd_backslashes_before(const uint8_t* __restrict__ buf, int64_t p) {
  uint32_t c = 0;
  int64_t q = p - 1;
  while (q >= 0 && __ldg(buf + q) == (uint8_t)'\\') {
    ++c;
    --q;
  }
  return (c & 1u) != 0;
}

__device__ __forceinline__ bool is_unescaped_newline(const uint8_t* __restrict__ buf, int64_t p) {
  return __ldg(buf + p) == (uint8_t)'\n' && !odd_backslashes_before(buf, p);
}

// Classifies the 16-byte chunk at position p0 (buf + p0 is 16-byte aligned; p0 may be negative for the
// first chunk of an unaligned buffer, and p0 + 16 may exceed n).  Only positions inside
// [vlo, vhi) intersected with [0, n) are reported.
__device__ __forceinline__ ChunkMasks classify_chunk(const uint8_t* __restrict__ buf, uint64_t n, int64_t p0, int64_t vlo,
                                                     int64_t vhi) {
  ChunkMasks r{0u, 0u, 0u};
  int64_t a = p0 > vlo ? p0 : vlo;
  if (a < 0) a = 0;
  int64_t b = p0 + 16 < vhi ? p0 + 16 : vhi;
  if (b > (int64_t)n) b = (int64_t)n;
  if (a >= b) return r;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(buf + p0));
  // bits for positions inside the buffer [0, n)
  const int64_t ia = p0 < 0 ? -p0 : 0;
  const int64_t ib = (p0 + 16 > (int64_t)n) ? (int64_t)n - p0 : 16;
  const uint32_t inbuf = ((1u << ib) - 1u) & ~((1u << ia) - 1u);
  uint32_t tab = chunk_mask(v, '\t') & inbuf;
  uint32_t nl = chunk_has(v, '\n') ? (chunk_mask(v, '\n') & inbuf) : 0u;  // newlines are rare: 1 per row
  if ((tab | nl) == 0u) return r;
  uint32_t prevbs = chunk_has(v, '\\') ? (chunk_mask(v, '\\') & inbuf) << 1 : 0u;
  if (((tab | nl) & 1u) && p0 > 0 && __ldg(buf + p0 - 1) == (uint8_t)'\\') prevbs |= 1u;
  uint32_t sus = (tab | nl) & prevbs;
  while (sus) {
    const int i = __ffs(sus) - 1;
    sus &= sus - 1;
    if (odd_backslashes_before(buf, p0 + i)) {
      tab &= ~(1u << i);
      nl &= ~(1u << i);
    }
  }
  uint32_t skip = 0;
  if (nl) {
    uint32_t prevnl = nl << 1;
    if (p0 <= 0) prevnl |= (1u << (-p0));                   // offset 0 behaves like "just after a newline"
    else if (is_unescaped_newline(buf, p0 - 1)) prevnl |= 1u;
    skip = nl & prevnl;
  }
  const uint32_t valid = ((1u << (b - p0)) - 1u) & ~((1u << (a - p0)) - 1u);
  r.tab = tab & valid;
  r.term = nl & ~skip & valid;
  r.skip = skip & valid;
  return r;
}

}  // namespace zdwb
