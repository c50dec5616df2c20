// carry.cuh -- "last explicit value per column" carried across strips of rows (used by the decoder for the
// reference's columnVal[] state, UnconvertFromZDW.cpp:985-986,1349-1453, and by the encoder's row-delta pass 2 for the
// previous-row values of ConvertToZDW.cpp:504-505,532,548,567).
//
// Input: for every strip s and used column u, shas[s*U+u] = the strip holds an explicit value for the column and
// sval[s*U+u] = the last one.  Output: cin[s*U+u] = the value the column has where strip s begins (0 at block start).
// "Select the last explicit value" is associative: strips are grouped into segments of S, reduced, scanned, applied.
#pragma once
#include <stdint.h>

namespace zdwb {
namespace {

// carry propagation over strips, per used column: "select the last explicit value" is associative
__global__ void k_carry_reduce(const unsigned long long* __restrict__ sval, const uint8_t* __restrict__ shas,
                               uint32_t nstrips, uint32_t U, uint32_t S, unsigned long long* __restrict__ seg_val,
                               uint8_t* __restrict__ seg_has) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x, seg = blockIdx.y;
  if (u >= U) return;
  const uint32_t sb = seg * S, se = min(nstrips, sb + S);
  unsigned long long v = 0;
  uint8_t h = 0;
  for (uint32_t s = sb; s < se; ++s) {
    if (shas[(size_t)s * U + u]) {
      v = sval[(size_t)s * U + u];
      h = 1;
    }
  }
  seg_val[(size_t)seg * U + u] = v;
  seg_has[(size_t)seg * U + u] = h;
}

__global__ void k_carry_scan(unsigned long long* __restrict__ seg_val, const uint8_t* __restrict__ seg_has, uint32_t nseg,
                             uint32_t U) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  unsigned long long v = 0;  // columnVal starts at 0 in every block: UnconvertFromZDW.cpp:985-986
  for (uint32_t s = 0; s < nseg; ++s) {
    const unsigned long long mine = seg_val[(size_t)s * U + u];
    const uint8_t h = seg_has[(size_t)s * U + u];
    seg_val[(size_t)s * U + u] = v;  // becomes the carry-in of the segment
    if (h) v = mine;
  }
}

__global__ void k_carry_apply(const unsigned long long* __restrict__ sval, const uint8_t* __restrict__ shas,
                              const unsigned long long* __restrict__ seg_cin, uint32_t nstrips, uint32_t U, uint32_t S,
                              unsigned long long* __restrict__ cin) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x, seg = blockIdx.y;
  if (u >= U) return;
  const uint32_t sb = seg * S, se = min(nstrips, sb + S);
  unsigned long long v = seg_cin[(size_t)seg * U + u];
  for (uint32_t s = sb; s < se; ++s) {
    cin[(size_t)s * U + u] = v;
    if (shas[(size_t)s * U + u]) v = sval[(size_t)s * U + u];
  }
}

}  // namespace
}  // namespace zdwb
