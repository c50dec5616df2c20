// decode.cu -- one ZDW block -> TSV rows, entirely on the GPU.
//
// Replaces UnconvertFromZDW_Base::parseBlockHeader (UnconvertFromZDW.cpp:758-1000) and the per-row loop
// UnconvertFromZDW<T>::readNextRow (:1270-1464) with GetWord (:359-371), llutoa/lltoa (:318-356) and
// outputDefault (:1224-1266).
//
// The encoded row stream has no row lengths: a row is ceil(U/8) flag bytes followed by the values of the
// flagged columns, so where row r+1 starts is only known after row r's flags have been read.  The kernels:
//   k_dec_tile_maps    for every tile of the stream and every possible entry offset e in [0, M) (M = max row
//                      bytes): where does a walk that enters the tile at e leave it, and how many rows start
//                      inside?  Row length at a byte position is a popcount over 4 size bit-planes.
//   k_dec_compose      composes those maps group by group (function composition is associative) ...
//   k_dec_descend      ... and hands every tile its true entry offset and first row number.
//   k_dec_row_starts   one warp per tile walks the true chain and records row_off[r].
//   k_dec_strip_summary / k_carry_*   last explicit value per (strip of rows, used column) and its
//                      propagation across strips: the value a column holds where a strip begins.
//   k_dec_format       per strip: scatter explicit values, fill forward down the rows, measure every
//                      field, decoupled look-back for the strip's output offset, write the TSV bytes.
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace zdwb {

namespace {

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr uint64_t LB_AGG = 1ull << 62, LB_PFX = 2ull << 62, LB_MASK = (1ull << 62) - 1ull;

struct DecMeta {
  uint32_t err;            // 1 = dictionary offset out of range (CORRUPTED_DATA_ERROR)
  uint32_t tile_ticket;
  unsigned long long out_bytes;
  uint32_t overflow;       // output buffer estimate too small
  uint32_t pad;
};

struct DecParams {
  const uint8_t* blk;   // device pointer to the block (numRows field)
  uint64_t avail;       // bytes readable from blk
  uint64_t rows_base;   // offset of the first encoded row
  uint64_t dict_base;   // offset of the dictionary origin byte
  uint64_t dict_total;  // dictionary bytes incl. the origin byte
  uint32_t nrows, U, F, M, W;
  const uint8_t* usz;          // [U] value width of every used column
  const unsigned long long* ubase;  // [U]
  const uint8_t* utype;        // [U]
  const uint8_t* lut;          // [F][256] bytes of values selected by a flag byte
  const uint32_t* planes;      // [4][W] bit k of the width of used column u, as flag-word masks
};

__device__ __forceinline__ uint32_t stream_byte(const DecParams& P, uint64_t s) {
  const uint64_t a = P.rows_base + s;
  return a < P.avail ? (uint32_t)__ldg(P.blk + a) : 0u;
}

// ---------------------------------------------------------------------------------------------
// row-boundary discovery
// ---------------------------------------------------------------------------------------------
// maps[tile*M + e] = (rows started << 32) | exit offset into the next tile
__global__ void __launch_bounds__(DEC_THREADS)
    k_dec_tile_maps(const DecParams P, uint64_t win_bytes, uint32_t T, uint64_t* __restrict__ maps) {
  extern __shared__ __align__(16) uint8_t dsm[];
  // layout: next[T] u32 | words[(T + 4W + 8)/4] u32 | planes[4W] u32
  uint32_t* next = reinterpret_cast<uint32_t*>(dsm);
  uint32_t* words = next + T;
  const uint32_t nwords = (T + 4 * P.W + 8) / 4;
  uint32_t* planes = words + nwords;
  const uint64_t t0 = (uint64_t)blockIdx.x * T;
  for (uint32_t i = threadIdx.x; i < nwords; i += DEC_THREADS) {
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint64_t s = t0 + (uint64_t)i * 4 + b;
      const uint32_t v = s < win_bytes ? stream_byte(P, s) : 0u;
      w |= v << (8 * b);
    }
    words[i] = w;
  }
  for (uint32_t i = threadIdx.x; i < 4 * P.W; i += DEC_THREADS) planes[i] = P.planes[i];
  __syncthreads();
  const uint32_t W = P.W;
  for (uint32_t s = threadIdx.x; s < T; s += DEC_THREADS) {
    const uint32_t a = s >> 2, sh = (s & 3u) * 8u;
    uint32_t prev = words[a];
    uint32_t acc = P.F;
    for (uint32_t w = 0; w < W; ++w) {
      const uint32_t nxt = words[a + w + 1];
      const uint32_t word = __funnelshift_r(prev, nxt, sh);
      prev = nxt;
      acc += __popc(word & planes[w]) + 2u * __popc(word & planes[W + w]) + 4u * __popc(word & planes[2 * W + w]) +
             8u * __popc(word & planes[3 * W + w]);
    }
    next[s] = s + acc;
  }
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < P.M; e += DEC_THREADS) {
    uint32_t p = e, c = 0;
    while (p < T) {
      p = next[p];
      ++c;
    }
    maps[(size_t)blockIdx.x * P.M + e] = ((uint64_t)c << 32) | (uint64_t)(p - T);
  }
}

// out[g][e] = in[gG + G-1] o ... o in[gG] (e)
__global__ void __launch_bounds__(DEC_THREADS)
    k_dec_compose(const uint64_t* __restrict__ in, uint32_t n_in, uint32_t M, uint32_t G, uint64_t* __restrict__ out) {
  const uint32_t g = blockIdx.x;
  const uint32_t tb = g * G, te = min(n_in, tb + G);
  for (uint32_t e = threadIdx.x; e < M; e += DEC_THREADS) {
    uint32_t x = e;
    uint32_t c = 0;
    for (uint32_t t = tb; t < te; ++t) {
      const uint64_t v = in[(size_t)t * M + x];
      x = (uint32_t)v;
      c += (uint32_t)(v >> 32);
    }
    out[(size_t)g * M + e] = ((uint64_t)c << 32) | x;
  }
}

// ent_child[t] = state on entering child t: (rows before << 32) | entry offset
__global__ void k_dec_descend(const uint64_t* __restrict__ in, uint32_t n_in, uint32_t M, uint32_t G,
                              const uint64_t* __restrict__ ent_parent, uint32_t n_parent, uint64_t* __restrict__ ent_child) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_parent) return;
  uint64_t st = ent_parent[g];
  const uint32_t tb = g * G, te = min(n_in, tb + G);
  for (uint32_t t = tb; t < te; ++t) {
    ent_child[t] = st;
    const uint64_t v = in[(size_t)t * M + (uint32_t)st];
    st = (((st >> 32) + (v >> 32)) << 32) | (uint32_t)v;
  }
}

// warp-cooperative row length at stream offset p
__device__ __forceinline__ uint32_t warp_row_len(const DecParams& P, uint64_t p) {
  const unsigned lane = lane_id();
  uint32_t acc = 0;
  for (uint32_t w0 = 0; w0 < P.W; w0 += 32) {
    const uint32_t w = w0 + lane;
    if (w < P.W) {
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) word |= stream_byte(P, p + 4ull * w + b) << (8 * b);
      acc += __popc(word & P.planes[w]) + 2u * __popc(word & P.planes[P.W + w]) +
             4u * __popc(word & P.planes[2 * P.W + w]) + 8u * __popc(word & P.planes[3 * P.W + w]);
    }
  }
  acc = __reduce_add_sync(0xffffffffu, acc);
  return P.F + acc;
}

__global__ void __launch_bounds__(DEC_THREADS)
    k_dec_row_starts(const DecParams P, uint32_t T, uint32_t ntiles, const uint64_t* __restrict__ ent,
                     uint32_t* __restrict__ row_off) {
  const uint32_t tile = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
  if (tile >= ntiles) return;
  const uint64_t st = ent[tile];
  uint64_t r = st >> 32;
  uint64_t p = (uint64_t)tile * T + (uint32_t)st;
  const uint64_t pend = (uint64_t)(tile + 1) * T;
  while (p < pend && r <= P.nrows) {
    if (lane_id() == 0) row_off[r] = (uint32_t)p;
    if (r == P.nrows) break;
    p += warp_row_len(P, p);
    ++r;
  }
}

// ---------------------------------------------------------------------------------------------
// row parser: one warp walks the flag bytes of a row and calls cb(u, value offset inside the row)
// for every flagged used column, lanes working on different flag bytes.
// ---------------------------------------------------------------------------------------------
template <class CB>
__device__ __forceinline__ void warp_parse_row(const DecParams& P, const uint8_t* __restrict__ lut,
                                               const uint8_t* __restrict__ rp, CB&& cb) {
  const unsigned lane = lane_id();
  uint32_t run = P.F;
  for (uint32_t j0 = 0; j0 < P.F; j0 += 32) {
    const uint32_t j = j0 + lane;
    uint32_t fb = 0, bl = 0;
    if (j < P.F) {
      fb = __ldg(rp + j);
      bl = lut[j * 256 + fb];
    }
    uint32_t inc = bl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (unsigned)o) inc += t;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t voff = run + inc - bl;
    while (fb) {
      const int b = __ffs(fb) - 1;
      fb &= fb - 1;
      const uint32_t u = j * 8 + b;
      if (u < P.U) {
        cb(u, voff);
        voff += P.usz[u];
      }
    }
    run += tot;
  }
}

__device__ __forceinline__ unsigned long long load_le(const uint8_t* __restrict__ p, uint32_t sz) {
  unsigned long long v = 0;
  for (uint32_t b = 0; b < sz; ++b) v |= (unsigned long long)__ldg(p + b) << (8 * b);
  return v;
}

__device__ __forceinline__ const uint8_t* stage_lut(const DecParams& P, uint8_t* s_lut, bool lut_in_smem) {
  if (!lut_in_smem) return P.lut;
  for (uint32_t i = threadIdx.x; i < P.F * 256; i += blockDim.x) s_lut[i] = P.lut[i];
  return s_lut;
}

// last explicit value of every used column inside a strip of R rows
__global__ void __launch_bounds__(DEC_THREADS)
    k_dec_strip_summary(const DecParams P, const uint32_t* __restrict__ row_off, uint32_t R, int lut_in_smem,
                        unsigned long long* __restrict__ sval, uint8_t* __restrict__ shas) {
  extern __shared__ __align__(16) uint8_t dsm[];
  int32_t* last_row = reinterpret_cast<int32_t*>(dsm);
  uint8_t* s_lut = dsm + (size_t)P.U * 4;
  const uint8_t* lut = stage_lut(P, s_lut, lut_in_smem != 0);
  for (uint32_t u = threadIdx.x; u < P.U; u += DEC_THREADS) last_row[u] = -1;
  __syncthreads();
  const uint32_t r0 = blockIdx.x * R, r1 = min(P.nrows, r0 + R);
  const unsigned warp = threadIdx.x >> 5;
  const uint8_t* rows = P.blk + P.rows_base;
  for (uint32_t r = r0 + warp; r < r1; r += DEC_WARPS) {
    const int32_t rl = (int32_t)(r - r0);
    warp_parse_row(P, lut, rows + row_off[r], [&](uint32_t u, uint32_t) { atomicMax(&last_row[u], rl); });
  }
  __syncthreads();
  for (uint32_t r = r0 + warp; r < r1; r += DEC_WARPS) {
    const int32_t rl = (int32_t)(r - r0);
    const uint8_t* rp = rows + row_off[r];
    warp_parse_row(P, lut, rp, [&](uint32_t u, uint32_t voff) {
      if (last_row[u] == rl) sval[(size_t)blockIdx.x * P.U + u] = load_le(rp + voff, P.usz[u]);
    });
  }
  for (uint32_t u = threadIdx.x; u < P.U; u += DEC_THREADS) shas[(size_t)blockIdx.x * P.U + u] = last_row[u] >= 0 ? 1 : 0;
}

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

// ---------------------------------------------------------------------------------------------
// formatting
// ---------------------------------------------------------------------------------------------
struct FmtTables {
  uint32_t n_items;             // dynamic items (used columns that are output), in output order
  const uint32_t* item_u;       // [n_items] used-column index
  const uint32_t* seg_off;      // [n_items + 1] static segment i precedes item i; the last one ends the row
  const uint32_t* seg_len;      // [n_items + 1]
  const uint32_t* seg_cum;      // [n_items + 1] static bytes before segment i
  const uint8_t* blob;          // static bytes (separators, defaults of unused columns, terminator)
  uint32_t static_total;
};

// Length of the text of used column u holding stored value v; *dict_ptr gets the dictionary string for
// text-like columns.  Mirrors the switch in readNextRow (UnconvertFromZDW.cpp:1349-1453).
__device__ __forceinline__ uint32_t value_text(const DecParams& P, uint32_t u, unsigned long long v, uint8_t* tmp24,
                                               const uint8_t** src, DecMeta* meta) {
  const uint8_t t = P.utype[u];
  if (is_text_like(t)) {
    if (v == 0) {
      if (t == ZDWB_DECIMAL) {  // outputDefault(DECIMAL)
        *src = reinterpret_cast<const uint8_t*>("0.000000000000");
        return 14;
      }
      *src = tmp24;
      return 0;
    }
    const uint32_t index = (uint32_t)(v + P.ubase[u]);  // ULONG index: :1363
    if ((uint64_t)index > P.dict_total) {                // :1364 (the reference allows index == dictionarySize)
      meta->err = 1;
      *src = tmp24;
      return 0;
    }
    const uint8_t* s = P.blk + P.dict_base + index;
    const uint64_t room = P.dict_total > index ? P.dict_total - index : 0;
    uint32_t l = 0;
    while (l < room && __ldg(s + l) != 0) ++l;
    *src = s;
    return l;
  }
  if (t == ZDWB_CHAR) {  // :1396-1420
    *src = tmp24;
    if (v == 0) return 0;
    const unsigned long long tu = v + P.ubase[u];
    tmp24[0] = (uint8_t)tu;
    if (tmp24[0] != (uint8_t)'\\') return tmp24[0] ? 1u : 0u;
    tmp24[1] = (uint8_t)(tu >> 8);
    return 2;
  }
  const unsigned long long full = v ? v + P.ubase[u] : 0ull;
  uint32_t l;
  if (is_signed_int_type(t)) l = fmt_i64((int64_t)full, tmp24 + 24);
  else l = fmt_u64(full, tmp24 + 24);
  *src = tmp24 + 24 - l;
  return l;
}

__global__ void __launch_bounds__(DEC_THREADS)
    k_dec_format(const DecParams P, const FmtTables FT, const uint32_t* __restrict__ row_off, uint32_t R, int lut_in_smem,
                 const unsigned long long* __restrict__ cin, uint64_t* __restrict__ strip_status, uint8_t* __restrict__ out,
                 uint64_t out_cap, uint64_t* __restrict__ out_row_off, DecMeta* __restrict__ meta) {
  extern __shared__ __align__(16) uint8_t dsm[];
  __shared__ uint32_t s_strip;
  __shared__ unsigned long long s_base;
  // layout: val[R*U] u64 | ilen[R*NI] u32 | ioff[R*NI] u32 | rowoff[R+1] u32 | sflag[R*F] u8 | lut[F*256]
  const uint32_t U = P.U, F = P.F, NI = FT.n_items;
  unsigned long long* val = reinterpret_cast<unsigned long long*>(dsm);
  uint32_t* ilen = reinterpret_cast<uint32_t*>(val + (size_t)R * U);
  uint32_t* ioff = ilen + (size_t)R * NI;
  uint32_t* rowoff = ioff + (size_t)R * NI;
  uint8_t* sflag = reinterpret_cast<uint8_t*>(rowoff + R + 1);
  uint8_t* s_lut = sflag + (size_t)R * F;
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) s_strip = atomicAdd(&meta->tile_ticket, 1u);
  const uint8_t* lut = stage_lut(P, s_lut, lut_in_smem != 0);
  __syncthreads();
  const uint32_t strip = s_strip;
  const uint32_t r0 = strip * R, r1 = min(P.nrows, r0 + R), Rn = r1 - r0;
  const uint8_t* rows = P.blk + P.rows_base;

  // ---- explicit values of this strip
  for (uint32_t j = warp; j < Rn; j += DEC_WARPS) {
    const uint8_t* rp = rows + row_off[r0 + j];
    for (uint32_t k = lane; k < F; k += 32) sflag[j * F + k] = __ldg(rp + k);
    warp_parse_row(P, lut, rp, [&](uint32_t u, uint32_t voff) { val[(size_t)j * U + u] = load_le(rp + voff, P.usz[u]); });
  }
  __syncthreads();
  // ---- fill forward: a column keeps its value until a row flags it again (:1339-1345)
  for (uint32_t u = tid; u < U; u += DEC_THREADS) {
    unsigned long long v = cin[(size_t)strip * U + u];
    const uint32_t fb = u >> 3, bit = u & 7u;
    for (uint32_t j = 0; j < Rn; ++j) {
      if ((sflag[j * F + fb] >> bit) & 1u) v = val[(size_t)j * U + u];
      else val[(size_t)j * U + u] = v;
    }
  }
  __syncthreads();
  // ---- field lengths and their prefix inside each row
  uint8_t tmp[24];
  for (uint32_t j = warp; j < Rn; j += DEC_WARPS) {
    uint32_t run = 0;
    for (uint32_t i0 = 0; i0 < NI; i0 += 32) {
      const uint32_t i = i0 + lane;
      uint32_t l = 0;
      if (i < NI) {
        const uint32_t u = FT.item_u[i];
        const uint8_t* src;
        l = value_text(P, u, val[(size_t)j * U + u], tmp, &src, meta);
      }
      uint32_t inc = l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
      }
      if (i < NI) {
        ilen[(size_t)j * NI + i] = l;
        ioff[(size_t)j * NI + i] = run + inc - l;
      }
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) rowoff[j] = FT.static_total + run;
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t run = 0;
    for (uint32_t j0 = 0; j0 < Rn; j0 += 32) {
      const uint32_t j = j0 + lane;
      const uint32_t v = j < Rn ? rowoff[j] : 0u;
      uint32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
      }
      if (j < Rn) rowoff[j] = run + inc - v;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) rowoff[Rn] = run;
  }
  __syncthreads();
  // ---- decoupled look-back over strips
  if (tid == 0) {
    const uint64_t total = rowoff[Rn];
    uint64_t run = 0;
    if (strip == 0) {
      st_release_u64(&strip_status[0], LB_PFX | total);
    } else {
      st_release_u64(&strip_status[strip], LB_AGG | total);
      int64_t q = (int64_t)strip - 1;
      for (;;) {
        uint64_t s;
        do {
          s = ld_acquire_u64(&strip_status[q]);
        } while ((s >> 62) == 0ull);
        run += s & LB_MASK;
        if ((s >> 62) == 2ull) break;
        --q;
      }
      st_release_u64(&strip_status[strip], LB_PFX | (run + total));
    }
    s_base = run;
    if (r1 == P.nrows) {
      meta->out_bytes = run + total;
      if (out_row_off) out_row_off[P.nrows] = run + total;
    }
  }
  __syncthreads();
  const uint64_t base = s_base;
  if (out_row_off)
    for (uint32_t j = tid; j < Rn; j += DEC_THREADS) out_row_off[r0 + j] = base + rowoff[j];
  if (base + rowoff[Rn] > out_cap) {  // the host's size estimate was too small: it reruns with the exact size
    if (tid == 0) meta->overflow = 1;
    return;
  }

  // ---- write: static segment i, then the text of item i; the last segment closes the row
  for (uint32_t j = warp; j < Rn; j += DEC_WARPS) {
    uint8_t* orow = out + base + rowoff[j];
    const uint32_t dyn_total = (j + 1 <= Rn ? rowoff[j + 1] - rowoff[j] : 0u) - FT.static_total;
    for (uint32_t i0 = 0; i0 <= NI; i0 += 32) {
      const uint32_t i = i0 + lane;
      if (i > NI) continue;
      const uint32_t doff = i < NI ? ioff[(size_t)j * NI + i] : dyn_total;
      uint8_t* d = orow + FT.seg_cum[i] + doff;
      const uint8_t* sb = FT.blob + FT.seg_off[i];
      const uint32_t sl = FT.seg_len[i];
      for (uint32_t k = 0; k < sl; ++k) d[k] = __ldg(sb + k);
      if (i < NI) {
        d += sl;
        const uint32_t u = FT.item_u[i];
        const uint8_t* src;
        const uint32_t l = value_text(P, u, val[(size_t)j * U + u], tmp, &src, meta);
        for (uint32_t k = 0; k < l; ++k) d[k] = src[k];
      }
    }
  }
}

struct HostBlockHeader {
  uint32_t nrows = 0, line_len = 0;
  uint8_t last = 0, idx_size = 0;
  uint64_t dict_total = 0;  // 0 = empty dictionary
  uint64_t dict_base = 0, stats_base = 0;
};

// reads `len` bytes at block offset `off` into dst (host), from host or device memory
int fetch(Ctx* ctx, const uint8_t* blk, bool on_device, uint64_t avail, uint64_t off, void* dst, size_t len) {
  if (off + len > avail) {
    ctx->err = "decode: block runs past the end of the data";
    return ZDWB_ERR_TRUNCATED;
  }
  if (!on_device) {
    memcpy(dst, blk + off, len);
    return ZDWB_OK;
  }
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(dst, blk + off, len, cudaMemcpyDeviceToHost, ctx->stream));
  ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return ZDWB_OK;
}

void append_default(std::string& blob, uint8_t type) {  // outputDefault, UnconvertFromZDW.cpp:1224-1266
  if (is_int_type(type)) blob += "0";
  else if (type == ZDWB_DECIMAL) blob += "0.000000000000";
}

template <typename T>
int upload(Ctx* ctx, DevBuf& d, const std::vector<T>& v) {
  ZDWB_TRY(d.alloc(ctx, v.size() * sizeof(T) + 16));
  if (!v.empty())
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(d.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return ZDWB_OK;
}

}  // namespace

int decode_block_impl(Ctx* ctx, const zdwb_schema* schema, const void* zdw, size_t avail, const zdwb_decode_opts* opts,
                      zdwb_rows_out* out) {
  memset(out, 0, sizeof(*out));
  cudaStream_t st = ctx->stream;
  const uint32_t nc = schema->ncols;
  if (nc == 0 || !schema->types) {
    ctx->err = "decode: schema has no columns";
    return ZDWB_ERR_BAD_ARG;
  }
  for (uint32_t c = 0; c < nc; ++c) {
    if (!is_known_type(schema->types[c])) {
      ctx->err = "decode: unsupported column type id " + std::to_string((int)schema->types[c]) + " (v1-v8 files are out of scope)";
      return ZDWB_ERR_UNSUPPORTED;
    }
  }
  const bool in_dev = opts->input_on_device != 0;
  const uint8_t* src = static_cast<const uint8_t*>(zdw);

  // ---- block header (readLineLength :782-810, readDictionary :812-846, readColumnFieldStats :966-1000)
  HostBlockHeader H;
  uint8_t head[16];
  ZDWB_TRY(fetch(ctx, src, in_dev, avail, 0, head, 10));
  memcpy(&H.nrows, head, 4);
  memcpy(&H.line_len, head + 4, 4);
  H.last = head[8];
  H.idx_size = head[9];
  if (H.idx_size > 4) {
    ctx->err = "decode: dictionary index size > 4";
    return ZDWB_ERR_CORRUPT;
  }
  if (H.idx_size) {
    uint8_t tb[4] = {0, 0, 0, 0};
    ZDWB_TRY(fetch(ctx, src, in_dev, avail, 10, tb, H.idx_size));
    uint32_t v;
    memcpy(&v, tb, 4);
    H.dict_total = v;
    H.dict_base = 10 + H.idx_size;
    H.stats_base = H.dict_base + H.dict_total;
  } else {
    H.dict_total = 0;
    H.dict_base = 10;
    H.stats_base = 10;
  }
  std::vector<uint8_t> csize(nc);
  ZDWB_TRY(fetch(ctx, src, in_dev, avail, H.stats_base, csize.data(), nc));
  std::vector<uint32_t> used_cols;
  for (uint32_t c = 0; c < nc; ++c)
    if (csize[c]) used_cols.push_back(c);
  const uint32_t U = (uint32_t)used_cols.size();
  std::vector<unsigned long long> ubase(U ? U : 1);
  if (U) ZDWB_TRY(fetch(ctx, src, in_dev, avail, H.stats_base + nc, ubase.data(), (size_t)U * 8));
  const uint64_t rows_base = H.stats_base + nc + 8ull * U;
  const uint32_t F = (U + 7) / 8, W = (F + 3) / 4;
  std::vector<uint8_t> usz(U ? U : 1), utype(U ? U : 1);
  uint32_t M = F;
  for (uint32_t u = 0; u < U; ++u) {
    usz[u] = csize[used_cols[u]];
    if (usz[u] > 8) {
      ctx->err = "decode: column size > 8";
      return ZDWB_ERR_CORRUPT;
    }
    utype[u] = schema->types[used_cols[u]];
    M += usz[u];
  }
  out->nrows = H.nrows;
  out->line_length = H.line_len;
  out->is_last = H.last;
  out->dict_bytes = H.dict_total;
  out->ncols_used = U;
  const uint32_t nrows = H.nrows;

  // zero-byte rows at the very end of the file are never read by the reference (input->eof(), :1577)
  if (U == 0 && nrows > 0 && opts->at_end_of_file && rows_base >= avail) {
    ctx->err = "Rows unpacked (0) does not match expected (" + std::to_string(nrows) + ")";
    return ZDWB_ERR_ROW_COUNT;
  }

  // ---- residency: the row stream cannot be longer than nrows * M
  const uint64_t win = std::min<uint64_t>(avail > rows_base ? avail - rows_base : 0, (uint64_t)nrows * M);
  if (win >= 0xfffffff0ull) {
    ctx->err = "decode: a block's row stream must be smaller than 4 GiB";
    return ZDWB_ERR_UNSUPPORTED;
  }
  DevBuf blk_dev;
  const uint8_t* blk;
  if (in_dev) {
    blk = src;
  } else {
    const uint64_t need = rows_base + win;
    ZDWB_TRY(blk_dev.alloc(ctx, need + 64));
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(blk_dev.p, src, need, cudaMemcpyHostToDevice, st));
    blk = blk_dev.as<uint8_t>();
  }
  const uint64_t dev_avail = in_dev ? avail : rows_base + win;

  // ---- tables
  std::vector<uint8_t> lut((size_t)(F ? F : 1) * 256, 0);
  std::vector<uint32_t> planes((size_t)4 * (W ? W : 1), 0);
  for (uint32_t u = 0; u < U; ++u) {
    const uint32_t j = u / 8, b = u % 8;
    for (uint32_t v = 0; v < 256; ++v)
      if (v & (1u << b)) lut[(size_t)j * 256 + v] += usz[u];
    for (uint32_t k = 0; k < 4; ++k)
      if (usz[u] & (1u << k)) planes[(size_t)k * W + u / 32] |= 1u << (u % 32);
  }
  DevBuf d_usz, d_ubase, d_utype, d_lut, d_planes, d_meta;
  ZDWB_TRY(upload(ctx, d_usz, usz));
  ZDWB_TRY(upload(ctx, d_ubase, ubase));
  ZDWB_TRY(upload(ctx, d_utype, utype));
  ZDWB_TRY(upload(ctx, d_lut, lut));
  ZDWB_TRY(upload(ctx, d_planes, planes));
  ZDWB_TRY(d_meta.alloc(ctx, sizeof(DecMeta)));
  ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(d_meta.p, 0, sizeof(DecMeta), st));
  DecMeta* meta = d_meta.as<DecMeta>();

  DecParams P;
  P.blk = blk;
  P.avail = dev_avail;
  P.rows_base = rows_base;
  P.dict_base = H.dict_base;
  P.dict_total = H.dict_total;
  P.nrows = nrows;
  P.U = U;
  P.F = F;
  P.M = M;
  P.W = W;
  P.usz = d_usz.as<uint8_t>();
  P.ubase = d_ubase.as<unsigned long long>();
  P.utype = d_utype.as<uint8_t>();
  P.lut = d_lut.as<uint8_t>();
  P.planes = d_planes.as<uint32_t>();

  // ---- output plan: static segments and dynamic items in output order
  const uint8_t sep = opts->separator;
  const uint8_t term = sep == '\t' ? (uint8_t)'\n' : (uint8_t)0;
  std::vector<int32_t> used_idx(nc, -1);
  for (uint32_t u = 0; u < U; ++u) used_idx[used_cols[u]] = (int32_t)u;
  std::vector<int32_t> pos_src;  // output position -> file column, -1 = blank
  if (opts->out_col) {
    pos_src.assign(opts->n_out, -1);
    for (uint32_t c = 0; c < nc; ++c) {
      const int32_t k = opts->out_col[c];
      if (k < 0) continue;
      if ((uint32_t)k >= opts->n_out) {
        ctx->err = "decode: out_col position out of range";
        return ZDWB_ERR_BAD_ARG;
      }
      pos_src[k] = (int32_t)c;
    }
  } else {
    pos_src.resize(nc);
    for (uint32_t c = 0; c < nc; ++c) pos_src[c] = (int32_t)c;
  }
  std::string blob;
  std::vector<uint32_t> seg_off, seg_len, seg_cum, item_u;
  {
    size_t seg_start = 0;
    uint32_t cum = 0;
    for (size_t k = 0; k < pos_src.size(); ++k) {
      if (k) blob.push_back((char)sep);
      const int32_t c = pos_src[k];
      if (c < 0) continue;
      if (used_idx[c] < 0) {
        append_default(blob, schema->types[c]);
      } else {
        seg_off.push_back((uint32_t)seg_start);
        seg_len.push_back((uint32_t)(blob.size() - seg_start));
        seg_cum.push_back(cum);
        cum += (uint32_t)(blob.size() - seg_start);
        item_u.push_back((uint32_t)used_idx[c]);
        seg_start = blob.size();
      }
    }
    blob.push_back((char)term);
    seg_off.push_back((uint32_t)seg_start);
    seg_len.push_back((uint32_t)(blob.size() - seg_start));
    seg_cum.push_back(cum);
  }
  const uint32_t NI = (uint32_t)item_u.size();
  std::vector<uint8_t> blobv(blob.begin(), blob.end());
  DevBuf d_item_u, d_seg_off, d_seg_len, d_seg_cum, d_blob;
  ZDWB_TRY(upload(ctx, d_item_u, item_u));
  ZDWB_TRY(upload(ctx, d_seg_off, seg_off));
  ZDWB_TRY(upload(ctx, d_seg_len, seg_len));
  ZDWB_TRY(upload(ctx, d_seg_cum, seg_cum));
  ZDWB_TRY(upload(ctx, d_blob, blobv));
  FmtTables FT;
  FT.n_items = NI;
  FT.item_u = d_item_u.as<uint32_t>();
  FT.seg_off = d_seg_off.as<uint32_t>();
  FT.seg_len = d_seg_len.as<uint32_t>();
  FT.seg_cum = d_seg_cum.as<uint32_t>();
  FT.blob = d_blob.as<uint8_t>();
  FT.static_total = (uint32_t)blob.size();

  if (nrows == 0) {
    out->consumed = rows_base;
    return ZDWB_OK;
  }

  // ---- row starts
  DevBuf row_off;
  ZDWB_TRY(row_off.alloc(ctx, ((size_t)nrows + 1) * 4));
  uint64_t consumed_stream = 0;
  if (U == 0) {
    ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(row_off.p, 0, ((size_t)nrows + 1) * 4, st));
  } else {
    ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(row_off.p, 0xff, ((size_t)nrows + 1) * 4, st));
    uint32_t T = (uint32_t)std::max<long long>(256, std::min<long long>(ctx->dec_tile_bytes, 32768));
    T &= ~3u;
    // every row occupies at least F bytes, so tiles beyond nrows*M never matter; +1 tile for the end mark
    const uint32_t ntiles = (uint32_t)(win / T) + 1;
    DevBuf maps0;
    ZDWB_TRY(maps0.alloc(ctx, (size_t)ntiles * M * 8));
    const size_t smem_maps = (size_t)T * 4 + ((T + 4 * W + 8) / 4) * 4 + (size_t)4 * W * 4 + 16;
    if (smem_maps > 200 * 1024) {
      ctx->err = "decode: too many used columns for the row-boundary kernel";
      return ZDWB_ERR_UNSUPPORTED;
    }
    ZDWB_CUDA_TRY(ctx, cudaFuncSetAttribute(k_dec_tile_maps, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
      KernelScope _ks(ctx, "k_dec_tile_maps");
      k_dec_tile_maps<<<ntiles, DEC_THREADS, smem_maps, st>>>(P, win, T, maps0.as<uint64_t>());
    }
    ZDWB_LAUNCH_CHECK(ctx);
    // up-sweep: compose groups of G maps until one remains
    const uint32_t G = 32;
    std::vector<DevBuf*> levels;  // levels[0] = tile maps
    std::vector<uint32_t> counts;
    levels.push_back(&maps0);
    counts.push_back(ntiles);
    std::vector<DevBuf*> owned;
    auto cleanup = [&]() {
      for (DevBuf* b : owned) delete b;
    };
    while (counts.back() > 1) {
      const uint32_t n_in = counts.back();
      const uint32_t n_out = (n_in + G - 1) / G;
      DevBuf* nb = new DevBuf();
      owned.push_back(nb);
      int rc = nb->alloc(ctx, (size_t)n_out * M * 8);
      if (rc) {
        cleanup();
        return rc;
      }
      {
        KernelScope _ks(ctx, "k_dec_compose");
        k_dec_compose<<<n_out, DEC_THREADS, 0, st>>>(levels.back()->as<uint64_t>(), n_in, M, G, nb->as<uint64_t>());
      }
      ctx->launches++;
      levels.push_back(nb);
      counts.push_back(n_out);
    }
    // down-sweep: entry state of every node, from the single root (rows = 0, offset = 0) to the tiles
    DevBuf* ent_parent = new DevBuf();
    owned.push_back(ent_parent);
    {
      int rc = ent_parent->alloc(ctx, 8);
      if (rc) {
        cleanup();
        return rc;
      }
      cudaMemsetAsync(ent_parent->p, 0, 8, st);
    }
    for (int lv = (int)levels.size() - 2; lv >= 0; --lv) {
      const uint32_t n_in = counts[lv], n_parent = counts[lv + 1];
      DevBuf* ec = new DevBuf();
      owned.push_back(ec);
      int rc = ec->alloc(ctx, (size_t)n_in * 8);
      if (rc) {
        cleanup();
        return rc;
      }
      {
        KernelScope _ks(ctx, "k_dec_descend");
        k_dec_descend<<<(n_parent + 127) / 128, 128, 0, st>>>(levels[lv]->as<uint64_t>(), n_in, M, G, ent_parent->as<uint64_t>(),
                                                          n_parent, ec->as<uint64_t>());
      }
      ctx->launches++;
      ent_parent = ec;
    }
    // with a single tile the root entry is the tile entry
    {
      KernelScope _ks(ctx, "k_dec_row_starts");
      k_dec_row_starts<<<(ntiles + DEC_WARPS - 1) / DEC_WARPS, DEC_THREADS, 0, st>>>(P, T, ntiles, ent_parent->as<uint64_t>(),
                                                                                  row_off.as<uint32_t>());
    }
    ctx->launches++;
    cudaError_t le = cudaGetLastError();
    uint32_t h_end = 0;
    cudaError_t ce = cudaMemcpyAsync(ctx->meta_host, row_off.as<uint32_t>() + nrows, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t se = cudaStreamSynchronize(st);
    cleanup();
    ZDWB_CUDA_TRY(ctx, le);
    ZDWB_CUDA_TRY(ctx, ce);
    ZDWB_CUDA_TRY(ctx, se);
    h_end = *static_cast<uint32_t*>(ctx->meta_host);
    if (h_end == 0xffffffffu || (uint64_t)h_end > win) {
      // the data ends before `nrows` rows do
      ctx->err = "decode: block truncated (fewer rows than the header promises)";
      return opts->at_end_of_file ? ZDWB_ERR_ROW_COUNT : ZDWB_ERR_TRUNCATED;
    }
    consumed_stream = h_end;
  }
  out->consumed = rows_base + consumed_stream;

  // ---- strips
  const int lut_in_smem = (size_t)F * 256 <= 48 * 1024 ? 1 : 0;
  const size_t lut_smem = lut_in_smem ? (size_t)F * 256 : 0;
  // rows per strip: shared memory of k_dec_format is R * (8U + 8NI + F) + lut
  const size_t per_row = (size_t)8 * U + (size_t)8 * NI + F + 4;
  const size_t budget = 64 * 1024;
  uint32_t R = (uint32_t)std::max<size_t>(1, std::min<size_t>(budget / std::max<size_t>(per_row, 1), 512));
  {
    // keep the TSV bytes per strip moderate (about 32 KiB)
    const uint64_t approx_row = (uint64_t)FT.static_total + (uint64_t)NI * 8;
    const uint32_t by_bytes = (uint32_t)std::max<uint64_t>(8, 32768 / std::max<uint64_t>(approx_row, 1));
    R = std::min(R, std::max(by_bytes, 8u));
    if (R == 0) R = 1;
  }
  const size_t smem_fmt = (size_t)R * per_row + 64 + lut_smem + (size_t)(R + 1) * 4;
  if (smem_fmt > 200 * 1024) {
    ctx->err = "decode: too many used columns for the format kernel";
    return ZDWB_ERR_UNSUPPORTED;
  }
  const uint32_t nstrips = (nrows + R - 1) / R;

  DevBuf cin;
  ZDWB_TRY(cin.alloc(ctx, (size_t)nstrips * std::max(U, 1u) * 8));
  if (U) {
    DevBuf sval, shas, seg_val, seg_has;
    ZDWB_TRY(sval.alloc(ctx, (size_t)nstrips * U * 8));
    ZDWB_TRY(shas.alloc(ctx, (size_t)nstrips * U));
    const size_t smem_sum = (size_t)U * 4 + lut_smem + 16;
    ZDWB_CUDA_TRY(ctx, cudaFuncSetAttribute(k_dec_strip_summary, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    {
      KernelScope _ks(ctx, "k_dec_strip_summary");
      k_dec_strip_summary<<<nstrips, DEC_THREADS, smem_sum, st>>>(P, row_off.as<uint32_t>(), R, lut_in_smem,
                                                               sval.as<unsigned long long>(), shas.as<uint8_t>());
    }
    ZDWB_LAUNCH_CHECK(ctx);
    const uint32_t S = (nstrips + 255) / 256;
    const uint32_t nseg = (nstrips + S - 1) / S;
    ZDWB_TRY(seg_val.alloc(ctx, (size_t)nseg * U * 8));
    ZDWB_TRY(seg_has.alloc(ctx, (size_t)nseg * U));
    dim3 g2((U + 127) / 128, nseg);
    {
      KernelScope _ks(ctx, "k_carry_reduce");
      k_carry_reduce<<<g2, 128, 0, st>>>(sval.as<unsigned long long>(), shas.as<uint8_t>(), nstrips, U, S,
                                       seg_val.as<unsigned long long>(), seg_has.as<uint8_t>());
    }
    ZDWB_LAUNCH_CHECK(ctx);
    {
      KernelScope _ks(ctx, "k_carry_scan");
      k_carry_scan<<<(U + 127) / 128, 128, 0, st>>>(seg_val.as<unsigned long long>(), seg_has.as<uint8_t>(), nseg, U);
    }
    ZDWB_LAUNCH_CHECK(ctx);
    {
      KernelScope _ks(ctx, "k_carry_apply");
      k_carry_apply<<<g2, 128, 0, st>>>(sval.as<unsigned long long>(), shas.as<uint8_t>(), seg_val.as<unsigned long long>(), nstrips,
                                      U, S, cin.as<unsigned long long>());
    }
    ZDWB_LAUNCH_CHECK(ctx);
  }

  // ---- output buffers.  The exact TSV size is only known once every field has been measured, which the
  // format kernel does anyway; so size the buffer by an estimate, let strips that would not fit skip their
  // writes (the look-back still yields the exact total) and rerun once with the exact size if needed.
  uint64_t est_row = FT.static_total;
  for (uint32_t i = 0; i < NI; ++i) est_row += is_text_like(utype[item_u[i]]) ? 16 : 8;
  if (ctx->last_out_per_row > est_row) est_row = ctx->last_out_per_row;
  uint64_t out_cap = (uint64_t)nrows * est_row + 4096;
  if (ctx->out_dev2) {
    cudaFreeAsync(ctx->out_dev2, st);
    ctx->out_dev2 = nullptr;
  }
  if (opts->want_row_offsets) {
    DevBuf ro;
    ZDWB_TRY(ro.alloc(ctx, ((size_t)nrows + 1) * 8));
    ctx->out_dev2 = ro.detach();
  }
  DevBuf status;
  ZDWB_TRY(status.alloc(ctx, (size_t)nstrips * 8));
  ZDWB_CUDA_TRY(ctx, cudaFuncSetAttribute(k_dec_format, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  DecMeta* hm = static_cast<DecMeta*>(ctx->meta_host);
  for (int attempt = 0;; ++attempt) {
    if (ctx->out_dev) {
      cudaFreeAsync(ctx->out_dev, st);
      ctx->out_dev = nullptr;
    }
    {
      DevBuf ob;
      ZDWB_TRY(ob.alloc(ctx, out_cap));
      ctx->out_dev = ob.detach();
    }
    ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(status.p, 0, (size_t)nstrips * 8, st));
    ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(d_meta.p, 0, sizeof(DecMeta), st));
    {
      KernelScope _ks(ctx, "k_dec_format");
      k_dec_format<<<nstrips, DEC_THREADS, smem_fmt, st>>>(P, FT, row_off.as<uint32_t>(), R, lut_in_smem,
                                                        cin.as<unsigned long long>(), status.as<uint64_t>(),
                                                        static_cast<uint8_t*>(ctx->out_dev), out_cap,
                                                        static_cast<uint64_t*>(ctx->out_dev2), meta);
    }
    ZDWB_LAUNCH_CHECK(ctx);
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(hm, meta, sizeof(DecMeta), cudaMemcpyDeviceToHost, st));
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (!hm->overflow) break;
    if (attempt >= 1) {
      ctx->err = "decode: output size estimate failed twice";
      return ZDWB_ERR_CUDA;
    }
    out_cap = hm->out_bytes + 64;
  }
  ctx->last_out_per_row = (hm->out_bytes / nrows) + (hm->out_bytes / nrows) / 8 + 16;
  if (hm->err) {
    ctx->err = "decode: dictionary offset out of range";
    return ZDWB_ERR_CORRUPT;  // CORRUPTED_DATA_ERROR, UnconvertFromZDW.cpp:1364-1365
  }
  const uint64_t out_len = hm->out_bytes;
  out->len = out_len;
  if (opts->output_on_device) {
    out->tsv = static_cast<const uint8_t*>(ctx->out_dev);
    out->row_off = static_cast<const uint64_t*>(ctx->out_dev2);
    return ZDWB_OK;
  }
  if (ctx->out_host_cap < out_len) {
    if (ctx->out_host) cudaFreeHost(ctx->out_host);
    ctx->out_host = nullptr;
    ctx->out_host_cap = 0;
    const size_t cap = std::max<size_t>(out_len, 1 << 20);
    ZDWB_CUDA_TRY(ctx, cudaHostAlloc(&ctx->out_host, cap, cudaHostAllocDefault));
    ctx->out_host_cap = cap;
  }
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->out_host, ctx->out_dev, out_len, cudaMemcpyDeviceToHost, st));
  if (opts->want_row_offsets) {
    const size_t rb = ((size_t)nrows + 1) * 8;
    if (ctx->out_host2_cap < rb) {
      if (ctx->out_host2) cudaFreeHost(ctx->out_host2);
      ctx->out_host2 = nullptr;
      ctx->out_host2_cap = 0;
      ZDWB_CUDA_TRY(ctx, cudaHostAlloc(&ctx->out_host2, std::max<size_t>(rb, 1 << 16), cudaHostAllocDefault));
      ctx->out_host2_cap = std::max<size_t>(rb, 1 << 16);
    }
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->out_host2, ctx->out_dev2, rb, cudaMemcpyDeviceToHost, st));
    out->row_off = static_cast<const uint64_t*>(ctx->out_host2);
  }
  ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  out->tsv = static_cast<const uint8_t*>(ctx->out_host);
  return ZDWB_OK;
}

}  // namespace zdwb
