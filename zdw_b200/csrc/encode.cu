// encode.cu -- TSV -> one ZDW block, entirely on the GPU.
//
// Kernel sequence for a block (reference functions replaced are cited per kernel):
//   k_rows_count / k_rows_write   GetNextRow                     getnextrow.cpp:26-84
//   k_row_longest                 m_LongestLine bookkeeping      getnextrow.cpp:57-65
//   k_pass1                       parseInput + get_next_column   ConvertToZDW.cpp:329-414, :1048-1067
//                                 + Dictionary::insert           dictionary.cpp:31-51
//   k_ht_compact, sort_strings, k_sorted_lens, k_dict_emit
//                                 Dictionary::write              dictionary.cpp:76-111
//   k_col_stats, k_block_header   writeLookupColumnStats         ConvertToZDW.cpp:417-483, :839-842
//   k_pass2                       writeBlockRows                 ConvertToZDW.cpp:486-606
//                                 + Dictionary::getOffset        dictionary.cpp:53-59
//
// Data layout in HBM: the TSV block stays where it is (one read per pass); the only per-row index is
// row_start/row_end (8 B/row).  Field boundaries are never materialised: both passes re-derive them from
// the bytes with a 16-byte-per-thread classifier and a block-wide prefix count, so work is proportional
// to bytes + non-empty fields, not to the (mostly empty) field count of wide analytics schemas.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tsv.cuh"

namespace zdwb {

namespace {

constexpr int ENC_THREADS = 256;
constexpr int IDX_CH = 4;                             // 16-byte chunks per thread in the row indexer
constexpr int IDX_TILE = ENC_THREADS * IDX_CH * 16;   // 16 KiB per CTA
constexpr uint32_t HT_MAX_PROBE = 2048;

struct EncMeta {
  uint32_t bad_row;        // first row whose field count != schema (0xffffffff = none)
  uint32_t max_line;       // longest logical line incl. '\n' among the block's rows (+ tail rule)
  uint32_t ht_overflow;    // pass 1 gave up: hash table too small
  uint32_t lookup_miss;    // pass 2 internal consistency counter (must stay 0)
  unsigned long long n_unique;
  unsigned long long dict_str_bytes;  // sum(len + 1)
  uint32_t max_str_len;
  uint32_t n_used;
  uint32_t idx_size;
  uint32_t nflag;
  uint32_t max_row_bytes;
  uint32_t pad0;
  uint64_t dict_total;   // Dictionary::getSize()
  uint64_t dict_base;    // offset of the dictionary origin byte inside the block
  uint64_t stats_base;
  uint64_t rows_base;
  unsigned long long rows_bytes;
  uint32_t compact_count;
  uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// row index
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_rows_count(const uint8_t* __restrict__ buf, uint64_t n, int64_t lo, uint64_t* __restrict__ tile_cnt) {
  __shared__ uint64_t sh[ENC_THREADS / 32];
  const int64_t p0 = lo + ((int64_t)blockIdx.x * ENC_THREADS + threadIdx.x) * (IDX_CH * 16);
  uint64_t cnt = 0;
#pragma unroll
  for (int c = 0; c < IDX_CH; ++c) {
    const int64_t p = p0 + c * 16;
    if (p < (int64_t)n) {
      ChunkMasks m = classify_chunk(buf, n, p, 0, (int64_t)n);
      cnt += ((uint64_t)__popc(m.term) << 32) | (uint64_t)__popc(m.tab);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < ENC_THREADS / 32; ++w) t += sh[w];
    tile_cnt[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(ENC_THREADS)
    k_rows_write(const uint8_t* __restrict__ buf, uint64_t n, int64_t lo, const uint64_t* __restrict__ tile_base,
                 uint32_t ncols, uint32_t* __restrict__ row_start, uint32_t* __restrict__ row_end,
                 EncMeta* __restrict__ meta) {
  __shared__ uint64_t ws[34];
  const int64_t p0 = lo + ((int64_t)blockIdx.x * ENC_THREADS + threadIdx.x) * (IDX_CH * 16);
  ChunkMasks m[IDX_CH];
  uint64_t cnt = 0;
#pragma unroll
  for (int c = 0; c < IDX_CH; ++c) {
    const int64_t p = p0 + c * 16;
    m[c] = ChunkMasks{0u, 0u, 0u};
    if (p < (int64_t)n) {
      m[c] = classify_chunk(buf, n, p, 0, (int64_t)n);
      cnt += ((uint64_t)__popc(m[c].term) << 32) | (uint64_t)__popc(m[c].tab);
    }
  }
  uint64_t excl = block_exclusive_scan64(cnt, ws, nullptr) + tile_base[blockIdx.x];
  uint64_t rows_before = excl >> 32, tabs_before = excl & 0xffffffffull;
  const uint64_t tabs_per_row = (uint64_t)ncols - 1;
#pragma unroll
  for (int c = 0; c < IDX_CH; ++c) {
    uint32_t d = m[c].tab | m[c].term;
    while (d) {
      const int i = __ffs(d) - 1;
      d &= d - 1;
      if ((m[c].tab >> i) & 1u) {
        ++tabs_before;
      } else {
        const uint64_t r = rows_before;
        const uint64_t pos = (uint64_t)(p0 + c * 16 + i);
        if (tabs_before != (r + 1) * tabs_per_row) atomicMin(&meta->bad_row, (uint32_t)r);
        row_end[r] = (uint32_t)pos;
        uint64_t s = pos + 1;
        while (s < n && __ldg(buf + s) == (uint8_t)'\n') ++s;  // blank lines directly after a row are skipped
        row_start[r + 1] = (uint32_t)s;
        ++rows_before;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint64_t s = 0;
    while (s < n && __ldg(buf + s) == (uint8_t)'\n') ++s;
    row_start[0] = (uint32_t)s;
  }
}

// Longest logical line (incl. its '\n') among rows [0, nrows); with `tail_bytes` >= 2 the unterminated
// tail participates as a line of tail_bytes + 1 (it fills the row buffer the same way before being dropped).
__global__ void k_row_longest(const uint32_t* __restrict__ row_start, const uint32_t* __restrict__ row_end,
                              uint32_t nrows, uint32_t tail_bytes, EncMeta* __restrict__ meta) {
  uint32_t mx = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    uint32_t L = row_end[r] - row_start[r] + 1;
    mx = L > mx ? L : mx;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && tail_bytes >= 2) mx = mx > tail_bytes + 1 ? mx : tail_bytes + 1;
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(&meta->max_line, mx);
}

// ---------------------------------------------------------------------------------------------
// field walker: calls f(row, col, start, len) for every NON-EMPTY field of rows [row0, row0+k) whose
// bytes are [b0, b1).
//
// Phase A (delimiter-parallel): one thread owns 32 bytes per iteration (two aligned 16-byte loads), builds the
// tab / terminator / blank-line bit masks and - with bit arithmetic only - the mask of delimiters that close a
// non-empty field (in analytics-shaped data most delimiters close empty fields and cost nothing further).  A
// block-wide prefix count of delimiters gives every delimiter its (row, col).
// Phase B (field-parallel): the non-empty fields of the iteration are queued in shared memory and handed out
// one per thread, so the expensive per-field work (hashing, probing, number parsing) runs with full warps no
// matter how the fields were distributed over the bytes.
// ---------------------------------------------------------------------------------------------
constexpr int WALK_SPAN = 32;                          // bytes per thread per iteration
constexpr int QCAP = 1024;                             // queued fields per round

struct WalkScratch {
  uint32_t scan_ws[34];
  int32_t last[ENC_THREADS];
  int32_t wmax[ENC_THREADS / 32];
  int32_t carry;
  uint4 queue[QCAP];  // (row, col, start, len)
};

constexpr int32_t NO_BOUNDARY = INT32_MIN;

template <class F>
__device__ __forceinline__ void walk_tile_fields(const uint8_t* __restrict__ buf, uint64_t n, int64_t lo, uint32_t b0,
                                                 uint32_t b1, uint32_t row0, uint32_t ncols, bool trim, WalkScratch& ts,
                                                 F&& f) {
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t pc0 = lo + ((((int64_t)b0 - lo) >> 4) << 4);
  const uint32_t nsp = (uint32_t)(((int64_t)b1 - pc0 + WALK_SPAN - 1) / WALK_SPAN);
  if (tid == 0) ts.carry = (int32_t)((int64_t)b0 - 1 - pc0);  // the byte before a row start is a boundary
  uint32_t kbase = 0;
  __syncthreads();
  for (uint32_t it0 = 0; it0 < nsp; it0 += ENC_THREADS) {
    const uint32_t j = it0 + tid;
    const int64_t p0 = pc0 + (int64_t)j * WALK_SPAN;
    uint32_t tab = 0, term = 0, skip = 0;
    if (j < nsp) {
      const ChunkMasks m0 = classify_chunk(buf, n, p0, (int64_t)b0, (int64_t)b1);
      const ChunkMasks m1 = classify_chunk(buf, n, p0 + 16, (int64_t)b0, (int64_t)b1);
      tab = m0.tab | (m1.tab << 16);
      term = m0.term | (m1.term << 16);
      skip = m0.skip | (m1.skip << 16);
    }
    const uint32_t bound = tab | term | skip;
    const uint32_t delims = tab | term;
    const int32_t rel0 = (int32_t)(j * WALK_SPAN);
    const int32_t mylast = bound ? rel0 + (31 - __clz(bound)) : NO_BOUNDARY;
    ts.last[tid] = mylast;
    const int32_t wm = __reduce_max_sync(0xffffffffu, mylast);
    if (lane == 0) ts.wmax[warp] = wm;
    __syncthreads();
    // last boundary before this thread's span
    int32_t pb = NO_BOUNDARY;
    if (delims) {
      int t = (int)tid - 1;
      while (t >= 0 && (pb = ts.last[t]) == NO_BOUNDARY) --t;
      if (t < 0) pb = ts.carry;
    }
    // delimiters whose preceding byte is not a boundary close a non-empty field
    const uint32_t ne = delims & ~((bound << 1) | (pb == rel0 - 1 ? 1u : 0u));
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(((uint32_t)__popc(ne) << 16) | (uint32_t)__popc(delims), ts.scan_ws, &total);
    const uint32_t total_ne = total >> 16;
    uint32_t row_f = 0, col_f = 0;
    if (ne) {
      const uint32_t k0 = kbase + (excl & 0xffffu);
      const uint32_t q = k0 / ncols;
      row_f = row0 + q;
      col_f = k0 - q * ncols;
    }
    for (uint32_t base = 0; base < total_ne; base += QCAP) {
      if (ne) {
        uint32_t qi = excl >> 16;
        uint32_t d = ne;
        while (d) {
          const int i = __ffs(d) - 1;
          d &= d - 1;
          if (qi >= base && qi < base + QCAP) {
            const uint32_t below = (1u << i) - 1u;
            uint32_t col = col_f + (uint32_t)__popc(delims & below);
            uint32_t row = row_f;
            while (col >= ncols) {
              col -= ncols;
              ++row;
            }
            const uint32_t lowb = bound & below;
            const int32_t srel = (lowb ? rel0 + (31 - __clz(lowb)) : pb) + 1;
            const uint32_t start = (uint32_t)(pc0 + srel);
            const uint32_t end = (uint32_t)(p0 + i);
            ts.queue[qi - base] = make_uint4(row, col, start, end - start);
          }
          ++qi;
        }
      }
      __syncthreads();
      const uint32_t cnt = min((uint32_t)QCAP, total_ne - base);
      for (uint32_t t = tid; t < cnt; t += ENC_THREADS) {
        const uint4 rec = ts.queue[t];
        uint32_t len = rec.w;
        if (trim) {  // -t: ConvertToZDW.cpp:295-313
          while (len && __ldg(buf + rec.z + len - 1) == (uint8_t)' ') --len;
        }
        if (len) f(rec.x, rec.y, rec.z, len);
      }
      __syncthreads();
    }
    if (tid == 0) {
      int32_t mx = NO_BOUNDARY;
#pragma unroll
      for (int w = 0; w < ENC_THREADS / 32; ++w) mx = ts.wmax[w] > mx ? ts.wmax[w] : mx;
      if (mx != NO_BOUNDARY) ts.carry = mx;
    }
    kbase += total & 0xffffu;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// string hash set (open addressing, 64-bit slots: (start+1) << 32 | len, 0 = empty)
// ---------------------------------------------------------------------------------------------
// Little-endian 32-bit words of the byte string that starts at an arbitrary address: aligned loads + funnel shift.
// Reads at most one aligned word past the word holding the last requested byte.
struct WordStream {
  const uint32_t* w;
  uint32_t sh, cur;
  __device__ __forceinline__ explicit WordStream(const uint8_t* a) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    w = reinterpret_cast<const uint32_t*>(u & ~(uintptr_t)3);
    sh = (uint32_t)(u & 3u) * 8u;
    cur = __ldg(w);
  }
  __device__ __forceinline__ uint32_t next() {
    const uint32_t nx = __ldg(++w);
    const uint32_t r = __funnelshift_r(cur, nx, sh);
    cur = nx;
    return r;
  }
  // the next `rem` (1..3) bytes, zero-extended
  __device__ __forceinline__ uint32_t tail(uint32_t rem) {
    uint32_t r;
    if (sh + rem * 8u <= 32u) r = cur >> sh;
    else r = __funnelshift_r(cur, __ldg(w + 1), sh);
    return r & ((1u << (rem * 8u)) - 1u);
  }
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// MurmurHash3 (x86_32) over the field bytes
__device__ __forceinline__ uint32_t hash_bytes(const uint8_t* __restrict__ p, uint32_t len) {
  WordStream ws(p);
  uint32_t h = 0x9747b28cu;
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4) {
    uint32_t k = ws.next();
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5u + 0xe6546b64u;
  }
  if (len - i) {
    uint32_t k = ws.tail(len - i);
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
  }
  h ^= len;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__device__ __forceinline__ bool bytes_equal(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint32_t len) {
  WordStream sa(a), sb(b);
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4)
    if (sa.next() != sb.next()) return false;
  if (len - i) return sa.tail(len - i) == sb.tail(len - i);
  return true;
}

struct HashTable {
  unsigned long long* slots;
  uint32_t mask;
};

// returns true when a new entry was created
__device__ __forceinline__ bool ht_insert(const HashTable ht, const uint8_t* __restrict__ buf, uint32_t start, uint32_t len,
                                          EncMeta* __restrict__ meta) {
  if (*reinterpret_cast<volatile uint32_t*>(&meta->ht_overflow)) return false;
  uint32_t i = hash_bytes(buf + start, len) & ht.mask;
  const unsigned long long mine = ((unsigned long long)(start + 1u) << 32) | len;
  for (uint32_t probe = 0; probe < HT_MAX_PROBE; ++probe) {
    unsigned long long cur = ht.slots[i];
    if (cur == 0ull) {
      cur = atomicCAS(&ht.slots[i], 0ull, mine);
      if (cur == 0ull) return true;
    }
    if ((uint32_t)cur == len) {
      const uint32_t cs = (uint32_t)(cur >> 32) - 1u;
      if (cs == start || bytes_equal(buf + cs, buf + start, len)) return false;
    }
    i = (i + 1) & ht.mask;
  }
  *reinterpret_cast<volatile uint32_t*>(&meta->ht_overflow) = 1u;
  return false;
}

// returns the slot holding the string or 0xffffffff
__device__ __forceinline__ uint32_t ht_find(const HashTable ht, const uint8_t* __restrict__ buf, uint32_t start,
                                            uint32_t len) {
  uint32_t i = hash_bytes(buf + start, len) & ht.mask;
  for (uint32_t probe = 0; probe <= ht.mask; ++probe) {
    const unsigned long long cur = ht.slots[i];
    if (cur == 0ull) return 0xffffffffu;
    if ((uint32_t)cur == len) {
      const uint32_t cs = (uint32_t)(cur >> 32) - 1u;
      if (cs == start || bytes_equal(buf + cs, buf + start, len)) return i;
    }
    i = (i + 1) & ht.mask;
  }
  return 0xffffffffu;
}

// ---------------------------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_pass1(const uint8_t* __restrict__ buf, uint64_t n, int64_t lo, const uint32_t* __restrict__ row_start,
            const uint32_t* __restrict__ row_end, uint32_t nrows, uint32_t rows_per_cta, uint32_t ncols,
            const uint8_t* __restrict__ types, int trim, HashTable ht, uint32_t* __restrict__ colset,
            unsigned long long* __restrict__ colmin, unsigned long long* __restrict__ colmax,
            EncMeta* __restrict__ meta) {
  __shared__ WalkScratch ts;
  __shared__ unsigned long long s_new_bytes;
  __shared__ uint32_t s_new_count, s_max_len;
  if (threadIdx.x == 0) {
    s_new_bytes = 0;
    s_new_count = 0;
    s_max_len = 0;
  }
  const uint32_t r0 = blockIdx.x * rows_per_cta;
  const uint32_t r1 = min(nrows, r0 + rows_per_cta);
  const uint32_t b0 = row_start[r0], b1 = row_end[r1 - 1] + 1;
  walk_tile_fields(buf, n, lo, b0, b1, r0, ncols, trim != 0, ts, [&](uint32_t, uint32_t col, uint32_t start, uint32_t len) {
    const uint8_t t = __ldg(types + col);
    const uint8_t* p = buf + start;
    if (is_text_like(t)) {
      if (__ldcg(colset + col) == 0u) colset[col] = 1u;
      if (ht_insert(ht, buf, start, len, meta)) {
        atomicAdd(&s_new_count, 1u);
        atomicAdd(&s_new_bytes, (unsigned long long)len + 1ull);
        atomicMax(&s_max_len, len);
      }
    } else {
      const uint64_t v = (t == ZDWB_CHAR) ? char_tuple(p, len, false) : parse_u64_field(p, len);
      if (v != 0) {  // zero / empty numeric cells take no part in min/max: ConvertToZDW.cpp:362,386
        if (__ldcg(colset + col) == 0u) colset[col] = 1u;
        if (v < __ldcg(colmin + col)) atomicMin(colmin + col, (unsigned long long)v);
        if (v > __ldcg(colmax + col)) atomicMax(colmax + col, (unsigned long long)v);
      }
    }
  });
  __syncthreads();
  if (threadIdx.x == 0 && s_new_count) {
    atomicAdd(&meta->n_unique, (unsigned long long)s_new_count);
    atomicAdd(&meta->dict_str_bytes, s_new_bytes);
    atomicMax(&meta->max_str_len, s_max_len);
  }
}

// ---------------------------------------------------------------------------------------------
// dictionary
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_ht_compact(const unsigned long long* __restrict__ slots, uint32_t capacity, uint32_t* __restrict__ ustart,
                 uint32_t* __restrict__ ulen, uint32_t* __restrict__ uslot, EncMeta* __restrict__ meta) {
  __shared__ uint32_t ws[34];
  __shared__ uint32_t s_base;
  const uint32_t i = blockIdx.x * ENC_THREADS + threadIdx.x;
  const unsigned long long cur = i < capacity ? slots[i] : 0ull;
  const uint32_t has = cur != 0ull ? 1u : 0u;
  uint32_t total;
  const uint32_t excl = block_exclusive_scan(has, ws, &total);
  if (threadIdx.x == 0) s_base = total ? atomicAdd(&meta->compact_count, total) : 0u;
  __syncthreads();
  if (has) {
    const uint32_t id = s_base + excl;
    ustart[id] = (uint32_t)(cur >> 32) - 1u;
    ulen[id] = (uint32_t)cur;
    uslot[id] = i;
  }
}

__global__ void k_sorted_lens(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ulen, uint32_t n,
                              uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ulen[order[i]] + 1u;
}

// dictionary offset of every hash-set slot: offs[i] = sum of (len+1) of the entries sorted before i; offset 0 is the
// origin byte (dictionary.cpp:96-98)
__global__ void k_dict_slots(const uint32_t* __restrict__ order, const uint32_t* __restrict__ uslot,
                             const uint32_t* __restrict__ offs, uint32_t n, uint32_t* __restrict__ slot_off) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) slot_off[uslot[order[g]]] = 1u + offs[g];
}

// 8 lanes copy one dictionary entry
__global__ void k_dict_emit(const uint8_t* __restrict__ buf, const uint32_t* __restrict__ order,
                            const uint32_t* __restrict__ ustart, const uint32_t* __restrict__ ulen,
                            const uint32_t* __restrict__ offs, uint32_t n, uint8_t* __restrict__ dict_origin) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
  if (g >= n) return;
  const uint32_t id = order[g];
  const uint32_t off = 1u + offs[g];
  const uint32_t len = ulen[id];
  const uint8_t* src = buf + ustart[id];
  uint8_t* dst = dict_origin + off;
  for (uint32_t k = sub; k < len; k += 8) dst[k] = __ldg(src + k);
  if (sub == 0) dst[len] = 0;
}

// ---------------------------------------------------------------------------------------------
// column statistics + block header
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENC_THREADS)
    k_col_stats(const uint8_t* __restrict__ types, uint32_t ncols, const uint32_t* __restrict__ colset,
                const unsigned long long* __restrict__ colmin, const unsigned long long* __restrict__ colmax,
                uint8_t* __restrict__ csize, unsigned long long* __restrict__ cbase, int32_t* __restrict__ used_idx,
                uint32_t* __restrict__ used_cols, EncMeta* __restrict__ meta) {
  __shared__ uint32_t ws[34];
  __shared__ uint32_t s_sum[ENC_THREADS / 32];
  // Dictionary::getSize() = size + 1; getBytesInOffset(): dictionary.cpp:62-73
  const uint64_t dict_total = meta->dict_str_bytes + 1ull;
  const uint32_t idx_size = bytes_needed(dict_total);
  uint32_t ubase = 0, size_sum = 0;
  for (uint32_t c0 = 0; c0 < ncols; c0 += ENC_THREADS) {
    const uint32_t c = c0 + threadIdx.x;
    uint32_t sz = 0;
    unsigned long long base = 0;
    if (c < ncols && colset[c]) {
      if (is_text_like(types[c])) {
        sz = idx_size;  // every text-like column uses the dictionary's offset width: ConvertToZDW.cpp:447
      } else {
        base = colmin[c] - 1ull;  // :455
        sz = bytes_needed(colmax[c] - base);
      }
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(sz ? 1u : 0u, ws, &total);
    if (c < ncols) {
      csize[c] = (uint8_t)sz;
      cbase[c] = base;
      used_idx[c] = sz ? (int32_t)(ubase + excl) : -1;
      if (sz) used_cols[ubase + excl] = c;
    }
    size_sum += sz;
    ubase += total;
  }
  size_sum = __reduce_add_sync(0xffffffffu, size_sum);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = size_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < ENC_THREADS / 32; ++w) tot += s_sum[w];
    const uint32_t nflag = (ubase + 7) / 8;
    meta->n_used = ubase;
    meta->nflag = nflag;
    meta->max_row_bytes = nflag + tot;
    meta->idx_size = idx_size;
    meta->dict_total = dict_total;
    if (meta->n_unique == 0ull) {  // empty dictionary: a single 0x00 byte (dictionary.cpp:80-84)
      meta->dict_base = 9;         // (unused)
      meta->stats_base = 10;
    } else {
      meta->dict_base = 9 + 1 + idx_size;  // idxSize byte, totalBytes, then the origin byte
      meta->stats_base = meta->dict_base + dict_total;
    }
    meta->rows_base = meta->stats_base + ncols + 8ull * ubase;
  }
}

__global__ void k_block_header(uint8_t* __restrict__ out, const EncMeta* __restrict__ meta, uint32_t nrows,
                               uint32_t longest_line, uint32_t is_last, uint32_t ncols,
                               const uint8_t* __restrict__ csize, const unsigned long long* __restrict__ cbase,
                               const uint32_t* __restrict__ used_cols) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (tid == 0) {
    for (int k = 0; k < 4; ++k) out[k] = (uint8_t)(nrows >> (8 * k));
    for (int k = 0; k < 4; ++k) out[4 + k] = (uint8_t)(longest_line >> (8 * k));
    out[8] = (uint8_t)is_last;
    if (meta->n_unique == 0ull) {
      out[9] = 0;
    } else {
      out[9] = (uint8_t)meta->idx_size;
      for (uint32_t k = 0; k < meta->idx_size; ++k) out[10 + k] = (uint8_t)(meta->dict_total >> (8 * k));
      out[meta->dict_base] = 0;  // origin byte
    }
  }
  uint8_t* st = out + meta->stats_base;
  for (uint32_t c = tid; c < ncols; c += nt) st[c] = csize[c];
  uint8_t* bs = st + ncols;
  const uint32_t nu = meta->n_used;
  for (uint32_t u = tid; u < nu; u += nt) {
    const unsigned long long b = cbase[used_cols[u]];
    for (int k = 0; k < 8; ++k) bs[(size_t)u * 8 + k] = (uint8_t)(b >> (8 * k));
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2
// ---------------------------------------------------------------------------------------------
// One CTA encodes R consecutive rows.  The values of the rows (and of the row before the first one, which the
// repeat flags of the first row compare against) are gathered into a shared-memory matrix by the field walker;
// flags and value bytes are then produced row by row, one warp per row.  The tile's bytes go to its own slot of
// a staging buffer (tile * tile_cap); k_gather_tiles packs the tiles once every tile length is known, which keeps
// the row stream free of any cross-CTA dependency while it is being produced.
__global__ void __launch_bounds__(ENC_THREADS)
    k_pass2(const uint8_t* __restrict__ buf, uint64_t n, int64_t lo, const uint32_t* __restrict__ row_start,
            const uint32_t* __restrict__ row_end, uint32_t nrows, uint32_t rows_per_cta, uint32_t ncols,
            const uint8_t* __restrict__ types, int trim, HashTable ht, const uint32_t* __restrict__ slot_off,
            const int32_t* __restrict__ used_idx, const uint32_t* __restrict__ used_cols,
            const uint8_t* __restrict__ csize, const unsigned long long* __restrict__ cbase, uint32_t U,
            uint32_t nflag, uint64_t tile_cap, uint8_t* __restrict__ staging, uint64_t* __restrict__ tile_bytes,
            EncMeta* __restrict__ meta) {
  extern __shared__ __align__(16) uint8_t dsm[];
  __shared__ WalkScratch ts;
  // dynamic smem: nval[(R+1)*U] u64 | rowoff[R+1] u32 | usz[U] u8
  unsigned long long* nval = reinterpret_cast<unsigned long long*>(dsm);
  uint32_t* rowoff = reinterpret_cast<uint32_t*>(dsm + (size_t)(rows_per_cta + 1) * U * 8);
  uint8_t* usz = reinterpret_cast<uint8_t*>(rowoff + rows_per_cta + 1);
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  const uint32_t tile = blockIdx.x;
  const uint32_t r0 = tile * rows_per_cta;
  const uint32_t r1 = min(nrows, r0 + rows_per_cta);
  const uint32_t R = r1 - r0;
  for (uint32_t k = tid; k < (R + 1) * U; k += ENC_THREADS) nval[k] = 0ull;
  for (uint32_t u = tid; u < U; u += ENC_THREADS) usz[u] = csize[used_cols[u]];
  // (walk_tile_fields starts with a __syncthreads)

  // ---- values of the tile's rows and of the row before it
  const uint32_t rw0 = r0 ? r0 - 1 : 0;
  const uint32_t b0 = row_start[rw0], b1 = row_end[r1 - 1] + 1;
  walk_tile_fields(buf, n, lo, b0, b1, rw0, ncols, trim != 0, ts, [&](uint32_t row, uint32_t col, uint32_t start, uint32_t len) {
    const int32_t u = __ldg(used_idx + col);
    if (u < 0) return;
    const uint8_t t = __ldg(types + col);
    const uint8_t* p = buf + start;
    unsigned long long v;
    if (is_text_like(t)) {
      const uint32_t slot = ht_find(ht, buf, start, len);
      if (slot == 0xffffffffu) {
        atomicAdd(&meta->lookup_miss, 1u);
        v = 0;
      } else {
        v = slot_off[slot];
      }
    } else if (t == ZDWB_CHAR) {
      v = char_tuple(p, len, true);  // ConvertToZDW.cpp:543-547
      if (v) v -= __ldg(cbase + col);
    } else {
      v = parse_u64_field(p, len);   // :564-566
      if (v) v -= __ldg(cbase + col);
    }
    nval[(size_t)(row + 1 - r0) * U + (uint32_t)u] = v;
  });
  __syncthreads();

  // ---- encoded length of every row: nflag + sum of the sizes of the changed columns
  for (uint32_t j = warp; j < R; j += ENC_THREADS / 32) {
    const unsigned long long* cur = nval + (size_t)(j + 1) * U;
    const unsigned long long* prv = nval + (size_t)j * U;
    uint32_t acc = 0;
    for (uint32_t u = lane; u < U; u += 32)
      if (cur[u] != prv[u]) acc += usz[u];
    acc = __reduce_add_sync(0xffffffffu, acc);
    if (lane == 0) rowoff[j] = nflag + acc;
  }
  __syncthreads();
  // exclusive scan of the row lengths (R is small; warp 0 walks it in chunks of 32)
  if (warp == 0) {
    uint32_t run = 0;
    for (uint32_t j0 = 0; j0 < R; j0 += 32) {
      const uint32_t j = j0 + lane;
      const uint32_t v = j < R ? rowoff[j] : 0u;
      uint32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
      }
      if (j < R) rowoff[j] = run + inc - v;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
      rowoff[R] = run;
      tile_bytes[tile] = run;
    }
  }
  __syncthreads();
  uint8_t* tile_out = staging + (uint64_t)tile * tile_cap;

  // ---- emit: flag bytes, then the low columnSize bytes of every changed value (little-endian)
  for (uint32_t j = warp; j < R; j += ENC_THREADS / 32) {
    const unsigned long long* cur = nval + (size_t)(j + 1) * U;
    const unsigned long long* prv = nval + (size_t)j * U;
    uint8_t* orow = tile_out + rowoff[j];
    uint32_t voff = nflag;
    for (uint32_t ub = 0; ub < U; ub += 32) {
      const uint32_t u = ub + lane;
      unsigned long long v = 0;
      bool flag = false;
      if (u < U) {
        v = cur[u];
        flag = v != prv[u];
      }
      const unsigned bal = __ballot_sync(0xffffffffu, flag);
      if (lane < 4 && (ub >> 3) + lane < nflag) orow[(ub >> 3) + lane] = (uint8_t)(bal >> (8 * lane));
      const uint32_t sz = flag ? usz[u] : 0u;
      uint32_t inc = sz;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
      }
      const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
      if (flag) {
        uint8_t* d = orow + voff + inc - sz;
        for (uint32_t b = 0; b < sz; ++b) d[b] = (uint8_t)(v >> (8 * b));
      }
      voff += tot;
    }
  }
}

// packs the staged tiles into the row stream: tile t's bytes go to out_rows + tile_off[t]
__global__ void __launch_bounds__(ENC_THREADS)
    k_gather_tiles(const uint8_t* __restrict__ staging, uint64_t tile_cap, const uint64_t* __restrict__ tile_off,
                   const uint64_t* __restrict__ tile_bytes, uint8_t* __restrict__ out_rows) {
  const uint32_t tile = blockIdx.x;
  const uint8_t* src = staging + (uint64_t)tile * tile_cap;
  uint8_t* dst = out_rows + tile_off[tile];
  const uint32_t nb = (uint32_t)tile_bytes[tile];
  // head bytes up to a 16-byte aligned destination, 16-byte body through a funnel of source words, byte tail
  const uint32_t head = min(nb, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
  for (uint32_t k = threadIdx.x; k < head; k += ENC_THREADS) dst[k] = src[k];
  const uint32_t body = (nb - head) >> 4;
  for (uint32_t q = threadIdx.x; q < body; q += ENC_THREADS) {
    const uint8_t* s = src + head + (size_t)q * 16;
    WordStream ws(s);
    uint4 v;
    v.x = ws.next();
    v.y = ws.next();
    v.z = ws.next();
    v.w = ws.next();
    *reinterpret_cast<uint4*>(dst + head + (size_t)q * 16) = v;
  }
  for (uint32_t k = head + body * 16 + threadIdx.x; k < nb; k += ENC_THREADS) dst[k] = src[k];
}

__global__ void k_init_minmax(unsigned long long* colmin, unsigned long long* colmax, uint32_t* colset, uint32_t ncols) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncols) {
    colmin[c] = ~0ull;
    colmax[c] = 0ull;
    colset[c] = 0u;
  }
}

uint32_t longest_line_field(uint32_t prev, uint32_t max_line) {
  // smallest 16384 * 2^k >= L + 1, cumulative over the file (getnextrow.cpp:57-65)
  uint64_t cap = prev ? prev : 16384u;
  while (cap < (uint64_t)max_line + 1ull) cap *= 2;
  return (uint32_t)cap;
}

}  // namespace

// =================================================================================================
// host driver
// =================================================================================================
int encode_block_impl(Ctx* ctx, const zdwb_schema* schema, const void* tsv, size_t n, const zdwb_encode_opts* opts,
                      zdwb_block_out* out) {
  memset(out, 0, sizeof(*out));
  cudaStream_t st = ctx->stream;
  const uint32_t ncols = schema->ncols;
  if (ncols == 0 || !schema->types) {
    ctx->err = "encode: schema has no columns";
    return ZDWB_ERR_BAD_ARG;
  }
  for (uint32_t c = 0; c < ncols; ++c) {
    if (!is_known_type(schema->types[c])) {
      ctx->err = "encode: unsupported column type id " + std::to_string((int)schema->types[c]);
      return ZDWB_ERR_UNSUPPORTED;
    }
  }
  if (n >= 0xffffff00ull) {
    ctx->err = "encode: a block's TSV must be smaller than 4 GiB (row offsets are 32-bit); split the input";
    return ZDWB_ERR_UNSUPPORTED;
  }
  if (n == 0) return ZDWB_OK;

  // ---- input residency
  DevBuf tsv_dev;
  const uint8_t* buf;
  if (opts->input_on_device) {
    buf = static_cast<const uint8_t*>(tsv);
  } else {
    ZDWB_TRY(tsv_dev.alloc(ctx, n + 64));
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(tsv_dev.p, tsv, n, cudaMemcpyHostToDevice, st));
    buf = tsv_dev.as<uint8_t>();
  }
  const int64_t lo = -(int64_t)(reinterpret_cast<uintptr_t>(buf) & 15u);

  DevBuf meta_d, types_d;
  ZDWB_TRY(meta_d.alloc(ctx, sizeof(EncMeta)));
  ZDWB_TRY(types_d.alloc(ctx, ncols));
  EncMeta* meta = meta_d.as<EncMeta>();
  EncMeta* hmeta = static_cast<EncMeta*>(ctx->meta_host);
  ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(meta, 0, sizeof(EncMeta), st));
  {
    const uint32_t ff = 0xffffffffu;
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(&meta->bad_row, &ff, 4, cudaMemcpyHostToDevice, st));
  }
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(types_d.p, schema->types, ncols, cudaMemcpyHostToDevice, st));

  // ---- row index: count, scan, write
  const uint64_t span = (uint64_t)(-lo) + n;
  const uint32_t idx_tiles = (uint32_t)((span + IDX_TILE - 1) / IDX_TILE);
  DevBuf tile_cnt, total_d;
  ZDWB_TRY(tile_cnt.alloc(ctx, (size_t)idx_tiles * 8));
  ZDWB_TRY(total_d.alloc(ctx, 8));
  {
    KernelScope _ks(ctx, "k_rows_count");
    k_rows_count<<<idx_tiles, ENC_THREADS, 0, st>>>(buf, n, lo, tile_cnt.as<uint64_t>());
  }
  ZDWB_LAUNCH_CHECK(ctx);
  ZDWB_TRY(exclusive_scan_u64(ctx, tile_cnt.as<uint64_t>(), tile_cnt.as<uint64_t>(), idx_tiles, total_d.as<uint64_t>()));
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->meta_host, total_d.p, 8, cudaMemcpyDeviceToHost, st));
  ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  const uint64_t packed = *static_cast<uint64_t*>(ctx->meta_host);
  const uint64_t rows_total = packed >> 32;
  out->rows_in_buffer = rows_total;
  if (rows_total == 0) {
    out->tsv_consumed = opts->more_input_follows ? 0 : n;  // a window without a complete row: the caller widens it
    return ZDWB_OK;  // "Empty data file -- nothing to process", ConvertToZDW.cpp:824-835
  }
  DevBuf row_start, row_end;
  ZDWB_TRY(row_start.alloc(ctx, (size_t)(rows_total + 1) * 4));
  ZDWB_TRY(row_end.alloc(ctx, (size_t)rows_total * 4));
  {
    KernelScope _ks(ctx, "k_rows_write");
    k_rows_write<<<idx_tiles, ENC_THREADS, 0, st>>>(buf, n, lo, tile_cnt.as<uint64_t>(), ncols, row_start.as<uint32_t>(),
                                                  row_end.as<uint32_t>(), meta);
  }
  ZDWB_LAUNCH_CHECK(ctx);

  const uint64_t nrows64 = (opts->max_rows && opts->max_rows < rows_total) ? opts->max_rows : rows_total;
  const uint32_t nrows = (uint32_t)nrows64;
  const bool took_all = nrows64 == rows_total;
  const bool is_last = took_all && !opts->more_input_follows;

  // block extent + validation result
  uint32_t h_last[2] = {0, 0};  // row_end[nrows-1], row_start[rows_total]
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(hmeta, meta, sizeof(EncMeta), cudaMemcpyDeviceToHost, st));
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ctx->meta_host) + 1024, row_end.as<uint32_t>() + (nrows - 1), 4,
                                     cudaMemcpyDeviceToHost, st));
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ctx->meta_host) + 1028, row_start.as<uint32_t>() + rows_total, 4,
                                     cudaMemcpyDeviceToHost, st));
  ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  memcpy(h_last, reinterpret_cast<uint8_t*>(ctx->meta_host) + 1024, 8);
  if (hmeta->bad_row < nrows) {
    out->bad_row = hmeta->bad_row + 1;  // "Row %u had the problem": one past the last good row, ConvertToZDW.cpp:811
    ctx->err = "Row " + std::to_string(out->bad_row) + " had the problem";
    return ZDWB_ERR_WRONG_COLUMNS;
  }
  uint32_t tail_bytes = 0;
  if (is_last) {
    out->tsv_consumed = n;
    tail_bytes = (uint32_t)(n - h_last[1]);
  } else if (took_all) {
    out->tsv_consumed = h_last[1];  // the unterminated tail belongs to the next window
  } else {
    // the next block starts at the first byte of row `nrows`
    uint32_t nxt;
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->meta_host, row_start.as<uint32_t>() + nrows, 4, cudaMemcpyDeviceToHost, st));
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    nxt = *static_cast<uint32_t*>(ctx->meta_host);
    out->tsv_consumed = nxt;
  }
  out->nrows = nrows;

  {
    KernelScope _ks(ctx, "k_row_longest");
    k_row_longest<<<std::min<uint32_t>((nrows + 255) / 256, 1024u), 256, 0, st>>>(row_start.as<uint32_t>(), row_end.as<uint32_t>(),
                                                                              nrows, tail_bytes, meta);
  }
  ZDWB_LAUNCH_CHECK(ctx);

  // ---- pass 1 (retry with a larger hash set when it fills up)
  const uint64_t block_bytes = (uint64_t)h_last[0] + 1;
  const uint64_t avg_row = std::max<uint64_t>(1, block_bytes / nrows);
  uint32_t rpc1 = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, 65536 / avg_row), 4096);
  const uint32_t tiles1 = (nrows + rpc1 - 1) / rpc1;

  DevBuf colset, colmin, colmax, slots;
  ZDWB_TRY(colset.alloc(ctx, (size_t)ncols * 4));
  ZDWB_TRY(colmin.alloc(ctx, (size_t)ncols * 8));
  ZDWB_TRY(colmax.alloc(ctx, (size_t)ncols * 8));
  uint32_t cap_log2 = (uint32_t)std::max<long long>(10, std::min<long long>(ctx->ht_initial_log2, 31));
  while (cap_log2 < 31 && (1ull << cap_log2) < ctx->last_unique * 4) ++cap_log2;  // blocks of one file look alike
  {
    // a block cannot hold more distinct non-empty strings than bytes / 2
    uint32_t need = 10;
    while (need < 31 && (1ull << need) < block_bytes) ++need;
    if (cap_log2 > need) cap_log2 = need;
  }
  HashTable ht{nullptr, 0};
  for (;;) {
    const uint64_t cap = 1ull << cap_log2;
    ZDWB_TRY(slots.alloc(ctx, cap * 8));
    ZDWB_CUDA_TRY(ctx, cudaMemsetAsync(slots.p, 0, cap * 8, st));
    ht.slots = slots.as<unsigned long long>();
    ht.mask = (uint32_t)(cap - 1);
    {
      KernelScope _ks(ctx, "k_init_minmax");
      k_init_minmax<<<(ncols + 255) / 256, 256, 0, st>>>(colmin.as<unsigned long long>(), colmax.as<unsigned long long>(),
                                                       colset.as<uint32_t>(), ncols);
    }
    ZDWB_LAUNCH_CHECK(ctx);
    {
      KernelScope _ks(ctx, "k_pass1");
      k_pass1<<<tiles1, ENC_THREADS, 0, st>>>(buf, n, lo, row_start.as<uint32_t>(), row_end.as<uint32_t>(), nrows, rpc1, ncols,
                                            types_d.as<uint8_t>(), opts->trim_trailing_spaces, ht, colset.as<uint32_t>(),
                                            colmin.as<unsigned long long>(), colmax.as<unsigned long long>(), meta);
    }
    ZDWB_LAUNCH_CHECK(ctx);
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(hmeta, meta, sizeof(EncMeta), cudaMemcpyDeviceToHost, st));
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (!hmeta->ht_overflow && hmeta->n_unique * 2 <= cap) break;
    if (cap_log2 >= 31) {
      ctx->err = "encode: dictionary hash set exceeds 2^31 slots";
      return ZDWB_ERR_UNSUPPORTED;
    }
    cap_log2 = std::min<uint32_t>(31, cap_log2 + 3);
    // reset the counters pass 1 accumulates
    EncMeta z = *hmeta;
    z.ht_overflow = 0;
    z.n_unique = 0;
    z.dict_str_bytes = 0;
    z.max_str_len = 0;
    *hmeta = z;
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(meta, hmeta, sizeof(EncMeta), cudaMemcpyHostToDevice, st));
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  }
  const uint64_t n_unique = hmeta->n_unique;
  ctx->last_unique = n_unique;
  const uint64_t dict_total = hmeta->dict_str_bytes + 1;
  if (dict_total >= 0xffffffffull) {
    ctx->err = "encode: block dictionary would reach 4 GiB (Dictionary::size is 32-bit, dictionary.h:62); use smaller blocks";
    return ZDWB_ERR_UNSUPPORTED;
  }
  const uint32_t longest_field = longest_line_field(opts->prev_longest_line, hmeta->max_line);

  // ---- column statistics
  DevBuf csize, cbase, used_idx, used_cols;
  ZDWB_TRY(csize.alloc(ctx, ncols));
  ZDWB_TRY(cbase.alloc(ctx, (size_t)ncols * 8));
  ZDWB_TRY(used_idx.alloc(ctx, (size_t)ncols * 4));
  ZDWB_TRY(used_cols.alloc(ctx, (size_t)ncols * 4));
  {
    KernelScope _ks(ctx, "k_col_stats");
    k_col_stats<<<1, ENC_THREADS, 0, st>>>(types_d.as<uint8_t>(), ncols, colset.as<uint32_t>(), colmin.as<unsigned long long>(),
                                         colmax.as<unsigned long long>(), csize.as<uint8_t>(), cbase.as<unsigned long long>(),
                                         used_idx.as<int32_t>(), used_cols.as<uint32_t>(), meta);
  }
  ZDWB_LAUNCH_CHECK(ctx);
  ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(hmeta, meta, sizeof(EncMeta), cudaMemcpyDeviceToHost, st));
  ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
  const uint32_t U = hmeta->n_used, nflag = hmeta->nflag;
  const uint64_t rows_base = hmeta->rows_base;

  // ---- dictionary order: compact the hash set, sort, offsets of the sorted entries, offset per slot
  DevBuf slot_off, ustart, ulen, uslot, order, offs;
  ZDWB_TRY(slot_off.alloc(ctx, ((size_t)ht.mask + 1) * 4));
  const uint32_t nu = (uint32_t)n_unique;
  if (nu) {
    ZDWB_TRY(ustart.alloc(ctx, (size_t)nu * 4));
    ZDWB_TRY(ulen.alloc(ctx, (size_t)nu * 4));
    ZDWB_TRY(uslot.alloc(ctx, (size_t)nu * 4));
    ZDWB_TRY(order.alloc(ctx, (size_t)nu * 4));
    ZDWB_TRY(offs.alloc(ctx, (size_t)nu * 4));
    const uint32_t cap = ht.mask + 1;
    {
      KernelScope _ks(ctx, "k_ht_compact");
      k_ht_compact<<<(cap + ENC_THREADS - 1) / ENC_THREADS, ENC_THREADS, 0, st>>>(ht.slots, cap, ustart.as<uint32_t>(),
                                                                               ulen.as<uint32_t>(), uslot.as<uint32_t>(), meta);
    }
    ZDWB_LAUNCH_CHECK(ctx);
    ZDWB_TRY(sort_strings(ctx, buf, ustart.as<uint32_t>(), ulen.as<uint32_t>(), nu, hmeta->max_str_len, order.as<uint32_t>()));
    {
      KernelScope _ks(ctx, "k_sorted_lens");
      k_sorted_lens<<<(nu + 255) / 256, 256, 0, st>>>(order.as<uint32_t>(), ulen.as<uint32_t>(), nu, offs.as<uint32_t>());
    }
    ZDWB_LAUNCH_CHECK(ctx);
    ZDWB_TRY(exclusive_scan_u32(ctx, offs.as<uint32_t>(), offs.as<uint32_t>(), nu, nullptr));
    {
      KernelScope _ks(ctx, "k_dict_slots");
      k_dict_slots<<<(nu + 255) / 256, 256, 0, st>>>(order.as<uint32_t>(), uslot.as<uint32_t>(), offs.as<uint32_t>(), nu,
                                                   slot_off.as<uint32_t>());
    }
    ZDWB_LAUNCH_CHECK(ctx);
  }

  // ---- pass 2 into the staging buffer
  uint64_t rows_bytes = 0;
  DevBuf staging, tile_bytes, tile_off, rows_total_d;
  uint32_t tiles2 = 0;
  uint64_t tile_cap = 0;
  if (U > 0) {
    // rows per CTA: bounded by the shared-memory value matrix ((R+1) * U * 8 bytes) and a byte target
    const size_t smem_budget = 64 * 1024;
    const uint64_t by_smem = smem_budget / ((uint64_t)U * 8);
    if (by_smem < 2) {
      // even one row + its predecessor do not fit the default budget: allow the full 200 KiB
      if ((uint64_t)U * 8 * 2 + 1024 > 180 * 1024) {
        ctx->err = "encode: more than 11400 used columns in one block are not supported";
        return ZDWB_ERR_UNSUPPORTED;
      }
    }
    uint32_t rpc2 = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1, 65536 / avg_row), 1024);
    if (by_smem >= 2) rpc2 = (uint32_t)std::min<uint64_t>(rpc2, by_smem - 1);
    else rpc2 = 1;
    const size_t smem = (size_t)(rpc2 + 1) * U * 8 + (size_t)(rpc2 + 1) * 4 + U + 16;
    ZDWB_CUDA_TRY(ctx, cudaFuncSetAttribute(k_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024));
    tiles2 = (nrows + rpc2 - 1) / rpc2;
    tile_cap = (((uint64_t)rpc2 * hmeta->max_row_bytes) + 15) & ~15ull;
    ZDWB_TRY(staging.alloc(ctx, (size_t)tiles2 * tile_cap + 64));
    ZDWB_TRY(tile_bytes.alloc(ctx, (size_t)tiles2 * 8));
    ZDWB_TRY(tile_off.alloc(ctx, (size_t)tiles2 * 8));
    ZDWB_TRY(rows_total_d.alloc(ctx, 8));
    {
      KernelScope _ks(ctx, "k_pass2");
      k_pass2<<<tiles2, ENC_THREADS, smem, st>>>(buf, n, lo, row_start.as<uint32_t>(), row_end.as<uint32_t>(), nrows, rpc2, ncols,
                                               types_d.as<uint8_t>(), opts->trim_trailing_spaces, ht, slot_off.as<uint32_t>(),
                                               used_idx.as<int32_t>(), used_cols.as<uint32_t>(), csize.as<uint8_t>(),
                                               cbase.as<unsigned long long>(), U, nflag, tile_cap, staging.as<uint8_t>(),
                                               tile_bytes.as<uint64_t>(), meta);
    }
    ZDWB_LAUNCH_CHECK(ctx);
    ZDWB_TRY(exclusive_scan_u64(ctx, tile_bytes.as<uint64_t>(), tile_off.as<uint64_t>(), tiles2, rows_total_d.as<uint64_t>()));
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(hmeta, meta, sizeof(EncMeta), cudaMemcpyDeviceToHost, st));
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<uint8_t*>(ctx->meta_host) + 1024, rows_total_d.p, 8, cudaMemcpyDeviceToHost, st));
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (hmeta->lookup_miss) {
      ctx->err = "encode: internal error, " + std::to_string(hmeta->lookup_miss) + " dictionary lookups missed in pass 2";
      return ZDWB_ERR_CUDA;
    }
    memcpy(&rows_bytes, reinterpret_cast<uint8_t*>(ctx->meta_host) + 1024, 8);
  }

  // ---- output buffer of the exact size: header + dictionary + column stats, then the packed rows
  if (ctx->out_dev) {
    cudaFreeAsync(ctx->out_dev, st);
    ctx->out_dev = nullptr;
  }
  {
    DevBuf ob;
    ZDWB_TRY(ob.alloc(ctx, rows_base + rows_bytes + 64));
    ctx->out_dev = ob.detach();
  }
  uint8_t* outp = static_cast<uint8_t*>(ctx->out_dev);
  {
    KernelScope _ks(ctx, "k_block_header");
    k_block_header<<<std::max(1u, std::min((ncols + 255) / 256, 64u)), 256, 0, st>>>(
      outp, meta, nrows, longest_field, is_last ? 1u : 0u, ncols, csize.as<uint8_t>(), cbase.as<unsigned long long>(),
      used_cols.as<uint32_t>());
  }
  ZDWB_LAUNCH_CHECK(ctx);
  if (nu) {
    const uint64_t threads = (uint64_t)nu * 8;
    KernelScope _ks(ctx, "k_dict_emit");
    k_dict_emit<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(buf, order.as<uint32_t>(), ustart.as<uint32_t>(),
                                                                ulen.as<uint32_t>(), offs.as<uint32_t>(), nu,
                                                                outp + hmeta->dict_base);
  }
  if (nu) ZDWB_LAUNCH_CHECK(ctx);
  if (tiles2) {
    KernelScope _ks(ctx, "k_gather_tiles");
    k_gather_tiles<<<tiles2, ENC_THREADS, 0, st>>>(staging.as<uint8_t>(), tile_cap, tile_off.as<uint64_t>(),
                                                 tile_bytes.as<uint64_t>(), outp + rows_base);
  }
  if (tiles2) ZDWB_LAUNCH_CHECK(ctx);

  const size_t total_len = (size_t)(rows_base + rows_bytes);
  out->len = total_len;
  out->longest_line = longest_field;
  out->ncols_used = U;
  out->dict_entries = n_unique;
  out->dict_bytes = dict_total;
  out->dict_index_size = bytes_needed(dict_total);
  if (opts->output_on_device) {
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    out->bytes = outp;
  } else {
    if (ctx->out_host_cap < total_len) {
      if (ctx->out_host) cudaFreeHost(ctx->out_host);
      ctx->out_host = nullptr;
      ctx->out_host_cap = 0;
      size_t cap = std::max<size_t>(total_len, 1 << 20);
      ZDWB_CUDA_TRY(ctx, cudaHostAlloc(&ctx->out_host, cap, cudaHostAllocDefault));
      ctx->out_host_cap = cap;
    }
    ZDWB_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->out_host, outp, total_len, cudaMemcpyDeviceToHost, st));
    ZDWB_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    out->bytes = static_cast<const uint8_t*>(ctx->out_host);
  }
  return ZDWB_OK;
}

}  // namespace zdwb
