/*
 * synth_gen.c -- synthetic "analytics-hits-shaped" TSV generator (BASELINE config C4).
 *
 * Benchmark input only; not part of the product or the oracle.  Takes the reference's own
 * analytics-hits fixture (3 768 rows x 2 086 columns, 256 populated) and resamples it COLUMN-WISE:
 * every column keeps its empirical value pool (cardinality, fill rate, value lengths) and its
 * empirical row-to-row change rate, via a per-column Markov "repeat previous value" bit.  Numerics
 * stay canonical and there are no backslashes (as in the fixture), so decode(encode(x)) == x.
 * Deterministic: xorshift64* seeded with (seed, block index).
 *
 *   synth_profile_create(fixture, n, ncols)  -> handle (parses the fixture once)
 *   synth_block(handle, seed, block, nrows, out, cap) -> bytes written (0 if cap too small)
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

typedef struct {
  uint32_t ncols, nrows;
  const uint8_t* data;  /* private copy of the fixture */
  uint32_t* off;        /* [nrows*ncols] field start */
  uint32_t* len;        /* [nrows*ncols] field length */
  uint32_t* change_thr; /* [ncols] P(change) as a 32-bit threshold */
  uint8_t* populated;   /* [ncols] column has any non-empty value */
  uint32_t max_len_sum; /* upper bound of a generated row's length */
} profile;

static inline uint64_t xs(uint64_t* s) {
  uint64_t x = *s;
  x ^= x >> 12;
  x ^= x << 25;
  x ^= x >> 27;
  *s = x;
  return x * 2685821657736338717ULL;
}

void* synth_profile_create(const uint8_t* fixture, size_t n, uint32_t ncols) {
  profile* p = (profile*)calloc(1, sizeof(profile));
  uint8_t* copy = (uint8_t*)malloc(n + 1);
  memcpy(copy, fixture, n);
  p->data = copy;
  p->ncols = ncols;
  uint32_t nrows = 0;
  for (size_t i = 0; i < n; ++i) nrows += fixture[i] == '\n';
  p->nrows = nrows;
  p->off = (uint32_t*)malloc((size_t)nrows * ncols * 4);
  p->len = (uint32_t*)malloc((size_t)nrows * ncols * 4);
  p->change_thr = (uint32_t*)calloc(ncols, 4);
  p->populated = (uint8_t*)calloc(ncols, 1);
  size_t pos = 0;
  for (uint32_t r = 0; r < nrows; ++r) {
    for (uint32_t c = 0; c < ncols; ++c) {
      size_t s = pos;
      while (pos < n && copy[pos] != '\t' && copy[pos] != '\n') ++pos;
      p->off[(size_t)r * ncols + c] = (uint32_t)s;
      p->len[(size_t)r * ncols + c] = (uint32_t)(pos - s);
      if (pos > s) p->populated[c] = 1;
      if (c + 1 < ncols) {
        if (pos >= n || copy[pos] != '\t') { /* malformed fixture */
          free(p->off); free(p->len); free(p->change_thr); free(p->populated); free(copy); free(p);
          return NULL;
        }
        ++pos;
      }
    }
    if (pos >= n || copy[pos] != '\n') {
      free(p->off); free(p->len); free(p->change_thr); free(p->populated); free(copy); free(p);
      return NULL;
    }
    ++pos;
  }
  uint32_t bound = ncols;
  for (uint32_t c = 0; c < ncols; ++c) {
    uint32_t changes = 0, mx = 0;
    for (uint32_t r = 0; r < nrows; ++r) {
      const uint32_t l = p->len[(size_t)r * ncols + c];
      if (l > mx) mx = l;
      if (r) {
        const uint32_t lp = p->len[(size_t)(r - 1) * ncols + c];
        if (l != lp || memcmp(copy + p->off[(size_t)r * ncols + c], copy + p->off[(size_t)(r - 1) * ncols + c], l)) ++changes;
      }
    }
    bound += mx;
    /* A resampled draw equals the previous value with the pool's collision probability, so the draw
     * rate that reproduces the observed change rate is a little higher; the plain rate is close enough. */
    double rate = nrows > 1 ? (double)changes / (double)(nrows - 1) : 0.0;
    p->change_thr[c] = rate >= 1.0 ? 0xffffffffu : (uint32_t)(rate * 4294967296.0);
  }
  p->max_len_sum = bound;
  return p;
}

void synth_profile_destroy(void* h) {
  profile* p = (profile*)h;
  if (!p) return;
  free((void*)p->data);
  free(p->off);
  free(p->len);
  free(p->change_thr);
  free(p->populated);
  free(p);
}

uint32_t synth_max_row_bytes(void* h) { return ((profile*)h)->max_len_sum; }

size_t synth_block(void* h, uint64_t seed, uint64_t block, uint32_t nrows, uint8_t* out, size_t cap) {
  const profile* p = (const profile*)h;
  const uint32_t nc = p->ncols;
  uint64_t st = seed * 0x9E3779B97F4A7C15ULL + (block + 1) * 0xD1B54A32D192ED03ULL;
  if (!st) st = 1;
  for (int i = 0; i < 8; ++i) xs(&st);
  uint32_t* cur = (uint32_t*)malloc((size_t)nc * 4); /* fixture row currently supplying column c */
  uint32_t* pop = (uint32_t*)malloc((size_t)nc * 4);
  uint32_t npop = 0;
  for (uint32_t c = 0; c < nc; ++c) {
    cur[c] = (uint32_t)(xs(&st) % p->nrows);
    if (p->populated[c]) pop[npop++] = c;
  }
  size_t w = 0;
  for (uint32_t r = 0; r < nrows; ++r) {
    if (w + p->max_len_sum + 1 > cap) {
      free(cur);
      free(pop);
      return 0;
    }
    /* a row is ncols-1 tabs and a newline, with the populated columns' values spliced in */
    uint32_t prev_col = 0;
    for (uint32_t k = 0; k < npop; ++k) {
      const uint32_t c = pop[k];
      const uint64_t rnd = xs(&st);
      if ((uint32_t)rnd < p->change_thr[c]) cur[c] = (uint32_t)((rnd >> 32) % p->nrows);
      const uint32_t gap = c - prev_col;
      memset(out + w, '\t', gap);
      w += gap;
      const size_t cell = (size_t)cur[c] * nc + c;
      const uint32_t l = p->len[cell];
      memcpy(out + w, p->data + p->off[cell], l);
      w += l;
      prev_col = c;
    }
    const uint32_t gap = nc - 1 - prev_col;
    memset(out + w, '\t', gap);
    w += gap;
    out[w++] = '\n';
  }
  free(cur);
  free(pop);
  return w;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Config C5: high-cardinality text.  Row i: i+1, 'K' + hex(sha256(str(i))), 'V' + the same hex reversed,
 * 'W' + (i mod 97) - byte for byte what tests/c5_check.py:make_rows builds in Python, fast enough for 5 x 10^7 rows.
 * --------------------------------------------------------------------------------------------------------------- */
static uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_short(const unsigned char* msg, size_t len, unsigned char out[32]) { /* len < 56: one block */
  static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  unsigned char blk[64];
  memset(blk, 0, sizeof(blk));
  memcpy(blk, msg, len);
  blk[len] = 0x80;
  const uint64_t bits = (uint64_t)len * 8;
  for (int k = 0; k < 8; ++k) blk[63 - k] = (unsigned char)(bits >> (8 * k));
  uint32_t w[64];
  for (int t = 0; t < 16; ++t)
    w[t] = ((uint32_t)blk[4 * t] << 24) | ((uint32_t)blk[4 * t + 1] << 16) | ((uint32_t)blk[4 * t + 2] << 8) | blk[4 * t + 3];
  for (int t = 16; t < 64; ++t) {
    const uint32_t s0 = rotr32(w[t - 15], 7) ^ rotr32(w[t - 15], 18) ^ (w[t - 15] >> 3);
    const uint32_t s1 = rotr32(w[t - 2], 17) ^ rotr32(w[t - 2], 19) ^ (w[t - 2] >> 10);
    w[t] = w[t - 16] + s0 + w[t - 7] + s1;
  }
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int t = 0; t < 64; ++t) {
    const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25), ch = (e & f) ^ (~e & g);
    const uint32_t t1 = hh + S1 + ch + K[t] + w[t];
    const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
    const uint32_t t2 = S0 + maj;
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  for (int k = 0; k < 8; ++k)
    for (int j = 0; j < 4; ++j) out[4 * k + j] = (unsigned char)(h[k] >> (24 - 8 * j));
}

/* rows first .. first + count - 1 into out; returns the bytes written, 0 if they do not fit cap */
size_t c5_rows(uint64_t first, uint64_t count, char* out, size_t cap) {
  static const char HEX[] = "0123456789abcdef";
  size_t at = 0;
  for (uint64_t i = first; i < first + count; ++i) {
    if (at + 200 > cap) return 0;
    char num[24];
    const int nl = snprintf(num, sizeof(num), "%llu", (unsigned long long)i);
    unsigned char dig[32];
    sha256_short((const unsigned char*)num, (size_t)nl, dig);
    char hex[64];
    for (int k = 0; k < 32; ++k) {
      hex[2 * k] = HEX[dig[k] >> 4];
      hex[2 * k + 1] = HEX[dig[k] & 15];
    }
    at += (size_t)snprintf(out + at, 24, "%llu", (unsigned long long)(i + 1));
    out[at++] = '\t';
    out[at++] = 'K';
    memcpy(out + at, hex, 64);
    at += 64;
    out[at++] = '\t';
    out[at++] = 'V';
    for (int k = 0; k < 64; ++k) out[at + k] = hex[63 - k];
    at += 64;
    out[at++] = '\t';
    out[at++] = 'W';
    at += (size_t)snprintf(out + at, 8, "%u", (unsigned)(i % 97));
    out[at++] = '\n';
  }
  return at;
}
