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
