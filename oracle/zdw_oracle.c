/*
 * zdw_oracle.c -- TEST INFRASTRUCTURE ONLY (see zdw_oracle.h).
 *
 * Plain-C CPU restatement of the adobe/zdw two-pass row transform.  Written from the behaviour of
 * the reference (cited per function, paths relative to /root/reference/cplusplus); it shares no
 * code with it.  Single-threaded, row-at-a-time, exactly like the reference, so that every quirk
 * (escape parity, strtoull semantics, signed char, lltoa(INT64_MIN), longest-line doubling) is
 * reproduced by construction rather than by a parallel re-derivation.
 *
 * Parity status: PINNED against the golden vectors and against oracle/_ref (see header).
 */
#define _GNU_SOURCE
#include "zdw_oracle.h"

#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

/* ------------------------------------------------------------------ small utilities */

typedef struct {
  uint8_t* d;
  size_t len, cap;
} obuf;

static void ob_reserve(obuf* b, size_t extra) {
  if (b->len + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 4096;
  while (nc < b->len + extra) nc *= 2;
  b->d = (uint8_t*)realloc(b->d, nc);
  if (!b->d) abort();
  b->cap = nc;
}
static void ob_put(obuf* b, const void* p, size_t n) {
  ob_reserve(b, n);
  memcpy(b->d + b->len, p, n);
  b->len += n;
}
static void ob_u8(obuf* b, uint8_t v) { ob_put(b, &v, 1); }
static void ob_u16(obuf* b, uint16_t v) { ob_put(b, &v, 2); } /* x86 little-endian, raw memory like the reference */
static void ob_u32(obuf* b, uint32_t v) { ob_put(b, &v, 4); }
static void ob_u64(obuf* b, uint64_t v) { ob_put(b, &v, 8); }

void zo_free(void* p) { free(p); }

static int is_text_like(uint8_t t) {
  /* the "dictionary" types: ConvertToZDW.cpp:345-352 */
  return t == ZO_DECIMAL || t == ZO_VARCHAR || t == ZO_TEXT || t == ZO_TINYTEXT || t == ZO_MEDIUMTEXT ||
         t == ZO_LONGTEXT || t == ZO_DATETIME || t == ZO_CHAR_2;
}

/* ------------------------------------------------------------------ .desc.sql */

/* ConvertToZDW::ReadDescFile, ConvertToZDW.cpp:91-162.  Lines are consumed like fgets(row, 1024). */
int zo_parse_desc(const char* text, size_t len, zo_schema* out) {
  enum { MAXL = 1024 };
  memset(out, 0, sizeof(*out));
  size_t cap = 0, pos = 0;
  char row[MAXL];
  while (pos < len) {
    /* fgets semantics: at most MAXL-1 bytes, stop after '\n' */
    size_t k = 0;
    while (pos < len && k < MAXL - 1) {
      char ch = text[pos++];
      row[k++] = ch;
      if (ch == '\n') break;
    }
    row[k] = 0;
    if (!strncasecmp(row, "Field", 5)) continue; /* :105 */
    char* tab = strchr(row, '\t');
    if (!tab) {
      zo_schema_free(out);
      return ZO_ENC_DESC_FILE_MISSING_TYPE_INFO; /* :109-110 */
    }
    *tab = 0;
    if (out->ncols == cap) {
      cap = cap ? cap * 2 : 64;
      out->names = (char**)realloc(out->names, cap * sizeof(char*));
      out->types = (uint8_t*)realloc(out->types, cap);
      out->charsize = (uint16_t*)realloc(out->charsize, cap * 2);
    }
    uint32_t c = out->ncols++;
    out->names[c] = strdup(row);
    int cs = 0;
    uint8_t ty;
    ++tab;
    if (!strncmp(tab, "varchar", 7)) {
      cs = atoi(tab + 8);
      ty = ZO_VARCHAR;
    } else if (!strncmp(tab, "char", 4)) {
      cs = atoi(tab + 5);
      ty = cs == 1 ? ZO_CHAR : cs == 2 ? ZO_CHAR_2 : ZO_VARCHAR;
    } else if (!strncmp(tab, "text", 4)) ty = ZO_TEXT;
    else if (!strncmp(tab, "tinytext", 8)) ty = ZO_TINYTEXT;
    else if (!strncmp(tab, "mediumtext", 10)) ty = ZO_MEDIUMTEXT;
    else if (!strncmp(tab, "longtext", 8)) ty = ZO_LONGTEXT;
    else if (!strncmp(tab, "datetime", 8)) ty = ZO_DATETIME;
    else if (!strncmp(tab, "decimal", 7) || !strncmp(tab + 1, "decimal", 7)) ty = ZO_DECIMAL;
    else {
      int sgn = strstr(tab, "unsigned") == NULL; /* :148 */
      if (!strncmp(tab, "tinyint", 7)) ty = sgn ? ZO_TINY_SIGNED : ZO_TINY;
      else if (!strncmp(tab, "smallint", 8)) ty = sgn ? ZO_SHORT_SIGNED : ZO_SHORT;
      else if (!strncmp(tab, "bigint", 6)) ty = sgn ? ZO_LONGLONG_SIGNED : ZO_LONGLONG;
      else ty = sgn ? ZO_LONG_SIGNED : ZO_LONG;
    }
    out->types[c] = ty;
    out->charsize[c] = (uint16_t)cs; /* static_cast<USHORT>, ConvertToZDW.cpp:733 */
  }
  return ZO_OK;
}

void zo_schema_free(zo_schema* s) {
  if (!s) return;
  for (uint32_t i = 0; i < s->ncols; ++i) free(s->names ? s->names[i] : NULL);
  free(s->names);
  free(s->types);
  free(s->charsize);
  memset(s, 0, sizeof(*s));
}

/* ------------------------------------------------------------------ row reader */

typedef struct {
  const uint8_t* p;
  size_t n, pos;
} memf;

/* fgets over memory: up to size-1 bytes, stops after '\n'.  Returns bytes stored (0 == NULL). */
static size_t m_fgets(memf* f, char* dst, size_t size) {
  size_t k = 0;
  while (f->pos < f->n && k + 1 < size) {
    char ch = (char)f->p[f->pos++];
    dst[k++] = ch;
    if (ch == '\n') break;
  }
  dst[k] = 0;
  return k;
}

/* GetNextRow, getnextrow.cpp:26-84.  *row is a malloc'd buffer of *rowSize bytes that doubles
 * exactly when the reference's does; *rowSize is the header's "longest line" field. */
static size_t next_row(memf* f, char** row, uint32_t* rowSize) {
  size_t len = 0, got;
  while ((got = m_fgets(f, *row, *rowSize))) {
    len = got; /* strlen(row): embedded NULs are unsupported in the reference too */
    if (len < 2) {
      len = 0;
      continue; /* :39-43 */
    }
    int endl = (*row)[len - 1] == '\n';
    size_t e = 2;
    while (e <= len && (*row)[len - e] == '\\') ++e;
    int eol = endl && (e % 2) == 0;
    while (!eol) {
      if (len == (size_t)*rowSize - 1) { /* :57-65 */
        char* t = (char*)malloc((size_t)*rowSize * 2);
        memcpy(t, *row, len + 1);
        free(*row);
        *row = t;
        *rowSize *= 2;
      }
      got = m_fgets(f, *row + len, *rowSize - len);
      if (!got) return 0; /* :67-69 eof inside a logical line: dropped */
      len += got;
      endl = (*row)[len - 1] == '\n';
      e = 2;
      while (e <= len && (*row)[len - e] == '\\') ++e;
      eol = endl && (e % 2) == 0;
    }
    (*row)[len - 1] = 0; /* :79 */
    break;
  }
  return len;
}

/* get_next_column, ConvertToZDW.cpp:1048-1067.  `lo` is the first byte of the row buffer: the
 * reference's backward scan can step to lo[-1]; that byte (heap metadata) is never a backslash. */
static char* next_column(char* col, const char* lo) {
  col = strchr(col, '\t');
  if (col && col > lo && col[-1] == '\\') {
    char* slash = col - 2;
    while (slash >= lo && *slash == '\\') --slash;
    while (col && ((col - slash) % 2) == 0) {
      col = strchr(col + 1, '\t');
      if (col) {
        slash = col - 1;
        while (slash >= lo && *slash == '\\') --slash;
      }
    }
  }
  return col;
}

typedef struct {
  char** v;
  size_t n, cap;
} fieldvec;

static void fv_push(fieldvec* f, char* p) {
  if (f->n == f->cap) {
    f->cap = f->cap ? f->cap * 2 : 256;
    f->v = (char**)realloc(f->v, f->cap * sizeof(char*));
  }
  f->v[f->n++] = p;
}

/* ConvertToZDW::GetDataRow, ConvertToZDW.cpp:265-323 (without the -i tee). */
static size_t get_data_row(memf* f, char** row, uint32_t* rowSize, fieldvec* cols, int trim) {
  cols->n = 0;
  if (next_row(f, row, rowSize)) {
    char *col = *row, *temp;
    while (col) {
      fv_push(cols, col);
      col = next_column(col, *row);
      if (col) {
        *col = 0;
        ++col;
        if (trim) { /* :295-301 */
          temp = col - 2;
          while (temp >= *row && *temp == ' ') {
            *temp = 0;
            --temp;
          }
        }
      }
    }
    if (trim) { /* :305-313 */
      char* ff = cols->v[cols->n - 1];
      size_t l = strlen(ff);
      temp = ff + l;
      while (temp > ff && temp[-1] == ' ') {
        --temp;
        *temp = 0;
      }
    }
  }
  return cols->n;
}

/* ------------------------------------------------------------------ dictionary */

/* Dictionary (dictionary.h:30-63, dictionary.cpp:31-111): a set of unique strings ordered by
 * strcmp; offsets are assigned in sorted order at write time.  Restated as an open-addressing
 * hash set + qsort (the observable behaviour - sorted bytes and offsets - is identical). */
typedef struct {
  char* s;
  uint32_t off;
} dent;

typedef struct {
  dent* ent;
  size_t n, cap;
  uint32_t* slots; /* index+1 into ent, 0 = empty */
  size_t nslots;
  uint64_t size;   /* sum(len+1), Dictionary::size */
} dict;

static uint64_t hash_str(const char* s) {
  uint64_t h = 1469598103934665603ULL;
  for (; *s; ++s) h = (h ^ (uint8_t)*s) * 1099511628211ULL;
  return h ^ (h >> 29);
}

static void dict_clear(dict* d) {
  for (size_t i = 0; i < d->n; ++i) free(d->ent[i].s);
  free(d->ent);
  free(d->slots);
  memset(d, 0, sizeof(*d));
}

static void dict_rehash(dict* d, size_t nslots) {
  free(d->slots);
  d->slots = (uint32_t*)calloc(nslots, 4);
  d->nslots = nslots;
  for (size_t i = 0; i < d->n; ++i) {
    size_t h = hash_str(d->ent[i].s) & (nslots - 1);
    while (d->slots[h]) h = (h + 1) & (nslots - 1);
    d->slots[h] = (uint32_t)i + 1;
  }
}

static dent* dict_find(const dict* d, const char* s) {
  if (!d->nslots) return NULL;
  size_t h = hash_str(s) & (d->nslots - 1);
  while (d->slots[h]) {
    dent* e = &d->ent[d->slots[h] - 1];
    if (!strcmp(e->s, s)) return e;
    h = (h + 1) & (d->nslots - 1);
  }
  return NULL;
}

/* returns strlen+1 when the string is new (it then goes to the string heap), 0 when it was known */
static size_t dict_insert(dict* d, const char* s) { /* Dictionary::insert, dictionary.cpp:31-51 */
  if (dict_find(d, s)) return 0;
  if ((d->n + 1) * 2 > d->nslots) dict_rehash(d, d->nslots ? d->nslots * 2 : 1024);
  if (d->n == d->cap) {
    d->cap = d->cap ? d->cap * 2 : 1024;
    d->ent = (dent*)realloc(d->ent, d->cap * sizeof(dent));
  }
  d->ent[d->n].s = strdup(s);
  d->ent[d->n].off = 0;
  d->size += strlen(s) + 1;
  size_t h = hash_str(s) & (d->nslots - 1);
  while (d->slots[h]) h = (h + 1) & (d->nslots - 1);
  d->slots[h] = (uint32_t)d->n + 1;
  d->n++;
  return strlen(s) + 1;
}

/* StringHeap as far as it decides where a block ends: stringheap.cpp:23,31-59,75-86 */
#define ZO_HEAP_BLOCK_SIZE ((size_t)64 * 1024 * 1024)
typedef struct {
  size_t free_bytes;   /* freeBytesInCurrentBlock */
  uint32_t allocs;     /* heap blocks opened for the current ZDW block */
  uint32_t limit;      /* the allocation that finds the process over --mem-limit (0 = never) */
} heapsim;

/* copyToHeap: returns 1 when this insert flagged low memory */
static int heap_copy(heapsim* h, size_t len) {
  if (h->free_bytes >= len) {
    h->free_bytes -= len;
    return 0;
  }
  h->free_bytes = (len > ZO_HEAP_BLOCK_SIZE ? len : ZO_HEAP_BLOCK_SIZE) - len; /* residual of the old block is wasted */
  h->allocs++;
  return h->limit && h->allocs >= h->limit;
}

static int dent_cmp(const void* a, const void* b) { /* cstringComp, dictionary.h:30-33 */
  return strcmp(((const dent*)a)->s, ((const dent*)b)->s);
}

static uint32_t bytes_needed_u32(uint32_t v) { /* Dictionary::getBytesInOffset, dictionary.cpp:62-73 */
  uint32_t k = 1;
  while (v >= 256) {
    ++k;
    v /= 256;
  }
  return k;
}

/* Dictionary::write, dictionary.cpp:76-111 */
static uint32_t dict_write(dict* d, obuf* out) {
  if (d->size == 0) {
    ob_u8(out, 0);
    return 1; /* getBytesInOffset() of an empty dictionary: getSize()==1 -> 1 byte */
  }
  uint32_t total = (uint32_t)(d->size + 1); /* ULONG arithmetic: dictionary.h:58,62 */
  uint32_t isz = bytes_needed_u32(total);
  ob_u8(out, (uint8_t)isz);
  ob_put(out, &total, isz);
  ob_u8(out, 0);
  qsort(d->ent, d->n, sizeof(dent), dent_cmp);
  dict_rehash(d, d->nslots); /* entries moved */
  uint32_t idx = 1;
  for (size_t i = 0; i < d->n; ++i) {
    d->ent[i].off = idx;
    size_t l = strlen(d->ent[i].s) + 1;
    ob_put(out, d->ent[i].s, l);
    idx += (uint32_t)l;
  }
  assert(idx == total);
  return isz;
}

/* ------------------------------------------------------------------ encoder */

/* file header: ConvertToZDW.cpp:673-737 */
static void write_file_header(const zo_schema* s, const zo_encode_opts* o, obuf* out) {
  ob_u16(out, 11); /* CONVERT_ZDW_CURRENT_VERSION, ConvertToZDW.cpp:75 */
  uint32_t mlen = 0;
  for (uint32_t i = 0; o && i < o->nmeta; ++i) mlen += (uint32_t)(strlen(o->meta_keys[i]) + strlen(o->meta_vals[i]) + 2);
  ob_u32(out, mlen);
  for (uint32_t i = 0; o && i < o->nmeta; ++i) {
    ob_put(out, o->meta_keys[i], strlen(o->meta_keys[i]) + 1);
    ob_put(out, o->meta_vals[i], strlen(o->meta_vals[i]) + 1);
  }
  for (uint32_t c = 0; c < s->ncols; ++c) ob_put(out, s->names[c], strlen(s->names[c]) + 1);
  ob_u8(out, 0);
  ob_put(out, s->types, s->ncols);
  ob_put(out, s->charsize, (size_t)s->ncols * 2);
}

int zo_write_file_header(const zo_schema* s, const zo_encode_opts* o, uint8_t** out, size_t* out_len) {
  obuf b = {0};
  write_file_header(s, o, &b);
  *out = b.d;
  *out_len = b.len;
  return ZO_OK;
}

/* (signed char) promoted to ULONGLONG as x86 g++ does: ConvertToZDW.cpp:359,543 */
static uint64_t sx(char ch) { return (uint64_t)(int64_t)(signed char)ch; }

/* one row (its first ncols_todo columns) through the per-column part of parseInput, ConvertToZDW.cpp:338-401 */
/* Returns -1, or the column whose insert ran the string heap out of memory (hadEnoughMemory = false, :334-355): the
   columns behind it have not been looked at. */
static int pass1_fields(const zo_schema* s, const fieldvec* cols, uint32_t ncols_todo, uint8_t* minmaxset, uint64_t* cmin,
                        uint64_t* cmax, dict* uniq, heapsim* heap) {
  for (uint32_t c = 0; c < ncols_todo; ++c) {
    const char* f = cols->v[c];
    if (!f[0]) continue;
    const uint8_t t = s->types[c];
    if (is_text_like(t)) {
      minmaxset[c] = 1;
      const size_t fresh = dict_insert(uniq, f);
      if (fresh && heap && heap_copy(heap, fresh)) return (int)c;
    } else {
      uint64_t val;
      if (t == ZO_CHAR) {
        val = sx(f[0]);
        if (f[0] == '\\') val += (uint64_t)(int64_t)((int)(signed char)f[1] * 256); /* :360-361 */
      } else {
        val = strtoull(f, NULL, 10); /* :385 */
      }
      if (val > 0) {
        if (minmaxset[c]) {
          if (val > cmax[c]) cmax[c] = val;
          else if (val < cmin[c]) cmin[c] = val;
        } else {
          cmax[c] = cmin[c] = val;
          minmaxset[c] = 1;
        }
      }
    }
  }
  return -1;
}

int zo_encode_file(const zo_schema* s, const uint8_t* tsv, size_t n, const zo_encode_opts* o,
                   uint8_t** out_p, size_t* out_len, zo_encode_info* info) {
  const uint32_t nc = s->ncols;
  const int trim = o ? o->trim_trailing_spaces : 0;
  const uint32_t nplan = o ? o->nplan : 0;
  const uint32_t rpb = (o && !nplan) ? o->rows_per_block : 0;
  heapsim heap = {0, 0, (o && !nplan && !rpb) ? o->heap_blocks : 0};
  obuf out = {0};
  zo_encode_info inf;
  memset(&inf, 0, sizeof(inf));
  write_file_header(s, o, &out);

  uint32_t rowSize = 16 * 1024; /* m_LongestLine, ConvertToZDW.cpp:965 */
  char* row = (char*)malloc(rowSize);
  fieldvec cols = {0};
  uint8_t* minmaxset = (uint8_t*)calloc(nc ? nc : 1, 1);
  uint64_t* cmin = (uint64_t*)calloc(nc ? nc : 1, 8);
  uint64_t* cmax = (uint64_t*)calloc(nc ? nc : 1, 8);
  uint8_t* csize = (uint8_t*)calloc(nc ? nc : 1, 1);
  uint32_t* used = (uint32_t*)calloc(nc ? nc : 1, 4);
  uint64_t* prev = (uint64_t*)calloc(nc ? nc : 1, 8);
  dict uniq;
  memset(&uniq, 0, sizeof(uniq));
  int rc = ZO_OK;

  /* explicit block policy: total row count decides which block is the last one */
  uint64_t total_rows = 0;
  if (rpb || nplan) {
    memf f = {tsv, n, 0};
    uint32_t rs = 16 * 1024;
    char* r2 = (char*)malloc(rs);
    while (next_row(&f, &r2, &rs)) ++total_rows;
    free(r2);
  }

  memf in = {tsv, n, 0};
  uint64_t rows_done = 0;
  for (;;) {
    /* ---- pass 1: parseInput, ConvertToZDW.cpp:329-414 */
    const size_t fbegin = in.pos; /* fgetpos, :802 */
    uint32_t numRows = 0;
    memset(minmaxset, 0, nc);
    int last = 1;
    uint32_t limit = 0; /* 0 = read to EOF */
    if (rpb && rows_done + rpb < total_rows) {
      limit = rpb;
      last = 0;
    }
    uint32_t spill = 0;
    if (nplan && inf.nblocks < nplan && rows_done + o->plan_rows[inf.nblocks] < total_rows) {
      limit = o->plan_rows[inf.nblocks];
      spill = o->plan_spill ? o->plan_spill[inf.nblocks] : 0;
      last = 0;
    }
    size_t k;
    heap.free_bytes = 0; /* uniques.clear() -> StringHeap::FreeMemory, stringheap.cpp:88-100 */
    heap.allocs = 0;
    int out_of_memory = 0;
    size_t row_start = in.pos;
    while ((!limit || numRows < limit) && (k = get_data_row(&in, &row, &rowSize, &cols, trim))) {
      if (k != nc) {
        inf.bad_row = numRows + 1; /* :811 */
        rc = ZO_ENC_WRONG_NUM_OF_COLUMNS_ON_A_ROW;
        goto done;
      }
      if (pass1_fields(s, &cols, nc, minmaxset, cmin, cmax, &uniq, heap.limit ? &heap : NULL) >= 0) {
        /* IS_NOT_ENOUGH_MEMORY: the row is not counted (:404-406) and is read again for the next block (fsetpos, :867) */
        out_of_memory = 1;
        in.pos = row_start;
        break;
      }
      ++numRows;
      row_start = in.pos;
    }
    if (out_of_memory) {
      if (!numRows) { /* low on memory before a single row fit: :824-834 */
        rc = ZO_ENC_OUT_OF_MEMORY;
        goto done;
      }
      last = 0;
    }
    if (limit && numRows == limit && !last) {
      /* the interrupted row: read in full (the row buffer grows with it, getnextrow.cpp:57-65), its leading columns
         feed the dictionary and the column ranges, then it is left for the next block (fsetpos, :867) */
      const size_t keep = in.pos;
      k = get_data_row(&in, &row, &rowSize, &cols, trim);
      if (k && k != nc) {
        inf.bad_row = numRows + 1;
        rc = ZO_ENC_WRONG_NUM_OF_COLUMNS_ON_A_ROW;
        goto done;
      }
      if (k) pass1_fields(s, &cols, spill < nc ? spill : nc, minmaxset, cmin, cmax, &uniq, NULL);
      in.pos = keep;
    }
    if (!numRows) break; /* :824-835 "Empty data file -- nothing to process" (only possible on block 1) */

    /* ---- block header, :839-842 */
    ob_u32(&out, numRows);
    ob_u32(&out, rowSize);
    ob_u8(&out, (uint8_t)last);

    /* ---- dictionary, :848 */
    inf.dict_entries = uniq.n;
    const uint32_t offsetSize = dict_write(&uniq, &out);

    /* ---- writeLookupColumnStats, :417-483 */
    uint32_t nused = 0;
    for (uint32_t c = 0; c < nc; ++c) {
      if (!minmaxset[c]) {
        csize[c] = 0;
        continue;
      }
      if (is_text_like(s->types[c])) {
        csize[c] = (uint8_t)offsetSize;
        cmin[c] = 0;
      } else {
        --cmin[c];
        uint64_t v = cmax[c] - cmin[c];
        csize[c] = 1;
        while (v >= 256) {
          ++csize[c];
          v /= 256;
        }
      }
      used[nused++] = c;
    }
    ob_put(&out, csize, nc);
    for (uint32_t u = 0; u < nused; ++u) ob_u64(&out, cmin[used[u]]);

    /* ---- pass 2: writeBlockRows, :486-606 */
    in.pos = fbegin; /* fsetpos, :867 */
    const size_t nflag = (nused + 7) / 8;
    uint8_t* flags = (uint8_t*)malloc(nflag ? nflag : 1);
    uint8_t* vals = (uint8_t*)malloc((size_t)nused * 8 + 1);
    memset(prev, 0, (size_t)nc * 8);
    uint32_t cnt = 0;
    while (cnt < numRows && get_data_row(&in, &row, &rowSize, &cols, trim) > 0) {
      size_t p = 0;
      memset(flags, 0, nflag);
      for (uint32_t u = 0; u < nused; ++u) {
        const uint32_t c = used[u];
        const char* f = cols.v[c];
        const uint8_t t = s->types[c];
        uint64_t v;
        if (is_text_like(t)) {
          if (f[0]) {
            dent* e = dict_find(&uniq, f);
            assert(e);
            v = e->off;
          } else v = 0;
        } else if (t == ZO_CHAR) {
          v = sx(f[0]);
          if (v) {
            v += (uint64_t)(int64_t)((int)(signed char)f[1] * 256); /* :545 */
            v -= cmin[c];
          }
        } else {
          v = strtoull(f, NULL, 10);
          if (v > 0) v -= cmin[c];
        }
        if (v != prev[c]) {
          flags[u / 8] |= (uint8_t)(1u << (u % 8));
          memcpy(vals + p, &v, csize[c]); /* low columnSize bytes, little-endian */
          p += csize[c];
        }
        prev[c] = v;
      }
      ob_put(&out, flags, nflag);
      ob_put(&out, vals, p);
      ++cnt;
    }
    free(flags);
    free(vals);
    dict_clear(&uniq); /* :880 */
    inf.total_rows += cnt;
    inf.nblocks++;
    rows_done += cnt;
    if (last) break;
  }
done:
  inf.longest_line = rowSize;
  if (info) *info = inf;
  dict_clear(&uniq);
  free(row);
  free(cols.v);
  free(minmaxset);
  free(cmin);
  free(cmax);
  free(csize);
  free(used);
  free(prev);
  if (rc != ZO_OK) {
    free(out.d);
    *out_p = NULL;
    *out_len = 0;
    return rc;
  }
  *out_p = out.d;
  *out_len = out.len;
  return ZO_OK;
}

/* ------------------------------------------------------------------ decoder */

typedef struct {
  const uint8_t* p;
  size_t n, pos;
  int fail;
} rd;

static int rd_bytes(rd* r, void* dst, size_t k) { /* readBytes, UnconvertFromZDW.cpp:286-302 */
  if (r->n - r->pos < k) {
    r->pos = r->n;
    r->fail = 1;
    return 0;
  }
  memcpy(dst, r->p + r->pos, k);
  r->pos += k;
  return 1;
}

/* llutoa / lltoa, UnconvertFromZDW.cpp:318-356 -- including the INT64_MIN behaviour of lltoa:
 * `value = -value` overflows and every `value % 10` is then negative. */
static size_t fmt_u64(uint64_t v, char* end) {
  char* p = end;
  do {
    *--p = (char)(v % 10 + 0x30);
    v /= 10;
  } while (v);
  return (size_t)(end - p);
}
static size_t fmt_i64(int64_t v, char* end) {
  char* p = end;
  int minus = 0;
  if (v < 0) {
    minus = 1;
    v = (int64_t)(0 - (uint64_t)v); /* wraps for INT64_MIN exactly like the reference's -value on x86 */
  }
  do {
    int64_t rem = v % 10; /* negative for INT64_MIN */
    v /= 10;
    *--p = (char)((size_t)rem + 0x30);
  } while (v != 0);
  if (minus) *--p = '-';
  return (size_t)(end - p);
}

int zo_read_file_header(const uint8_t* zdw, size_t n, zo_schema* s, uint16_t* version, size_t* hdr_len) {
  rd r = {zdw, n, 0, 0};
  memset(s, 0, sizeof(*s));
  uint16_t ver = 0;
  if (!rd_bytes(&r, &ver, 2)) return ZO_DEC_GZREAD_FAILED;
  if (version) *version = ver;
  if (ver > 11) return ZO_DEC_UNSUPPORTED_ZDW_VERSION_ERR; /* :1044 */
  if (ver < 9) return ZO_DEC_UNSUPPORTED_ZDW_VERSION_ERR;  /* legacy v1-v8: out of scope (SURVEY s.2 row 2) */
  if (ver >= 11) { /* :1051-1070 */
    uint32_t ml = 0;
    if (!rd_bytes(&r, &ml, 4)) return ZO_DEC_GZREAD_FAILED;
    if (r.n - r.pos < ml) return ZO_DEC_GZREAD_FAILED;
    r.pos += ml;
  }
  /* column names: NUL-terminated, list ends with an empty name (:1086-1096) */
  size_t cap = 0;
  for (;;) {
    const uint8_t* z = (const uint8_t*)memchr(zdw + r.pos, 0, n - r.pos);
    if (!z) {
      zo_schema_free(s);
      return ZO_DEC_GZREAD_FAILED;
    }
    size_t l = (size_t)(z - (zdw + r.pos));
    if (l == 0) {
      r.pos++;
      break;
    }
    if (s->ncols == cap) {
      cap = cap ? cap * 2 : 64;
      s->names = (char**)realloc(s->names, cap * sizeof(char*));
    }
    s->names[s->ncols++] = strdup((const char*)zdw + r.pos);
    r.pos += l + 1;
  }
  s->types = (uint8_t*)malloc(s->ncols ? s->ncols : 1);
  s->charsize = (uint16_t*)malloc(s->ncols ? (size_t)s->ncols * 2 : 2);
  if (!rd_bytes(&r, s->types, s->ncols) || !rd_bytes(&r, s->charsize, (size_t)s->ncols * 2)) {
    zo_schema_free(s);
    return ZO_DEC_GZREAD_FAILED;
  }
  if (hdr_len) *hdr_len = r.pos;
  return ZO_OK;
}

static void out_default(obuf* b, uint8_t type) { /* outputDefault, UnconvertFromZDW.cpp:1224-1266 */
  switch (type) {
    case ZO_TINY: case ZO_TINY_SIGNED: case ZO_SHORT: case ZO_SHORT_SIGNED:
    case ZO_LONG: case ZO_LONG_SIGNED: case ZO_LONGLONG: case ZO_LONGLONG_SIGNED:
    case ZO_VISID_HIGH:
      ob_put(b, "0", 1);
      break;
    case ZO_DECIMAL:
      ob_put(b, "0.000000000000", 14);
      break;
    default: /* text-like and CHAR: empty */
      break;
  }
}

int zo_decode_file(const uint8_t* zdw, size_t n, const zo_decode_opts* o, uint8_t** out_p, size_t* out_len,
                   zo_decode_info* info) {
  zo_schema s;
  uint16_t ver;
  size_t hl;
  zo_decode_info inf;
  memset(&inf, 0, sizeof(inf));
  int rc = zo_read_file_header(zdw, n, &s, &ver, &hl);
  if (rc) return rc;
  inf.version = ver;
  inf.ncols = s.ncols;
  const uint32_t nc = s.ncols;
  const char sep = o ? o->sep : '\t';
  const int* ocol = o ? o->out_col : NULL;
  const uint32_t n_out = ocol ? o->n_out : nc;
  rd r = {zdw, n, hl, 0};
  obuf out = {0};
  uint8_t* csize = (uint8_t*)malloc(nc ? nc : 1);
  uint64_t* cbase = (uint64_t*)calloc(nc ? nc : 1, 8);
  uint64_t* cval = (uint64_t*)calloc(nc ? nc : 1, 8);
  obuf* colbuf = NULL; /* BufferedOrderedOutput column buffers, BufferedOutput.cpp:101-165 */
  if (ocol) colbuf = (obuf*)calloc(n_out ? n_out : 1, sizeof(obuf));
  uint8_t* flags = NULL;
  char tmp[64];
  uint8_t last = 0;

  do {
    /* ---- parseBlockHeader, :758-1000 */
    uint32_t numLines = 0, lineLen = 0;
    const size_t block_at = r.pos;
    if (!rd_bytes(&r, &numLines, 4) || !rd_bytes(&r, &lineLen, 4) || !rd_bytes(&r, &last, 1)) {
      rc = ZO_DEC_GZREAD_FAILED;
      goto done;
    }
    inf.line_length = lineLen;
    uint8_t isz = 0;
    uint64_t dsize = 0;
    if (!rd_bytes(&r, &isz, 1)) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
    if (isz) {
      uint32_t v = 0;
      if (isz > 4 || !rd_bytes(&r, &v, isz)) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
      dsize = v;
    }
    if (r.n - r.pos < dsize) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
    const char* dictp = (const char*)zdw + r.pos; /* offset 0 = the origin byte */
    r.pos += dsize;
    if (!rd_bytes(&r, csize, nc)) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
    uint32_t nused = 0;
    for (uint32_t c = 0; c < nc; ++c) {
      if (csize[c]) {
        /* the reference reads columnSize[c] bytes into an 8-byte columnVal (:1299-1308) whatever the size says: undefined
           behaviour for a size above 8, which only a corrupt file has.  Reported, not restated. */
        if (csize[c] > 8) { rc = ZO_DEC_CORRUPTED_DATA_ERROR; goto done; }
        if (!rd_bytes(&r, &cbase[c], 8)) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
        ++nused;
      } else cbase[c] = 0;
    }
    const size_t nflag = (nused + 7) / 8;
    free(flags);
    flags = (uint8_t*)malloc(nflag ? nflag : 1);
    memset(cval, 0, (size_t)nc * 8);

    /* ---- rows: readNextRow, :1270-1464 */
    if (inf.nblocks < 16) {
      inf.block_rows[inf.nblocks] = numLines;
      inf.block_offset[inf.nblocks] = block_at;
    }
    for (uint32_t rr = 0; rr < numLines; ++rr) {
      /* `while (rowsRead < numLines && !isFinished())`, :1577: input->eof() is true once every byte has been
       * consumed, so rows of zero bytes (no used column) at the very end of the file are never read */
      if (r.pos >= r.n) { rc = ZO_DEC_ROW_COUNT_ERR; goto done; }
      if (!rd_bytes(&r, flags, nflag)) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
      uint32_t u = 0;
      int written = 0;
      for (uint32_t c = 0; c < nc; ++c) {
        const uint8_t ct = s.types[c];
        const int ignore = ocol && ocol[c] < 0;
        obuf* dst = ocol ? (ignore ? NULL : &colbuf[ocol[c]]) : &out;
        if (csize[c]) {
          if (flags[u / 8] & (1u << (u % 8))) {
            cval[c] = 0;
            if (!rd_bytes(&r, &cval[c], csize[c])) { rc = ZO_DEC_GZREAD_FAILED; goto done; }
          }
          ++u;
        }
        if (ignore) continue;
        if (!ocol && written) ob_u8(&out, (uint8_t)sep);
        if (ocol) dst->len = 0;
        written = 1;
        if (!csize[c]) {
          out_default(dst, ct);
          continue;
        }
        const uint64_t v = cval[c];
        switch (ct) {
          case ZO_VARCHAR: case ZO_TEXT: case ZO_TINYTEXT: case ZO_MEDIUMTEXT: case ZO_LONGTEXT:
          case ZO_DATETIME: case ZO_CHAR_2: case ZO_DECIMAL:
            if (v) {
              uint32_t index = (uint32_t)(v + cbase[c]); /* ULONG index, :1363 */
              if (index > dsize) { rc = ZO_DEC_CORRUPTED_DATA_ERROR; goto done; }
              ob_put(dst, dictp + index, strnlen(dictp + index, (size_t)(dsize - index)));
            } else out_default(dst, ct);
            break;
          case ZO_CHAR: /* :1396-1420 */
            if (v) {
              const uint64_t t = v + cbase[c];
              tmp[0] = (char)t;
              if (tmp[0] != '\\') {
                if (tmp[0]) ob_put(dst, tmp, 1);
              } else {
                tmp[1] = (char)(t / 256);
                ob_put(dst, tmp, 2);
              }
            }
            break;
          case ZO_TINY: case ZO_SHORT: case ZO_LONG: case ZO_LONGLONG: {
            size_t l = fmt_u64(v ? v + cbase[c] : 0, tmp + 63);
            ob_put(dst, tmp + 63 - l, l);
          } break;
          case ZO_TINY_SIGNED: case ZO_SHORT_SIGNED: case ZO_LONG_SIGNED: case ZO_LONGLONG_SIGNED: {
            size_t l = fmt_i64((int64_t)(v ? v + cbase[c] : 0), tmp + 63);
            ob_put(dst, tmp + 63 - l, l);
          } break;
          default:
            rc = ZO_DEC_CORRUPTED_DATA_ERROR; /* VISID_* only exist before v8 */
            goto done;
        }
      }
      if (ocol) { /* BufferedOrderedOutput::writeEndline, BufferedOutput.cpp:139-165 */
        for (uint32_t k = 0; k < n_out; ++k) {
          if (k) ob_u8(&out, (uint8_t)'\t');
          ob_put(&out, colbuf[k].d, colbuf[k].len);
        }
        ob_u8(&out, '\n');
      } else {
        ob_u8(&out, sep == '\t' ? '\n' : 0);
      }
      inf.total_rows++;
    }
    inf.nblocks++;
  } while (!last);
  inf.consumed = r.pos;
  if (r.pos != r.n) rc = ZO_DEC_ZDW_LONGER_THAN_EXPECTED_ERR; /* :1823-1834 */
done:
  if (info) *info = inf;
  free(csize);
  free(cbase);
  free(cval);
  free(flags);
  if (colbuf) {
    for (uint32_t k = 0; k < n_out; ++k) free(colbuf[k].d);
    free(colbuf);
  }
  zo_schema_free(&s);
  if (rc != ZO_OK && rc != ZO_DEC_ZDW_LONGER_THAN_EXPECTED_ERR) {
    free(out.d);
    *out_p = NULL;
    *out_len = 0;
    return rc;
  }
  *out_p = out.d;
  *out_len = out.len;
  return rc;
}
