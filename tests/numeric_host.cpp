// numeric_host -- TEST INFRASTRUCTURE: the numeric helpers the kernels share with the host (zdw_b200/csrc/common.cuh)
// against libc, on the CPU.  Exit code 0 = all good; mismatches are printed.
//   parse_u64_field  vs strtoull on a NUL-terminated copy of the field       (ConvertToZDW.cpp:385,564)
//   char_tuple       vs the two CHAR rules of the reference                  (ConvertToZDW.cpp:358-361, :543-547)
//   bytes_needed     vs the loop of writeLookupColumnStats                   (ConvertToZDW.cpp:458-465)
//   fmt_u64 / fmt_i64 / render_int vs llutoa / lltoa incl. the INT64_MIN quirk (UnconvertFromZDW.cpp:318-356; App. B-22)
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "common.cuh"

using namespace zdwb;

static int bad = 0;
#define CHECK(cond, ...)            \
  do {                              \
    if (!(cond)) {                  \
      if (++bad <= 20) {            \
        printf("MISMATCH: ");       \
        printf(__VA_ARGS__);        \
        printf("\n");               \
      }                             \
    }                               \
  } while (0)

// lltoa as the reference wrote it (value = -value overflows for INT64_MIN, remainders then stay negative)
static std::string ref_lltoa(long long sv) {
  unsigned char tmp[48];
  int p = 48;
  bool minus = false;
  if (sv < 0) {
    minus = true;
    sv = (long long)(0 - (unsigned long long)sv);
  }
  do {
    long long rem = sv % 10;
    sv /= 10;
    tmp[--p] = (unsigned char)(rem + 0x30);
  } while (sv != 0);
  if (minus) tmp[--p] = '-';
  return std::string((const char*)tmp + p, 48 - p);
}

static uint32_t digits10(unsigned long long v) {
  uint32_t n = 1;
  while (v >= 10) {
    v /= 10;
    ++n;
  }
  return n;
}

static void check_render(unsigned long long full, bool is_signed) {
  const bool neg = is_signed && (long long)full < 0;
  std::string want;
  if (is_signed) {
    want = ref_lltoa((long long)full);
  } else {
    char b[32];
    snprintf(b, sizeof b, "%llu", full);
    want = b;
  }
  const uint32_t len = neg ? 1u + digits10(0ull - full) : digits10(full);
  CHECK(len == want.size(), "length of %llu signed=%d: %u vs %zu", full, (int)is_signed, len, want.size());
  uint32_t w[5];
  render_int(full, neg, len, w);
  CHECK(!memcmp(w, want.data(), want.size()), "render_int(%llu, signed=%d) = %.*s, want %s", full, (int)is_signed, (int)len,
        (const char*)w, want.c_str());
  uint8_t buf[32];
  const uint32_t n = is_signed ? fmt_i64((int64_t)full, buf + 32) : fmt_u64(full, buf + 32);
  CHECK(n == want.size() && !memcmp(buf + 32 - n, want.data(), n), "fmt of %llu signed=%d", full, (int)is_signed);
}

// the register fast path of pass 1: the field as five masked little-endian words
static void check_fast(const char* s, size_t n) {
  if (n == 0 || n > NUM_FAST_MAX || memchr(s, 0, n)) return;
  uint32_t x[5] = {0, 0, 0, 0, 0};
  memcpy(x, s, n);
  unsigned long long got = 0;
  const bool ok = fast_number(x, (uint32_t)n, &got);
  std::string z(s, n);
  const unsigned long long want = strtoull(z.c_str(), NULL, 10);
  // the fast path may decline (returns false) but must never accept and disagree; plain digit strings must be accepted
  if (ok) CHECK(got == want, "fast_number(\"%s\") = %llu, strtoull = %llu", z.c_str(), got, want);
  bool digits = true;
  for (size_t k = 0; k < n; ++k) digits = digits && s[k] >= '0' && s[k] <= '9';
  if (digits) CHECK(ok, "fast_number declined the digit string \"%s\"", z.c_str());
}

static void check_parse(const char* s, size_t n) {
  check_fast(s, n);
  std::string z(s, n);  // the reference NUL-terminates every field
  const unsigned long long want = strtoull(z.c_str(), NULL, 10);
  const unsigned long long got = parse_u64_field((const uint8_t*)s, (uint32_t)n);
  // strtoull stops at an embedded NUL; fields never contain one
  if (memchr(s, 0, n)) return;
  CHECK(got == want, "parse_u64_field(\"%s\") = %llu, strtoull = %llu", z.c_str(), got, want);
}

int main() {
  // ---- ascii4
  for (uint32_t x = 0; x < 10000; ++x) {
    char b[8];
    snprintf(b, sizeof b, "%04u", x);
    const uint32_t w = ascii4(x);
    CHECK(!memcmp(&w, b, 4), "ascii4(%u)", x);
  }
  // ---- decimal text: powers of ten and their neighbours, both signs, the extremes, random values of every width
  unsigned long long p = 1;
  for (int d = 0; d < 20; ++d) {
    for (long long k = -3; k <= 3; ++k) {
      const unsigned long long v = p + (unsigned long long)k;
      check_render(v, false);
      check_render(v, true);
      check_render(0ull - v, false);
      check_render(0ull - v, true);
    }
    if (d < 19) p *= 10;
  }
  const unsigned long long edge[] = {0ull, 1ull, 0x7fffffffffffffffull, 0x8000000000000000ull, 0x8000000000000001ull, ~0ull,
                                     0xffffffffull, 0x100000000ull, 99999999ull, 100000000ull, 9999999999999999ull, 10000000000000000ull};
  for (unsigned long long v : edge) {
    check_render(v, false);
    check_render(v, true);
  }
  srand(20190901);
  for (int i = 0; i < 1000000; ++i) {
    unsigned long long v = ((unsigned long long)rand() << 40) ^ ((unsigned long long)rand() << 20) ^ (unsigned long long)rand();
    v >>= rand() % 64;
    check_render(v, false);
    check_render(v, true);
  }
  // ---- strtoull semantics (corpus d5 and friends)
  const char* fields[] = {"", "0", "007", "+4", "-5", "  12", "12abc", "3.9", "18446744073709551615", "18446744073709551616",
                          "-18446744073709551616", "-0", "+", "-", "0x10", "1e3", " -7 ", "\v13", "1\r", "\t\n 42", "--5", "+-5",
                          "99999999999999999999999", "-9223372036854775808", "9223372036854775808", "18446744073709551614",
                          "1844674407370955161", "18446744073709551609", "00000000000000000000001", " ", "\f\r9"};
  for (const char* f : fields) check_parse(f, strlen(f));
  for (int nd = 1; nd <= 20; ++nd) {  // every length, random digits, with and without a sign; the 2^64 boundary
    for (int i = 0; i < 20000; ++i) {
      char b[32];
      int at = 0;
      const int sg = rand() % 3;
      if (sg && nd < 20) b[at++] = sg == 1 ? '-' : '+';
      for (int k = 0; k < nd && at < 20; ++k) b[at++] = (char)('0' + rand() % 10);
      if (i % 7 == 0 && at == 20) memcpy(b, "1844674407370955161", 19), b[19] = (char)('0' + rand() % 10);
      check_parse(b, (size_t)at);
    }
  }
  for (int i = 0; i < 300000; ++i) {
    char b[32];
    const int n = rand() % 24;
    for (int k = 0; k < n; ++k) {
      const int r = rand() % 16;
      b[k] = r < 11 ? (char)('0' + rand() % 10) : r == 11 ? ' ' : r == 12 ? '-' : r == 13 ? '+' : r == 14 ? 'x' : '\t';
    }
    check_parse(b, (size_t)n);
  }
  // ---- CHAR cells: sign-extended first byte, second byte * 256 after a backslash (pass 1) or always (pass 2)
  for (int a = 1; a < 256; ++a) {
    for (int b2 = 0; b2 < 256; b2 += 5) {
      const uint8_t f[2] = {(uint8_t)a, (uint8_t)b2};
      const uint32_t len = b2 ? 2u : 1u;
      const long long v0 = (long long)(signed char)f[0];
      const long long v1 = v0 + (f[0] == '\\' ? (long long)((int)(signed char)f[1] * 256) : 0);
      const long long v2 = v0 + (long long)((int)(signed char)f[1] * 256);
      CHECK(char_tuple(f, len, false) == (uint64_t)v1, "char_tuple pass 1 %d %d", a, b2);
      CHECK(char_tuple(f, len, true) == (uint64_t)v2, "char_tuple pass 2 %d %d", a, b2);
    }
  }
  // ---- bytes needed for a value
  for (int sh = 0; sh < 64; ++sh) {
    for (long long k = -1; k <= 1; ++k) {
      const unsigned long long v = (1ull << sh) + (unsigned long long)k;
      uint32_t want = 1;
      for (unsigned long long t = v; t >= 256; t /= 256) ++want;
      CHECK(bytes_needed(v) == want, "bytes_needed(%llu)", v);
    }
  }
  printf("numeric_host: %d mismatches\n", bad);
  return bad != 0;
}
