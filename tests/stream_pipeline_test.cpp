// stream_pipeline_test -- TEST INFRASTRUCTURE: the read-ahead window and the writer thread of convertDWfile
// (zdw_b200/host/stream_pipeline.h) on the CPU, with malloc as the allocator.  Exit code 0 = all good.
//   ReadAheadInput vs the sequential window it replaced (fill to `cap` or end of stream, memmove on consume): same
//   window bytes, length and end-of-stream flag at every step, for random consumption, random use of prefetch() and
//   widen(), streams that end exactly on a window boundary, and a tee that must receive the stream once, in order.
//   The same over a pipe with a slow producer, and close() with a read-ahead waiting on a pipe that stays silent.
//   AsyncWriter: blocks arrive in order and complete.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "stream_pipeline.h"

using adobe::zdw::AsyncWriter;
using adobe::zdw::ReadAheadInput;

static int bad = 0;
#define CHECK(cond, ...)      \
  do {                        \
    if (!(cond)) {            \
      if (++bad <= 20) {      \
        printf("MISMATCH: "); \
        printf(__VA_ARGS__);  \
        printf("\n");         \
      }                       \
    }                         \
  } while (0)

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return rng_state;
}

// the window ConvertToZDW.cpp used before: one buffer, topped up to cap, memmove on consume
struct SeqWindow {
  std::vector<char> buf;
  size_t cap = 0, len = 0;
  bool eof = false;
  void reserve(size_t want) {
    if (want > cap) {
      buf.resize(want);
      cap = want;
    }
  }
  void fill(FILE* in) {
    while (!eof && len < cap) {
      const size_t got = fread(buf.data() + len, 1, cap - len, in);
      len += got;
      if (got == 0) eof = true;
    }
  }
  void consume(size_t n) {
    if (n >= len) {
      len = 0;
    } else {
      memmove(buf.data(), buf.data() + n, len - n);
      len -= n;
    }
  }
};

static FILE* file_with(const std::string& bytes) {
  FILE* f = tmpfile();
  fwrite(bytes.data(), 1, bytes.size(), f);
  rewind(f);
  return f;
}

// mode: 0 = consume nearly everything (the block-bytes case), 1 = anything, 2 = little (planned rows), 3 = nothing
// now and then (the caller widens the window)
static void run_case(size_t total, size_t cap, int mode, int prefetch_pct) {
  std::string bytes(total, 0);
  for (size_t i = 0; i < total; ++i) bytes[i] = (char)(rnd() >> 56);
  FILE* fa = file_with(bytes);
  FILE* fb = file_with(bytes);
  FILE* tee = tmpfile();
  SeqWindow seq;
  seq.reserve(cap);
  ReadAheadInput win;
  CHECK(win.open(fb, tee, cap, malloc, free), "open failed");
  size_t steps = 0, delivered = 0;
  for (;;) {
    seq.fill(fa);
    win.fill();
    CHECK(seq.len == win.len(), "len %zu vs %zu (total %zu cap %zu mode %d step %zu)", seq.len, win.len(), total, cap, mode, steps);
    CHECK(seq.eof == win.eof(), "eof %d vs %d (total %zu cap %zu mode %d step %zu)", (int)seq.eof, (int)win.eof(), total, cap, mode, steps);
    if (seq.len != win.len()) break;
    CHECK(memcmp(seq.buf.data(), win.data(), seq.len) == 0, "window bytes differ (total %zu cap %zu mode %d step %zu)", total, cap, mode, steps);
    CHECK(memcmp(bytes.data() + delivered, win.data(), win.len()) == 0, "window is not the stream at %zu", delivered);
    if (seq.len == 0 && seq.eof) break;
    if ((int)(rnd() % 100) < prefetch_pct) win.prefetch();
    size_t n;
    const size_t len = seq.len;
    switch (mode) {
      case 0: n = len - std::min<size_t>(len, rnd() % (cap / 16 + 2)); break;
      case 1: n = rnd() % (len + 1); break;
      case 2: n = std::min<size_t>(len, 1 + rnd() % (cap / 8 + 1)); break;
      default: n = (rnd() % 4 == 0) ? 0 : len - std::min<size_t>(len, rnd() % (cap / 4 + 2)); break;
    }
    if (seq.eof && n == 0) n = len;  // the callers stop at a window that ends the stream and yields nothing
    if (n == 0) {  // "not one complete row": widen and try again
      cap = cap * 2;
      seq.reserve(cap);
      CHECK(win.widen(cap), "widen failed");
      CHECK(win.windowBytes() == cap, "windowBytes");
    } else {
      seq.consume(n);
      win.consume(n);
      delivered += n;
    }
    if (++steps > 100000) {
      CHECK(false, "no progress");
      break;
    }
  }
  CHECK(delivered == total, "delivered %zu of %zu (cap %zu mode %d)", delivered, total, cap, mode);
  win.close();
  // the tee saw the stream exactly once, in order
  fflush(tee);
  rewind(tee);
  std::string teed(total + 16, 0);
  const size_t got = fread(&teed[0], 1, teed.size(), tee);
  CHECK(got == total && memcmp(teed.data(), bytes.data(), total) == 0, "tee: %zu of %zu bytes or wrong bytes", got, total);
  fclose(fa);
  fclose(fb);
  fclose(tee);
}

static void writer_case(size_t nblocks) {
  int fds[2];
  if (pipe(fds)) {
    CHECK(false, "pipe");
    return;
  }
  FILE* out = fdopen(fds[1], "w");
  std::string want, got;
  std::thread reader([&]() {  // a slow consumer, like a compressor
    char buf[4096];
    ssize_t n;
    while ((n = read(fds[0], buf, sizeof(buf))) > 0) {
      got.append(buf, (size_t)n);
      if (rnd() % 64 == 0) usleep(200);
    }
  });
  {
    AsyncWriter w;
    w.start(out);
    for (size_t b = 0; b < nblocks; ++b) {
      std::vector<unsigned char> blk(1 + (b * 7919u) % 300000u);
      for (size_t i = 0; i < blk.size(); ++i) blk[i] = (unsigned char)(b + i * 31u);
      want.append(reinterpret_cast<const char*>(blk.data()), blk.size());
      w.push(blk);
      CHECK(blk.empty(), "push leaves the vector empty");
    }
    w.finish();
    w.finish();  // idempotent
  }
  fclose(out);
  reader.join();
  close(fds[0]);
  CHECK(want == got, "writer: %zu bytes written, %zu expected", got.size(), want.size());
}

// a pipe with a producer that dawdles: read-ahead must deliver the stream unchanged while the producer runs on
static void pipe_case(size_t total, size_t cap) {
  int fds[2];
  if (pipe(fds)) {
    CHECK(false, "pipe");
    return;
  }
  std::string bytes(total, 0);
  for (size_t i = 0; i < total; ++i) bytes[i] = (char)(rnd() >> 56);
  std::thread producer([&]() {
    size_t at = 0;
    while (at < total) {
      const size_t n = std::min<size_t>(total - at, 1 + (at * 2654435761u) % 70000u);
      const ssize_t w = write(fds[1], bytes.data() + at, n);
      if (w <= 0) break;
      at += (size_t)w;
      if ((at >> 12) % 7 == 0) usleep(300);
    }
    close(fds[1]);
  });
  FILE* in = fdopen(fds[0], "r");
  {
    ReadAheadInput win;
    CHECK(win.open(in, NULL, cap, malloc, free), "open failed");
    size_t delivered = 0, steps = 0;
    for (;;) {
      win.fill();
      CHECK(delivered + win.len() <= total && memcmp(bytes.data() + delivered, win.data(), win.len()) == 0,
            "pipe: window is not the stream at %zu", delivered);
      if (win.len() == 0 && win.eof()) break;
      win.prefetch();
      size_t n = win.len() - std::min<size_t>(win.len(), rnd() % (cap / 16 + 2));
      if (n == 0) n = win.len();
      win.consume(n);
      delivered += n;
      if (++steps > 100000) {
        CHECK(false, "pipe: no progress");
        break;
      }
    }
    CHECK(delivered == total, "pipe: delivered %zu of %zu", delivered, total);
  }
  producer.join();
  fclose(in);
}

// a read-ahead that waits for a producer which never writes again: close() must come back (the error paths of
// convertDWfile -i rely on it)
static void pipe_cancel_case() {
  int fds[2];
  if (pipe(fds)) {
    CHECK(false, "pipe");
    return;
  }
  const size_t cap = 4096;
  std::string first(cap, 'x');
  CHECK(write(fds[1], first.data(), cap) == (ssize_t)cap, "pipe write");
  FILE* in = fdopen(fds[0], "r");
  struct timespec t0, t1;
  {
    ReadAheadInput win;
    CHECK(win.open(in, NULL, cap, malloc, free), "open failed");
    win.fill();
    CHECK(win.len() == cap && !win.eof(), "cancel: first window");
    win.prefetch();  // nothing more will ever arrive, and the write end stays open
    usleep(50000);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    win.close();
    clock_gettime(CLOCK_MONOTONIC, &t1);
  }
  const double s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  CHECK(s < 2.0, "cancel: close() took %.2f s", s);
  close(fds[1]);
  fclose(in);
}

int main() {
  const size_t caps[] = {1, 2, 7, 8, 9, 64, 100, 1000, 4096, 65536, 1000003};
  int cases = 0;
  for (size_t cap : caps) {
    const size_t totals[] = {0, 1, cap - 1, cap, cap + 1, 2 * cap, 2 * cap + 1, 3 * cap - 1, 5 * cap + cap / 2, 17 * cap + 3};
    for (size_t total : totals) {
      if (total > (size_t)40 << 20) continue;
      for (int mode = 0; mode < 4; ++mode) {
        for (int pct : {0, 50, 100}) {
          run_case(total, cap, mode, pct);
          ++cases;
        }
      }
    }
  }
  pipe_case(0, 4096);
  pipe_case(3000000, 65536);
  pipe_case(1 << 20, 1 << 18);
  pipe_cancel_case();
  writer_case(0);
  writer_case(1);
  writer_case(40);
  printf("%d window cases, %d mismatches\n", cases, bad);
  return bad ? 1 : 0;
}
