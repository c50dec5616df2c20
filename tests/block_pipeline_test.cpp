// block_pipeline_test.cpp -- the host-side window cuts of the multi-GPU encoder (zdw_b200/host/block_pipeline.h) against a
// plain sequential walk over the same bytes, the ordered hand-over of results that finish out of order, and the sink that
// puts decoded blocks out in file order (pipe: in turn; regular file: claimed places written side by side).
// CPU only (g++), no CUDA.  Prints "<n> mismatches".
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <random>
#include <string>
#include <thread>

#include "block_pipeline.h"

using namespace adobe::zdw;

// reference: the byte behind the last newline preceded by an even number of backslashes, walking forward
static size_t slowLastBreak(const std::string& s, size_t a, size_t n) {
  size_t best = 0, run = 0;
  for (size_t i = 0; i < n; ++i) {
    const char c = s[a + i];
    if (c == '\\') {
      ++run;
      continue;
    }
    if (c == '\n' && (run & 1u) == 0) best = i + 1;
    run = 0;
  }
  return best;
}

int main() {
  int bad = 0;
  std::mt19937_64 rng(20190901);
  char name[] = "/tmp/zdw_bp_XXXXXX";
  const int fd = mkstemp(name);
  if (fd < 0) return 2;
  for (int round = 0; round < 300; ++round) {
    // rows with escaped newlines / tabs and runs of backslashes, a few blank lines, sometimes no final newline
    std::string data;
    const size_t target = 2000 + rng() % 60000;
    while (data.size() < target) {
      const size_t len = rng() % 400;
      for (size_t i = 0; i < len; ++i) {
        const unsigned r = rng() % 40;
        if (r == 0) data += "\\\n";
        else if (r == 1) data += "\\\\";
        else if (r == 2) data += std::string(1 + rng() % 5, '\\') + (rng() % 2 ? "\n" : "x");
        else if (r == 3) data += '\t';
        else data += (char)('a' + rng() % 26);
      }
      data += '\n';
      if (rng() % 17 == 0) data += "\n\n";
    }
    if (rng() % 3 == 0) data += "tail without newline";
    if (ftruncate(fd, 0) != 0 || pwrite(fd, data.data(), data.size(), 0) != (ssize_t)data.size()) return 2;
    size_t cap = 64 + rng() % 3000;
    if (round % 50 == 0) cap = data.size();  // the file ends exactly at the window's last byte
    std::vector<FileWindow> wins;
    size_t finalCap = 0;
    const bool ok = planFileWindows(fd, data.size(), cap, (size_t)1 << 30, wins, &finalCap);
    // the sequential walk
    std::vector<FileWindow> want;
    size_t c = cap, s = 0;
    for (;;) {
      const size_t left = data.size() - s, len = std::min(c, left);
      const bool eof = left < c;
      if (len == 0) break;
      FileWindow w = {s, len, len, !eof};
      if (eof) {
        want.push_back(w);
        break;
      }
      const size_t cut = slowLastBreak(data, s, len);
      if (!cut) {
        c *= 2;
        continue;
      }
      w.consumed = cut;
      want.push_back(w);
      s += cut;
    }
    bool same = ok && wins.size() == want.size() && finalCap == c;
    for (size_t k = 0; same && k < wins.size(); ++k)
      same = wins[k].offset == want[k].offset && wins[k].len == want[k].len && wins[k].consumed == want[k].consumed &&
             wins[k].more == want[k].more;
    if (!same) {
      ++bad;
      printf("round %d: cap %zu: %zu windows, wanted %zu\n", round, cap, wins.size(), want.size());
    }
    // findLastRowBreak on whole windows
    for (size_t k = 0; k < want.size(); ++k)
      if (want[k].more && findLastRowBreak(data.data() + want[k].offset, want[k].len) != want[k].consumed) ++bad;
  }
  close(fd);
  unlink(name);

  // results put out of order by four threads are taken in order; nobody runs more than `ahead` jobs in front
  {
    const size_t n = 500, ahead = 6;
    OrderedResults<std::string> res(n);
    std::atomic<size_t> next(0), maxLead(0), taken(0);
    auto work = [&](unsigned seed) {
      std::mt19937 r(seed);
      for (;;) {
        const size_t k = next.fetch_add(1);
        if (k >= n) break;
        res.waitTurn(k, ahead);
        const size_t lead = k - taken.load();
        size_t m = maxLead.load();
        while (lead > m && !maxLead.compare_exchange_weak(m, lead)) {}
        if (r() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(r() % 300));
        res.put(k, "job " + std::to_string(k));
      }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < 4; ++t) th.emplace_back(work, 100 + t);
    for (size_t k = 0; k < n; ++k) {
      if (res.take(k) != "job " + std::to_string(k)) ++bad;
      taken = k + 1;
    }
    for (auto& t : th) t.join();
    if (maxLead.load() > ahead + 4) {  // (taken is updated a moment after take() returns)
      ++bad;
      printf("lead %zu\n", maxLead.load());
    }
  }
  // ---- OrderedSink: blocks decoded by several workers in any order leave in file order
  for (int mode = 0; mode < 2; ++mode) {  // 0: a regular file (claim a place, write side by side), 1: a pipe (written in turn)
    for (int round = 0; round < 20; ++round) {
      const size_t nblocks = 1 + rng() % 40;
      const size_t skipAt = (round % 5 == 4) ? rng() % nblocks : nblocks;  // one worker "fails": nothing behind it counts
      std::vector<std::string> prefix(nblocks), rows(nblocks);
      std::string want = "HEAD";
      for (size_t k = 0; k < nblocks; ++k) {
        if (rng() % 3 == 0) prefix[k] = "#header of block " + std::to_string(k) + "\n";
        rows[k].assign(rng() % 5000, (char)('a' + k % 26));
        rows[k] += "\n";
        if (k < skipAt) want += prefix[k] + rows[k];
      }
      FILE* f = NULL;
      int pfd[2] = {-1, -1};
      std::string got;
      std::thread drain;
      char oname[] = "/tmp/zdw_os_XXXXXX";
      if (mode == 0) {
        const int ofd = mkstemp(oname);
        f = fdopen(ofd, "w+");
      } else {
        if (pipe(pfd)) return 2;
        f = fdopen(pfd[1], "w");
        drain = std::thread([&]() {
          char buf[4096];
          ssize_t n;
          while ((n = read(pfd[0], buf, sizeof(buf))) > 0) got.append(buf, (size_t)n);
        });
      }
      fputs("HEAD", f);  // what was written in front of the blocks stays in front
      bool finished = false;
      {
        OrderedSink sink(f);
        if (sink.seekable() != (mode == 0)) {
          ++bad;
          printf("seekable() wrong in mode %d\n", mode);
        }
        std::atomic<size_t> next(0);
        std::vector<std::thread> th;
        for (int w = 0; w < 5; ++w) {
          th.push_back(std::thread([&]() {
            std::mt19937_64 r2(next.load() * 7919u + 17u);
            for (;;) {
              const size_t k = next.fetch_add(1);
              if (k >= nblocks) break;
              if (r2() % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(r2() % 300));
              if (k == skipAt) {
                sink.skip(k);
              } else if (sink.seekable() && (k & 1)) {  // the direct path of the decode workers: claim, then write on their own
                uint64_t at = 0;
                const bool ok = sink.claim(k, prefix[k].size() + rows[k].size(), &at);
                if (ok) {
                  sink.writePrefix(prefix[k], at);
                  if (pwrite(sink.fd(), rows[k].data(), rows[k].size(), (off_t)(at + prefix[k].size())) != (ssize_t)rows[k].size()) sink.fail();
                }
              } else {
                sink.deliver(k, prefix[k], rows[k].data(), rows[k].size());
              }
            }
          }));
        }
        for (auto& t : th) t.join();
        finished = sink.finish();
      }
      if (finished != (skipAt == nblocks)) {
        ++bad;
        printf("finish() = %d with skipAt %zu of %zu (mode %d)\n", (int)finished, skipAt, nblocks, mode);
      }
      if (mode == 0) {
        fputs("TAIL", f);  // the FILE* stands behind the rows afterwards
        fflush(f);
        const long size = lseek(fileno(f), 0, SEEK_END);
        got.resize((size_t)size);
        if (pread(fileno(f), &got[0], got.size(), 0) != (ssize_t)got.size()) return 2;
        fclose(f);
        unlink(oname);
        // (behind a failed block a regular file may hold later blocks' bytes at their offsets: only the front counts)
        if (skipAt == nblocks) {
          want += "TAIL";
          if (got != want) {
            ++bad;
            printf("file sink: %zu bytes, expected %zu (round %d)\n", got.size(), want.size(), round);
          }
        } else if (got.compare(0, want.size(), want) != 0) {
          ++bad;
          printf("file sink: front differs after a failed block (round %d)\n", round);
        }
      } else {
        fclose(f);
        drain.join();
        close(pfd[0]);
        if (got != want) {
          ++bad;
          printf("pipe sink: %zu bytes, expected %zu (round %d, skipAt %zu of %zu)\n", got.size(), want.size(), round, skipAt, nblocks);
        }
      }
    }
  }
  printf("%d mismatches\n", bad);
  return bad ? 1 : 0;
}
