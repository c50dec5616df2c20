// block_pipeline.h -- host side of the multi-GPU block fan-out: ZDW blocks are self-contained, so whole
// blocks go to different GPUs / contexts and the host only has to (1) cut the input into the windows the sequential
// loop would have used, before any of them is encoded, and (2) put the encoded blocks out in file order with the two
// header fields that depend on the neighbours patched (isLast, cumulative longestLine: reference
// ConvertToZDW.cpp:841-842,965).  Plain C++ (no CUDA) so that the cutting rules are unit-tested on a CPU
// (tests/block_pipeline_test.cpp).
//
// The cut the GPU reports for a window that is followed by more input (zdwb_block_out.tsv_consumed) is the byte behind
// the last row break of the window: the last '\n' preceded by an even number of backslashes (getnextrow.cpp:44-53).
// findLastRowBreak() restates exactly that on the host; the encode workers check their block's tsv_consumed against the
// planned window and fail loudly on any difference.
#ifndef ZDWB_HOST_BLOCK_PIPELINE_H
#define ZDWB_HOST_BLOCK_PIPELINE_H

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

namespace adobe {
namespace zdw {

// Position behind the last unescaped newline of p[0..n), 0 = there is none.  p[0] is the first byte of a row, so a run
// of backslashes never reaches in from the left.  `needMore` (may be NULL) is set when the answer depends on bytes in
// front of p (only when the caller passed a suffix of the window, see planFileWindows).
inline size_t findLastRowBreak(const char* p, size_t n, bool suffix = false, bool* needMore = NULL) {
  if (needMore) *needMore = false;
  size_t end = n;
  while (end > 0) {
    const void* hit = memrchr(p, '\n', end);
    if (!hit) return 0;
    const size_t at = (size_t)(static_cast<const char*>(hit) - p);
    size_t k = at, slashes = 0;
    while (k > 0 && p[k - 1] == '\\') {
      --k;
      ++slashes;
    }
    if (k == 0 && suffix && slashes > 0) {  // the run of backslashes may go on in front of the suffix
      if (needMore) *needMore = true;
      return 0;
    }
    if ((slashes & 1u) == 0) return at + 1;
    end = at;  // escaped: part of a field, look further left
  }
  return 0;
}

struct FileWindow {
  uint64_t offset;  // first byte (a row boundary)
  size_t len;       // bytes handed to zdwb_encode_block
  size_t consumed;  // bytes the block is expected to cover (the next window starts behind them)
  bool more;        // more input follows the window (zdwb_encode_opts.more_input_follows)
};

// The windows the sequential loop of ConvertToZDW::processFile walks through, computed up front from the file alone:
// window k starts where block k-1 ended and holds `cap` bytes or the rest of the file; a window without a row break is
// doubled (up to maxCap) and stays doubled, as in the sequential loop.  Only the tail of every window is read.
// Returns false when a single row exceeds maxCap or a read fails.
inline bool planFileWindows(int fd, uint64_t fileSize, size_t cap, size_t maxCap, std::vector<FileWindow>& out, size_t* finalCap) {
  out.clear();
  uint64_t s = 0;
  std::vector<char> chunk;
  for (;;) {
    const uint64_t left = fileSize - s;
    const size_t len = (size_t)std::min<uint64_t>(cap, left);
    const bool eof = left < cap;  // the stream ends INSIDE the window (a window that ends exactly at the last byte learns it one call later)
    if (len == 0) break;          // (eof is implied)
    FileWindow w;
    w.offset = s;
    w.len = len;
    w.more = !eof;
    if (eof) {
      w.consumed = len;
      out.push_back(w);
      break;
    }
    // last row break of [s, s + len): read suffixes of growing size
    size_t cut = 0;
    size_t want = std::min<size_t>(len, (size_t)1 << 20);
    for (;;) {
      chunk.resize(want);
      size_t got = 0;
      while (got < want) {
        const ssize_t r = pread(fd, chunk.data() + got, want - got, (off_t)(s + len - want + got));
        if (r <= 0) return false;
        got += (size_t)r;
      }
      bool needMore = false;
      const size_t at = findLastRowBreak(chunk.data(), want, want < len, &needMore);
      if (at) {
        cut = len - want + at;
        break;
      }
      if (want == len) break;  // (needMore cannot be set for the whole window)
      want = std::min<size_t>(len, want * 4);
    }
    if (cut == 0) {  // not one complete row in the window: widen it, like the sequential loop
      if (cap >= maxCap) return false;
      cap = std::min(cap * 2, maxCap);
      continue;
    }
    w.consumed = cut;
    out.push_back(w);
    s += cut;
  }
  if (finalCap) *finalCap = cap;
  return true;
}

// Results of jobs that finish out of order, handed to one consumer in order.  T needs a default constructor.
template <typename T>
class OrderedResults {
 public:
  explicit OrderedResults(size_t n) : slots_(n), ready_(n, 0), retired_(0) {}
  void put(size_t k, T&& v) {
    std::lock_guard<std::mutex> lk(m_);
    slots_[k] = std::move(v);
    ready_[k] = 1;
    cv_.notify_all();
  }
  // blocks until result k is there; the slot is emptied
  T take(size_t k) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&]() { return ready_[k] != 0; });
    T v = std::move(slots_[k]);
    slots_[k] = T();
    retired_ = k + 1;
    cv_.notify_all();
    return v;
  }
  // producers call this before starting job k: at most `ahead` finished or running jobs in front of the consumer
  void waitTurn(size_t k, size_t ahead) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&]() { return k < retired_ + ahead; });
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  std::vector<T> slots_;
  std::vector<char> ready_;
  size_t retired_;
};

// Rows of decoded blocks leave in file order although their workers finish in any order (UnconvertFromZDW.cpp,
// decodeBlocksFanOut): through the FILE* for pipes / stdout; when the sink is a regular file a block only CLAIMS its
// place in turn (its offset is known once the blocks in front have been decoded, not written) and is written there side
// by side with the others (pwrite, or zdwb_device_to_fd straight from the device).
class OrderedSink {
 public:
  explicit OrderedSink(FILE* f) : fp_(f), fd_(fileno(f)), seekable_(false), base_(0), nextSeq_(0), total_(0), failed_(false) {
    struct stat st;
    fflush(fp_);
    if (fstat(fd_, &st) == 0 && S_ISREG(st.st_mode)) {
      const off_t at = lseek(fd_, 0, SEEK_CUR);
      if (at >= 0) {
        seekable_ = true;
        base_ = (uint64_t)at;
      }
    }
  }
  // rows of block `seq` (with the optional line that precedes them); returns false once anything failed to be written
  bool deliver(size_t seq, const std::string& prefix, const void* rows, size_t len) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&]() { return nextSeq_ == seq; });
    bool ok = !failed_;
    if (!seekable_) {
      if (ok && !prefix.empty()) ok = fwrite(prefix.data(), 1, prefix.size(), fp_) == prefix.size();
      if (ok && len) ok = fwrite(rows, 1, len, fp_) == len;
      if (!ok) failed_ = true;
      ++nextSeq_;
      cv_.notify_all();
      return ok;
    }
    const uint64_t at = base_ + total_;
    total_ += prefix.size() + len;
    ++nextSeq_;  // the next block may take its offset: the writes themselves run side by side
    cv_.notify_all();
    lk.unlock();
    ok = ok && writeAt(prefix.data(), prefix.size(), at) && writeAt(rows, len, at + prefix.size());
    if (!ok) {
      std::lock_guard<std::mutex> g(m_);
      failed_ = true;
    }
    return ok;
  }
  // a regular file: the place of block `seq` (prefix + rows, `bytes` in all) - the caller writes there itself (the rows
  // go from the device to the file through zdwb_device_to_fd).  false once anything failed.
  bool seekable() const { return seekable_; }
  int fd() const { return fd_; }
  bool claim(size_t seq, size_t bytes, uint64_t* at) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&]() { return nextSeq_ == seq; });
    *at = base_ + total_;
    total_ += bytes;
    ++nextSeq_;
    cv_.notify_all();
    return !failed_;
  }
  void fail() {
    std::lock_guard<std::mutex> g(m_);
    failed_ = true;
  }
  bool writePrefix(const std::string& prefix, uint64_t at) { return writeAt(prefix.data(), prefix.size(), at); }
  // a block that produced nothing (its worker failed): later blocks must not wait for it for ever
  // (nothing behind it is written either: the output ends where the reference's would, in front of the bad block)
  void skip(size_t seq) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&]() { return nextSeq_ == seq; });
    failed_ = true;
    ++nextSeq_;
    cv_.notify_all();
  }
  // after every worker is done: leaves the FILE* positioned behind the rows
  bool finish() {
    if (seekable_ && lseek(fd_, (off_t)(base_ + total_), SEEK_SET) < 0) return false;
    return !failed_;
  }

 private:
  bool writeAt(const void* p, size_t n, uint64_t at) {
    const char* c = static_cast<const char*>(p);
    while (n) {
      const ssize_t w = pwrite(fd_, c, n, (off_t)at);
      if (w <= 0) return false;
      c += w;
      n -= (size_t)w;
      at += (uint64_t)w;
    }
    return true;
  }
  FILE* fp_;
  int fd_;
  bool seekable_;
  uint64_t base_;
  std::mutex m_;
  std::condition_variable cv_;
  size_t nextSeq_;
  uint64_t total_;
  bool failed_;
};


}  // namespace zdw
}  // namespace adobe
#endif
