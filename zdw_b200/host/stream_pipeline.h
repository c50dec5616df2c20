// stream_pipeline.h -- the two host-side helpers that keep the GPU fed by convertDWfile (overlap of the
// input read, the copies + kernels of a block and the compressor pipe).  Plain C++ (no CUDA): the buffers come from
// the allocator the caller passes in (pinned memory of the C ABI, or malloc for a one-window input), so the window
// logic is exercised on a CPU by tests/stream_pipeline_test.cpp.
//
//   ReadAheadInput  a window of at most `cap` bytes over a FILE*, always starting at a row boundary.  It behaves like
//                   the sequential window it replaces - fill() tops the window up to exactly `cap` bytes or the end of
//                   the stream, so block cuts do not move - but prefetch() lets a helper thread read the bytes of the
//                   NEXT window into a second buffer while the caller encodes the current one.
//   AsyncWriter     a helper thread that drains finished blocks into the compressor pipe (gzip takes longer over a
//                   block than the GPU does), at most two blocks behind.
//
// The reference reads and writes from its one thread (getnextrow.cpp:26-84, ConvertToZDW.cpp:486-606); there is
// nothing to overlap there because parsing is the bottleneck.
#ifndef ZDWB_HOST_STREAM_PIPELINE_H
#define ZDWB_HOST_STREAM_PIPELINE_H

#include <stdio.h>
#include <string.h>

#include <errno.h>
#include <fcntl.h>
#include <poll.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace adobe {
namespace zdw {

class ReadAheadInput {
 public:
  typedef void* (*AllocFn)(size_t);
  typedef void (*FreeFn)(void*);

  ReadAheadInput() : alloc_(NULL), free_(NULL), in_(NULL), tee_(NULL), cap_(0), gap_(0), streamEof_(false), pending_(false),
                     stop_(false), aheadLen_(0), aheadEof_(false), aheadFailed_(false) {}
  ~ReadAheadInput() { close(); }

  // `windowBytes` = the most the caller takes per block; buffers are allocated with `alloc` on first use
  bool open(FILE* in, FILE* tee, size_t windowBytes, AllocFn alloc, FreeFn release) {
    close();
    in_ = in;
    tee_ = tee;
    alloc_ = alloc;
    free_ = release;
    streamEof_ = false;
    setCap(windowBytes);
#ifdef F_SETPIPE_SZ
    // a pipe (-i): 1 MiB instead of 64 KiB between the producer and us - fewer hand-overs per gigabyte (no effect, and
    // no error we care about, on anything that is not a pipe)
    (void)fcntl(fileno(in), F_SETPIPE_SZ, 1 << 20);
#endif
    return grow(cur_, cap_ + gap_);
  }

  void close() {
    stop_ = true;  // a read-ahead that waits for a pipe gives up (its bytes are not wanted any more)
    joinReader();
    stop_ = false;
    if (free_) {
      if (cur_.buf) free_(cur_.buf);
      if (next_.buf) free_(next_.buf);
    }
    cur_ = Buf();
    next_ = Buf();
    pending_ = false;
  }

  // Tops the window up to `cap` bytes or the end of the stream (reads on the calling thread).
  void fill() {
    settle(0);
    if (streamEof_ || cur_.len >= cap_) return;
    if (cur_.start + cap_ > cur_.capacity) {  // the window slid to the end of its buffer: back to the front
      memmove(cur_.buf, cur_.buf + cur_.start, cur_.len);
      cur_.start = 0;
    }
    cur_.len += readInto(cur_.buf + cur_.start + cur_.len, cap_ - cur_.len, streamEof_, false);
  }

  char* data() const { return cur_.buf + cur_.start; }
  size_t len() const { return std::min(cur_.len, cap_); }
  // true when the stream ended inside this window (the sequential window's flag: a read came back empty while the
  // window still had room); a stream that ends exactly at the window's last byte reports it one window later
  bool eof() const { return streamEof_ && cur_.len < cap_; }
  size_t windowBytes() const { return cap_; }

  // Starts reading the bytes that follow the current window into the second buffer.  Call it right before the encode
  // call, and only when that call will consume (nearly) the whole window: the tail it leaves over has to fit in front
  // of the read-ahead bytes (other tails work too, through a slower merge).
  void prefetch() {
    if (pending_ || streamEof_ || cur_.len != cap_ || cap_ <= gap_) return;
    aheadLen_ = 0;
    aheadEof_ = false;
    aheadFailed_ = false;
    pending_ = true;
    reader_ = std::thread([this]() {
      if (next_.capacity < cap_ + gap_ && !grow(next_, cap_ + gap_)) {
        aheadFailed_ = true;  // no second buffer: the caller carries on sequentially
        return;
      }
      aheadLen_ = readInto(next_.buf + gap_, cap_ - gap_, aheadEof_, true);
    });
  }

  // The first `n` bytes of the window are done with.
  void consume(size_t n) { settle(std::min(n, cur_.len)); }

  // A bigger window (nothing of the current one is lost).
  bool widen(size_t windowBytes) {
    settle(0);
    if (windowBytes <= cap_) return true;
    const size_t old = cap_;
    setCap(windowBytes);
    if (cur_.start + cap_ + gap_ > cur_.capacity) {
      Buf nb;
      if (!grow(nb, std::max(cap_ + gap_, cur_.len))) {
        setCap(old);
        return false;
      }
      memcpy(nb.buf, cur_.buf + cur_.start, cur_.len);
      nb.len = cur_.len;
      free_(cur_.buf);
      cur_ = nb;
    }
    if (next_.buf) {  // re-made at the new size by the next prefetch
      free_(next_.buf);
      next_ = Buf();
    }
    return true;
  }

 private:
  struct Buf {
    Buf() : buf(NULL), capacity(0), start(0), len(0) {}
    char* buf;
    size_t capacity, start, len;
  };

  void setCap(size_t windowBytes) {
    cap_ = std::max<size_t>(windowBytes, 1);
    gap_ = std::min<size_t>(cap_ / 8, (size_t)64 << 20);
  }

  bool grow(Buf& b, size_t capacity) {
    char* p = static_cast<char*>(alloc_(capacity + 64));
    if (!p) return false;
    if (b.buf) free_(b.buf);
    b.buf = p;
    b.capacity = capacity;
    b.start = b.len = 0;
    return true;
  }

  // reads until `want` bytes arrived or a read comes back empty (then eof = true); everything read goes to the tee.
  // Reads the descriptor directly (nobody else reads this stream, so stdio's buffer stays empty): the helper thread
  // can then wait for a pipe with poll() and notice close() within a tenth of a second instead of sitting in a read
  // that may never return.
  size_t readInto(char* dst, size_t want, bool& eof, bool helper) {
    const int fd = fileno(in_);
    size_t got = 0;
    while (got < want) {
      if (helper) {
        if (stop_) break;
        struct pollfd pf;
        pf.fd = fd;
        pf.events = POLLIN;
        pf.revents = 0;
        const int pr = poll(&pf, 1, 100);
        if (pr == 0 || (pr < 0 && errno == EINTR)) continue;
      }
      const ssize_t n = read(fd, dst + got, std::min<size_t>(want - got, (size_t)1 << 30));
      if (n < 0 && errno == EINTR) continue;
      if (n <= 0) {
        eof = true;
        break;
      }
      if (tee_) fwrite(dst + got, 1, (size_t)n, tee_);
      got += (size_t)n;
    }
    return got;
  }

  void joinReader() {
    if (reader_.joinable()) reader_.join();
  }

  // Drops the first n bytes of the current buffer and, when a read-ahead is under way, waits for it and makes
  // [what is left of the current buffer][the bytes read ahead] the new current buffer.
  void settle(size_t n) {
    cur_.start += n;
    cur_.len -= n;
    if (!pending_) return;
    joinReader();
    pending_ = false;
    if (aheadFailed_) return;
    const size_t t = cur_.len, r = aheadLen_;
    const char* tail = cur_.buf + cur_.start;
    if (t <= gap_) {
      memcpy(next_.buf + gap_ - t, tail, t);
      next_.start = gap_ - t;
    } else if (t + r <= next_.capacity) {
      memmove(next_.buf + t, next_.buf + gap_, r);
      memcpy(next_.buf, tail, t);
      next_.start = 0;
    } else {
      Buf nb;
      if (!grow(nb, t + r)) throw std::bad_alloc();  // convertFile reports OUT_OF_MEMORY, like any other allocation
      memcpy(nb.buf, tail, t);
      memcpy(nb.buf + t, next_.buf + gap_, r);
      free_(next_.buf);
      next_ = nb;
    }
    next_.len = t + r;
    streamEof_ = aheadEof_;
    std::swap(cur_, next_);
    next_.start = next_.len = 0;
  }

  AllocFn alloc_;
  FreeFn free_;
  FILE* in_;
  FILE* tee_;
  size_t cap_, gap_;
  Buf cur_, next_;
  bool streamEof_;
  bool pending_;
  std::atomic<bool> stop_;
  std::thread reader_;
  size_t aheadLen_;
  bool aheadEof_, aheadFailed_;
};

class AsyncWriter {
 public:
  AsyncWriter() : out_(NULL), done_(false), started_(false), failed_(false) {}
  ~AsyncWriter() { finish(); }

  void start(FILE* out) {
    out_ = out;
    done_ = false;
    started_ = true;
    worker_ = std::thread([this]() { run(); });
  }

  // hands a block over (the vector is emptied); waits while two blocks are still queued
  void push(std::vector<unsigned char>& block) {
    std::unique_lock<std::mutex> lk(m_);
    roomCv_.wait(lk, [this]() { return queue_.size() < 2; });
    queue_.emplace_back();
    queue_.back().swap(block);
    workCv_.notify_one();
  }

  // false once a write to the pipe came up short (the compressor died, the disk is full)
  bool ok() {
    std::lock_guard<std::mutex> lk(m_);
    return !failed_;
  }

  // everything pushed so far has been written when this returns; the FILE* is the caller's to close
  void finish() {
    if (!started_) return;
    {
      std::lock_guard<std::mutex> lk(m_);
      done_ = true;
    }
    workCv_.notify_one();
    worker_.join();
    started_ = false;
  }

 private:
  void run() {
    for (;;) {
      std::vector<unsigned char> blk;
      {
        std::unique_lock<std::mutex> lk(m_);
        workCv_.wait(lk, [this]() { return done_ || !queue_.empty(); });
        if (queue_.empty()) return;
        blk.swap(queue_.front());
      }
      const bool wrote = fwrite(blk.data(), 1, blk.size(), out_) == blk.size();
      {
        std::lock_guard<std::mutex> lk(m_);
        if (!wrote) failed_ = true;
        queue_.pop_front();  // only now: "queued" counts the block being written
      }
      roomCv_.notify_one();
    }
  }

  FILE* out_;
  std::mutex m_;
  std::condition_variable workCv_, roomCv_;
  std::deque<std::vector<unsigned char> > queue_;
  bool done_, started_, failed_;
  std::thread worker_;
};

}  // namespace zdw
}  // namespace adobe
#endif
