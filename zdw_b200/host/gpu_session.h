// gpu_session.h -- RAII owner of one zdwb_ctx (include/zdw_b200.h).  No CPU fallback: creation fails without a GPU.
#ifndef ZDWB_HOST_GPU_SESSION_H
#define ZDWB_HOST_GPU_SESSION_H

#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "zdw_b200.h"

namespace adobe {
namespace zdw {

// --gpus=<N | all | a,b,...> (and $ZDW_GPUS): the CUDA device of every worker group.  A count means the devices
// first, first + 1, ...; a device may be listed more than once.  Returns false for a malformed spec.
inline bool parseGpuSpec(const std::string& spec, int first, std::vector<int>& out) {
  out.clear();
  if (spec.empty()) return true;
  if (spec == "all") {
    const int n = zdwb_device_count();
    for (int d = 0; d < n; ++d) out.push_back(d);
    return true;
  }
  if (spec.find_first_not_of("0123456789,") != std::string::npos) return false;
  if (spec.find(',') == std::string::npos) {
    const int n = atoi(spec.c_str());
    if (n < 1) return false;
    for (int d = 0; d < n; ++d) out.push_back((first >= 0 ? first : 0) + d);
    return true;
  }
  size_t at = 0;
  for (;;) {
    const size_t comma = spec.find(',', at);
    const std::string item = spec.substr(at, comma == std::string::npos ? std::string::npos : comma - at);
    if (item.empty()) return false;
    out.push_back(atoi(item.c_str()));
    if (comma == std::string::npos) break;
    at = comma + 1;
  }
  return true;
}

class GpuSession {
 public:
  GpuSession() : ctx_(NULL), rc_(ZDWB_OK), device_(-1), pending_(false) {}
  ~GpuSession() { close(); }
  static int resolve(int device) {  // device < 0: take $ZDW_GPU or device 0
    if (device >= 0) return device;
    const char* e = getenv("ZDW_GPU");
    return e ? atoi(e) : 0;
  }
  // Starts creating the context on a helper thread (CUDA start-up takes a few hundred milliseconds): the caller goes
  // on reading its input and meets the context again in open().
  void prefetch(int device = -1) {
    if (ctx_ || pending_) return;
    device_ = resolve(device);
    pending_ = true;
    worker_ = std::thread([this]() { rc_ = zdwb_ctx_create(device_, 0, &ctx_); });
  }
  bool open(int device = -1) {
    device = resolve(device);
    if (pending_) {
      worker_.join();
      pending_ = false;
      if (device_ != device) close();  // the device was changed after the prefetch
    }
    if (ctx_) return true;
    device_ = device;
    rc_ = zdwb_ctx_create(device, 0, &ctx_);
    return rc_ == ZDWB_OK;
  }
  void close() {
    if (pending_) {
      worker_.join();
      pending_ = false;
    }
    // (tearing a context down - unpinning its buffers, freeing its arena - takes tenths of a second; a command line tool
    // that is about to exit leaves that to the process exit)
    if (ctx_ && !processExiting()) zdwb_ctx_destroy(ctx_);
    ctx_ = NULL;
  }
  // set by the command line tools before their last file: contexts are no longer torn down one by one
  static bool& processExiting() {
    static bool flag = false;
    return flag;
  }
  zdwb_ctx* get() const { return ctx_; }
  int status() const { return rc_; }
  std::string lastError() const { return ctx_ ? std::string(zdwb_last_error(ctx_)) : std::string("no CUDA device"); }

 private:
  GpuSession(const GpuSession&);
  GpuSession& operator=(const GpuSession&);
  zdwb_ctx* ctx_;
  int rc_;
  int device_;
  bool pending_;
  std::thread worker_;
};

}  // namespace zdw
}  // namespace adobe
#endif
