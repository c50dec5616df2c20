// gpu_session.h -- RAII owner of one zdwb_ctx (include/zdw_b200.h).  No CPU fallback: creation fails without a GPU.
#ifndef ZDWB_HOST_GPU_SESSION_H
#define ZDWB_HOST_GPU_SESSION_H

#include <cstdlib>
#include <string>
#include <thread>

#include "zdw_b200.h"

namespace adobe {
namespace zdw {

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
    if (ctx_) zdwb_ctx_destroy(ctx_);
    ctx_ = NULL;
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
