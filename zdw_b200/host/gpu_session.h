// gpu_session.h -- RAII owner of one zdwb_ctx (include/zdw_b200.h).  No CPU fallback: creation fails without a GPU.
#ifndef ZDWB_HOST_GPU_SESSION_H
#define ZDWB_HOST_GPU_SESSION_H

#include <cstdlib>
#include <string>

#include "zdw_b200.h"

namespace adobe {
namespace zdw {

class GpuSession {
 public:
  GpuSession() : ctx_(NULL), rc_(ZDWB_OK) {}
  ~GpuSession() { close(); }
  // device < 0: take $ZDW_GPU or device 0
  bool open(int device = -1) {
    if (ctx_) return true;
    if (device < 0) {
      const char* e = getenv("ZDW_GPU");
      device = e ? atoi(e) : 0;
    }
    rc_ = zdwb_ctx_create(device, 0, &ctx_);
    return rc_ == ZDWB_OK;
  }
  void close() {
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
};

}  // namespace zdw
}  // namespace adobe
#endif
