// rpe/session.hpp — RAII owner of an rpe_ctx for the header-only host API.
//
// The reference is stateless between calls (everything lives in the adapter). Here the device buffers live
// in an rpe_ctx (include/rpe_c_api.h); estimators use a thread-local Session so that the reference's call
// pattern — construct adapter, call <method>_ransac(adapter, ...), call <method>_ls(adapter) — needs no
// extra arguments. One Session per host thread and GPU; select the GPU with rpe::Session::setDevice().
#ifndef RPE_SESSION_HPP_
#define RPE_SESSION_HPP_

#include <stdexcept>
#include <string>
#include <vector>

#include "../rpe_c_api.h"

namespace rpe {

struct Failure : std::runtime_error {
  int status;
  Failure(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

class Session {
 public:
  explicit Session(int device = 0) : ctx_(nullptr), device_(device), token_(0) {
    const int rc = rpe_create(device, &ctx_);
    if (rc != RPE_OK) throw Failure(rc, std::string("rpe_create: ") + rpe_status_string(rc));
    // the header API is blocking and the reference's loops usually stop after a few dozen to a few hundred iterations:
    // start with a short pass (256, 512, 1024, ... iterations), the result does not depend on it
    rpe_set_first_pass_iters(ctx_, 256);
  }
  ~Session() {
    if (ctx_) rpe_destroy(ctx_);
  }
  Session(const Session&) = delete;
  Session& operator=(const Session&) = delete;

  static int& defaultDevice() {
    static thread_local int dev = 0;
    return dev;
  }
  static void setDevice(int device) { defaultDevice() = device; }
  // thread-local session on the default device, created on first use
  static Session& local() {
    static thread_local Session* s = nullptr;
    static thread_local int s_dev = -1;
    if (!s || s_dev != defaultDevice()) {
      delete s;
      s = new Session(defaultDevice());
      s_dev = defaultDevice();
    }
    return *s;
  }

  rpe_ctx* ctx() { return ctx_; }
  // Reproduce the reference's stale sample columns on frames with invalid depth (rpe_set_stale_sample_columns;
  // /root/reference/pose/AbsoluteOrientationNormal.hpp:48-75, 299-315). Off by default; RPE_STALE_SAMPLE_COLUMNS=1 in the
  // environment switches it on for every session.
  void setReferenceStaleSampleColumns(bool on) { check(rpe_set_stale_sample_columns(ctx_, on ? 1 : 0), "rpe_set_stale_sample_columns"); }
  // Overlap the upload of page-locked 3-D / 3-D frames with generation and scoring (rpe_set_upload_overlap); 0 = off.
  void setUploadOverlap(int chunks) { check(rpe_set_upload_overlap(ctx_, chunks), "rpe_set_upload_overlap"); }
  void check(int rc, const char* what) {
    if (rc != RPE_OK) throw Failure(rc, std::string(what) + ": " + rpe_status_string(rc) + " (" + rpe_last_error(ctx_) + ")");
  }

  // Tp = double: the arrays go up in binary64 and the next RANSAC decides in binary64 (rpe_upload_f64).
  void upload(const double* bv, const double* xc, const double* nc, const double* xw, const double* nw, int n) {
    check(rpe_upload_f64(ctx_, bv, xc, nc, xw, nw, n), "rpe_upload_f64");
    ++token_;
  }
  // Upload column-major 3 x n arrays of any other scalar type (converted to the float the kernels compute in).
  template <class Tp>
  void upload(const Tp* bv, const Tp* xc, const Tp* nc, const Tp* xw, const Tp* nw, int n) {
    const Tp* src[5] = {bv, xc, nc, xw, nw};
    const float* f[5];
    for (int k = 0; k < 5; ++k) f[k] = stage(src[k], n, k);
    check(rpe_upload(ctx_, f[0], f[1], f[2], f[3], f[4], n), "rpe_upload");
    ++token_;
  }
  // identifies "the device holds this adapter's inlier mask and pose" (see detail::run in Estimators.hpp)
  unsigned long long token() const { return token_; }
  unsigned long long bump() { return ++token_; }

 private:
  const float* stage(const float* p, int, int) { return p; }
  template <class Tp>
  const float* stage(const Tp* p, int n, int k) {
    if (!p) return nullptr;
    staging_[k].resize((size_t)3 * n);
    for (size_t i = 0; i < (size_t)3 * n; ++i) staging_[k][i] = (float)p[i];
    return staging_[k].data();
  }
  rpe_ctx* ctx_;
  int device_;
  unsigned long long token_;
  std::vector<float> staging_[5];
};

}  // namespace rpe

#endif  // RPE_SESSION_HPP_
