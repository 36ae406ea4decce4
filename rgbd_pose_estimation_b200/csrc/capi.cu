// capi.cu — implementation of include/rpe_c_api.h: contexts, device memory, and the per-frame
// launch sequence  reset -> pack -> generate -> score (fast + exact fix-up) -> replay -> mask(+Kabsch)
// [-> LM iterations], all asynchronous on one stream with a single synchronisation at the end.
//
// There is no CPU fallback in this file: every compute entry point needs a live CUDA context and
// returns RPE_ERR_NO_DEVICE / RPE_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <vector>

#include "kernels.cuh"

namespace rpe {
int score_variant();
void set_use_packed(bool v);
bool use_packed();
void set_score_variant(int v);
void set_nosync(int v);
constexpr int kDefaultScoreVariant = 14;  // must match g_variant's initial value in score.cu
}

using namespace rpe;

namespace {

enum { A_BV = 0, A_XC = 1, A_NC = 2, A_XW = 3, A_NW = 4 };
constexpr int kNumStaging = 256;  // results that may be in flight; when full only the oldest one is waited for
enum { ST_UPLOAD = 0, ST_GEN = 1, ST_SCORE = 2, ST_REPLAY = 3, ST_MASK = 4, ST_GN = 5, ST_TOTAL = 6, ST_FAST = 7, ST_COUNT = 8 };

}  // namespace

struct rpe_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  std::string err;
  long long launches = 0;

  // correspondences
  int n = 0;
  size_t cap_n = 0;
  float* d_raw[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const float* view[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // what kernels read (own copy or caller's)
  float4* d_pk = nullptr;
  size_t pk_cap_bytes = 0;
  int pk_kind = -1;  // modalities currently packed, -1 = stale
  int npairs_pad = 0;

  // hypotheses
  int cap_slots = 0;
  int cap_H = 0;
  int n_slots = 0;
  int cur_method = -1;
  HypGen* d_gen = nullptr;
  HypFast* d_fast = nullptr;
  int32_t* d_votes = nullptr;
  int32_t* d_samples = nullptr;

  FrameStats* d_stats = nullptr;
  bool stats_clean = true;       // per-pass counters are zero (the replay kernel leaves them so)
  ReplayState* d_rs = nullptr;
  ReplayState* h_rs = nullptr;   // pinned
  ReplayOut* d_pose = nullptr;    // current pose = the adapter's (R_cw, t_w, max_votes) state
  ReplayOut* d_kabsch = nullptr;  // Kabsch refit computed by the mask kernel's last CTA
  ReplayOut* h_pose = nullptr;    // pinned, kNumStaging slots; [0] doubles as scratch for set_pose
  bool kabsch_valid = false;
  bool suff_valid = false;  // rb.suff matches the inlier columns in d_mask
  // overlap of the host-to-device upload with generation and scoring (rpe_set_upload_overlap; 3-D / 3-D family,
  // page-locked host arrays): the frame is copied in chunks on copy_stream, the generator reads its sample points
  // straight from the page-locked host arrays, the scorer is launched once per chunk as the chunk lands
  static constexpr int kMaxChunks = 8;
  int overlap_chunks = 0;          // 0 = off
  cudaStream_t copy_stream = nullptr, early_stream = nullptr;
  cudaEvent_t ev_chunk[kMaxChunks] = {};
  cudaEvent_t ev_prev = nullptr, ev_early = nullptr;
  // the inlier mask (1.2 MB for a dense frame) goes back on its own stream so that the refits enqueued behind the RANSAC
  // do not wait for it (own streams only; with a borrowed stream the caller synchronises that stream and nothing else)
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t ev_mask_ready = nullptr, ev_mask_copied = nullptr;
  bool mask_copy_pending = false;
  bool deferred = false;           // rpe_upload has only taken note of the host arrays: nothing copied yet
  const float* host_src[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // the caller's page-locked arrays (deferred)
  int n_chunks = 0, chunk_corr = 0;
  const float* host_view[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // device-visible addresses of the host arrays
  bool stale_cols = false;       // rpe_set_stale_sample_columns
  int32_t* d_stale_eff = nullptr;  // [cap_stale x 2] effective camera-side sample indices of the current pass
  int32_t* d_stale_carry = nullptr;
  int cap_stale = 0;
  int first_pass = 1024;  // iterations of the first device pass (kFirstPassIters; rpe_set_first_pass_iters)
  // peer-memory vote exchange (hypothesis-sharded single frame)
  unsigned char* d_peer_block = nullptr;  // own block (exported)
  PeerTable peers = {};
  bool peer_opened[kMaxPeers] = {};
  int peer_rank = -1, peer_world = 0;
  unsigned int peer_epoch = 0;
  unsigned int* h_peer_err = nullptr;   // pinned word the exchange kernel sets when a peer timed out (read after a sync)
  unsigned long long peer_timeout_ns = 2000000000ull;
  // binary64 path (rpe_upload_f64)
  bool f64 = false;
  double* d_raw64[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const double* view64[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t cap_n64 = 0;
  HypGen64* d_gen64 = nullptr;
  int cap_slots64 = 0;
  ReplayState64* d_rs64 = nullptr;
  ReplayState64* h_rs64 = nullptr;  // pinned
  Pose64* d_pose64 = nullptr;
  Pose64* h_pose64 = nullptr;       // pinned, kNumStaging slots
  int32_t* h_samples = nullptr;  // pinned staging for rpe_ransac_stream
  int h_samples_cap = 0;
  Worklist wl = {nullptr, nullptr, 0};
  unsigned int wl_allocated = 0;  // entries allocated (wl.capacity may be lowered by the test hook)
  unsigned int wl_want = 0;       // grow to this many entries before the next scoring call (set after an overflow)
  bool wl_fixed = false;          // test hook in force: no automatic growth
  int16_t* d_mask = nullptr;
  uint32_t* d_maskbits = nullptr;  // the mask as one bit per flag (rpe_set_mask_transfer(1)), 3 x ceil(n / 32) words
  size_t maskbits_cap = 0;
  int mask_transfer = 0;           // 0: the int16 matrix goes to the host as it is, 1: bits + expansion on the host, 2: constant column 0 stays behind
  std::vector<uint32_t> bits_scratch;
  size_t mask_cap = 0;
  int mask_cols = 0;
  RefitBuffers rb;
  GnState* d_gn = nullptr;
  NlskState* d_nlsk = nullptr;
  float* d_weights3 = nullptr;
  size_t weights_cap = 0;
  double* d_gn_cost = nullptr;
  int32_t* d_gn_evals = nullptr;
  double* h_gn_cost = nullptr;  // pinned (cost + evals packed)
  int32_t* h_gn_evals = nullptr;

  // results of enqueued-but-not-yet-synchronised calls, in stream order
  struct Pending {
    rpe_result* out;
    int slot;
    bool is_refit, gn;
    int16_t* mask_expand = nullptr;  // bit form waiting in the tail of this host buffer (see expand_mask)
    int mask_n = 0, mask_cols = 0;
    int16_t* mask_fill0 = nullptr;   // rpe_set_mask_transfer(2): column 0 was not copied, the collector writes its constant
  };
  std::deque<Pending> pending;
  int next_slot = 0;                      // staging slots are handed out round-robin
  cudaEvent_t ev_lane[2] = {};            // fences between this context's stream and the device's scorer lane
  cudaEvent_t ev_lane_done[4] = {};       // chunked frames: one fence per lane used (kMaxLanes)
  cudaEvent_t ev_slot[kNumStaging] = {};  // recorded behind each result's device-to-host copies
  Thresh last_th = {0.f, 0.f, 0.f};

  // stage timing
  bool timing = false;       // record the per-stage events
  bool timing_fast = false;  // record the two events around the tiled scoring kernel
  cudaEvent_t ev[ST_COUNT + 1] = {};
  cudaEvent_t ev_fast[2] = {};
  bool ev_fast_recorded = false;
  // every scorer launch of an asynchronous stream of frames gets its own event pair out of a small ring; elapsed times
  // are folded into (fast_sum_ms, fast_count) when a pair is reused or at the next synchronisation (rpe_scorer_time_stats)
  static constexpr int kFastRing = 32;
  cudaEvent_t ev_ring[2 * kFastRing] = {};
  bool ev_ring_live[kFastRing] = {};
  int ev_ring_next = 0;
  double fast_sum_ms = 0.0;
  long long fast_count = 0;
  bool upload_stamped = false;
  bool ev_ok = false;
  bool ev_recorded[ST_COUNT + 1] = {};
  float stage_ms[ST_COUNT] = {};
};

namespace {

int fail(rpe_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) {
      c->err += ": ";
      c->err += cudaGetErrorString(e);
    }
  }
  return code;
}
#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return fail(ctx, RPE_ERR_CUDA, #call, e__); \
  } while (0)

// Every entry point that replaces the frame with binary32 arrays calls this: a context that was in binary64 mode
// (rpe_upload_f64) must not route the next rpe_ransac to the binary64 arrays of an earlier, possibly smaller frame.
void leave_f64_mode(rpe_ctx* c) {
  c->f64 = false;
  for (int k = 0; k < 5; ++k) c->view64[k] = nullptr;
  c->deferred = false;  // (rpe_upload sets it again after this call when it defers the copies)
}

int kind_for_method(int method) {
  return (method_uses_2d(method) ? 1 : 0) | (method_uses_3d(method) ? 2 : 0) | (method_uses_nl(method) ? 4 : 0);
}
int f4_per_pair(int kind) {
  int arrays = 1 + ((kind & 1) ? 1 : 0) + ((kind & 2) ? 1 : 0) + ((kind & 4) ? 2 : 0);
  return (arrays * 6 + 3) / 4;
}
bool method_ok(int m) { return m >= RPE_SHINJI && m <= RPE_KNEIP_QUAT; }

// host-side wait for everything the context has enqueued, including a mask copy on the side stream
cudaError_t sync_stream(rpe_ctx* ctx) {
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && ctx->mask_copy_pending) {
    e = cudaEventSynchronize(ctx->ev_mask_copied);
    ctx->mask_copy_pending = false;
  }
  return e;
}
// stream-side: whoever overwrites d_mask next waits for the copy of the previous mask
void order_after_mask_copy(rpe_ctx* ctx) {
  if (ctx->mask_copy_pending) cudaStreamWaitEvent(ctx->stream, ctx->ev_mask_copied, 0);
}

// A deferred upload (rpe_set_upload_overlap) that the next call cannot overlap with anything is simply carried out now,
// on the context's stream, the plain way.
int flush_deferred(rpe_ctx* ctx) {
  if (!ctx->deferred) return RPE_OK;
  ctx->deferred = false;
  for (int k = 0; k < 5; ++k)
    if (ctx->host_src[k] && ctx->view[k])
      CK(cudaMemcpyAsync(ctx->d_raw[k], ctx->host_src[k], (size_t)ctx->n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  return RPE_OK;
}
// first thing every entry point that works on the context's frame does
#define ENTER(ctx)                                     \
  do {                                                 \
    CK(cudaSetDevice((ctx)->device));                  \
    if (int rcf__ = flush_deferred(ctx)) return rcf__; \
  } while (0)

int check_arrays(rpe_ctx* ctx, int method) {
  if (ctx->n <= 0) return fail(ctx, RPE_ERR_STATE, "no correspondences uploaded");
  if (!ctx->view[A_XW]) return fail(ctx, RPE_ERR_STATE, "world points missing");
  if (method_uses_2d(method) && !ctx->view[A_BV]) return fail(ctx, RPE_ERR_STATE, "bearing vectors missing for this method");
  if ((method_uses_3d(method) || method_uses_nl(method)) && !ctx->view[A_XC])
    return fail(ctx, RPE_ERR_STATE, "camera points missing for this method");
  if (method_uses_nl(method) && (!ctx->view[A_NC] || !ctx->view[A_NW]))
    return fail(ctx, RPE_ERR_STATE, "normals missing for this method");
  return RPE_OK;
}

// The tiled scorers stream the caller's arrays themselves when they are 16-byte aligned (bulk TMA); see
// score3d_raw_kernel / score_multi_fast_kernel<.., RAW>.
bool g_raw_tiles = true;  // test hook: false = always pack
unsigned raw_aligned_bits(const rpe_ctx* c) {
  if (!g_raw_tiles || !use_packed()) return 0u;
  unsigned bits = 0;
  for (int k = 0; k < 5; ++k)
    if (c->view[k] && (reinterpret_cast<uintptr_t>(c->view[k]) & 15) == 0) bits |= 1u << k;
  return bits;
}

FrameView make_view(const rpe_ctx* c) {
  FrameView f;
  f.bv = c->view[A_BV];
  f.xc = c->view[A_XC];
  f.nc = c->view[A_NC];
  f.xw = c->view[A_XW];
  f.nw = c->view[A_NW];
  f.n = c->n;
  f.npairs_pad = c->npairs_pad;
  f.pk = c->d_pk;
  f.pk_kind = c->pk_kind;
  f.pk_f4_per_pair = c->pk_kind >= 0 ? f4_per_pair(c->pk_kind) : 0;
  f.raw_aligned = raw_aligned_bits(c);
  if (c->pk_kind < 0) {  // nothing packed (raw-array scorer): the pair count still rounds up to whole rescan groups
    const int npairs = (c->n + 1) / 2;
    f.npairs_pad = ((npairs + kSubPairs - 1) / kSubPairs) * kSubPairs;
  }
  return f;
}

int ensure_corr_capacity(rpe_ctx* ctx, int n, bool own_copy) {
  if (own_copy && (size_t)n > ctx->cap_n) {
    for (int k = 0; k < 5; ++k) {
      if (ctx->d_raw[k]) cudaFree(ctx->d_raw[k]);
      ctx->d_raw[k] = nullptr;
    }
    const size_t cap = (size_t)n + (size_t)n / 8 + 64;
    for (int k = 0; k < 5; ++k) CK(cudaMalloc(&ctx->d_raw[k], cap * 3 * sizeof(float)));
    ctx->cap_n = cap;
  }
  const size_t mask_need = (size_t)n * 3;
  if (mask_need > ctx->mask_cap) {
    if (ctx->d_mask) cudaFree(ctx->d_mask);
    CK(cudaMalloc(&ctx->d_mask, mask_need * sizeof(int16_t)));
    ctx->mask_cap = mask_need;
  }
  const size_t bits_need = 3 * (((size_t)n + 31) / 32);
  if (bits_need > ctx->maskbits_cap) {
    if (ctx->d_maskbits) cudaFree(ctx->d_maskbits);
    ctx->d_maskbits = nullptr;
    CK(cudaMalloc(&ctx->d_maskbits, bits_need * sizeof(uint32_t)));
    ctx->maskbits_cap = bits_need;
  }
  const int blocks = (n + 255) / 256;
  if (blocks > ctx->rb.max_blocks) {
    if (ctx->rb.partials) cudaFree(ctx->rb.partials);
    CK(cudaMalloc(&ctx->rb.partials, (size_t)blocks * kMomentCount * sizeof(double)));
    ctx->rb.max_blocks = blocks;
  }
  return RPE_OK;
}

int ensure_hyp_capacity(rpe_ctx* ctx, int H, int slots) {
  if (slots > ctx->cap_slots) {
    if (ctx->d_gen) cudaFree(ctx->d_gen);
    if (ctx->d_fast) cudaFree(ctx->d_fast);
    if (ctx->d_votes) cudaFree(ctx->d_votes);
    const int cap = slots + slots / 4 + 256;
    CK(cudaMalloc(&ctx->d_gen, (size_t)cap * sizeof(HypGen)));
    CK(cudaMalloc(&ctx->d_fast, (size_t)cap * sizeof(HypFast)));
    CK(cudaMalloc(&ctx->d_votes, (size_t)cap * sizeof(int32_t)));
    ctx->cap_slots = cap;
  }
  if (H > ctx->cap_H) {
    if (ctx->d_samples) cudaFree(ctx->d_samples);
    const int cap = H + H / 4 + 256;
    CK(cudaMalloc(&ctx->d_samples, (size_t)cap * 4 * sizeof(int32_t)));
    ctx->cap_H = cap;
  }
  return RPE_OK;
}

// opt-in stale sample columns: run the prefix scan for this pass and return the table the generator should read
// (nullptr when the option is off or the family has no nl_2p slot)
int prepare_stale(rpe_ctx* ctx, int method, const int32_t* samples_dev, int hc, bool first_pass_of_frame, const int32_t** eff) {
  *eff = nullptr;
  if (!ctx->stale_cols || !(method == RPE_NL_SHINJI || method == RPE_NL_SHINJI_KNEIP)) return RPE_OK;
  if (hc > ctx->cap_stale) {
    if (ctx->d_stale_eff) cudaFree(ctx->d_stale_eff);
    ctx->d_stale_eff = nullptr;
    CK(cudaMalloc(&ctx->d_stale_eff, (size_t)(hc + 256) * 2 * sizeof(int32_t)));
    ctx->cap_stale = hc + 256;
  }
  if (!ctx->d_stale_carry) CK(cudaMalloc(&ctx->d_stale_carry, 2 * sizeof(int32_t)));
  launch_stale_cols(samples_dev, hc, ctx->view[A_XC], ctx->f64 ? ctx->view64[A_XC] : nullptr, ctx->n, ctx->d_stale_carry,
                    first_pass_of_frame, ctx->d_stale_eff, ctx->stream);
  ctx->launches++;
  *eff = ctx->d_stale_eff;
  return RPE_OK;
}

int ensure_packed(rpe_ctx* ctx, int kind) {
  if (ctx->pk_kind == kind) return RPE_OK;
  {
    FrameView fv = make_view(ctx);
    if (frame_raw_ok(fv, kind)) return RPE_OK;  // scored straight from the arrays
  }
  const int npairs = (ctx->n + 1) / 2;
  const int npad = ((npairs + kSubPairs - 1) / kSubPairs) * kSubPairs;
  const size_t bytes = (size_t)npad * f4_per_pair(kind) * sizeof(float4);
  if (bytes > ctx->pk_cap_bytes) {
    if (ctx->d_pk) cudaFree(ctx->d_pk);
    CK(cudaMalloc(&ctx->d_pk, bytes + bytes / 8));
    ctx->pk_cap_bytes = bytes + bytes / 8;
  }
  ctx->npairs_pad = npad;
  ctx->pk_kind = kind;
  FrameView f = make_view(ctx);
  launch_reset_corr_bound(ctx->d_stats, ctx->stream);
  launch_pack(f, kind, ctx->d_pk, ctx->d_stats, ctx->stream);
  ctx->launches += 2;
  return RPE_OK;
}

void stamp(rpe_ctx* ctx, int which) {
  if (ctx->timing && ctx->ev_ok) {
    cudaEventRecord(ctx->ev[which], ctx->stream);
    ctx->ev_recorded[which] = true;
  }
}

// ---- per-launch timing of the tiled scorer over a whole asynchronous region ------------------------------------
void fast_ring_fold(rpe_ctx* ctx, int k) {
  if (!ctx->ev_ring_live[k]) return;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_ring[2 * k], ctx->ev_ring[2 * k + 1]) == cudaSuccess) {
    ctx->fast_sum_ms += ms;
    ctx->fast_count += 1;
  } else {
    (void)cudaGetLastError();  // not finished yet (cannot happen 32 frames later on the same stream) or never run
  }
  ctx->ev_ring_live[k] = false;
}
int fast_ring_claim(rpe_ctx* ctx) {
  const int k = ctx->ev_ring_next;
  ctx->ev_ring_next = (k + 1) % rpe_ctx::kFastRing;
  fast_ring_fold(ctx, k);
  ctx->ev_ring_live[k] = true;
  return k;
}
void fast_ring_drain(rpe_ctx* ctx) {
  for (int k = 0; k < rpe_ctx::kFastRing; ++k) fast_ring_fold(ctx, k);
}

void quat_to_R_rowmajor(const float q[4], float R[9]) {
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.f - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1.f - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1.f - (txx + tyy);
}

void fill_result(const rpe_ctx* ctx, rpe_result* out, int slot, bool is_refit, bool gn) {
  const ReplayOut& p = ctx->h_pose[slot];
  memset(out, 0, sizeof(*out));
  for (int k = 0; k < 4; ++k) out->q[k] = p.q[k];
  for (int k = 0; k < 3; ++k) out->t[k] = p.t[k];
  quat_to_R_rowmajor(p.q, out->R);
  out->max_votes = p.max_votes;
  out->iter_final = p.iter_final;
  out->winner = p.winner;
  out->n_slots = p.n_slots;
  out->n_borderline = p.n_borderline;
  out->flags = p.flags;
  for (int k = 0; k < 3; ++k) out->n_inliers[k] = p.n_inliers[k];
  out->refit_ok = is_refit ? p.refit_ok : 0;
  if (!is_refit && (p.flags & 2)) {  // binary64 path: the accepted hypothesis as the double CPU path holds it
    for (int k = 0; k < 4; ++k) out->qd[k] = ctx->h_pose64[slot].q[k];
    for (int k = 0; k < 3; ++k) out->td[k] = ctx->h_pose64[slot].t[k];
  } else {
    for (int k = 0; k < 4; ++k) out->qd[k] = (double)p.q[k];
    for (int k = 0; k < 3; ++k) out->td[k] = (double)p.t[k];
  }
  if (gn) {
    out->refit_cost = ctx->h_gn_cost[slot];
    out->refit_evals = ctx->h_gn_evals[slot];
  }
}

// a delivered result that reports a worklist overflow asks for a larger list
void note_overflow(rpe_ctx* ctx, int slot) {
  if ((ctx->h_pose[slot].flags & 1) && !ctx->wl_fixed && ctx->wl_allocated < (1u << 26)) {
    const unsigned int want = ctx->wl_allocated * 4u;
    if (want > ctx->wl_want) ctx->wl_want = want;
  }
}

// rpe_set_mask_transfer(1): the device sent cols x ceil(n / 32) words of flag bits into the TAIL of the caller's mask
// buffer; turn them into the reference's n x cols matrix of 16-bit flags (setInlier layout) in place. The words are
// moved to a scratch vector first (77 KB for a dense frame), so the expansion may overwrite where they were.
struct ExpandLut {
  alignas(16) int16_t v[256][8];
  ExpandLut() {
    for (int b = 0; b < 256; ++b)
      for (int k = 0; k < 8; ++k) v[b][k] = (int16_t)((b >> k) & 1);
  }
};
void expand_mask(rpe_ctx* ctx, int16_t* mask, int n, int cols) {
  static const ExpandLut lut;
  const size_t wpc = ((size_t)n + 31) / 32, words = wpc * (size_t)cols;
  const size_t bytes = (size_t)n * cols * sizeof(int16_t);
  ctx->bits_scratch.resize(words);
  memcpy(ctx->bits_scratch.data(), reinterpret_cast<const char*>(mask) + bytes - words * sizeof(uint32_t), words * sizeof(uint32_t));
  for (int col = 0; col < cols; ++col) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(ctx->bits_scratch.data() + (size_t)col * wpc);
    int16_t* dst = mask + (size_t)col * n;
    const int full = n / 8;
    for (int i = 0; i < full; ++i) memcpy(dst + 8 * (size_t)i, lut.v[src[i]], 16);
    for (int c = 8 * full; c < n; ++c) dst[c] = (int16_t)((src[c >> 3] >> (c & 7)) & 1);
  }
}
void deliver(rpe_ctx* ctx, const rpe_ctx::Pending& p) {
  fill_result(ctx, p.out, p.slot, p.is_refit, p.gn);
  note_overflow(ctx, p.slot);
  if (p.mask_expand) expand_mask(ctx, p.mask_expand, p.mask_n, p.mask_cols);
  if (p.mask_fill0) {
    // column 0 of a family without the 2-D test: 0 once a hypothesis has been accepted, the adapters' initial 1 otherwise
    // (mask_kernel writes exactly this constant; it did not have to cross the bus)
    std::fill_n(p.mask_fill0, (size_t)p.mask_n, (int16_t)(ctx->h_pose[p.slot].winner >= 0 ? 0 : 1));
  }
}

int finish_pending(rpe_ctx* ctx) {
  for (const rpe_ctx::Pending& p : ctx->pending) deliver(ctx, p);
  ctx->pending.clear();
  if (ctx->timing_fast && ctx->ev_ok) fast_ring_drain(ctx);
  if ((ctx->timing || ctx->timing_fast) && ctx->ev_ok) {
    for (int k = 0; k < ST_COUNT; ++k) ctx->stage_ms[k] = 0.f;
    for (int k = 0; k < ST_TOTAL; ++k)
      if (ctx->ev_recorded[k] && ctx->ev_recorded[k + 1]) cudaEventElapsedTime(&ctx->stage_ms[k], ctx->ev[k], ctx->ev[k + 1]);
    if (ctx->ev_recorded[0]) {
      int last = 0;
      for (int k = 0; k <= ST_TOTAL; ++k)
        if (ctx->ev_recorded[k]) last = k;
      if (last > 0) cudaEventElapsedTime(&ctx->stage_ms[ST_TOTAL], ctx->ev[0], ctx->ev[last]);
    }
    if (ctx->ev_fast_recorded) cudaEventElapsedTime(&ctx->stage_ms[ST_FAST], ctx->ev_fast[0], ctx->ev_fast[1]);
    ctx->ev_fast_recorded = false;
    for (int k = 0; k <= ST_COUNT; ++k) ctx->ev_recorded[k] = false;
  }
  return RPE_OK;
}

// After a synchronisation: did an exchange kernel of the sharded mode give up on a peer? The latch (host word + the
// flag in the own block) is cleared so that the next frame starts clean.
int check_comm(rpe_ctx* ctx) {
  if (!ctx->h_peer_err || !*(volatile unsigned int*)ctx->h_peer_err) return RPE_OK;
  *ctx->h_peer_err = 0;
  if (ctx->d_peer_block) {
    cudaMemsetAsync(peer_flags(ctx->d_peer_block) + kPeerErrSlot, 0, sizeof(unsigned int), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  return fail(ctx, RPE_ERR_COMM, "a peer did not publish its votes within the time-out (the frame has no winner)");
}

// claim a pinned staging slot for an enqueued result; when all are in flight, wait for the OLDEST result only
// (its event) and deliver it, so that a long asynchronous stream of frames never drains the GPU queue
int claim_slot(rpe_ctx* ctx, int* slot) {
  if ((int)ctx->pending.size() >= kNumStaging) {
    const rpe_ctx::Pending p = ctx->pending.front();
    CK(cudaEventSynchronize(ctx->ev_slot[p.slot]));
    deliver(ctx, p);
    ctx->pending.pop_front();
  }
  *slot = ctx->next_slot;
  ctx->next_slot = (ctx->next_slot + 1) % kNumStaging;
  return RPE_OK;
}
// the result whose copies were just enqueued becomes pending
int push_pending(rpe_ctx* ctx, rpe_result* out, int slot, bool is_refit, bool gn, int16_t* mask_expand = nullptr, int mask_n = 0,
                 int mask_cols = 0, int16_t* mask_fill0 = nullptr) {
  CK(cudaEventRecord(ctx->ev_slot[slot], ctx->stream));
  rpe_ctx::Pending p{out, slot, is_refit, gn};
  p.mask_expand = mask_expand;
  p.mask_n = mask_n;
  p.mask_cols = mask_cols;
  p.mask_fill0 = mask_fill0;
  ctx->pending.push_back(p);
  return RPE_OK;
}

bool g_force_exact_multi = false;  // test hook: run the exact-order kernel for the non-AO families

// The borderline worklist starts at 4 Mi entries (32 MiB). A frame that overflows it is still answered exactly (the
// whole frame is rescored in the reference's operation order), and the list is grown x4 (up to 64 Mi entries) before
// the next scoring call so that dense frames with many near-threshold 2-D evaluations stay on the fast path.
constexpr unsigned int kWorklistInitial = 1u << 22, kWorklistMax = 1u << 26;
int alloc_worklist(rpe_ctx* ctx, unsigned int entries) {
  if (ctx->wl.entries) cudaFree(ctx->wl.entries);
  ctx->wl.entries = nullptr;
  CK(cudaMalloc(&ctx->wl.entries, (size_t)entries * sizeof(uint2)));
  ctx->wl_allocated = entries;
  ctx->wl.capacity = entries;
  return RPE_OK;
}
int grow_worklist(rpe_ctx* ctx) {
  if (ctx->wl_fixed || ctx->wl_want <= ctx->wl_allocated) return RPE_OK;
  CK(sync_stream(ctx));  // kernels in flight may still read the old list
  finish_pending(ctx);
  return alloc_worklist(ctx, ctx->wl_want);
}

// ---- the scorer lane -----------------------------------------------------------------------------------
// The tiled scorer fills every SM (one 512-thread CTA with > half of the shared memory per SM), so two of them never
// run side by side: when several contexts (streams) are in flight their scorers are launched into ONE per-device
// stream, in submission order, fenced against the owning context's stream by events. Other frames' small kernels
// (pack, generation, replay, mask, refinement) keep overlapping the running scorer from their own streams, and the
// CUDA events recorded around a scorer launch bracket its execution instead of its wait for the SMs.
// RPE_SCORER_LANE=0 in the environment launches scorers on the context's own stream instead.
struct ScorerLane {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  bool tried = false;
};
constexpr int kMaxLanes = 4;
ScorerLane g_lane[64][kMaxLanes];
std::atomic<unsigned int> g_lane_rr[64];
const bool g_lane_enabled = !(getenv("RPE_SCORER_LANE") && getenv("RPE_SCORER_LANE")[0] == '0');
// Scorers of consecutive calls alternate between TWO lane streams: the next scorer's CTAs take over an SM the moment the
// previous scorer's CTA leaves it, which hides the launch gap and the spread of the CTAs' finishing times (resident
// frames 0.1813 -> 0.1789 ms, host frames 0.208 -> 0.193 ms per frame, round 2: under host-to-device DMA traffic a
// launch into a busy lane starts ~30 us late). RPE_SCORER_LANES=k overrides (1 = strictly back to back, up to 4).
const int g_lane_count = getenv("RPE_SCORER_LANES") ? std::max(1, std::min(kMaxLanes, atoi(getenv("RPE_SCORER_LANES")))) : 2;

ScorerLane* lane_for(rpe_ctx* ctx) {
  if (!g_lane_enabled || ctx->device < 0 || ctx->device >= 64 || !ctx->ev_ok) return nullptr;
  const unsigned int which = g_lane_count > 1 ? g_lane_rr[ctx->device].fetch_add(1u, std::memory_order_relaxed) % (unsigned int)g_lane_count : 0u;
  ScorerLane* L = &g_lane[ctx->device][which];
  std::lock_guard<std::mutex> g(L->mu);
  if (!L->tried) {
    L->tried = true;
    if (cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking) != cudaSuccess) {
      (void)cudaGetLastError();
      L->stream = nullptr;
    }
  }
  return L->stream ? L : nullptr;
}

// ---- time the device spends in scorer launches (two lanes overlap the head of one launch with the tail of the
// previous one, so a launch's own event pair no longer is its cost): per device, every timed launch leaves its
// [start, end] event interval (start = the moment the launch could begin: its inputs were ready and its lane free), and
// rpe_scorer_busy_stats returns the length of the UNION of those intervals and their number. Idle time between
// launches is not counted, time a launch spends queued behind another one is counted once. Events live in a
// per-device ring and are turned into times (relative to one reference event) 64 launches later — far more than
// can be in flight.
struct LaneClock {
  static constexpr int kRing = 128;
  std::mutex mu;  // also orders the submissions of all lanes of the device
  cudaEvent_t start[kRing] = {}, end[kRing] = {}, ref = nullptr;
  bool live[kRing] = {};
  bool ref_set = false;
  int next = 0;
  bool tried = false, ok = false;
  std::vector<std::pair<float, float>> iv;  // ms since ref
};
LaneClock g_clock[64];

void lane_clock_fold(LaneClock& c, int j) {
  if (!c.live[j]) return;
  float a = 0.f, b = 0.f;
  if (cudaEventElapsedTime(&a, c.ref, c.start[j]) != cudaSuccess || cudaEventElapsedTime(&b, c.ref, c.end[j]) != cudaSuccess) {
    (void)cudaGetLastError();  // not finished: leave it for the next drain
    return;
  }
  if (c.iv.size() < ((size_t)1 << 22)) c.iv.emplace_back(a, b);
  c.live[j] = false;
}
// caller holds c.mu; `stream` is the lane the launch goes to
int lane_clock_claim(LaneClock& c, cudaStream_t stream) {
  if (!c.tried) {
    c.tried = true;
    c.ok = cudaEventCreate(&c.ref) == cudaSuccess;
    for (int k = 0; k < LaneClock::kRing && c.ok; ++k)
      c.ok = cudaEventCreate(&c.start[k]) == cudaSuccess && cudaEventCreate(&c.end[k]) == cudaSuccess;
    if (!c.ok) (void)cudaGetLastError();
  }
  if (!c.ok) return -1;
  if (!c.ref_set) {
    cudaEventRecord(c.ref, stream);
    c.ref_set = true;
  }
  const int k = c.next;
  c.next = (k + 1) % LaneClock::kRing;
  lane_clock_fold(c, (k + LaneClock::kRing / 2) % LaneClock::kRing);
  if (c.live[k]) lane_clock_fold(c, k);  // (only if the half-ring fold found it unfinished)
  c.live[k] = true;
  return k;
}

// score the slot range with the best available kernel, including the exact fix-up
int score_range(rpe_ctx* ctx, int method, int slot_begin, int slot_end, Thresh th) {
  FrameView f = make_view(ctx);
  if (method == RPE_SHINJI || !g_force_exact_multi) {
    const bool tm = ctx->timing_fast && ctx->ev_ok;
    ScorerLane* lane = lane_for(ctx);
    int nseg = 0;
    if (int rcw = grow_worklist(ctx)) return rcw;
    const int rk = tm ? fast_ring_claim(ctx) : 0;
    if (lane) {
      CK(cudaEventRecord(ctx->ev_lane[0], ctx->stream));  // everything the scorer reads has been enqueued before this
      LaneClock& clk = g_clock[ctx->device];
      std::lock_guard<std::mutex> gd(clk.mu);
      std::lock_guard<std::mutex> g(lane->mu);
      CK(cudaStreamWaitEvent(lane->stream, ctx->ev_lane[0], 0));
      const int ck = tm ? lane_clock_claim(clk, lane->stream) : -1;
      if (tm) cudaEventRecord(ctx->ev_ring[2 * rk], lane->stream);
      if (ck >= 0) cudaEventRecord(clk.start[ck], lane->stream);
      nseg = launch_score_fast(method, f, ctx->d_gen, ctx->d_fast, slot_begin, slot_end, th, ctx->d_votes, ctx->d_stats,
                               ctx->wl, ctx->num_sms, lane->stream, 0, 0, (int)(lane - &g_lane[ctx->device][0]));
      if (ck >= 0) cudaEventRecord(clk.end[ck], lane->stream);
      if (tm) cudaEventRecord(ctx->ev_ring[2 * rk + 1], lane->stream);
      CK(cudaEventRecord(ctx->ev_lane[1], lane->stream));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane[1], 0));
    } else {
      if (tm) cudaEventRecord(ctx->ev_ring[2 * rk], ctx->stream);
      nseg = launch_score_fast(method, f, ctx->d_gen, ctx->d_fast, slot_begin, slot_end, th, ctx->d_votes, ctx->d_stats,
                               ctx->wl, ctx->num_sms, ctx->stream);
      if (tm) cudaEventRecord(ctx->ev_ring[2 * rk + 1], ctx->stream);
    }
    if (tm) {  // rpe_last_stage_ms[7] keeps reporting the LAST launch
      ctx->ev_fast[0] = ctx->ev_ring[2 * rk];
      ctx->ev_fast[1] = ctx->ev_ring[2 * rk + 1];
      ctx->ev_fast_recorded = true;
    }
    launch_fixup(method, f, ctx->d_gen, th, ctx->d_votes, ctx->d_stats, ctx->wl, nseg, slot_begin, slot_end, ctx->stream);
    launch_score_exact(method, f, ctx->d_gen, slot_begin, slot_end, th, ctx->d_votes, ctx->d_stats, true, ctx->num_sms,
                       ctx->stream);
    ctx->launches += 3;
  } else {
    launch_score_exact(method, f, ctx->d_gen, slot_begin, slot_end, th, ctx->d_votes, ctx->d_stats, false, ctx->num_sms,
                       ctx->stream);
    ctx->launches += 2;
  }
  return RPE_OK;
}

// mask of the accepted hypothesis (+ fused Kabsch), result and mask copies. The replay must already be enqueued.
int do_finish(rpe_ctx* ctx, int method, Thresh th, rpe_result* out, int16_t* mask, bool blocking) {
  FrameView f = make_view(ctx);
  stamp(ctx, ST_MASK);
  ctx->mask_cols = method_mask_cols(method);
  order_after_mask_copy(ctx);
  const size_t mask_bytes = (size_t)ctx->n * ctx->mask_cols * sizeof(int16_t);
  // bit form for the host (16 x fewer bytes on the bus, expanded by whoever collects the result): asynchronous calls only —
  // a blocking caller would wait for the expansion, while the plain copy hides behind the refits on the side stream
  const bool as_bits = mask && ctx->mask_transfer == 1 && !blocking && mask_bytes >= (size_t)256 * 1024 && ctx->d_maskbits;
  launch_mask(method, f, ctx->d_pose, th, ctx->d_mask, ctx->d_kabsch, ctx->rb, ctx->d_stats, ctx->stream,
              as_bits ? ctx->d_maskbits : nullptr);
  ctx->launches += 1;
  ctx->kabsch_valid = method_uses_3d(method);
  ctx->suff_valid = true;
  ctx->last_th = th;
  stamp(ctx, ST_GN);
  int slot = 0;
  int rc = claim_slot(ctx, &slot);
  if (rc) return rc;
  CK(cudaMemcpyAsync(&ctx->h_pose[slot], ctx->d_pose, sizeof(ReplayOut), cudaMemcpyDeviceToHost, ctx->stream));
  if (as_bits) {
    const size_t wbytes = (((size_t)ctx->n + 31) / 32) * ctx->mask_cols * sizeof(uint32_t);
    CK(cudaMemcpyAsync(reinterpret_cast<char*>(mask) + mask_bytes - wbytes, ctx->d_maskbits, wbytes, cudaMemcpyDeviceToHost, ctx->stream));
    return push_pending(ctx, out, slot, false, false, mask, ctx->n, ctx->mask_cols);
  }
  int16_t* fill0 = nullptr;
  if (mask) {
    // rpe_set_mask_transfer(2): column 0 of a family without the 2-D test is a constant the collector can write itself
    const bool skip0 = ctx->mask_transfer == 2 && ctx->mask_cols >= 2 && !method_uses_2d(method) && mask_bytes >= (size_t)256 * 1024;
    const size_t off = skip0 ? (size_t)ctx->n : 0;
    const size_t bytes = mask_bytes - off * sizeof(int16_t);
    if (skip0) fill0 = mask;
    if (ctx->d2h_stream && bytes >= (size_t)256 * 1024) {
      CK(cudaEventRecord(ctx->ev_mask_ready, ctx->stream));
      CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_mask_ready, 0));
      // (one transfer: cutting it into pieces so that the refits' small result copies can slip in between was measured
      // slower — 0.438 against 0.412 ms per blocking frame, round 2)
      CK(cudaMemcpyAsync(mask + off, ctx->d_mask + off, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
      CK(cudaEventRecord(ctx->ev_mask_copied, ctx->d2h_stream));
      ctx->mask_copy_pending = true;
    } else {
      CK(cudaMemcpyAsync(mask + off, ctx->d_mask + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  if (int rcp = push_pending(ctx, out, slot, false, false, nullptr, ctx->n, ctx->mask_cols, fill0)) return rcp;
  if (blocking) {
    CK(sync_stream(ctx));
    finish_pending(ctx);
    return check_comm(ctx);
  }
  return RPE_OK;
}

constexpr int kMaxPassIters = 8192;  // iterations generated + scored per device pass, at most
constexpr int kFirstPassIters = 1024;  // a longer Iter is scored progressively: 1024, 2048, 4096, 8192, 8192, ... iterations,
                                       // looking at the adaptive bound in between (the reference rarely gets past a few hundred)

// ---- binary64 path ---------------------------------------------------------------------------------------
bool g_f64_exact_only = false;  // test hook: score every evaluation in binary64 (no binary32 prefilter)
FrameView64 make_view64(const rpe_ctx* c) {
  FrameView64 f;
  f.bv = c->view64[A_BV];
  f.xc = c->view64[A_XC];
  f.nc = c->view64[A_NC];
  f.xw = c->view64[A_XW];
  f.nw = c->view64[A_NW];
  f.n = c->n;
  return f;
}

int do_ransac64(rpe_ctx* ctx, int method, const int32_t* samples, rpe_sample_fn fn, void* fn_user, int H, double thr3d,
                double cos_thr2d, double cos_thrN, double confidence, rpe_result* out, int16_t* mask) {
  if (!method_ok(method) || (!samples && !fn) || H <= 0 || !out)
    return fail(ctx, RPE_ERR_ARG, "bad argument to rpe_ransac");
  int rc = check_arrays(ctx, method);
  if (rc) return rc;
  if (!ctx->f64 || (size_t)ctx->n > ctx->cap_n64) return fail(ctx, RPE_ERR_STATE, "no binary64 arrays of this frame on the device");
  for (int k = 0; k < 5; ++k)
    if (ctx->view[k] && !ctx->view64[k]) return fail(ctx, RPE_ERR_STATE, "binary64 copy of an array is missing");
  ENTER(ctx);
  const int S = method_slots(method);
  const Thresh64 th = {thr3d, cos_thr2d, cos_thrN};
  const int pass_cap = H < kMaxPassIters ? H : kMaxPassIters;
  rc = ensure_hyp_capacity(ctx, pass_cap, pass_cap * S);
  if (rc) return rc;
  if (pass_cap * S > ctx->cap_slots64) {
    if (ctx->d_gen64) cudaFree(ctx->d_gen64);
    ctx->d_gen64 = nullptr;
    const int cap = pass_cap * S + 256;
    CK(cudaMalloc(&ctx->d_gen64, (size_t)cap * sizeof(HypGen64)));
    ctx->cap_slots64 = cap;
  }
  bool samples_on_device = false;
  if (fn) {
    if (ctx->h_samples_cap < pass_cap) {
      if (ctx->h_samples) cudaFreeHost(ctx->h_samples);
      ctx->h_samples = nullptr;
      CK(cudaMallocHost(&ctx->h_samples, (size_t)pass_cap * 4 * sizeof(int32_t)));
      ctx->h_samples_cap = pass_cap;
    }
  } else {
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, samples);
    if (pe != cudaSuccess) (void)cudaGetLastError();
    samples_on_device = pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  }
  const FrameView64 f = make_view64(ctx);
  if (!ctx->stats_clean) {
    launch_reset_stats(ctx->d_stats, ctx->stream);
    ctx->launches++;
  }
  if (!g_f64_exact_only) {
    rc = ensure_packed(ctx, kind_for_method(method));  // pair records from the float copies
    if (rc) return rc;
  }
  launch_replay64_begin(ctx->d_rs64, H, ctx->stream);
  ctx->launches++;
  const bool single = H <= ctx->first_pass;
  int pass = single ? H : ctx->first_pass;
  for (int base = 0; base < H; base += pass, pass = (2 * pass < kMaxPassIters ? 2 * pass : kMaxPassIters)) {
    const int hc = (H - base) < pass ? (H - base) : pass;
    const int32_t* chunk = samples ? samples + (size_t)base * 4 : nullptr;
    if (fn) {
      if (fn(fn_user, base, hc, ctx->h_samples) != 0) return fail(ctx, RPE_ERR_ARG, "the sample callback failed");
      chunk = ctx->h_samples;
    }
    const int32_t* samples_dev = chunk;
    if (!samples_on_device) {
      CK(cudaMemcpyAsync(ctx->d_samples, chunk, (size_t)hc * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      samples_dev = ctx->d_samples;
    }
    const int32_t* stale_eff = nullptr;
    if (int rcs = prepare_stale(ctx, method, samples_dev, hc, base == 0, &stale_eff)) return rcs;
    launch_hypgen64(method, f, samples_dev, hc, ctx->d_gen64, ctx->d_votes, ctx->stream, stale_eff);
    ctx->launches++;
    if (g_f64_exact_only) {
      launch_score64(method, f, ctx->d_gen64, hc * S, th, ctx->d_votes, ctx->num_sms, nullptr, ctx->stream);
      ctx->launches++;
    } else {
      // binary32 prefilter on the float copies + binary64 evaluation of what falls inside the guard bands
      const FrameView f32 = make_view(ctx);
      const Thresh thf = {(float)thr3d, (float)cos_thr2d, (float)cos_thrN};
      launch_derive_fast64(ctx->d_gen64, ctx->d_gen, ctx->d_fast, hc * S, ctx->stream);
      if (int rcw = grow_worklist(ctx)) return rcw;
      int nseg = 0;
      ScorerLane* lane = lane_for(ctx);
      if (lane) {
        CK(cudaEventRecord(ctx->ev_lane[0], ctx->stream));
        std::lock_guard<std::mutex> g(lane->mu);
        CK(cudaStreamWaitEvent(lane->stream, ctx->ev_lane[0], 0));
        nseg = launch_score_fast(method, f32, ctx->d_gen, ctx->d_fast, 0, hc * S, thf, ctx->d_votes, ctx->d_stats, ctx->wl,
                                 ctx->num_sms, lane->stream);
        CK(cudaEventRecord(ctx->ev_lane[1], lane->stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane[1], 0));
      } else {
        nseg = launch_score_fast(method, f32, ctx->d_gen, ctx->d_fast, 0, hc * S, thf, ctx->d_votes, ctx->d_stats, ctx->wl,
                                 ctx->num_sms, ctx->stream);
      }
      launch_fixup64(method, f, ctx->d_gen64, th, ctx->d_votes, ctx->d_stats, ctx->wl, nseg, ctx->stream);
      launch_score64(method, f, ctx->d_gen64, hc * S, th, ctx->d_votes, ctx->num_sms, ctx->d_stats, ctx->stream);
      ctx->launches += 5;
    }
    launch_replay64(method, ctx->d_gen64, ctx->d_votes, hc, base, ctx->n, confidence, ctx->d_stats, ctx->d_rs64, ctx->d_pose,
                    ctx->d_pose64, single, ctx->stream);
    ctx->launches++;
    ctx->n_slots = hc * S;
    ctx->cur_method = method;
    if (!single) {
      CK(cudaMemcpyAsync(ctx->h_rs64, ctx->d_rs64, sizeof(ReplayState64), cudaMemcpyDeviceToHost, ctx->stream));
      CK(sync_stream(ctx));
      if (ctx->h_rs64->stop != 0 || base + hc >= H) {
        launch_replay64(method, ctx->d_gen64, ctx->d_votes, 0, base + hc, ctx->n, confidence, ctx->d_stats, ctx->d_rs64,
                        ctx->d_pose, ctx->d_pose64, true, ctx->stream);
        ctx->launches++;
        break;
      }
    }
  }
  ctx->mask_cols = method_mask_cols(method);
  order_after_mask_copy(ctx);
  launch_mask64(method, f, ctx->d_pose, ctx->d_pose64, th, ctx->d_mask, ctx->num_sms, ctx->stream);
  ctx->launches++;
  ctx->kabsch_valid = false;  // refits gather their statistics from the float copies when asked
  ctx->suff_valid = false;
  ctx->stats_clean = true;  // replay64 leaves the per-pass counters clean
  int slot = 0;
  rc = claim_slot(ctx, &slot);
  if (rc) return rc;
  CK(cudaMemcpyAsync(&ctx->h_pose[slot], ctx->d_pose, sizeof(ReplayOut), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&ctx->h_pose64[slot], ctx->d_pose64, sizeof(Pose64), cudaMemcpyDeviceToHost, ctx->stream));
  if (mask)
    CK(cudaMemcpyAsync(mask, ctx->d_mask, (size_t)ctx->n * ctx->mask_cols * sizeof(int16_t), cudaMemcpyDeviceToHost,
                       ctx->stream));
  if (int rcp = push_pending(ctx, out, slot, false, false)) return rcp;
  CK(sync_stream(ctx));
  finish_pending(ctx);
  return RPE_OK;
}

int do_ransac(rpe_ctx* ctx, int method, const int32_t* samples, rpe_sample_fn fn, void* fn_user, int H, float thr3d,
              float cos_thr2d, float cos_thrN, float confidence, rpe_result* out, int16_t* mask, bool blocking) {
  if (!ctx) return RPE_ERR_ARG;
  if (!method_ok(method) || (!samples && !fn) || H <= 0 || !out)
    return fail(ctx, RPE_ERR_ARG, "bad argument to rpe_ransac");
  if (ctx->f64)  // arrays uploaded in binary64: thresholds given as float are widened (use rpe_ransac_f64 for exact ones)
    return do_ransac64(ctx, method, samples, fn, fn_user, H, (double)thr3d, (double)cos_thr2d, (double)cos_thrN,
                       (double)confidence, out, mask);
  int rc = check_arrays(ctx, method);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  const int S = method_slots(method);
  const Thresh th = {thr3d, cos_thr2d, cos_thrN};
  if (!ctx->upload_stamped) stamp(ctx, ST_UPLOAD);
  ctx->upload_stamped = false;
  const int pass_cap = H < kMaxPassIters ? H : kMaxPassIters;
  rc = ensure_hyp_capacity(ctx, pass_cap, pass_cap * S);
  if (rc) return rc;
  // sample table: device pointers are used in place; host memory is copied on the stream (pageable memory is
  // staged by the driver before the call returns, page-locked memory must stay alive until rpe_sync)
  bool samples_on_device = false;
  if (fn) {  // rows are produced pass by pass into a page-locked staging buffer
    if (ctx->h_samples_cap < pass_cap) {
      if (ctx->h_samples) cudaFreeHost(ctx->h_samples);
      ctx->h_samples = nullptr;
      CK(cudaMallocHost(&ctx->h_samples, (size_t)pass_cap * 4 * sizeof(int32_t)));
      ctx->h_samples_cap = pass_cap;
    }
  } else {
    cudaPointerAttributes attr;
    const cudaError_t pe = cudaPointerGetAttributes(&attr, samples);
    if (pe != cudaSuccess) (void)cudaGetLastError();
    samples_on_device = pe == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  }
  const bool single = H <= ctx->first_pass;
  // ---- deferred upload (rpe_set_upload_overlap): generate from the page-locked host arrays, then copy and score chunk by chunk
  if (ctx->deferred) {
    ScorerLane* lane = lane_for(ctx);
    FrameView fv = make_view(ctx);
    if (!(method == RPE_SHINJI && single && !fn && H * S > 512 && lane && frame_raw_ok(fv, 2) && !ctx->timing && !ctx->timing_fast &&
          !ctx->stale_cols && ctx->wl_want <= ctx->wl_allocated && rpe::score_variant() == rpe::kDefaultScoreVariant &&
          (ctx->overlap_chunks != 1 || H * S <= 1024))) {  // (stream mode: one hypothesis column, the frame crosses the bus once)
      if (int rcf = flush_deferred(ctx)) return rcf;
    } else {
      ctx->deferred = false;
      cudaStream_t es = ctx->early_stream;
      if (ctx->n_chunks == 1 && ctx->overlap_chunks == 1) {
        // ---- stream mode: nothing is copied by the copy engine. The generator reads its sample points from the host
        // arrays, the scorer's bulk-TMA loads read the frame itself from the page-locked host arrays while it scores,
        // and it leaves the device copy behind that the fix-up, mask and refit kernels use.
        CK(cudaEventRecord(ctx->ev_prev, ctx->stream));
        CK(cudaStreamWaitEvent(es, ctx->ev_prev, 0));
        if (!ctx->stats_clean) {
          launch_reset_stats(ctx->d_stats, es);
          ctx->launches++;
        }
        const int32_t* sdev = samples;
        if (!samples_on_device) {
          CK(cudaMemcpyAsync(ctx->d_samples, samples, (size_t)H * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, es));
          sdev = ctx->d_samples;
        }
        FrameView fh = fv;
        fh.xw = ctx->host_view[A_XW];
        fh.xc = ctx->host_view[A_XC];
        launch_hypgen(method, fh, sdev, H, ctx->d_gen, ctx->d_fast, ctx->d_votes, ctx->d_stats, es);
        ctx->launches++;
        ctx->n_slots = H * S;
        ctx->cur_method = method;
        CK(cudaEventRecord(ctx->ev_early, es));
        int nseg = 0;
        {
          const int li = (int)(lane - &g_lane[ctx->device][0]);
          std::lock_guard<std::mutex> g(lane->mu);
          CK(cudaStreamWaitEvent(lane->stream, ctx->ev_early, 0));
          nseg = launch_score3d_stream(fh, ctx->d_raw[A_XW], ctx->d_raw[A_XC], ctx->d_gen, ctx->d_fast, 0, H * S, th.thr3d,
                                       ctx->d_votes, ctx->d_stats, ctx->wl, ctx->num_sms, lane->stream);
          ctx->launches++;
          CK(cudaEventRecord(ctx->ev_lane_done[li], lane->stream));
          CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane_done[li], 0));
        }
        launch_fixup(method, fv, ctx->d_gen, th, ctx->d_votes, ctx->d_stats, ctx->wl, nseg, 0, H * S, ctx->stream);
        launch_score_exact(method, fv, ctx->d_gen, 0, H * S, th, ctx->d_votes, ctx->d_stats, true, ctx->num_sms, ctx->stream);
        launch_replay(method, ctx->d_gen, ctx->d_votes, H, 0, ctx->n, confidence, ctx->d_stats, ctx->d_rs, ctx->d_pose, true, H,
                      ctx->stream);
        ctx->launches += 3;
        ctx->stats_clean = true;
        return do_finish(ctx, method, th, out, mask, blocking);
      }
      static const bool trace = getenv("RPE_OVERLAP_TRACE") != nullptr;  // debugging aid: device timeline of this path on stderr
      static thread_local cudaEvent_t tev[40] = {};
      int ntev = 0;
      if (trace && !tev[0])
        for (int i = 0; i < 40; ++i) cudaEventCreate(&tev[i]);
      auto mark = [&](cudaStream_t st) {
        if (trace && ntev < 40) cudaEventRecord(tev[ntev++], st);
      };
      CK(cudaEventRecord(ctx->ev_prev, ctx->stream));  // everything enqueued so far (it may still read the arrays / the tables)
      CK(cudaStreamWaitEvent(es, ctx->ev_prev, 0));
      mark(es);
      if (!ctx->stats_clean) {
        launch_reset_stats(ctx->d_stats, es);
        ctx->launches++;
      }
      const int32_t* samples_dev = samples;
      if (!samples_on_device) {
        CK(cudaMemcpyAsync(ctx->d_samples, samples, (size_t)H * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, es));
        samples_dev = ctx->d_samples;
      }
      const int C = ctx->n_chunks;
      const unsigned int seg_cap = ctx->wl.capacity / (unsigned int)(C * ctx->num_sms);
      CK(cudaMemsetAsync(ctx->wl.counts, 0, (size_t)C * ctx->num_sms * sizeof(unsigned int), es));
      FrameView fh = fv;  // the generator reads its 3 H sample points over PCIe from the host arrays
      fh.xw = ctx->host_view[A_XW];
      fh.xc = ctx->host_view[A_XC];
      launch_hypgen(method, fh, samples_dev, H, ctx->d_gen, ctx->d_fast, ctx->d_votes, ctx->d_stats, es);
      ctx->launches++;
      ctx->n_slots = H * S;
      ctx->cur_method = method;
      CK(cudaEventRecord(ctx->ev_early, es));
      mark(es);  // 1: generator done
      // Now the frame, chunk by chunk on the copy stream. The copies start BESIDE the generator (its reads get slower,
      // 40 -> 64 us, but the first chunk has landed when it ends; RPE_OVERLAP_EAGER=0: behind it), and every chunk's scorer is
      // enqueued right after its copy, so that the host's issue order never holds a launch back (blocking frame 0.413 ->
      // 0.396 ms, profiles/r02_latency_breakdown.md). The chunk scorers alternate between the device's lanes like the
      // scorers of consecutive frames do: the head of one chunk's launch overlaps the tail of the previous one (they add
      // into the same vote table with atomics).
      static const bool eager = !(getenv("RPE_OVERLAP_EAGER") && getenv("RPE_OVERLAP_EAGER")[0] == '0');
      CK(cudaStreamWaitEvent(ctx->copy_stream, eager ? ctx->ev_prev : ctx->ev_early, 0));
      bool lane_used[kMaxLanes] = {};
      for (int c = 0; c < C; ++c) {
        const int c0 = c * ctx->chunk_corr;
        const int cnt = (ctx->n - c0) < ctx->chunk_corr ? (ctx->n - c0) : ctx->chunk_corr;
        for (int k = 0; k < 5; ++k)
          if (ctx->host_src[k] && ctx->view[k])
            CK(cudaMemcpyAsync(ctx->d_raw[k] + 3 * (size_t)c0, ctx->host_src[k] + 3 * (size_t)c0, (size_t)cnt * 3 * sizeof(float),
                               cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
        mark(ctx->copy_stream);  // chunk c landed
        ScorerLane* L = c == 0 ? lane : lane_for(ctx);
        if (!L) L = lane;
        const int li = (int)(L - &g_lane[ctx->device][0]);
        std::lock_guard<std::mutex> g(L->mu);
        if (!lane_used[li]) CK(cudaStreamWaitEvent(L->stream, ctx->ev_early, 0));
        lane_used[li] = true;
        cudaStream_t lane_stream = L->stream;
        CK(cudaStreamWaitEvent(lane_stream, ctx->ev_chunk[c], 0));
        FrameView fc = fv;
        fc.xw = fv.xw + 3 * (size_t)c0;
        fc.xc = fv.xc + 3 * (size_t)c0;
        fc.n = cnt;
        const int npairs = (cnt + 1) / 2;
        fc.npairs_pad = ((npairs + kSubPairs - 1) / kSubPairs) * kSubPairs;
        Worklist wc = ctx->wl;
        wc.entries = ctx->wl.entries + (size_t)c * ctx->num_sms * seg_cap;
        wc.counts = ctx->wl.counts + (size_t)c * ctx->num_sms;
        mark(lane_stream);  // chunk c: the lane has passed its waits
        launch_score_fast(method, fc, ctx->d_gen, ctx->d_fast, 0, H * S, th, ctx->d_votes, ctx->d_stats, wc, ctx->num_sms,
                          lane_stream, c0, seg_cap);
        ctx->launches++;
        mark(lane_stream);  // chunk c scored
        CK(cudaEventRecord(ctx->ev_lane_done[li], lane_stream));  // (the last record of a lane is the one that counts)
      }
      CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[C - 1], 0));  // the context's own stream sees the whole frame
      for (int li = 0; li < kMaxLanes; ++li)
        if (lane_used[li]) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane_done[li], 0));
      Worklist wall = ctx->wl;
      wall.capacity = seg_cap * (unsigned int)(C * ctx->num_sms);
      launch_fixup(method, fv, ctx->d_gen, th, ctx->d_votes, ctx->d_stats, wall, C * ctx->num_sms, 0, H * S, ctx->stream);
      launch_score_exact(method, fv, ctx->d_gen, 0, H * S, th, ctx->d_votes, ctx->d_stats, true, ctx->num_sms, ctx->stream);
      launch_replay(method, ctx->d_gen, ctx->d_votes, H, 0, ctx->n, confidence, ctx->d_stats, ctx->d_rs, ctx->d_pose, true, H,
                    ctx->stream);
      ctx->launches += 3;
      ctx->stats_clean = true;
      mark(ctx->stream);  // fix-up + replay done
      const int rcf = do_finish(ctx, method, th, out, mask, blocking);
      if (trace) {
        mark(ctx->stream);  // mask + result copies enqueued behind it
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[rpe overlap trace] us after the call: generator, %d x (chunk landed, scorer may start, chunk scored), fix-up+replay, mask:", C);
        for (int i = 1; i < ntev; ++i) {
          float ms = 0.f;
          cudaEventElapsedTime(&ms, tev[0], tev[i]);
          fprintf(stderr, " %.1f", ms * 1e3f);
        }
        fprintf(stderr, "\n");
      }
      return rcf;
    }
  }
  if (!ctx->stats_clean) {
    launch_reset_stats(ctx->d_stats, ctx->stream);
    ctx->launches++;
  }
  if (method == RPE_SHINJI || !g_force_exact_multi) {
    rc = ensure_packed(ctx, kind_for_method(method));
    if (rc) return rc;
  }
  int pass = single ? H : ctx->first_pass;
  for (int base = 0; base < H; base += pass, pass = (2 * pass < kMaxPassIters ? 2 * pass : kMaxPassIters)) {
    const int hc = (H - base) < pass ? (H - base) : pass;
    const int32_t* chunk = samples ? samples + (size_t)base * 4 : nullptr;
    if (fn) {
      if (fn(fn_user, base, hc, ctx->h_samples) != 0) return fail(ctx, RPE_ERR_ARG, "the sample callback failed");
      chunk = ctx->h_samples;
    }
    const int32_t* samples_dev = chunk;
    if (!samples_on_device) {
      CK(cudaMemcpyAsync(ctx->d_samples, chunk, (size_t)hc * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      samples_dev = ctx->d_samples;
    }
    if (base == 0) stamp(ctx, ST_GEN);
    FrameView f = make_view(ctx);
    const int32_t* stale_eff = nullptr;
    if (int rcs = prepare_stale(ctx, method, samples_dev, hc, base == 0, &stale_eff)) return rcs;
    launch_hypgen(method, f, samples_dev, hc, ctx->d_gen, ctx->d_fast, ctx->d_votes, ctx->d_stats, ctx->stream, stale_eff);
    ctx->launches++;
    ctx->n_slots = hc * S;
    ctx->cur_method = method;
    if (base == 0) stamp(ctx, ST_SCORE);
    rc = score_range(ctx, method, 0, hc * S, th);
    if (rc) return rc;
    if (base == 0) stamp(ctx, ST_REPLAY);
    launch_replay(method, ctx->d_gen, ctx->d_votes, hc, base, ctx->n, confidence, ctx->d_stats, ctx->d_rs, ctx->d_pose,
                  single, base == 0 ? H : -1, ctx->stream);
    ctx->launches++;
    ctx->stats_clean = true;
    if (!single) {
      // the caller's Iter exceeds one pass: look at the adaptive bound before generating more hypotheses
      CK(cudaMemcpyAsync(ctx->h_rs, ctx->d_rs, sizeof(ReplayState), cudaMemcpyDeviceToHost, ctx->stream));
      CK(sync_stream(ctx));
      const bool last = ctx->h_rs->stop != 0 || base + hc >= H;
      if (last) {
        launch_replay(method, ctx->d_gen, ctx->d_votes, 0, base + hc, ctx->n, confidence, ctx->d_stats, ctx->d_rs,
                      ctx->d_pose, true, -1, ctx->stream);
        ctx->launches++;
        break;
      }
    }
  }
  return do_finish(ctx, method, th, out, mask, blocking);
}

}  // namespace

extern "C" {

int rpe_version(void) { return RPE_API_VERSION; }

const char* rpe_status_string(int status) {
  switch (status) {
    case RPE_OK: return "ok";
    case RPE_ERR_ARG: return "invalid argument";
    case RPE_ERR_CUDA: return "CUDA error";
    case RPE_ERR_STATE: return "invalid state";
    case RPE_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case RPE_ERR_COMM: return "communication error";
    case RPE_ERR_NOMEM: return "out of memory";
    default: return "unknown status";
  }
}

int rpe_device_count(int* count) {
  if (!count) return RPE_ERR_ARG;
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    (void)cudaGetLastError();
    return RPE_ERR_NO_DEVICE;
  }
  *count = n;
  return RPE_OK;
}

static int create_common(int device, void* stream, bool own, rpe_ctx** out) {
  if (!out) return RPE_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    return RPE_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) return RPE_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return RPE_ERR_CUDA;
  rpe_ctx* ctx = new rpe_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    (void)cudaGetLastError();
    delete ctx;
    return RPE_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("RPE_KABSCH_POLAR"))  // measurement aid: 0 = Jacobi SVD in every Kabsch refit of this device
    rpe::set_kabsch_polar(e[0] != '0');
  ctx->rb.num_sms = ctx->num_sms;
  if (prop.major < 10) {
    delete ctx;
    return RPE_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  }
  if (own) {
    // The context's own stream carries the small kernels of a frame (generation, fix-up, replay, mask, refits). It gets
    // the highest priority, the scorer lanes the lowest: when a scorer CTA leaves an SM, the waiting CTAs of another
    // frame's small kernels go first and the next scorer never waits for its inputs (RPE_STREAM_PRIORITY=0: all equal).
    int prio_lo = 0, prio_hi = 0;
    static const bool use_prio = !(getenv("RPE_STREAM_PRIORITY") && getenv("RPE_STREAM_PRIORITY")[0] == '0');
    if (use_prio && cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) {
      (void)cudaGetLastError();
      prio_lo = prio_hi = 0;
    }
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, use_prio ? prio_hi : 0) != cudaSuccess) {
      delete ctx;
      return RPE_ERR_CUDA;
    }
  } else {
    ctx->stream = (cudaStream_t)stream;
  }
  ctx->own_stream = own;
  {
    const char* e = getenv("RPE_STALE_SAMPLE_COLUMNS");  // the header-only adapters have no knob of their own for it
    ctx->stale_cols = e && e[0] == '1';
  }
  bool ok = true;
  ok = ok && cudaMalloc(&ctx->d_stats, sizeof(FrameStats)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_pose, sizeof(ReplayOut)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_rs, sizeof(ReplayState)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_rs, sizeof(ReplayState)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_kabsch, sizeof(ReplayOut)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_pose, kNumStaging * sizeof(ReplayOut)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->rb.moments, kMomentCount * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->rb.suff, kMomentCount * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_gn, sizeof(GnState)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_rs64, sizeof(ReplayState64)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_rs64, sizeof(ReplayState64)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_pose64, sizeof(Pose64)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_pose64, kNumStaging * sizeof(Pose64)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_nlsk, sizeof(NlskState)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_gn_cost, sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_gn_evals, sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_gn_cost, kNumStaging * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_gn_evals, kNumStaging * sizeof(int32_t)) == cudaSuccess;
  ok = ok && alloc_worklist(ctx, kWorklistInitial) == RPE_OK;
  ok = ok && cudaMalloc(&ctx->wl.counts, kMaxWorklistSegments * sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMemsetAsync(ctx->wl.counts, 0, kMaxWorklistSegments * sizeof(unsigned int), ctx->stream) == cudaSuccess;
  if (ok) {
    ok = ok && cudaMemsetAsync(ctx->d_stats, 0, sizeof(FrameStats), ctx->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(ctx->d_pose, 0, sizeof(ReplayOut), ctx->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(ctx->d_kabsch, 0, sizeof(ReplayOut), ctx->stream) == cudaSuccess;
  }
  for (int k = 0; k <= ST_COUNT && ok; ++k) ok = ok && cudaEventCreate(&ctx->ev[k]) == cudaSuccess;
  for (int k = 0; k < 2 * rpe_ctx::kFastRing && ok; ++k) ok = ok && cudaEventCreate(&ctx->ev_ring[k]) == cudaSuccess;
  for (int k = 0; k < kNumStaging && ok; ++k)
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_slot[k], cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < 2 && ok; ++k)
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_lane[k], cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < 4 && ok; ++k)
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_lane_done[k], cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; k < rpe_ctx::kMaxChunks && ok; ++k)
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_chunk[k], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_prev, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_early, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->early_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_mask_ready, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->ev_mask_copied, cudaEventDisableTiming) == cudaSuccess;
  if (own) ok = ok && cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
  ctx->ev_ok = ok;
  if (!ok) {
    rpe_destroy(ctx);
    return RPE_ERR_CUDA;
  }
  memset(ctx->h_pose, 0, kNumStaging * sizeof(ReplayOut));
  ctx->h_pose->q[3] = 1.f;
  ctx->h_pose->winner = -1;
  ctx->h_pose->max_votes = -1;
  cudaMemcpyAsync(ctx->d_pose, ctx->h_pose, sizeof(ReplayOut), cudaMemcpyHostToDevice, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  *out = ctx;
  return RPE_OK;
}

int rpe_create(int device, rpe_ctx** ctx) { return create_common(device, nullptr, true, ctx); }
int rpe_create_on_stream(int device, void* cuda_stream, rpe_ctx** ctx) {
  return create_common(device, cuda_stream, false, ctx);
}

int rpe_destroy(rpe_ctx* ctx) {
  if (!ctx) return RPE_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int k = 0; k < 5; ++k)
    if (ctx->d_raw[k]) cudaFree(ctx->d_raw[k]);
  cudaFree(ctx->d_pk);
  cudaFree(ctx->d_gen);
  cudaFree(ctx->d_fast);
  cudaFree(ctx->d_votes);
  cudaFree(ctx->d_samples);
  cudaFree(ctx->d_stale_eff);
  cudaFree(ctx->d_stale_carry);
  cudaFree(ctx->d_stats);
  cudaFree(ctx->d_pose);
  cudaFree(ctx->d_rs);
  if (ctx->h_rs) cudaFreeHost(ctx->h_rs);
  cudaFree(ctx->d_kabsch);
  if (ctx->h_pose) cudaFreeHost(ctx->h_pose);
  cudaFree(ctx->wl.entries);
  cudaFree(ctx->wl.counts);
  cudaFree(ctx->d_mask);
  cudaFree(ctx->d_maskbits);
  cudaFree(ctx->rb.partials);
  cudaFree(ctx->rb.moments);
  cudaFree(ctx->rb.suff);
  cudaFree(ctx->d_gn);
  cudaFree(ctx->d_nlsk);
  cudaFree(ctx->d_weights3);
  cudaFree(ctx->d_gn_cost);
  cudaFree(ctx->d_gn_evals);
  for (int r = 0; r < kMaxPeers; ++r)
    if (ctx->peer_opened[r]) cudaIpcCloseMemHandle(ctx->peers.block[r]);
  if (ctx->d_peer_block) cudaFree(ctx->d_peer_block);
  if (ctx->h_peer_err) cudaFreeHost(ctx->h_peer_err);
  if (ctx->h_samples) cudaFreeHost(ctx->h_samples);
  for (int k = 0; k < 5; ++k)
    if (ctx->d_raw64[k]) cudaFree(ctx->d_raw64[k]);
  if (ctx->d_gen64) cudaFree(ctx->d_gen64);
  if (ctx->d_rs64) cudaFree(ctx->d_rs64);
  if (ctx->h_rs64) cudaFreeHost(ctx->h_rs64);
  if (ctx->d_pose64) cudaFree(ctx->d_pose64);
  if (ctx->h_pose64) cudaFreeHost(ctx->h_pose64);
  if (ctx->h_gn_cost) cudaFreeHost(ctx->h_gn_cost);
  if (ctx->h_gn_evals) cudaFreeHost(ctx->h_gn_evals);
  for (int k = 0; k <= ST_COUNT; ++k)
    if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
  for (int k = 0; k < 2 * rpe_ctx::kFastRing; ++k)
    if (ctx->ev_ring[k]) cudaEventDestroy(ctx->ev_ring[k]);
  for (int k = 0; k < kNumStaging; ++k)
    if (ctx->ev_slot[k]) cudaEventDestroy(ctx->ev_slot[k]);
  for (int k = 0; k < 2; ++k)
    if (ctx->ev_lane[k]) cudaEventDestroy(ctx->ev_lane[k]);
  for (int k = 0; k < 4; ++k)
    if (ctx->ev_lane_done[k]) cudaEventDestroy(ctx->ev_lane_done[k]);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->early_stream) cudaStreamSynchronize(ctx->early_stream);
  for (int k = 0; k < rpe_ctx::kMaxChunks; ++k)
    if (ctx->ev_chunk[k]) cudaEventDestroy(ctx->ev_chunk[k]);
  if (ctx->ev_prev) cudaEventDestroy(ctx->ev_prev);
  if (ctx->ev_early) cudaEventDestroy(ctx->ev_early);
  if (ctx->d2h_stream) {
    cudaStreamSynchronize(ctx->d2h_stream);
    cudaStreamDestroy(ctx->d2h_stream);
  }
  if (ctx->ev_mask_ready) cudaEventDestroy(ctx->ev_mask_ready);
  if (ctx->ev_mask_copied) cudaEventDestroy(ctx->ev_mask_copied);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->early_stream) cudaStreamDestroy(ctx->early_stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  (void)cudaGetLastError();
  delete ctx;
  return RPE_OK;
}

const char* rpe_last_error(const rpe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* rpe_stream(rpe_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long long rpe_launch_count(const rpe_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rpe_sync(rpe_ctx* ctx) {
  if (!ctx) return RPE_ERR_ARG;
  ENTER(ctx);
  CK(sync_stream(ctx));
  finish_pending(ctx);
  return check_comm(ctx);
}

int rpe_poll(rpe_ctx* ctx) {
  if (!ctx) return RPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  while (!ctx->pending.empty()) {
    const rpe_ctx::Pending p = ctx->pending.front();
    const cudaError_t e = cudaEventQuery(ctx->ev_slot[p.slot]);
    if (e == cudaErrorNotReady) {
      (void)cudaGetLastError();
      break;
    }
    CK(e);
    deliver(ctx, p);
    ctx->pending.pop_front();
  }
  return RPE_OK;
}

int rpe_host_alloc(size_t bytes, void** ptr) {
  if (!ptr) return RPE_ERR_ARG;
  const cudaError_t e = cudaMallocHost(ptr, bytes);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return RPE_ERR_NOMEM;
  }
  return RPE_OK;
}
int rpe_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
  return RPE_OK;
}

static int upload_common(rpe_ctx* ctx, const float* const src[5], int n, bool from_device) {
  if (!ctx) return RPE_ERR_ARG;
  if (n <= 0 || !src[A_XW]) return fail(ctx, RPE_ERR_ARG, "rpe_upload needs n > 0 and world points");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_corr_capacity(ctx, n, !from_device);
  if (rc) return rc;
  ctx->n = n;
  bool defer = false;
  if (!from_device && ctx->overlap_chunks >= 1 && n >= 65536 && src[A_XC] && !src[A_BV] && !src[A_NC] && !src[A_NW]) {
    // page-locked (and device-visible) host arrays?
    defer = true;
    for (int k = 0; k < 5 && defer; ++k) {
      if (!src[k]) continue;
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, src[k]) != cudaSuccess) {
        (void)cudaGetLastError();
        defer = false;
      } else if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
        defer = false;
      } else {
        ctx->host_view[k] = static_cast<const float*>(attr.devicePointer);
      }
    }
  }
  if (defer) {
    // Nothing is copied yet. If the next call is an rpe_ransac that can run ahead of the copy, it first generates its
    // hypotheses from the host arrays (3 H points over an idle PCIe bus: a few microseconds; behind the frame's DMA
    // traffic the same reads took longer than the whole upload), then starts the chunk copies and scores chunk by
    // chunk. Any other call carries the upload out the plain way (flush_deferred).
    const int C = ctx->overlap_chunks;
    const int cs = (((n + C - 1) / C) + 1023) & ~1023;  // whole scorer stages; 16-byte aligned chunk starts
    ctx->chunk_corr = cs;
    ctx->n_chunks = (n + cs - 1) / cs;
    for (int k = 0; k < 5; ++k) {
      ctx->host_src[k] = src[k];
      ctx->view[k] = src[k] ? ctx->d_raw[k] : nullptr;
    }
  } else {
    for (int k = 0; k < 5; ++k) {
      if (!src[k]) {
        ctx->view[k] = nullptr;
        continue;
      }
      if (from_device) {
        ctx->view[k] = src[k];
      } else {
        CK(cudaMemcpyAsync(ctx->d_raw[k], src[k], (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ctx->view[k] = ctx->d_raw[k];
      }
    }
  }
  ctx->pk_kind = -1;
  ctx->kabsch_valid = false;
  ctx->suff_valid = false;
  ctx->n_slots = 0;
  leave_f64_mode(ctx);
  ctx->deferred = defer;
  return RPE_OK;
}

int rpe_upload(rpe_ctx* ctx, const float* bv, const float* xc, const float* nc, const float* xw, const float* nw, int n) {
  const float* src[5] = {bv, xc, nc, xw, nw};
  if (ctx) {
    stamp(ctx, ST_UPLOAD);
    ctx->upload_stamped = ctx->timing;
  }
  return upload_common(ctx, src, n, false);
}
int rpe_upload_device(rpe_ctx* ctx, const float* bv, const float* xc, const float* nc, const float* xw, const float* nw,
                      int n) {
  const float* src[5] = {bv, xc, nc, xw, nw};
  return upload_common(ctx, src, n, true);
}
int rpe_num_correspondences(const rpe_ctx* ctx) { return ctx ? ctx->n : 0; }

int rpe_ransac(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d, float cos_thrN,
               float confidence, rpe_result* out, int16_t* mask) {
  return do_ransac(ctx, method, samples, nullptr, nullptr, H, thr3d, cos_thr2d, cos_thrN, confidence, out, mask, true);
}
int rpe_ransac_async(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d,
                     float cos_thrN, float confidence, rpe_result* out, int16_t* mask) {
  return do_ransac(ctx, method, samples, nullptr, nullptr, H, thr3d, cos_thr2d, cos_thrN, confidence, out, mask, false);
}
int rpe_set_upload_overlap(rpe_ctx* ctx, int chunks) {
  if (!ctx || chunks < 0) return RPE_ERR_ARG;
  ctx->overlap_chunks = chunks > rpe_ctx::kMaxChunks ? rpe_ctx::kMaxChunks : chunks;  // 0 off, 1 stream from the host arrays, >= 2 chunked copy
  return RPE_OK;
}
int rpe_set_stale_sample_columns(rpe_ctx* ctx, int on) {
  if (!ctx) return RPE_ERR_ARG;
  ctx->stale_cols = on != 0;
  return RPE_OK;
}
int rpe_set_first_pass_iters(rpe_ctx* ctx, int iters) {
  if (!ctx || iters < 1) return RPE_ERR_ARG;
  ctx->first_pass = iters < kMaxPassIters ? iters : kMaxPassIters;
  return RPE_OK;
}
int rpe_ransac_stream(rpe_ctx* ctx, int method, rpe_sample_fn fn, void* user, int H, float thr3d, float cos_thr2d,
                      float cos_thrN, float confidence, rpe_result* out, int16_t* mask) {
  if (!fn) return RPE_ERR_ARG;
  return do_ransac(ctx, method, nullptr, fn, user, H, thr3d, cos_thr2d, cos_thrN, confidence, out, mask, true);
}
int rpe_ransac_f64(rpe_ctx* ctx, int method, const int32_t* samples, rpe_sample_fn fn, void* user, int H, double thr3d,
                   double cos_thr2d, double cos_thrN, double confidence, rpe_result* out, int16_t* mask) {
  if (!ctx) return RPE_ERR_ARG;
  if (!ctx->f64) return fail(ctx, RPE_ERR_STATE, "rpe_ransac_f64 needs arrays uploaded with rpe_upload_f64");
  return do_ransac64(ctx, method, samples, fn, user, H, thr3d, cos_thr2d, cos_thrN, confidence, out, mask);
}
int rpe_upload_f64(rpe_ctx* ctx, const double* bv, const double* xc, const double* nc, const double* xw, const double* nw,
                   int n) {
  if (!ctx || n <= 0 || !xw) return RPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_corr_capacity(ctx, n, true);
  if (rc) return rc;
  if ((size_t)n > ctx->cap_n64) {
    for (int k = 0; k < 5; ++k) {
      if (ctx->d_raw64[k]) cudaFree(ctx->d_raw64[k]);
      ctx->d_raw64[k] = nullptr;
    }
    const size_t cap = (size_t)n + (size_t)n / 8 + 64;
    for (int k = 0; k < 5; ++k) CK(cudaMalloc(&ctx->d_raw64[k], cap * 3 * sizeof(double)));
    ctx->cap_n64 = cap;
  }
  const double* src[5] = {bv, xc, nc, xw, nw};
  ctx->n = n;
  for (int k = 0; k < 5; ++k) {
    if (!src[k]) {
      ctx->view[k] = nullptr;
      ctx->view64[k] = nullptr;
      continue;
    }
    CK(cudaMemcpyAsync(ctx->d_raw64[k], src[k], (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launch_f64_to_f32(ctx->d_raw64[k], ctx->d_raw[k], (size_t)n * 3, ctx->stream);
    ctx->launches++;
    ctx->view64[k] = ctx->d_raw64[k];
    ctx->view[k] = ctx->d_raw[k];
  }
  ctx->pk_kind = -1;
  ctx->kabsch_valid = false;
  ctx->suff_valid = false;
  ctx->n_slots = 0;
  ctx->f64 = true;
  ctx->deferred = false;
  return RPE_OK;
}
int rpe_get_hypotheses_f64(rpe_ctx* ctx, int n_slots, double* hyps7, int32_t* valid) {
  if (!ctx || !hyps7 || n_slots <= 0) return RPE_ERR_ARG;
  if (!ctx->f64 || n_slots > ctx->n_slots) return fail(ctx, RPE_ERR_STATE, "no binary64 hypotheses of that many slots");
  ENTER(ctx);
  std::vector<HypGen64> h((size_t)n_slots);
  CK(cudaMemcpyAsync(h.data(), ctx->d_gen64, (size_t)n_slots * sizeof(HypGen64), cudaMemcpyDeviceToHost, ctx->stream));
  CK(sync_stream(ctx));
  for (int i = 0; i < n_slots; ++i) {
    for (int k = 0; k < 4; ++k) hyps7[7 * (size_t)i + k] = h[i].q[k];
    for (int k = 0; k < 3; ++k) hyps7[7 * (size_t)i + 4 + k] = h[i].t[k];
    if (valid) valid[i] = h[i].valid;
  }
  return RPE_OK;
}

static int do_refit(rpe_ctx* ctx, int kind, const float* weights, int max_iters, rpe_result* out, bool blocking) {
  if (!ctx || !out) return RPE_ERR_ARG;
  if (ctx->n <= 0) return fail(ctx, RPE_ERR_STATE, "no correspondences uploaded");
  ENTER(ctx);
  FrameView f = make_view(ctx);
  // every state / argument check comes before a staging slot is claimed: a failed call must not advance the ring
  if (kind == RPE_REFIT_KABSCH_INLIERS || kind == RPE_REFIT_KABSCH_ALL) {
    if (!f.xc) return fail(ctx, RPE_ERR_STATE, "Kabsch refit needs camera points");
    if (kind == RPE_REFIT_KABSCH_INLIERS && !ctx->kabsch_valid && ctx->mask_cols < 2)
      return fail(ctx, RPE_ERR_STATE, "no 3-D inlier column available");
  } else if (kind == RPE_REFIT_GN) {
    if (ctx->mask_cols <= 0) return fail(ctx, RPE_ERR_STATE, "no inlier mask: run rpe_ransac or rpe_set_mask first");
  } else if (kind == RPE_REFIT_NL_SK_LS) {
    if (!f.bv || !f.xc || !f.nc || !f.nw) return fail(ctx, RPE_ERR_STATE, "nl_shinji_kneip_ls needs all five arrays");
    if (ctx->mask_cols < 3) return fail(ctx, RPE_ERR_STATE, "nl_shinji_kneip_ls needs the three inlier columns");
  } else {
    return fail(ctx, RPE_ERR_ARG, "unknown refit kind");
  }
  int slot = 0;
  int rcs = claim_slot(ctx, &slot);
  if (rcs) return rcs;
  bool gn = false;
  if (kind == RPE_REFIT_KABSCH_INLIERS || kind == RPE_REFIT_KABSCH_ALL) {
    if (kind == RPE_REFIT_KABSCH_INLIERS && ctx->kabsch_valid) {
      // already computed by the mask kernel's last CTA; adopt it as the current pose
      CK(cudaMemcpyAsync(ctx->d_pose, ctx->d_kabsch, sizeof(ReplayOut), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      const int16_t* flags = kind == RPE_REFIT_KABSCH_INLIERS ? ctx->d_mask + ctx->n : nullptr;
      const int used = launch_kabsch_moments(f, flags, ctx->rb, ctx->d_stats, ctx->stream);
      launch_kabsch_solve(ctx->rb, used, ctx->d_pose, nullptr, ctx->stream);
      ctx->launches += 2;
    }
  } else if (kind == RPE_REFIT_GN) {
    const float w2 = weights ? weights[0] : 1.f, w3 = weights ? weights[1] : 1.f, wn = weights ? weights[2] : 1.f;
    const int iters = max_iters > 0 ? max_iters : 6;
    stamp(ctx, ST_GN);
    const bool generic = ctx->mask_cols >= 1 && f.bv != nullptr && w2 > 0.f;
    if (generic) {  // 2-D rows are not polynomial in the pose: one pass over the correspondences per evaluation
      launch_gn_init(ctx->d_pose, ctx->d_gn, ctx->stream);
      for (int it = 0; it < iters; ++it)
        launch_gn_iteration(f, ctx->d_mask, ctx->mask_cols, w2, w3, wn, ctx->rb, ctx->d_gn, ctx->d_stats, ctx->d_pose,
                            ctx->d_gn_cost, ctx->d_gn_evals, ctx->stream);
      ctx->launches += iters + 1;
    } else {
      const bool m3 = ctx->mask_cols >= 2 && f.xc && w3 > 0.f;
      const bool mn = ctx->mask_cols >= 3 && f.nc && f.nw && wn > 0.f;
      int used = 0;
      if (!ctx->suff_valid) {
        used = launch_suffstats(f, m3 ? ctx->d_mask + ctx->n : nullptr, false, mn ? ctx->d_mask + 2 * (size_t)ctx->n : nullptr,
                                ctx->rb, ctx->stream);
        ctx->launches += 1;
        ctx->suff_valid = m3 == (ctx->mask_cols >= 2 && f.xc != nullptr) && mn == (ctx->mask_cols >= 3 && f.nc && f.nw);
      }
      launch_gn_from_stats(ctx->rb, used, m3 ? w3 : 0.f, mn ? wn : 0.f, iters, ctx->d_pose, ctx->d_gn, ctx->d_gn_cost,
                           ctx->d_gn_evals, ctx->stream);
      ctx->launches += 1;
    }
    CK(cudaMemcpyAsync(&ctx->h_gn_cost[slot], ctx->d_gn_cost, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&ctx->h_gn_evals[slot], ctx->d_gn_evals, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    stamp(ctx, ST_TOTAL);
    gn = true;
  } else {  // RPE_REFIT_NL_SK_LS
    const float* w3 = nullptr;
    if (weights) {
      const size_t need = (size_t)ctx->n * 3;
      if (need > ctx->weights_cap) {
        if (ctx->d_weights3) cudaFree(ctx->d_weights3);
        CK(cudaMalloc(&ctx->d_weights3, need * sizeof(float)));
        ctx->weights_cap = need;
      }
      CK(cudaMemcpyAsync(ctx->d_weights3, weights, need * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      w3 = ctx->d_weights3;
    }
    launch_nlsk_prepass(f, ctx->d_mask, w3, ctx->d_pose, ctx->rb, ctx->d_nlsk, ctx->d_stats, ctx->stream);
    for (int it = 0; it < 3; ++it)
      launch_nlsk_iteration(f, ctx->d_mask, w3, ctx->rb, ctx->d_nlsk, ctx->d_stats, ctx->d_pose, ctx->stream);
    ctx->launches += 4;
  }
  ctx->kabsch_valid = false;
  CK(cudaMemcpyAsync(&ctx->h_pose[slot], ctx->d_pose, sizeof(ReplayOut), cudaMemcpyDeviceToHost, ctx->stream));
  if (int rcp = push_pending(ctx, out, slot, true, gn)) return rcp;
  if (blocking) {
    CK(sync_stream(ctx));
    finish_pending(ctx);
  }
  return RPE_OK;
}

int rpe_refit(rpe_ctx* ctx, int kind, const float* weights, int max_iters, rpe_result* out) {
  return do_refit(ctx, kind, weights, max_iters, out, true);
}
int rpe_refit_async(rpe_ctx* ctx, int kind, const float* weights, int max_iters, rpe_result* out) {
  return do_refit(ctx, kind, weights, max_iters, out, false);
}

int rpe_set_pose(rpe_ctx* ctx, const float q_xyzw[4], const float t[3], int max_votes) {
  if (!ctx || !q_xyzw || !t) return RPE_ERR_ARG;
  ENTER(ctx);
  CK(sync_stream(ctx));
  finish_pending(ctx);
  ReplayOut p;
  memset(&p, 0, sizeof(p));
  for (int k = 0; k < 4; ++k) p.q[k] = q_xyzw[k];
  for (int k = 0; k < 3; ++k) p.t[k] = t[k];
  p.max_votes = max_votes;
  p.winner = 0;
  *ctx->h_pose = p;
  CK(cudaMemcpyAsync(ctx->d_pose, ctx->h_pose, sizeof(ReplayOut), cudaMemcpyHostToDevice, ctx->stream));
  CK(sync_stream(ctx));
  ctx->kabsch_valid = false;
  return RPE_OK;
}

int rpe_set_mask(rpe_ctx* ctx, const int16_t* mask, int cols) {
  if (!ctx || !mask || cols < 1 || cols > 3) return RPE_ERR_ARG;
  if (ctx->n <= 0) return fail(ctx, RPE_ERR_STATE, "no correspondences uploaded");
  ENTER(ctx);
  order_after_mask_copy(ctx);
  CK(cudaMemcpyAsync(ctx->d_mask, mask, (size_t)ctx->n * cols * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(sync_stream(ctx));
  ctx->mask_cols = cols;
  ctx->kabsch_valid = false;
  ctx->suff_valid = false;
  return RPE_OK;
}

// ---- stage access ----------------------------------------------------------------------------------
int rpe_generate(rpe_ctx* ctx, int method, const int32_t* samples, int H) {
  if (ctx && ctx->f64) return fail(ctx, RPE_ERR_STATE, "the stage API is binary32 only (arrays were uploaded with rpe_upload_f64)");
  if (!ctx) return RPE_ERR_ARG;
  if (!method_ok(method) || !samples || H <= 0) return fail(ctx, RPE_ERR_ARG, "bad argument to rpe_generate");
  int rc = check_arrays(ctx, method);
  if (rc) return rc;
  ENTER(ctx);
  const int S = method_slots(method);
  rc = ensure_hyp_capacity(ctx, H, H * S);
  if (rc) return rc;
  // host (pageable or page-locked) or device memory: the unified address space tells them apart
  CK(cudaMemcpyAsync(ctx->d_samples, samples, (size_t)H * 4 * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
  launch_reset_stats(ctx->d_stats, ctx->stream);
  FrameView f = make_view(ctx);
  const int32_t* stale_eff = nullptr;
  if (int rcs = prepare_stale(ctx, method, ctx->d_samples, H, true, &stale_eff)) return rcs;
  launch_hypgen(method, f, ctx->d_samples, H, ctx->d_gen, ctx->d_fast, ctx->d_votes, ctx->d_stats, ctx->stream, stale_eff);
  ctx->launches += 2;
  ctx->n_slots = H * S;
  ctx->cur_method = method;
  ctx->stats_clean = false;
  return RPE_OK;  // asynchronous; rpe_get_hypotheses / rpe_get_votes / rpe_finish synchronise
}

int rpe_get_hypotheses(rpe_ctx* ctx, float* hyps, int32_t* valid, int n_slots) {
  if (ctx && ctx->f64) return fail(ctx, RPE_ERR_STATE, "the stage API is binary32 only (arrays were uploaded with rpe_upload_f64)");
  if (!ctx || n_slots <= 0 || n_slots > ctx->n_slots) return ctx ? fail(ctx, RPE_ERR_ARG, "bad slot count") : RPE_ERR_ARG;
  ENTER(ctx);
  std::vector<HypGen> h(n_slots);
  CK(cudaMemcpyAsync(h.data(), ctx->d_gen, (size_t)n_slots * sizeof(HypGen), cudaMemcpyDeviceToHost, ctx->stream));
  CK(sync_stream(ctx));
  for (int i = 0; i < n_slots; ++i) {
    if (hyps) {
      for (int k = 0; k < 4; ++k) hyps[7 * i + k] = h[i].q[k];
      for (int k = 0; k < 3; ++k) hyps[7 * i + 4 + k] = h[i].t[k];
    }
    if (valid) valid[i] = h[i].valid;
  }
  return RPE_OK;
}

int rpe_set_hypotheses(rpe_ctx* ctx, int method, const float* hyps, const int32_t* valid, int n_slots) {
  if (!ctx) return RPE_ERR_ARG;
  if (!method_ok(method) || !hyps || n_slots <= 0) return fail(ctx, RPE_ERR_ARG, "bad argument to rpe_set_hypotheses");
  ENTER(ctx);
  int rc = ensure_hyp_capacity(ctx, 1, n_slots);
  if (rc) return rc;
  std::vector<HypGen> h(n_slots);
  for (int i = 0; i < n_slots; ++i) {
    for (int k = 0; k < 4; ++k) h[i].q[k] = hyps[7 * i + k];
    for (int k = 0; k < 3; ++k) h[i].t[k] = hyps[7 * i + 4 + k];
    h[i].valid = valid ? valid[i] : 1;
  }
  CK(cudaMemcpyAsync(ctx->d_gen, h.data(), (size_t)n_slots * sizeof(HypGen), cudaMemcpyHostToDevice, ctx->stream));
  launch_reset_stats(ctx->d_stats, ctx->stream);
  launch_derive_fast(ctx->d_gen, ctx->d_fast, ctx->d_votes, n_slots, ctx->d_stats, ctx->stream);
  ctx->launches += 2;
  CK(sync_stream(ctx));
  ctx->n_slots = n_slots;
  ctx->cur_method = method;
  ctx->stats_clean = false;
  return RPE_OK;
}

int rpe_score(rpe_ctx* ctx, int method, int slot_begin, int slot_end, float thr3d, float cos_thr2d, float cos_thrN) {
  if (ctx && ctx->f64) return fail(ctx, RPE_ERR_STATE, "the stage API is binary32 only (arrays were uploaded with rpe_upload_f64)");
  if (!ctx) return RPE_ERR_ARG;
  if (!method_ok(method) || slot_begin < 0 || slot_end > ctx->n_slots || slot_begin > slot_end)
    return fail(ctx, RPE_ERR_ARG, "bad slot range");
  int rc = check_arrays(ctx, method);
  if (rc) return rc;
  ENTER(ctx);
  if (method == RPE_SHINJI || !g_force_exact_multi) {
    rc = ensure_packed(ctx, kind_for_method(method));
    if (rc) return rc;
  }
  const Thresh th = {thr3d, cos_thr2d, cos_thrN};
  stamp(ctx, ST_SCORE);
  rc = score_range(ctx, method, slot_begin, slot_end, th);
  launch_consume_worklist(ctx->d_stats, ctx->stream);
  ctx->launches++;
  stamp(ctx, ST_REPLAY);
  ctx->stats_clean = false;
  return rc;
}

int rpe_get_votes(rpe_ctx* ctx, int32_t* votes, int n_slots) {
  if (!ctx || !votes || n_slots <= 0 || n_slots > ctx->n_slots) return RPE_ERR_ARG;
  ENTER(ctx);
  CK(cudaMemcpyAsync(votes, ctx->d_votes, (size_t)n_slots * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(sync_stream(ctx));
  return RPE_OK;
}
int rpe_set_votes(rpe_ctx* ctx, const int32_t* votes, int n_slots) {
  if (!ctx || !votes || n_slots <= 0 || n_slots > ctx->n_slots) return RPE_ERR_ARG;
  ENTER(ctx);
  CK(cudaMemcpyAsync(ctx->d_votes, votes, (size_t)n_slots * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CK(sync_stream(ctx));
  return RPE_OK;
}
int32_t* rpe_votes_device_ptr(rpe_ctx* ctx) { return ctx ? ctx->d_votes : nullptr; }

// ---- peer-memory exchange of the vote table (one process per GPU, CUDA IPC over NVLink) ----------------------
// (Re-)arming the own block: flags, error latch and epoch start from zero. Export is the right place: a peer can only
// write into this block for the new group after it has imported the handle, i.e. after this call returned (the handles
// are gathered in between). All ranks must have synchronised their contexts before a group is set up again.
static int peer_arm(rpe_ctx* ctx) {
  if (!ctx->d_peer_block) {
    CK(cudaMalloc(&ctx->d_peer_block, kPeerBlockBytes));
    CK(cudaMemset(ctx->d_peer_block, 0, kPeerBlockBytes));
  } else {
    CK(sync_stream(ctx));
    CK(cudaMemset(ctx->d_peer_block, 0, kPeerFlagWords * sizeof(unsigned int)));
  }
  if (!ctx->h_peer_err) CK(cudaMallocHost(&ctx->h_peer_err, sizeof(unsigned int)));
  *ctx->h_peer_err = 0;
  ctx->peer_epoch = 0;
  return RPE_OK;
}
static void peer_close(rpe_ctx* ctx) {
  for (int r = 0; r < kMaxPeers; ++r) {
    if (ctx->peer_opened[r]) cudaIpcCloseMemHandle(ctx->peers.block[r]);
    ctx->peer_opened[r] = false;
    ctx->peers.block[r] = nullptr;
  }
  (void)cudaGetLastError();
  ctx->peer_world = 0;
  ctx->peer_rank = -1;
}
int rpe_peer_export(rpe_ctx* ctx, unsigned char handle[64]) {
  if (!ctx || !handle) return RPE_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  ENTER(ctx);
  peer_close(ctx);  // mappings of an earlier group
  if (int rc = peer_arm(ctx)) return rc;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->d_peer_block));
  memcpy(handle, &h, 64);
  return RPE_OK;
}
int rpe_peer_import(rpe_ctx* ctx, int rank, int world, const unsigned char* handles) {
  if (!ctx || !handles || world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return RPE_ERR_ARG;
  if (!ctx->d_peer_block) return fail(ctx, RPE_ERR_STATE, "call rpe_peer_export first");
  ENTER(ctx);
  peer_close(ctx);
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ctx->peers.block[r] = ctx->d_peer_block;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)r, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peers.block[r] = (unsigned char*)p;
    ctx->peer_opened[r] = true;
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  return RPE_OK;  // the epoch was reset by rpe_peer_export, together with the flags it is compared with
}
int rpe_peer_set_timeout_ms(rpe_ctx* ctx, int ms) {
  if (!ctx || ms < 1) return RPE_ERR_ARG;
  ctx->peer_timeout_ns = (unsigned long long)ms * 1000000ull;
  return RPE_OK;
}
// Same for contexts that live in ONE process (one per GPU): no IPC, the peers' blocks are addressed directly after
// enabling peer access. With blocking calls a single host thread would wait for itself: use the _async entry points
// (or one host thread per context).
int rpe_peer_import_local(rpe_ctx* ctx, int rank, int world, rpe_ctx* const* ctxs) {
  if (!ctx || !ctxs || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || ctxs[rank] != ctx) return RPE_ERR_ARG;
  ENTER(ctx);
  peer_close(ctx);
  // own block: (re-)armed here; the caller imports on every context before the first frame, so nobody writes into it
  // yet. Peers' blocks that do not exist yet are created (armed) now and left alone when their own import runs.
  if (int rc = peer_arm(ctx)) return rc;
  for (int r = 0; r < world; ++r) {
    rpe_ctx* p = ctxs[r];
    if (!p) return RPE_ERR_ARG;
    if (!p->d_peer_block) {
      CK(cudaSetDevice(p->device));
      CK(cudaMalloc(&p->d_peer_block, kPeerBlockBytes));
      CK(cudaMemset(p->d_peer_block, 0, kPeerBlockBytes));
      ENTER(ctx);
    }
    if (p->device != ctx->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, ctx->device, p->device));
      if (!can) return fail(ctx, RPE_ERR_STATE, "no peer access between the two devices");
      const cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, RPE_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
      (void)cudaGetLastError();
    }
    ctx->peers.block[r] = p->d_peer_block;
  }
  ctx->peer_rank = rank;
  ctx->peer_world = world;
  return RPE_OK;
}
int rpe_exchange_votes(rpe_ctx* ctx, int slot_begin, int slot_end) {
  if (!ctx) return RPE_ERR_ARG;
  if (ctx->peer_world < 1) return fail(ctx, RPE_ERR_STATE, "call rpe_peer_import first");
  if (slot_begin < 0 || slot_end < slot_begin || slot_end > ctx->n_slots || ctx->n_slots > kPeerSlots)
    return fail(ctx, RPE_ERR_ARG, "bad slot range for rpe_exchange_votes");
  ENTER(ctx);
  ctx->peer_epoch += 1;
  launch_exchange_votes(ctx->peers, ctx->peer_rank, ctx->peer_world, ctx->peer_epoch, slot_begin, slot_end, ctx->n_slots,
                        ctx->d_votes, ctx->h_peer_err, ctx->peer_timeout_ns, ctx->stream);
  ctx->launches++;
  return RPE_OK;
}
// One frame of the hypothesis-sharded mode as a single asynchronous stream of work: generate all H (replicated,
// cheap), score this rank's slots, exchange through peer memory, replay + mask on every rank (identical result).
static int do_ransac_sharded(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d,
                             float cos_thrN, float confidence, rpe_result* out, int16_t* mask, bool blocking) {
  if (!ctx) return RPE_ERR_ARG;
  if (!method_ok(method) || !samples || H <= 0 || !out) return fail(ctx, RPE_ERR_ARG, "bad argument to rpe_ransac_sharded");
  if (ctx->f64) return fail(ctx, RPE_ERR_STATE, "the sharded mode is binary32 only");
  if (ctx->peer_world < 1) return fail(ctx, RPE_ERR_STATE, "call rpe_peer_import first");
  const int S = method_slots(method);
  if (H > kMaxPassIters || H * S > kPeerSlots) return fail(ctx, RPE_ERR_ARG, "rpe_ransac_sharded: at most 8192 iterations");
  int rc = check_arrays(ctx, method);
  if (rc) return rc;
  ENTER(ctx);
  const Thresh th = {thr3d, cos_thr2d, cos_thrN};
  rc = ensure_hyp_capacity(ctx, H, H * S);
  if (rc) return rc;
  if (!ctx->stats_clean) {
    launch_reset_stats(ctx->d_stats, ctx->stream);
    ctx->launches++;
  }
  rc = ensure_packed(ctx, kind_for_method(method));
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_samples, samples, (size_t)H * 4 * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
  FrameView f = make_view(ctx);
  const int32_t* stale_eff = nullptr;
  if (int rcs = prepare_stale(ctx, method, ctx->d_samples, H, true, &stale_eff)) return rcs;
  launch_hypgen(method, f, ctx->d_samples, H, ctx->d_gen, ctx->d_fast, ctx->d_votes, ctx->d_stats, ctx->stream, stale_eff);
  ctx->launches++;
  ctx->n_slots = H * S;
  ctx->cur_method = method;
  // balanced contiguous slot ranges, the same split on every rank
  const int n_slots = H * S, G = ctx->peer_world, r = ctx->peer_rank;
  const int base = n_slots / G, rem = n_slots % G;
  const int sb = r * base + (r < rem ? r : rem), se = sb + base + (r < rem ? 1 : 0);
  rc = score_range(ctx, method, sb, se, th);
  if (rc) return rc;
  ctx->peer_epoch += 1;
  launch_exchange_votes(ctx->peers, r, G, ctx->peer_epoch, sb, se, n_slots, ctx->d_votes, ctx->h_peer_err,
                        ctx->peer_timeout_ns, ctx->stream);
  launch_replay(method, ctx->d_gen, ctx->d_votes, H, 0, ctx->n, confidence, ctx->d_stats, ctx->d_rs, ctx->d_pose, true, H,
                ctx->stream);
  ctx->launches += 2;
  ctx->stats_clean = true;
  return do_finish(ctx, method, th, out, mask, blocking);
}
int rpe_ransac_sharded(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d, float cos_thrN,
                       float confidence, rpe_result* out, int16_t* mask) {
  return do_ransac_sharded(ctx, method, samples, H, thr3d, cos_thr2d, cos_thrN, confidence, out, mask, true);
}
int rpe_ransac_sharded_async(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d,
                             float cos_thrN, float confidence, rpe_result* out, int16_t* mask) {
  return do_ransac_sharded(ctx, method, samples, H, thr3d, cos_thr2d, cos_thrN, confidence, out, mask, false);
}
int rpe_peer_status(rpe_ctx* ctx) {
  if (!ctx || !ctx->d_peer_block) return RPE_ERR_ARG;
  ENTER(ctx);
  CK(sync_stream(ctx));
  return check_comm(ctx);
}

int rpe_finish(rpe_ctx* ctx, int method, int H, float thr3d, float cos_thr2d, float cos_thrN, float confidence,
               rpe_result* out, int16_t* mask) {
  if (!ctx) return RPE_ERR_ARG;
  if (!method_ok(method) || H <= 0 || H * method_slots(method) > ctx->n_slots || !out)
    return fail(ctx, RPE_ERR_ARG, "bad argument to rpe_finish");
  ENTER(ctx);
  const Thresh th = {thr3d, cos_thr2d, cos_thrN};
  launch_replay(method, ctx->d_gen, ctx->d_votes, H, 0, ctx->n, confidence, ctx->d_stats, ctx->d_rs, ctx->d_pose, true, H,
                ctx->stream);
  ctx->launches += 1;
  ctx->stats_clean = true;
  return do_finish(ctx, method, th, out, mask, true);
}

// ---- device-side Simulator (frames are generated straight into the context's own device arrays) ------------
static unsigned long long host_mix64(unsigned long long x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
static unsigned int gcd_u(unsigned int a, unsigned int b) {
  while (b) {
    const unsigned int t = a % b;
    a = b;
    b = t;
  }
  return a;
}
static unsigned int modinv_u(unsigned int a, unsigned int n) {  // a^-1 mod n, gcd(a,n) = 1
  long long t = 0, nt = 1, r = n, nr = a % n;
  while (nr != 0) {
    const long long q = r / nr;
    long long tmp = t - q * nt;
    t = nt;
    nt = tmp;
    tmp = r - q * nr;
    r = nr;
    nr = tmp;
  }
  if (t < 0) t += n;
  return (unsigned int)t;
}

static SimParams make_sim_params(uint64_t seed, const float q[4], const float t[3], int n, float n2d, float or2d, float n3d,
                                 float or3d, float nnl, float ornl, float min_depth, float max_depth, float f,
                                 int use_gaussian, int kinect) {
  SimParams p;
  quat_to_R_rowmajor(q, p.R);
  for (int k = 0; k < 3; ++k) p.t[k] = t[k];
  p.n = n;
  p.noise2d = n2d;
  p.noise3d = n3d;
  p.noise_nl = nnl;
  p.out2d = (int)(or2d * n + .5);
  p.out3d = (int)(or3d * n + .5);
  p.outnl = (int)(ornl * n + .5f);
  for (int k = 0; k < 3; ++k) {
    unsigned int a = (unsigned int)(host_mix64(seed * 3 + k) % (unsigned long long)n);
    if (a < 2) a = 2;
    while (gcd_u(a, (unsigned int)n) != 1) ++a;
    a %= (unsigned int)n;
    if (a == 0) a = 1;
    p.ainv[k] = n > 1 ? modinv_u(a, (unsigned int)n) : 0;
    p.b[k] = (unsigned int)(host_mix64(seed * 7 + k + 11) % (unsigned long long)n);
  }
  p.min_depth = min_depth;
  p.max_depth = max_depth;
  p.f = f;
  p.gaussian = use_gaussian;
  p.kinect = kinect;
  p.seed = seed;
  return p;
}

static int sim_device_common(rpe_ctx* ctx, uint64_t seed, const float q[4], const float t[3], int n, float n2d, float or2d,
                             float n3d, float or3d, float nnl, float ornl, float min_depth, float max_depth, float f,
                             int use_gaussian, int mode_3d3d, int kinect = 0) {
  if (!ctx || !q || !t || n <= 0) return RPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_corr_capacity(ctx, n, true);
  if (rc) return rc;
  const SimParams p = make_sim_params(seed, q, t, n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth, f, use_gaussian, kinect);
  ctx->n = n;
  for (int k = 0; k < 5; ++k) ctx->view[k] = nullptr;
  leave_f64_mode(ctx);  // the frame replaces whatever rpe_upload_f64 left behind
  ctx->view[A_XW] = ctx->d_raw[A_XW];
  ctx->view[A_XC] = ctx->d_raw[A_XC];
  if (!mode_3d3d) {
    ctx->view[A_BV] = ctx->d_raw[A_BV];
    ctx->view[A_NW] = ctx->d_raw[A_NW];
    ctx->view[A_NC] = ctx->d_raw[A_NC];
  }
  launch_simulate(p, ctx->d_raw[A_XW], ctx->d_raw[A_XC], mode_3d3d ? nullptr : ctx->d_raw[A_BV],
                  mode_3d3d ? nullptr : ctx->d_raw[A_NW], mode_3d3d ? nullptr : ctx->d_raw[A_NC], mode_3d3d, ctx->stream);
  ctx->launches++;
  ctx->pk_kind = -1;
  ctx->kabsch_valid = false;
  ctx->suff_valid = false;
  ctx->n_slots = 0;
  return RPE_OK;
}

// The same generator into device buffers of the caller (a resident sequence of distinct frames: config #5). The context
// only lends its stream; its own frame is untouched.
int rpe_sim_3d_3d_device_to(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise,
                            float outlier_ratio, float min_depth, float max_depth, float f, int use_gaussian, float* d_xw,
                            float* d_xc) {
  if (!ctx || !q_xyzw || !t || n <= 0 || !d_xw || !d_xc) return RPE_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const SimParams p = make_sim_params(seed, q_xyzw, t, n, 0.f, 0.f, noise, outlier_ratio, 0.f, 0.f, min_depth, max_depth, f,
                                      use_gaussian, 0);
  launch_simulate(p, d_xw, d_xc, nullptr, nullptr, nullptr, 1, ctx->stream);
  ctx->launches++;
  return RPE_OK;
}

int rpe_sim_3d_3d_device(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise,
                         float outlier_ratio, float min_depth, float max_depth, float f, int use_gaussian) {
  return sim_device_common(ctx, seed, q_xyzw, t, n, 0.f, 0.f, noise, outlier_ratio, 0.f, 0.f, min_depth, max_depth, f,
                           use_gaussian, 1);
}
int rpe_sim_2d_3d_nl_device(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d,
                            float or2d, float n3d, float or3d, float nnl, float ornl, float min_depth, float max_depth,
                            float f, int use_gaussian) {
  return sim_device_common(ctx, seed, q_xyzw, t, n, n2d, or2d, n3d, or3d, nnl, ornl, min_depth, max_depth, f, use_gaussian, 0);
}
int rpe_sim_kinect_2d_3d_nl_device(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d,
                                   float or2d, float or3d, float nnl, float ornl, float min_depth, float max_depth, float f) {
  return sim_device_common(ctx, seed, q_xyzw, t, n, n2d, or2d, 0.f, or3d, nnl, ornl, min_depth, max_depth, f, 1, 0, 1);
}
// Copy the context's current correspondence arrays back to the host (NULL = skip). For tests / inspection.
int rpe_download(rpe_ctx* ctx, float* bv, float* xc, float* nc, float* xw, float* nw) {
  if (!ctx || ctx->n <= 0) return RPE_ERR_ARG;
  ENTER(ctx);
  float* dst[5] = {bv, xc, nc, xw, nw};
  for (int k = 0; k < 5; ++k)
    if (dst[k]) {
      if (!ctx->view[k]) return fail(ctx, RPE_ERR_STATE, "array not present on the device");
      CK(cudaMemcpyAsync(dst[k], ctx->view[k], (size_t)ctx->n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
  CK(sync_stream(ctx));
  return RPE_OK;
}

// ---- MinimalSolvers.hpp on the device: batches, one problem per thread ------------------------------------------
static int minsolv_common(rpe_ctx* ctx, const float* in, int in_stride, int count, float* out_a, int a_stride, float* out_b,
                          int b_stride, bool is_ms) {
  if (!ctx || !in || count <= 0 || !out_a || (is_ms && !out_b)) return RPE_ERR_ARG;
  ENTER(ctx);
  float *d_in = nullptr, *d_a = nullptr, *d_b = nullptr;
  CK(cudaMallocAsync(&d_in, (size_t)count * in_stride * sizeof(float), ctx->stream));
  CK(cudaMallocAsync(&d_a, (size_t)count * a_stride * sizeof(float), ctx->stream));
  if (is_ms) CK(cudaMallocAsync(&d_b, (size_t)count * b_stride * sizeof(float), ctx->stream));
  CK(cudaMemcpyAsync(d_in, in, (size_t)count * in_stride * sizeof(float), cudaMemcpyDefault, ctx->stream));
  if (is_ms)
    launch_minsolv_ms(d_in, count, d_a, d_b, ctx->stream);
  else
    launch_minsolv_ev(d_in, count, d_a, ctx->stream);
  ctx->launches++;
  CK(cudaMemcpyAsync(out_a, d_a, (size_t)count * a_stride * sizeof(float), cudaMemcpyDefault, ctx->stream));
  if (is_ms) CK(cudaMemcpyAsync(out_b, d_b, (size_t)count * b_stride * sizeof(float), cudaMemcpyDefault, ctx->stream));
  cudaFreeAsync(d_in, ctx->stream);
  cudaFreeAsync(d_a, ctx->stream);
  if (d_b) cudaFreeAsync(d_b, ctx->stream);
  CK(sync_stream(ctx));
  return RPE_OK;
}
int rpe_min_ev(rpe_ctx* ctx, const float* M9, int count, float* E3) {
  return minsolv_common(ctx, M9, 9, count, E3, 3, nullptr, 0, false);
}
int rpe_min_ms(rpe_ctx* ctx, const float* in24, int count, float* q4, float* t3) {
  return minsolv_common(ctx, in24, 24, count, q4, 4, t3, 3, true);
}

// ---- Library.cpp shim -----------------------------------------------------------------------------
// One context, created on first use and kept for the life of the process (a context owns a 32 MiB worklist, a ring of
// pinned staging slots and ~270 events: too much to build and tear down per call). Calls are serialised.
static std::mutex g_ao_mu;
static rpe_ctx* g_ao_ctx = nullptr;
static int ao_common(const float* x_w, const float* x_c, int n, float* R_cw, float* t, bool ransac) {
  if (!x_w || !x_c || n < 3 || !R_cw || !t) return RPE_ERR_ARG;
  std::lock_guard<std::mutex> g(g_ao_mu);
  if (!g_ao_ctx) {
    const int rcc = rpe_create(0, &g_ao_ctx);
    if (rcc) {
      g_ao_ctx = nullptr;
      return rcc;
    }
  }
  rpe_ctx* ctx = g_ao_ctx;
  rpe_result res;
  int rc = rpe_upload(ctx, nullptr, x_c, nullptr, x_w, nullptr, n);
  if (!rc && ransac) {
    // Library.cpp:54-64: thr 0.1, 1000 iterations, confidence 0.99999, unseeded rand() (seed 1)
    const int H = 1000;
    std::vector<int32_t> samples((size_t)H * 4);
    rc = rpe_sample_table(1u, n, 3, H, samples.data());
    if (!rc) rc = rpe_ransac(ctx, RPE_SHINJI, samples.data(), H, 0.1f, 0.f, 0.f, 0.99999f, &res, nullptr);
    if (!rc) rc = rpe_refit(ctx, RPE_REFIT_KABSCH_INLIERS, nullptr, 0, &res);
  } else if (!rc) {
    rc = rpe_refit(ctx, RPE_REFIT_KABSCH_ALL, nullptr, 0, &res);  // shinji_ls2 (Library.cpp:35)
  }
  if (!rc) {
    for (int i = 0; i < 9; ++i) R_cw[i] = res.R[i];
    for (int i = 0; i < 3; ++i) t[i] = res.t[i];
  }
  return rc;
}
int rpe_ao(const float* x_w, const float* x_c, int n, float* R_cw, float* t) { return ao_common(x_w, x_c, n, R_cw, t, false); }
int rpe_ao_ransac(const float* x_w, const float* x_c, int n, float* R_cw, float* t) {
  return ao_common(x_w, x_c, n, R_cw, t, true);
}

// ---- microbenchmark / timing -------------------------------------------------------------------------
int rpe_measure_ffma_tflops(rpe_ctx* ctx, int ms_target, double* tflops_scalar, double* tflops_packed) {
  if (!ctx) return RPE_ERR_ARG;
  ENTER(ctx);
  float* sink = nullptr;
  CK(cudaMalloc(&sink, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int blocks = ctx->num_sms * 8;
  double* outs[2] = {tflops_scalar, tflops_packed};
  for (int mode = 0; mode < 2; ++mode) {
    int iters = 2048;
    float ms = 0.f;
    for (int rep = 0; rep < 6; ++rep) {
      launch_ffma_bench(sink, iters, mode == 1, blocks, ctx->stream);  // warm / calibrate
      CK(cudaEventRecord(e0, ctx->stream));
      launch_ffma_bench(sink, iters, mode == 1, blocks, ctx->stream);
      CK(cudaEventRecord(e1, ctx->stream));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ctx->launches += 2;
      if (ms >= (float)ms_target || iters > (1 << 24)) break;
      const double scale = ms > 0.01f ? (double)ms_target / ms * 1.2 : 8.0;
      iters = (int)(iters * (scale > 8.0 ? 8.0 : (scale < 1.5 ? 1.5 : scale)));
    }
    // per thread per iteration: 4 rounds x 8 accumulators x 2 lanes FMAs
    const double fma = (double)blocks * 256.0 * (double)iters * 4.0 * 8.0 * 2.0;
    if (outs[mode]) *outs[mode] = 2.0 * fma / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return RPE_OK;
}

int rpe_set_mask_transfer(rpe_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return RPE_ERR_ARG;
  ctx->mask_transfer = mode;
  return RPE_OK;
}
int rpe_enable_stage_timing(rpe_ctx* ctx, int enable) {
  if (!ctx) return RPE_ERR_ARG;
  ctx->timing = enable == 1;
  ctx->timing_fast = enable != 0;
  return RPE_OK;
}
int rpe_scorer_time_stats(rpe_ctx* ctx, double* sum_ms, long long* count, int reset) {
  if (!ctx) return RPE_ERR_ARG;
  if (sum_ms) *sum_ms = ctx->fast_sum_ms;
  if (count) *count = ctx->fast_count;
  if (reset) {
    ctx->fast_sum_ms = 0.0;
    ctx->fast_count = 0;
  }
  return RPE_OK;
}
int rpe_scorer_busy_stats(int device, double* sum_ms, long long* count, int reset) {
  if (device < 0 || device >= 64) return RPE_ERR_ARG;
  LaneClock& c = g_clock[device];
  std::lock_guard<std::mutex> g(c.mu);
  if (c.ok)
    for (int i = 0; i < LaneClock::kRing; ++i) lane_clock_fold(c, (c.next + i) % LaneClock::kRing);
  std::vector<std::pair<float, float>> v = c.iv;
  std::sort(v.begin(), v.end());
  double sum = 0.0;
  float covered = -1e30f;
  for (const auto& e : v) {
    const float lo = e.first > covered ? e.first : covered;
    if (e.second > lo) sum += (double)(e.second - lo);
    if (e.second > covered) covered = e.second;
  }
  if (sum_ms) *sum_ms = sum;
  if (count) *count = (long long)v.size();
  if (reset) {
    c.iv.clear();
    for (int i = 0; i < LaneClock::kRing; ++i) c.live[i] = false;  // (a launch still running is dropped)
    c.ref_set = false;
  }
  return RPE_OK;
}
int rpe_last_stage_ms(rpe_ctx* ctx, float ms[8]) {
  if (!ctx || !ms) return RPE_ERR_ARG;
  for (int k = 0; k < ST_COUNT; ++k) ms[k] = ctx->stage_ms[k];
  return RPE_OK;
}

// test hook: choose the packed (FFMA2) or scalar (FFMA) fast kernel
int rpe_debug_set_packed(int packed) {
  rpe::set_use_packed(packed != 0);
  return RPE_OK;
}
// test hook: shrink the borderline worklist so that the overflow -> whole-frame exact rescoring path can be exercised
int rpe_debug_set_worklist_capacity(rpe_ctx* ctx, unsigned int cap) {
  if (!ctx) return RPE_ERR_ARG;
  ctx->wl.capacity = cap < ctx->wl_allocated ? cap : ctx->wl_allocated;
  ctx->wl_fixed = cap < ctx->wl_allocated;
  return RPE_OK;
}
int rpe_debug_set_raw_tiles(int v) {
  g_raw_tiles = v != 0;
  return RPE_OK;
}
int rpe_debug_f64_exact_only(int v) {
  g_f64_exact_only = v != 0;
  return RPE_OK;
}
int rpe_debug_force_exact_multi(int v) {
  g_force_exact_multi = v != 0;
  return RPE_OK;
}
int rpe_debug_set_nosync(int v) {
  rpe::set_nosync(v);
  return RPE_OK;
}
int rpe_debug_set_score_variant(int v) {
  rpe::set_score_variant(v);
  return RPE_OK;
}
// Host logic of the opt-in uniform-register scorer's launcher, callable without a device (tests): CTA shape for a frame of
// `n` correspondences against `nslots` hypothesis slots on `num_sms` SMs; out6 = {pairs per thread, threads, hypothesis
// rows, correspondence columns, pairs per column, tail pairs}. Returns 1 if the kernel takes the frame.
int rpe_debug_ur_shape(int n, int nslots, int num_sms, int* out6) {
  if (!out6 || n <= 0) return 0;
  const int npairs = (n + 1) / 2;
  const int npairs_pad = ((npairs + rpe::kSubPairs - 1) / rpe::kSubPairs) * rpe::kSubPairs;
  return rpe::ur_shape_for(npairs_pad, nslots, num_sms, out6);
}
// 0: the Kabsch refits of the CURRENT device always take the Jacobi SVD instead of the polar iteration (A/B in the tests)
int rpe_debug_set_kabsch_polar(int on) {
  cudaDeviceSynchronize();
  rpe::set_kabsch_polar(on);
  return cudaGetLastError() == cudaSuccess ? RPE_OK : RPE_ERR_CUDA;
}
// Back to the shipped configuration: every process-global test hook, and (ctx may be NULL) the per-context ones.
int rpe_debug_reset(rpe_ctx* ctx) {
  rpe::set_use_packed(true);
  rpe::set_score_variant(rpe::kDefaultScoreVariant);
  rpe::set_nosync(0);
  g_raw_tiles = true;
  g_f64_exact_only = false;
  g_force_exact_multi = false;
  if (ctx) {
    ctx->wl.capacity = ctx->wl_allocated;
    ctx->wl_fixed = false;
    ctx->first_pass = kFirstPassIters;
    ctx->timing = ctx->timing_fast = false;
  }
  return RPE_OK;
}

}  // extern "C"
