// seq.cpp — rpe_seq_*: a batched sequence of frames (BASELINE config #5) driven from native host threads.
//
// The reference runs one estimator call per frame from a single thread (SimpleMain.cpp:30-49: simulate, adapter,
// shinji_ransac2, shinji_ls1). A sequence of dense frames keeps a B200 busy only if frame k+1 is enqueued while frame k
// is still being scored, so the per-frame call sequence
//     draw the sample table (Utility.hpp:138-155)  ->  rpe_upload  ->  rpe_ransac_async  ->  rpe_refit_async ...
// is issued here by a small pool of C++ threads, each round-robin over its own contexts (one CUDA stream per context),
// with nothing but the public C-ABI of rpe_c_api.h underneath. No Python in the per-frame path, no device code here.
//
// Sample tables: frame i draws exactly rpe_sample_table(sample_seed + i, n, m, H) — a persistent RandomElements whose
// generator is re-seeded per frame — so results do not depend on the number of threads or contexts and every frame
// can be checked against the CPU path on its own. (The reference continues one global rand() stream across calls; how
// far a frame advances it depends on that frame's early stop, which no pipelined implementation can know in time.)
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rpe_c_api.h"

namespace {

constexpr int kMaxInflight = 4;  // shared mode: unfinished frames per context, at most (RPE_SEQ_INFLIGHT, default 2)
constexpr int kTableSlots = 4;  // pinned sample tables per context; a slot is reused only after its frame's copy ran

struct SeqContext {
  rpe_ctx* ctx = nullptr;
  int32_t* tables = nullptr;  // pinned, kTableSlots x H x 4
  cudaEvent_t ev[kTableSlots] = {};
  bool ev_used[kTableSlots] = {};
  int next = 0;
  // shared mode: at most kInflight frames enqueued and unfinished per context, so that a sequence takes frames at the
  // pace its GPU (and its PCIe path) works them off instead of at the pace its host thread can enqueue them
  cudaEvent_t done[kMaxInflight] = {};
  bool done_used[kMaxInflight] = {};
  int done_next = 0;
};

struct Job {
  const rpe_seq_frame* ring = nullptr;
  int ring_len = 0;
  long long first = 0;
  int n_frames = 0;  // static mode: frames of this call; shared mode: capacity of the result arrays
  rpe_result* ransac_out = nullptr;
  rpe_result* final_out = nullptr;
  // shared mode (rpe_seq_run_shared): frame indices come from a counter that several sequences — one per GPU, usually
  // one per process, the counter in shared memory — advance together
  long long* shared_next = nullptr;
  long long total = 0;
  long long* frame_index_out = nullptr;
};

}  // namespace

struct rpe_seq {
  rpe_seq_params p;
  std::vector<SeqContext> ctxs;
  std::vector<std::thread> workers;
  std::vector<rpe_sampler*> samplers;  // one per worker
  std::vector<int> sampler_n;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  Job job;
  int inflight = 2;
  std::atomic<int> local_done{0};  // shared mode: frames this sequence has taken in the current run
  unsigned long long generation = 0;
  int running = 0;
  bool quit = false;
  std::vector<int> status;  // per worker, of the last run
  std::vector<std::string> errors;
  std::string err;
};

namespace {

int method_sample_size(int method) { return method == RPE_SHINJI ? 3 : 4; }

// One frame on one context. Returns an rpe status.
int issue_frame(rpe_seq* s, int worker, SeqContext& sc, const rpe_seq_frame& f, long long frame_index, rpe_result* r_out,
                rpe_result* f_out) {
  const rpe_seq_params& p = s->p;
  int rc = rpe_poll(sc.ctx);  // hand over what has finished on this context (bit-form masks are expanded here, as we go)
  if (rc) return rc;
  if (f.on_device)
    rc = rpe_upload_device(sc.ctx, f.bv, f.xc, f.nc, f.xw, f.nw, f.n);
  else
    rc = rpe_upload(sc.ctx, f.bv, f.xc, f.nc, f.xw, f.nw, f.n);
  if (rc) return rc;
  const int32_t* samples = f.samples;
  if (!samples) {
    // draw inside the timed path, like the reference does inside its loop
    const int slot = sc.next;
    sc.next = (sc.next + 1) % kTableSlots;
    if (sc.ev_used[slot] && cudaEventSynchronize(sc.ev[slot]) != cudaSuccess) return RPE_ERR_CUDA;
    int32_t* tab = sc.tables + (size_t)slot * p.H * 4;
    if (s->sampler_n[worker] != f.n) {
      if (s->samplers[worker]) rpe_sampler_destroy(s->samplers[worker]);
      s->samplers[worker] = nullptr;
      rc = rpe_sampler_create(1u, f.n, &s->samplers[worker]);
      if (rc) return rc;
      s->sampler_n[worker] = f.n;
    }
    rc = rpe_sampler_reseed(s->samplers[worker], p.sample_seed + (uint32_t)frame_index);
    if (!rc) rc = rpe_sampler_rows(s->samplers[worker], method_sample_size(p.method), p.H, tab);
    if (rc) return rc;
    samples = tab;
    rc = rpe_ransac_async(sc.ctx, p.method, samples, p.H, p.thr3d, p.cos_thr2d, p.cos_thrN, p.confidence, r_out, f.mask);
    if (rc) return rc;
    // the table's host-to-device copy is stream-ordered in front of this event
    if (cudaEventRecord(sc.ev[slot], (cudaStream_t)rpe_stream(sc.ctx)) != cudaSuccess) return RPE_ERR_CUDA;
    sc.ev_used[slot] = true;
  } else {
    rc = rpe_ransac_async(sc.ctx, p.method, samples, p.H, p.thr3d, p.cos_thr2d, p.cos_thrN, p.confidence, r_out, f.mask);
    if (rc) return rc;
  }
  if (p.refit & RPE_SEQ_REFIT_KABSCH) {
    rc = rpe_refit_async(sc.ctx, RPE_REFIT_KABSCH_INLIERS, nullptr, 0, f_out);
    if (rc) return rc;
  }
  if (p.refit & RPE_SEQ_REFIT_NL_SK_LS) {
    rc = rpe_refit_async(sc.ctx, RPE_REFIT_NL_SK_LS, nullptr, 0, f_out);
    if (rc) return rc;
  }
  if (p.refit & RPE_SEQ_REFIT_GN) {
    rc = rpe_refit_async(sc.ctx, RPE_REFIT_GN, nullptr, p.gn_iters, f_out);
    if (rc) return rc;
  }
  return RPE_OK;
}

void worker_main(rpe_seq* s, int worker) {
  cudaSetDevice(s->p.device);
  const int T = s->p.n_threads;  // NOT workers.size(): the vector is still being filled while the first threads start
  unsigned long long seen = 0;
  // contexts of this worker: worker, worker + T, ...
  std::vector<int> mine;
  for (int c = worker; c < (int)s->ctxs.size(); c += T) mine.push_back(c);
  std::vector<rpe_result> scratch(mine.size() * 2);  // results nobody asked for still need a landing place
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(s->mu);
      s->cv_go.wait(lk, [&] { return s->quit || s->generation != seen; });
      if (s->quit) return;
      seen = s->generation;
      job = s->job;
    }
    int rc = RPE_OK;
    size_t k = 0;
    if (job.shared_next) {
      for (; rc == RPE_OK; ++k) {
        const size_t ci = k % mine.size();
        SeqContext& sc = s->ctxs[mine[ci]];
        // backpressure first, THEN take a frame: a frame is claimed only when this context can start it
        const int d = sc.done_next;
        if (sc.done_used[d] && cudaEventSynchronize(sc.done[d]) != cudaSuccess) {
          rc = RPE_ERR_CUDA;
          break;
        }
        const long long fi = __atomic_fetch_add(job.shared_next, 1LL, __ATOMIC_RELAXED);
        if (fi >= job.total) break;
        const int i = s->local_done.fetch_add(1);
        if (i >= job.n_frames) {  // result arrays full (cannot happen when the caller sized them for `total`)
          rc = RPE_ERR_ARG;
          s->errors[worker] = "rpe_seq_run_shared: result arrays too small";
          break;
        }
        const rpe_seq_frame& f = job.ring[(size_t)(fi % job.ring_len)];
        rpe_result* r_out = job.ransac_out ? &job.ransac_out[i] : &scratch[2 * ci];
        rpe_result* f_out = job.final_out ? &job.final_out[i] : &scratch[2 * ci + 1];
        if (job.frame_index_out) job.frame_index_out[i] = fi;
        rc = issue_frame(s, worker, sc, f, fi, r_out, f_out);
        if (rc) {
          s->errors[worker] = rpe_last_error(sc.ctx);
          break;
        }
        if (cudaEventRecord(sc.done[d], (cudaStream_t)rpe_stream(sc.ctx)) != cudaSuccess) rc = RPE_ERR_CUDA;
        sc.done_used[d] = true;
        sc.done_next = (d + 1) % s->inflight;
      }
    } else
    for (int i = worker; i < job.n_frames && rc == RPE_OK; i += T, ++k) {
      const size_t ci = k % mine.size();
      SeqContext& sc = s->ctxs[mine[ci]];
      const long long fi = job.first + i;
      const rpe_seq_frame& f = job.ring[(size_t)(fi % job.ring_len)];
      rpe_result* r_out = job.ransac_out ? &job.ransac_out[i] : &scratch[2 * ci];
      rpe_result* f_out = job.final_out ? &job.final_out[i] : &scratch[2 * ci + 1];
      rc = issue_frame(s, worker, sc, f, fi, r_out, f_out);
      if (rc) s->errors[worker] = rpe_last_error(sc.ctx);
    }
    for (int c : mine) {
      const int rs = rpe_sync(s->ctxs[c].ctx);
      if (rs && rc == RPE_OK) {
        rc = rs;
        s->errors[worker] = rpe_last_error(s->ctxs[c].ctx);
      }
    }
    {
      std::lock_guard<std::mutex> lk(s->mu);
      s->status[worker] = rc;
      if (--s->running == 0) s->cv_done.notify_all();
    }
  }
}

}  // namespace

extern "C" {

int rpe_seq_create(const rpe_seq_params* params, rpe_seq** out) {
  if (!params || !out) return RPE_ERR_ARG;
  *out = nullptr;
  if (params->n_contexts < 1 || params->n_contexts > 64 || params->n_threads < 1 || params->H < 1) return RPE_ERR_ARG;
  rpe_seq* s = new rpe_seq();
  s->p = *params;
  if (const char* e = getenv("RPE_SEQ_INFLIGHT")) {
    const int v = atoi(e);
    s->inflight = v < 1 ? 1 : (v > kMaxInflight ? kMaxInflight : v);
  }
  if (s->p.n_threads > s->p.n_contexts) s->p.n_threads = s->p.n_contexts;
  if (cudaSetDevice(s->p.device) != cudaSuccess) {
    (void)cudaGetLastError();
    delete s;
    return RPE_ERR_NO_DEVICE;
  }
  s->ctxs.resize(s->p.n_contexts);
  int rc = RPE_OK;
  for (SeqContext& sc : s->ctxs) {
    rc = rpe_create(s->p.device, &sc.ctx);
    if (rc) break;
    // RPE_SEQ_MASK_BITS=1: masks travel as bits and are expanded by the issuing threads (rpe_set_mask_transfer). Off by
    // default: on the 8-GPU box (4 host cores per GPU) the expansion costs the issuing threads more than the bus gains —
    // 21.1 k against 22.7 k frames/s, although the concurrent-upload ceiling rises from 190 to 226 GB/s (round 2).
    // RPE_SEQ_MASK_BITS=2: the constant column 0 of the 3-D / 3-D family stays on the device (rpe_set_mask_transfer(2)).
    static const int mask_mode = getenv("RPE_SEQ_MASK_BITS") ? getenv("RPE_SEQ_MASK_BITS")[0] - '0' : 0;
    if (mask_mode == 1 || mask_mode == 2) rpe_set_mask_transfer(sc.ctx, mask_mode);
    if (cudaMallocHost(&sc.tables, (size_t)kTableSlots * s->p.H * 4 * sizeof(int32_t)) != cudaSuccess) {
      rc = RPE_ERR_NOMEM;
      break;
    }
    for (int k = 0; k < kTableSlots && !rc; ++k)
      if (cudaEventCreateWithFlags(&sc.ev[k], cudaEventDisableTiming) != cudaSuccess) rc = RPE_ERR_CUDA;
    for (int k = 0; k < kMaxInflight && !rc; ++k)
      if (cudaEventCreateWithFlags(&sc.done[k], cudaEventDisableTiming) != cudaSuccess) rc = RPE_ERR_CUDA;
    if (rc) break;
  }
  if (rc) {
    rpe_seq_destroy(s);
    return rc;
  }
  const int T = s->p.n_threads;
  s->samplers.assign(T, nullptr);
  s->sampler_n.assign(T, 0);
  s->status.assign(T, RPE_OK);
  s->errors.assign(T, std::string());
  for (int t = 0; t < T; ++t) s->workers.emplace_back(worker_main, s, t);
  *out = s;
  return RPE_OK;
}

static int seq_run_job(rpe_seq* s, const Job& job) {
  {
    std::unique_lock<std::mutex> lk(s->mu);
    s->job = job;
    s->local_done.store(0);
    for (SeqContext& sc : s->ctxs) {
      for (int k = 0; k < kMaxInflight; ++k) sc.done_used[k] = false;
      sc.done_next = 0;
    }
    s->running = s->p.n_threads;
    ++s->generation;
    s->cv_go.notify_all();
    s->cv_done.wait(lk, [&] { return s->running == 0; });
  }
  for (size_t t = 0; t < s->status.size(); ++t)
    if (s->status[t]) {
      s->err = s->errors[t];
      return s->status[t];
    }
  return RPE_OK;
}

int rpe_seq_run(rpe_seq* s, const rpe_seq_frame* ring, int ring_len, long long first_frame, int n_frames,
                rpe_result* ransac_out, rpe_result* final_out) {
  if (!s || !ring || ring_len < 1 || n_frames < 0 || first_frame < 0) return RPE_ERR_ARG;
  if (n_frames == 0) return RPE_OK;
  Job job;
  job.ring = ring;
  job.ring_len = ring_len;
  job.first = first_frame;
  job.n_frames = n_frames;
  job.ransac_out = ransac_out;
  job.final_out = final_out;
  return seq_run_job(s, job);
}

int rpe_seq_run_shared(rpe_seq* s, const rpe_seq_frame* ring, int ring_len, long long* shared_next, long long total,
                       int capacity, rpe_result* ransac_out, rpe_result* final_out, long long* frame_index_out, int* n_done) {
  if (!s || !ring || ring_len < 1 || !shared_next || total < 0 || capacity < 1) return RPE_ERR_ARG;
  Job job;
  job.ring = ring;
  job.ring_len = ring_len;
  job.n_frames = capacity;
  job.ransac_out = ransac_out;
  job.final_out = final_out;
  job.shared_next = shared_next;
  job.total = total;
  job.frame_index_out = frame_index_out;
  const int rc = seq_run_job(s, job);
  if (n_done) {
    const int d = s->local_done.load();
    *n_done = d < capacity ? d : capacity;
  }
  return rc;
}

rpe_ctx* rpe_seq_context(rpe_seq* s, int index) {
  return (s && index >= 0 && index < (int)s->ctxs.size()) ? s->ctxs[index].ctx : nullptr;
}
int rpe_seq_num_contexts(const rpe_seq* s) { return s ? (int)s->ctxs.size() : 0; }
const char* rpe_seq_last_error(const rpe_seq* s) { return s ? s->err.c_str() : "null sequence"; }

int rpe_seq_destroy(rpe_seq* s) {
  if (!s) return RPE_OK;
  {
    std::lock_guard<std::mutex> lk(s->mu);
    s->quit = true;
    s->cv_go.notify_all();
  }
  for (std::thread& t : s->workers)
    if (t.joinable()) t.join();
  cudaSetDevice(s->p.device);
  for (SeqContext& sc : s->ctxs) {
    if (sc.ctx) rpe_destroy(sc.ctx);
    if (sc.tables) cudaFreeHost(sc.tables);
    for (int k = 0; k < kTableSlots; ++k)
      if (sc.ev[k]) cudaEventDestroy(sc.ev[k]);
    for (int k = 0; k < kMaxInflight; ++k)
      if (sc.done[k]) cudaEventDestroy(sc.done[k]);
  }
  for (rpe_sampler* sm : s->samplers)
    if (sm) rpe_sampler_destroy(sm);
  (void)cudaGetLastError();
  delete s;
  return RPE_OK;
}

}  // extern "C"
