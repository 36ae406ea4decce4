/* rpe_c_api.h — C-ABI of the B200-native robust absolute-pose hot path.
 *
 * This is the drop-in boundary. The reference (ShudaLi/rgbd_pose_estimation) is a
 * header-only C++ template library; the host side of this project keeps its class and function
 * names in include/rpe/ (PoseAdapterBase.hpp, PnPPoseAdapter.hpp, AOPoseAdapter.hpp,
 * NormalAOPoseAdapter.hpp, AOOnlyPoseAdapter.hpp, AbsoluteOrientation.hpp, P3P.hpp,
 * AbsoluteOrientationNormal.hpp) and every estimator body marshals into the entry points
 * declared here. Plain pointers and sizes only; no C++/torch/Eigen types cross this line.
 *
 * Reference interface each entry point replaces (paths into /root/reference):
 *   rpe_upload            adapters' const-reference 3xN matrices         PnPPoseAdapter.hpp:96-99, AOPoseAdapter.hpp:88,
 *                                                                        NormalAOPoseAdapter.hpp:83-84, AOOnlyPoseAdapter.hpp:94-95
 *   rpe_ransac            shinji_ransac / shinji_ransac2                  AbsoluteOrientation.hpp:101-213
 *                         shinji_prosac / shinji_kneip_ransac / _prosac  AbsoluteOrientation.hpp:215-271, 367-515
 *                         kneip_ransac / kneip_prosac                    P3P.hpp:320-469
 *                         nl_kneip_ransac / nl_shinji_ransac / nl_shinji_kneip_ransac
 *                                                                        AbsoluteOrientationNormal.hpp:215-445
 *                         (result -> setRcw/sett/setMaxVotes/setInlier: PoseAdapterBase.hpp:109-122,
 *                          PnPPoseAdapter.hpp:196-202, AOPoseAdapter.hpp:171-184, NormalAOPoseAdapter.hpp:179-195)
 *   rpe_refit             shinji_ls / shinji_ls1 / shinji_ls2            AbsoluteOrientation.hpp:273-342
 *                         nl_shinji_kneip_ls                             AbsoluteOrientationNormal.hpp:447-552
 *                         Gauss-Newton/LM on SE3 (north-star addition; no reference code)
 *   rpe_update_num_iters  RANSACUpdateNumIters                           P3P.hpp:296-318
 *   rpe_sample_table      RandomElements<int>::run                       Utility.hpp:125-156
 *   rpe_prosac_table      ProsacSampler::sample + getSortedIdx           Utility.hpp:161-250, PnPPoseAdapter.hpp:239-255
 *   rpe_sim_*             Simulator.hpp generators                       Simulator.hpp:158-367
 *   rpe_ao / rpe_ao_ransac  extern "C" ao() / ao_ransac()                Library.cpp:17-75
 *   rpe_min_ev / rpe_min_ms ev() / ms()                                    MinimalSolvers.hpp:49-104, 10-46
 *   rpe_seq_*             the per-frame call sequence of SimpleMain.cpp:30-49 for a sequence of frames
 *
 * Error behaviour: the reference returns void, asserts in debug builds and std::abort()s inside
 * SOPHUS_ENSURE (sophus/common.hpp:115-132). Every function here returns an int status instead
 * (0 = RPE_OK, negative = error, text via rpe_last_error); a hypothesis whose rotation would have
 * tripped SOPHUS_ENSURE is dropped (its vote slot is -1), nothing aborts.
 *
 * There is NO CPU fallback: rpe_create fails with RPE_ERR_NO_DEVICE when no CUDA device is
 * usable, and every compute entry point requires a context.
 */
#ifndef RPE_C_API_H_
#define RPE_C_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPE_API_VERSION 1

/* status codes */
#define RPE_OK 0
#define RPE_ERR_ARG (-1)
#define RPE_ERR_CUDA (-2)
#define RPE_ERR_STATE (-3)
#define RPE_ERR_NO_DEVICE (-4)
#define RPE_ERR_COMM (-5)
#define RPE_ERR_NOMEM (-6)

/* estimator families; numbering shared with the oracle (oracle/ransac.hpp) */
#define RPE_SHINJI 0          /* 3-D/3-D only, 3-point sample            shinji_ransac(2)            */
#define RPE_KNEIP 1           /* 2-D/3-D only, matrix-form rotation      kneip_ransac (P3P.hpp:365)  */
#define RPE_SHINJI_KNEIP 2    /* 3-D + 2-D votes, slots {AO, P3P}        shinji_kneip_ransac         */
#define RPE_NL_KNEIP 3        /* normal + 2-D votes, slot {P3P}          nl_kneip_ransac             */
#define RPE_NL_SHINJI 4       /* normal + 3-D votes, slots {AO, nl_2p}   nl_shinji_ransac            */
#define RPE_NL_SHINJI_KNEIP 5 /* all three votes, slots {AO, P3P, nl_2p} nl_shinji_kneip_ransac      */
#define RPE_KNEIP_QUAT 6      /* 2-D only, quaternion-form rotation      kneip_prosac (P3P.hpp:442)  */

/* refit kinds for rpe_refit */
#define RPE_REFIT_KABSCH_INLIERS 0 /* shinji_ls / shinji_ls1: Kabsch over 3-D inliers of the last RANSAC  */
#define RPE_REFIT_KABSCH_ALL 1     /* shinji_ls2: Kabsch over all correspondences                          */
#define RPE_REFIT_GN 2             /* LM/Gauss-Newton on SE3 over the inliers of every modality            */
#define RPE_REFIT_NL_SK_LS 3       /* nl_shinji_kneip_ls (3 weighted Kabsch + ray-intersection passes)     */

typedef struct rpe_ctx rpe_ctx; /* opaque: one per (host thread, GPU, stream) */
/* sample-row producer of rpe_ransac_stream / rpe_ransac_f64 (documented there) */
typedef int (*rpe_sample_fn)(void* user, int first_iteration, int count, int32_t* rows);

typedef struct rpe_result {
  float R[9];       /* R_cw, row-major (same convention Library.cpp:66-69 writes out) */
  float q[4];       /* unit quaternion x,y,z,w (Sophus/Eigen coefficient order)       */
  float t[3];       /* t_w: x_c = R_cw x_w + t                                        */
  int32_t max_votes;   /* adapter.getMaxVotes()                                       */
  int32_t iter_final;  /* the in/out `Iter` after adaptive shrinking                  */
  int32_t winner;      /* iteration*slots + slot of the accepted hypothesis, -1 none  */
  int32_t n_slots;     /* hypothesis slots generated and scored (H * slots)           */
  int32_t n_borderline;/* evaluations resolved by the exact-order path                */
  int32_t flags;       /* bit0: worklist overflow -> whole frame rescored exactly; bit1: binary64 path */
  int32_t n_inliers[3];/* per modality column (2-D, 3-D, normal) of the winner        */
  int32_t refit_ok;    /* 1 if the last rpe_refit produced a valid rotation           */
  double  refit_cost;  /* GN: final weighted sum of squared residuals                  */
  int32_t refit_evals; /* GN: cost/Jacobian evaluations executed                      */
  int32_t reserved;
  double  qd[4];       /* binary64 path (rpe_upload_f64): the accepted hypothesis exactly as the double CPU path */
  double  td[3];       /* holds it; q/t/R above are its float roundings. Zero on the float path.              */
} rpe_result;

/* ---- library / device ---------------------------------------------------------------- */
int rpe_version(void);
const char* rpe_status_string(int status);
int rpe_device_count(int* count);

int rpe_create(int device, rpe_ctx** ctx);                          /* owns a non-blocking stream */
int rpe_create_on_stream(int device, void* cuda_stream, rpe_ctx** ctx); /* borrows caller's stream */
int rpe_destroy(rpe_ctx* ctx);
const char* rpe_last_error(const rpe_ctx* ctx);
void* rpe_stream(rpe_ctx* ctx);
int rpe_sync(rpe_ctx* ctx);
/* Non-blocking companion of rpe_sync: hands over (fills the caller's rpe_result structs, expands bit-form masks) every
 * result of the asynchronous calls whose device work has finished, oldest first, and returns without waiting for the rest. */
int rpe_poll(rpe_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
long long rpe_launch_count(const rpe_ctx* ctx);

/* page-locked host memory for callers that want overlap of H2D with compute */
int rpe_host_alloc(size_t bytes, void** ptr);
int rpe_host_free(void* ptr);

/* ---- correspondences -------------------------------------------------------------------
 * Every array is column-major 3 x n float (n contiguous xyz triples: the memory layout of the
 * Eigen::Matrix<float,Dynamic,Dynamic> the reference adapters hold). NULL = modality absent.
 *   bv: bearing vectors (camera)   xc: points (camera)   nc: normals (camera)
 *   xw: points (world)             nw: normals (world)
 * rpe_upload copies from host memory (pinned or pageable) and repacks on the device;
 * rpe_upload_device takes device pointers (no PCIe traffic). Both are asynchronous on the
 * context's stream. `focal` is adapter.getFocal() (PoseAdapterBase.hpp:126).            */
int rpe_upload(rpe_ctx* ctx, const float* bv, const float* xc, const float* nc, const float* xw, const float* nw, int n);
int rpe_upload_device(rpe_ctx* ctx, const float* bv, const float* xc, const float* nc, const float* xw, const float* nw,
                      int n);
/* Latency knob for a caller that waits for every frame: with chunks >= 2 (at most 8), rpe_upload of PAGE-LOCKED host
 * arrays of the 3-D / 3-D family (x_c and x_w only, n >= 65 536) copies the frame in `chunks` pieces on a copy stream,
 * and the next rpe_ransac(_async) with RPE_SHINJI whose Iter fits one device pass starts before the copy has finished:
 * the generator reads its 3 H sample points straight from the host arrays over PCIe while the first chunk is already
 * on the bus, the scorer runs once per chunk as the chunk lands (votes accumulate). Results are identical. A dense
 * frame's blocking latency drops by about half of the 0.15 ms its upload takes (0.47 -> 0.40 ms); a pipeline of
 * asynchronous frames gains nothing (and pays three more scorer launches per frame), hence off by default (chunks = 0).
 * chunks = 1 ("stream mode", H <= 1024): no copy at all — the scorer's bulk-TMA loads read the frame from the page-locked
 * host arrays while it scores and leave the device copy behind for the fix-up, mask and refit kernels (same latency as
 * 4 chunks, no copy-engine work). Pageable host memory is always copied the plain way. */
int rpe_set_upload_overlap(rpe_ctx* ctx, int chunks);
/* Tp = double adapters (the reference's TestMain.cpp runs its estimators as <double>): host arrays in binary64. The
 * next rpe_ransac on this context generates, scores, replays and masks in binary64 in the reference's operation order
 * (no binary32 fast path), so winner / votes / Iter / masks equal the double CPU path's; the accepted hypothesis comes
 * back in rpe_result.qd / td. Refits run on float copies of the arrays with binary64 accumulation. */
int rpe_upload_f64(rpe_ctx* ctx, const double* bv, const double* xc, const double* nc, const double* xw, const double* nw,
                   int n);
/* rpe_ransac / rpe_ransac_stream with binary64 thresholds, for a context in binary64 mode. Exactly one of `samples`
 * (whole table) and `fn` (rows on demand) is non-NULL. Blocking. */
int rpe_ransac_f64(rpe_ctx* ctx, int method, const int32_t* samples, rpe_sample_fn fn, void* user, int H, double thr3d,
                   double cos_thr2d, double cos_thrN, double confidence, rpe_result* out, int16_t* mask);
/* binary64 path: the hypotheses of the last pass, (q.x q.y q.z q.w t.x t.y t.z) per slot + valid flags */
int rpe_get_hypotheses_f64(rpe_ctx* ctx, int n_slots, double* hyps7, int32_t* valid);
int rpe_num_correspondences(const rpe_ctx* ctx);

/* ---- robust estimation ------------------------------------------------------------------
 * samples: int32 [H x 4] (host, page-locked host, or device memory) rows of correspondence indices (3 used by RPE_SHINJI), exactly the
 *          draws RandomElements::run / ProsacSampler::sample would produce (see rpe_sample_table). A row with an
 *          index outside [0, n) makes that iteration an empty slot (the reference's samplers cannot produce one).
 * H      : the caller's `Iter` on entry. Up to 1024 iterations are generated and scored on the GPU
 *          in one go; the reference's sequential rule (strict `votes > max`, Iter =
 *          RANSACUpdateNumIters(..)) is then replayed on the device, so winner / max_votes /
 *          iter_final / mask are exactly what the early-stopping CPU loop returns. A longer Iter is
 *          scored progressively (1024, 2048, 4096, then 8192 iterations per pass) and the host looks
 *          at the replayed bound between passes — those calls synchronise even when `_async`.
 * thr3d  : dist_thre_3d_ (metres). cos_thr2d: cos(atan(thre_2d_/focal)) (P3P.hpp:323).
 * cos_thrN: cos(nl_thre) (AbsoluteOrientationNormal.hpp:223). Unused ones are ignored.
 * mask   : host int16 [n x cols] column-major, cols = 1 (KNEIP*), 2 (SHINJI, SHINJI_KNEIP), 3 (NL_*);
 *          col 0 = 2-D, col 1 = 3-D, col 2 = normal flags — the matrix the reference hands to
 *          setInlier(). May be NULL.
 * rpe_ransac blocks until the result is on the host; rpe_ransac_async only enqueues (out/mask must
 * then be page-locked and stay alive until rpe_sync).                                       */
int rpe_ransac(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d, float cos_thrN,
               float confidence, rpe_result* out, int16_t* mask);
int rpe_ransac_async(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d,
                     float cos_thrN, float confidence, rpe_result* out_pinned, int16_t* mask_pinned);

/* Iterations of the first device pass (default 1024; later passes double up to 8192). An `Iter` that fits the first
 * pass is scored in one go with no host round trip — what a pipeline of asynchronous frames wants; a latency-bound
 * blocking caller whose loops usually stop after a few dozen iterations (low outlier ratios) can start smaller. The
 * result does not depend on it. */
int rpe_set_first_pass_iters(rpe_ctx* ctx, int iters);
/* Opt-in: reproduce the reference's STALE SAMPLE COLUMNS on frames with invalid depth. Its nl_shinji_ransac /
 * nl_shinji_kneip_ransac hoist the sample buffers out of the loop and assign_sample skips the camera-side columns of an
 * invalid (all-NaN) sample point (AbsoluteOrientationNormal.hpp:48-75, 299-315), so nl_2p (:315, :389) pairs the
 * current world point / normal with the camera point / normal an EARLIER iteration left in that column — and such a
 * hypothesis can win. Off (default): the invalid point's NaN goes into nl_2p and the hypothesis scores nothing. On: the
 * device reproduces the reference (columns never written read as zeros; the reference reads uninitialised memory
 * there). Also switched on for every new context by RPE_STALE_SAMPLE_COLUMNS=1 in the environment. Frames without
 * invalid camera points (all Simulator inputs) give identical results either way. */
int rpe_set_stale_sample_columns(rpe_ctx* ctx, int on);
/* Same as rpe_ransac, with the sample rows produced on demand: `fn(user, first_iteration, count, rows)` must write
 * rows [first_iteration, first_iteration + count) of the table (count x 4 int32) and return 0. It is called once per
 * device pass (1024, 2048, 4096, 8192, ... iterations) in increasing order, so a caller whose Iter is 100 000 — the
 * reference's SimpleMain.cpp:45 — only draws the rows of the passes that run before the adaptive bound stops the loop.
 * The reference draws inside its loop (AbsoluteOrientation.hpp:124, Utility.hpp:139-152): the iterations it executes
 * are max(iter_final, winner / slots + 1) (all H when winner < 0), see rpe/Estimators.hpp. Blocking. */
int rpe_ransac_stream(rpe_ctx* ctx, int method, rpe_sample_fn fn, void* user, int H, float thr3d, float cos_thr2d,
                      float cos_thrN, float confidence, rpe_result* out, int16_t* mask);

/* Refit / refinement starting from the pose and inlier mask of the last rpe_ransac on ctx.
 * weights: modality weights {w2d, w3d, wN} for RPE_REFIT_GN (NULL = {1,1,1}); per-correspondence
 * n x 3 column-major weights for RPE_REFIT_NL_SK_LS (NULL = the adapters' default 1).
 * max_iters: LM evaluations for RPE_REFIT_GN (<=0 -> 6).                                   */
int rpe_refit(rpe_ctx* ctx, int kind, const float* weights, int max_iters, rpe_result* out);
int rpe_refit_async(rpe_ctx* ctx, int kind, const float* weights, int max_iters, rpe_result* out_pinned);
/* Override the pose/mask a refit starts from (setRcw/sett + setInlier on the adapter). */
int rpe_set_pose(rpe_ctx* ctx, const float q_xyzw[4], const float t[3], int max_votes);
int rpe_set_mask(rpe_ctx* ctx, const int16_t* mask, int cols);

/* ---- stage access (parity tests, profiling, hypothesis-sharded multi-GPU) ---------------- */
/* Generate hypotheses for iterations [0,H) from `samples`; nothing is scored. */
int rpe_generate(rpe_ctx* ctx, int method, const int32_t* samples, int H);
/* Copy back generated hypotheses: hyps [n_slots x 7] = qx,qy,qz,qw,tx,ty,tz ; valid [n_slots]. */
int rpe_get_hypotheses(rpe_ctx* ctx, float* hyps, int32_t* valid, int n_slots);
/* Replace the hypothesis set (scoring arbitrary poses). */
int rpe_set_hypotheses(rpe_ctx* ctx, int method, const float* hyps, const int32_t* valid, int n_slots);
/* Score slots [slot_begin, slot_end) of the current hypothesis set; votes stay on the device. */
int rpe_score(rpe_ctx* ctx, int method, int slot_begin, int slot_end, float thr3d, float cos_thr2d, float cos_thrN);
/* votes [n_slots] int32, -1 = empty slot. */
int rpe_get_votes(rpe_ctx* ctx, int32_t* votes, int n_slots);
int rpe_set_votes(rpe_ctx* ctx, const int32_t* votes, int n_slots);
/* Device pointer of the votes array (for NCCL all-gather by the caller in sharded mode). */
int32_t* rpe_votes_device_ptr(rpe_ctx* ctx);
/* The exchange step of the hypothesis-sharded mode over peer memory instead of a collective library: every rank
 * (one process per GPU of a node) exports one small block (rpe_peer_export -> 64-byte CUDA IPC handle), the handles
 * are gathered by any host-side means (MPI, torch.distributed, a file) and imported (rpe_peer_import, `handles` =
 * world x 64 bytes in rank order). rpe_exchange_votes then enqueues one kernel that writes this rank's slots
 * [slot_begin, slot_end) of the vote table into every rank's table over NVLink, publishes a flag, waits (at most 2 s)
 * for all ranks, and leaves the complete table in the context — asynchronous, no host round trip; every rank must call
 * it once per frame. rpe_peer_status synchronises and reports a time-out. */
int rpe_peer_export(rpe_ctx* ctx, unsigned char handle[64]);
int rpe_peer_import(rpe_ctx* ctx, int rank, int world, const unsigned char* handles);
/* contexts of one process (one per GPU): direct peer pointers instead of IPC; ctxs[rank] must be ctx itself */
int rpe_peer_import_local(rpe_ctx* ctx, int rank, int world, rpe_ctx* const* ctxs);
int rpe_exchange_votes(rpe_ctx* ctx, int slot_begin, int slot_end);
int rpe_peer_status(rpe_ctx* ctx);
/* How long the exchange kernel waits for a peer (default 2000 ms). The wait also absorbs host-side skew between the
 * per-GPU processes (first-frame allocations), so callers that do not barrier before their first sharded frame may want
 * more. On a time-out the frame has no winner (votes -1) and the blocking call / the next rpe_sync returns RPE_ERR_COMM
 * (the latch is cleared by that report). */
int rpe_peer_set_timeout_ms(rpe_ctx* ctx, int ms);
/* rpe_ransac for one frame whose hypotheses are sharded over the ranks set up with rpe_peer_import: every rank holds the
 * same correspondences and sample table (H <= 8192), generates all hypotheses, scores its own contiguous slice of
 * slots, exchanges the slices through peer memory and replays the rule — all ranks return the same result. One
 * asynchronous stream of work per frame; every rank must make the same sequence of calls. */
int rpe_ransac_sharded(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d, float cos_thrN,
                       float confidence, rpe_result* out, int16_t* mask);
int rpe_ransac_sharded_async(rpe_ctx* ctx, int method, const int32_t* samples, int H, float thr3d, float cos_thr2d,
                             float cos_thrN, float confidence, rpe_result* out, int16_t* mask);
/* Replay the adaptive rule over the current votes, build the winner's mask. */
int rpe_finish(rpe_ctx* ctx, int method, int H, float thr3d, float cos_thr2d, float cos_thrN, float confidence,
               rpe_result* out, int16_t* mask);

/* ---- batched sequences of frames (BASELINE config #5) -----------------------------------------
 * The reference runs one estimator call per frame from a single thread (SimpleMain.cpp:30-49). rpe_seq_* issues the
 * same per-frame sequence — draw the sample table (Utility.hpp:138-155), rpe_upload, rpe_ransac_async, refits — from
 * `n_threads` native host threads over `n_contexts` contexts (streams) of one device, so that frame k+1 is enqueued and
 * uploaded while frame k is scored. Frames shard across GPUs by giving every process/GPU its own rpe_seq and its own
 * frame range; there is no data-path collective.
 * Frame i (global index first_frame + i) reads ring[(first_frame + i) % ring_len] and, when the frame carries no
 * sample table, draws rpe_sample_table(sample_seed + first_frame + i, n, m, H) inside the call — results do not depend
 * on n_threads / n_contexts. Host arrays and masks must be page-locked (rpe_host_alloc) for the copies to overlap. */
#define RPE_SEQ_REFIT_KABSCH 1   /* shinji_ls / shinji_ls1 after the RANSAC                      */
#define RPE_SEQ_REFIT_GN 2       /* LM on SE3 over the inliers (after the Kabsch refit if both)   */
#define RPE_SEQ_REFIT_NL_SK_LS 4 /* nl_shinji_kneip_ls                                             */
typedef struct rpe_seq rpe_seq;
typedef struct rpe_seq_params {
  int device, n_contexts, n_threads;
  int method, H;
  float thr3d, cos_thr2d, cos_thrN, confidence;
  int refit;    /* RPE_SEQ_REFIT_* bits */
  int gn_iters; /* <= 0 -> 6 */
  uint32_t sample_seed;
} rpe_seq_params;
typedef struct rpe_seq_frame {
  const float *bv, *xc, *nc, *xw, *nw; /* as rpe_upload / rpe_upload_device */
  int n;
  int on_device;          /* 0: host arrays (copied, PCIe inside the call), 1: device pointers */
  const int32_t* samples; /* H x 4 table (host page-locked or device), NULL = draw inside      */
  int16_t* mask;          /* host page-locked n x cols, or NULL                                */
} rpe_seq_frame;
int rpe_seq_create(const rpe_seq_params* params, rpe_seq** out);
/* Blocking: returns when all n_frames results are on the host. ransac_out / final_out: n_frames entries each (the
 * RANSAC result and the result of the last refit of the frame) or NULL. */
int rpe_seq_run(rpe_seq* seq, const rpe_seq_frame* ring, int ring_len, long long first_frame, int n_frames,
                rpe_result* ransac_out, rpe_result* final_out);
/* The same with the frame indices taken from a COUNTER SHARED by several sequences (one per GPU; with one process per
 * GPU the counter lives in shared memory): every issuing thread fetch-adds *shared_next and processes frame `index`
 * (reading ring[index % ring_len], table seed sample_seed + index) until the counter reaches `total`, taking a new frame
 * only when one of its contexts has fewer than two unfinished frames — so a GPU behind a slower PCIe path simply takes
 * fewer frames. Still no data-path collective. frame_index_out[i] (may be NULL) = the frame whose results are in
 * ransac_out[i] / final_out[i]; *n_done = frames this sequence processed; capacity = length of the three arrays. */
int rpe_seq_run_shared(rpe_seq* seq, const rpe_seq_frame* ring, int ring_len, long long* shared_next, long long total,
                       int capacity, rpe_result* ransac_out, rpe_result* final_out, long long* frame_index_out, int* n_done);
rpe_ctx* rpe_seq_context(rpe_seq* seq, int index); /* for rpe_enable_stage_timing / rpe_launch_count */
int rpe_seq_num_contexts(const rpe_seq* seq);
const char* rpe_seq_last_error(const rpe_seq* seq);
int rpe_seq_destroy(rpe_seq* seq);

/* ---- host-side helpers that the reference computes on the CPU too ------------------------- */
/* RANSACUpdateNumIters<float> with the bit-reproducible log of include/rpe/det_math.h. */
int rpe_update_num_iters(float p, float ep, int model_points, int max_iters);
/* H consecutive RandomElements<int>::run(m) draws. The generator restates glibc's rand()
 * (TYPE_3, what the reference's ::rand() is) from `seed`; seed 1 == a process that never
 * called srand(). Rows are padded to 4 with -1. */
int rpe_sample_table(uint32_t seed, int n, int m, int H, int32_t* samples);
/* H consecutive ProsacSampler draws mapped through the descending-weight order. */
int rpe_prosac_table(uint32_t seed, int n, int m, int H, const float* weights, int32_t* samples);
/* A RandomElements<int> sampler (Utility.hpp:125-156) that lives across calls, for per-frame draws: the reference builds
 * one per estimator call and re-initialises its n-entry permutation on every draw; this one keeps the permutation and
 * undoes its swaps, so H draws cost O(H m) instead of O(H n) (27 us for n = 307 200, H = 1 024). Successive
 * rpe_sampler_rows calls continue the same rand() stream: two calls of H/2 rows give the rows of one call of H rows,
 * and a fresh sampler gives what rpe_sample_table gives for the same seed. Host only; one sampler per thread. */
typedef struct rpe_sampler rpe_sampler;
int rpe_sampler_create(uint32_t seed, int n, rpe_sampler** out);
int rpe_sampler_rows(rpe_sampler* s, int m, int H, int32_t* samples);
void rpe_sampler_destroy(rpe_sampler* s);
/* Restart the sampler's generator from `seed` (what srand(seed) does to ::rand()); the permutation is untouched, so
 * the next rows are exactly rpe_sample_table(seed, ...)'s at O(1) cost — per-frame tables of a sequence. */
int rpe_sampler_reseed(rpe_sampler* s, uint32_t seed);

/* ---- synthetic correspondences (Simulator.hpp) -------------------------------------------- */
/* pose: R = Rz*Ry*Rx from uniform angles (Simulator.hpp:23-83, use_gaussian=false), t = size*U(-1,1)^3 (:16-21) */
int rpe_sim_pose(uint64_t seed, float max_angle_rad, float t_size, float q_xyzw[4], float t[3]);
/* simulate_3d_3d_correspondences (Simulator.hpp:268-314). Q = noisy/outlier world points, P = clean camera points. */
int rpe_sim_3d_3d(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise, float outlier_ratio,
                  float min_depth, float max_depth, float f, int use_gaussian, float* Q_xw, float* P_xc, float* weights3);
/* simulate_2d_3d_correspondences (:175-233). U = unit bearing vectors. */
int rpe_sim_2d_3d(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise_px, float outlier_ratio,
                  float min_depth, float max_depth, float f, int use_gaussian, float* Q_xw, float* U_bv, float* P_gt,
                  float* weights3);
/* simulate_2d_3d_nl_correspondences (:316-367). */
int rpe_sim_2d_3d_nl(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d, float or2d, float n3d,
                     float or3d, float nnl, float ornl, float min_depth, float max_depth, float f, int use_gaussian,
                     float* Q_xw, float* M_nw, float* P_xc, float* N_nc, float* U_bv, float* weights3);

/* simulate_kinect_2d_3d_nl_correspondences (:389-436): camera points perturbed by the Kinect lateral / axial noise model
 * of Nguyen, Izadi & Lovell (:368-387); weights3 column 1 = sigma_axial(0, min_depth) / sigma_axial. */
int rpe_sim_kinect_2d_3d_nl(uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d, float or2d, float or3d,
                            float nnl, float ornl, float min_depth, float max_depth, float f, float* Q_xw, float* M_nw,
                            float* P_xc, float* N_nc, float* U_bv, float* weights3);

/* Device-side generators: the same distributions produced straight into the context's device arrays (no PCIe
 * traffic; counter-based random stream, so the values differ from the host generators for the same seed).
 * After the call the context holds n correspondences exactly as after rpe_upload. */
int rpe_sim_3d_3d_device(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise,
                         float outlier_ratio, float min_depth, float max_depth, float f, int use_gaussian);
int rpe_sim_2d_3d_nl_device(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d,
                            float or2d, float n3d, float or3d, float nnl, float ornl, float min_depth, float max_depth,
                            float f, int use_gaussian);
/* rpe_sim_3d_3d_device into device buffers of the caller (3 x n floats each); the context's own frame is untouched.
 * Asynchronous on the context's stream. For sequences of distinct frames resident in HBM (config #5). */
int rpe_sim_3d_3d_device_to(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float noise,
                            float outlier_ratio, float min_depth, float max_depth, float f, int use_gaussian, float* d_xw,
                            float* d_xc);
/* simulate_kinect_2d_3d_nl_correspondences on the device (Kinect lateral / axial noise on the camera points). */
int rpe_sim_kinect_2d_3d_nl_device(rpe_ctx* ctx, uint64_t seed, const float q_xyzw[4], const float t[3], int n, float n2d,
                                   float or2d, float or3d, float nnl, float ornl, float min_depth, float max_depth, float f);
/* Copy the context's current correspondence arrays back to host memory (NULL = skip). */
int rpe_download(rpe_ctx* ctx, float* bv, float* xc, float* nc, float* xw, float* nw);

/* ---- MinimalSolvers.hpp (ev: :49-83, ms: :10-46) ------------------------------------------------
 * Batches on the device, one problem per thread; the *_host forms run the same templates (rpe/solvers_min.h) on the
 * CPU and return identical bits. M9: count x 9 row-major symmetric 3x3 -> E3: count x 3 eigenvalues, descending.
 * in24: count x (Aw Bw Nw Mw Ac Bc Nc Mc), the argument order of ms() -> q4 (x,y,z,w of R_cw), t3 (t_w).
 * Pointers may be host or device memory. Blocking. */
int rpe_min_ev(rpe_ctx* ctx, const float* M9, int count, float* E3);
int rpe_min_ms(rpe_ctx* ctx, const float* in24, int count, float* q4, float* t3);
int rpe_min_ev_host(const float* M9, int count, float* E3);
int rpe_min_ev_host_f64(const double* M9, int count, double* E3);
int rpe_min_ms_host(const float* in24, int count, float* q4, float* t3);

/* ---- the reference's own C shim (Library.cpp:17-75), same argument meaning ------------------ */
/* x_w, x_c: 3 x n column-major; R_cw out row-major 9; t out 3. ao = shinji_ls2; ao_ransac =
 * shinji_ransac2(thr 0.1, 1000 iterations, confidence 0.99999) + shinji_ls1. */
int rpe_ao(const float* x_w, const float* x_c, int n, float* R_cw, float* t);
int rpe_ao_ransac(const float* x_w, const float* x_c, int n, float* R_cw, float* t);

/* ---- microbenchmarks used by bench.py for the roofline denominators ------------------------ */
/* Sustained FP32 FFMA throughput of the device in TFLOP/s (2 flop per FMA), measured with CUDA
 * events over `ms_target` milliseconds of dependent-chain-free FFMA work. */
int rpe_measure_ffma_tflops(rpe_ctx* ctx, int ms_target, double* tflops_scalar, double* tflops_packed);
/* Device time in ms of the last call's stages, measured with CUDA events on the context's stream:
 * [0] upload+pack [1] generate [2] score (fast + exact fix-up) [3] replay [4] mask+refit [5] GN [6] total
 * [7] the tiled fast scoring kernel alone (the roofline kernel)
 * rpe_enable_stage_timing: 0 = off, 1 = every stage (ten event records per frame; they cost a few percent of
 * throughput when several contexts overlap), 2 = only the two events around the tiled scoring kernel ([7]). */
int rpe_last_stage_ms(rpe_ctx* ctx, float ms[8]);
/* With stage timing on (1 or 2): sum and count of the tiled scoring kernel's CUDA-event durations over EVERY launch of
 * this context since the last reset, complete after rpe_sync (each launch gets its own event pair on the stream the
 * kernel runs on; nothing is synchronised to collect them). The roofline's `kernel_ms` over a timed region. */
int rpe_scorer_time_stats(rpe_ctx* ctx, double* sum_ms, long long* count, int reset);
/* Device-wide companion (all contexts of `device`, launches timed with stage timing on): the scorers of consecutive
 * frames are launched into two alternating "lane" streams, so the next launch takes over each SM the moment the previous
 * launch's CTA leaves it and the event pairs of neighbouring launches overlap. This call returns the length of the UNION
 * of the launches' [start, end] event intervals (start = inputs ready and lane free) and the number of launches: sum / count is the device time one launch costs inside a pipelined region. Complete after the contexts
 * have been synchronised. */
int rpe_scorer_busy_stats(int device, double* sum_ms, long long* count, int reset);
int rpe_enable_stage_timing(rpe_ctx* ctx, int enable);
/* How the inlier matrix of the ASYNCHRONOUS calls (rpe_ransac_async, rpe_ransac_sharded_async, the sequence runner) reaches
 * the caller's `mask` buffer. 0 (default): the n x cols matrix of 16-bit flags is copied as it is (the reference's
 * setInlier layout, PnPPoseAdapter.hpp:196-237: 2 bytes per flag, 1.2 MB for a dense 3-D / 3-D frame). 1: the device
 * sends one BIT per flag (77 KB) into the tail of the same buffer and the thread that collects the result (rpe_sync, or
 * the next call once 256 results are in flight) expands it in place — the buffer holds exactly the same matrix afterwards,
 * with 1/16 of the device-to-host bytes on the bus, for 0.04-0.1 ms of host work per dense frame (measured, round 2: on
 * the 8-GPU box the concurrent-upload ceiling rises from 190 to 226 GB/s, but with 4 host cores per GPU the expansion
 * costs the issuing threads more than that — 21.1 k against 22.7 k frames/s — so this is an opt-in for hosts with
 * cores to spare). Blocking calls always copy the matrix.
 * mode 2: column 0 of a family without the 2-D test (RPE_SHINJI, RPE_NL_SHINJI) is a constant — 0 once a hypothesis has
 * been accepted, the adapters' initial 1 otherwise (AOPoseAdapter.hpp: setInlier is fed a matrix whose 2-D column the loop
 * never touches) — so it stays on the device and the collecting thread writes it (one std::fill of n shorts): half the
 * device-to-host bytes of a dense 3-D / 3-D frame, blocking and asynchronous calls alike; the buffer holds the same matrix.
 * Measured on the 8-GPU box: the concurrent-upload ceiling rises from 188 to 207 GB/s, the frame rate does not (22.4 k
 * against 22.6 k frames/s: with 4 host cores per GPU the fill costs the issuing threads what the bus gains) — an opt-in too. */
int rpe_set_mask_transfer(rpe_ctx* ctx, int mode);

#ifdef __cplusplus
}
#endif

#endif /* RPE_C_API_H_ */
