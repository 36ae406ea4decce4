// kernels.cuh — launch interface between the C-ABI layer (capi.cu) and the kernel TUs.
#ifndef RPE_KERNELS_CUH_
#define RPE_KERNELS_CUH_

#include "rpe_device.cuh"

namespace rpe {

// Device view of one frame's correspondences. Raw arrays are the caller's column-major 3 x n
// layout (n contiguous xyz triples); `pk_*` are the pair-interleaved copies the tiled scorer streams
// through shared memory.
struct FrameView {
  const float* bv;
  const float* xc;
  const float* nc;
  const float* xw;
  const float* nw;
  int n;
  int npairs_pad;      // pairs, padded to a multiple of kSubPairs with NaN correspondences
  const float4* pk;    // packed pair records, `pk_f4_per_pair` float4 per pair
  int pk_f4_per_pair;  // 3 (AO) ...
  int pk_kind;         // which modalities are packed (bit0 2-D, bit1 3-D, bit2 normal)
  unsigned raw_aligned;  // bit k: array k (bv, xc, nc, xw, nw) is present, 16-byte aligned and may be streamed by bulk TMA;
                         // 0 when raw streaming is switched off (test hook): the tiled scorers then need the packed copy
};
// can the tiled scorer of modality set `kind` (bit0 2-D, bit1 3-D, bit2 normal) stream the caller's arrays themselves?
inline bool frame_raw_ok(const FrameView& f, int kind) {
  unsigned need = 1u << 3;                  // x_w
  if (kind & 1) need |= 1u << 0;            // b
  if (kind & 2) need |= 1u << 1;            // x_c
  if (kind & 4) need |= (1u << 2) | (1u << 4) | (1u << 1);  // n_c, n_w and x_c (isValid gate)
  return (f.raw_aligned & need) == need;
}

struct Thresh {
  float thr3d, cos_thr, cos_nl;
};

// Borderline evaluations queued for the exact fix-up. The list is segmented: CTA k of a scoring launch (k =
// blockIdx.y * gridDim.x + blockIdx.x, nseg CTAs in all) owns entries [k * seg, (k + 1) * seg) with seg = capacity / nseg,
// counts them in shared memory and publishes counts[k] when it ends — no global same-address atomics in the scorer.
struct Worklist {
  uint2* entries;        // (slot, modality << 30 | correspondence)
  unsigned int* counts;  // [kMaxWorklistSegments]
  unsigned int capacity;
};
constexpr int kMaxWorklistSegments = 8192;

constexpr int kSubPairs = 8;       // correspondences are rescanned in groups of 16
constexpr int kTilePairs = 256;    // pairs per shared-memory stage
constexpr int kScoreThreads = 256; // threads per scoring CTA
constexpr int kHypPerThread = 2;   // hypotheses held in registers by one thread

// -- prep ---------------------------------------------------------------------------------------
void launch_reset_stats(FrameStats* st, cudaStream_t s);
void launch_pack(const FrameView& f, int kind, float4* pk_out, FrameStats* st, cudaStream_t s);

void set_kabsch_polar(int on);  // debug: 0 = the Kabsch refits always take the Jacobi SVD (per device, synchronous)

// -- generation ----------------------------------------------------------------------------------
void launch_hypgen(int method, const FrameView& f, const int32_t* samples_dev, int H, HypGen* gen, HypFast* fast,
                   int32_t* votes, FrameStats* st, cudaStream_t s, const int32_t* stale_eff = nullptr);
// opt-in reproduction of the reference's stale sample columns (see hypgen_kernel): eff = [H x 2] int32
void launch_stale_cols(const int32_t* samples_dev, int H, const float* xc, const double* xc64, int n, int32_t* carry,
                       bool reset_carry, int32_t* eff, cudaStream_t s);
void launch_derive_fast(const HypGen* gen, HypFast* fast, int32_t* votes, int n_slots, FrameStats* st, cudaStream_t s);

// MinimalSolvers.hpp batches (one problem per thread); device pointers
void launch_minsolv_ev(const float* M9, int count, float* E3, cudaStream_t s);
void launch_minsolv_ms(const float* in24, int count, float* q4, float* t3, cudaStream_t s);

// -- scoring -------------------------------------------------------------------------------------
// corr_base / seg_cap (3-D raw-array scorer only): a launch may score ONE CHUNK of a frame that is still being uploaded
// — f then views the chunk, corr_base is the frame index of its first correspondence and seg_cap fixes the size of a
// worklist segment so that the launches of all chunks fill consecutive parts of one list (wl points at this chunk's part).
int launch_score_fast(int method, const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin,
                       int slot_end, Thresh th, int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s,
                       int corr_base = 0, unsigned int seg_cap = 0, int ur_lane = -1);
// The correspondence-stationary 3-D / 3-D scorer (score_ur.cu: hypotheses in uniform registers, fed from a __constant__
// buffer that belongs to scorer lane 0 / 1). ur_lane of launch_score_fast: the lane stream `s` is, -1 if `s` is no lane
// (the buffer of a lane must not be rewritten while another launch that reads it may still run: stream order does that
// for a lane's own stream only). Returns 0 if the frame / slot range is not for this kernel.
int ur_shape_for(int npairs_pad, int nslots, int num_sms, int* out6);  // host logic of the launcher (tests)
int launch_score3d_ur_lane0(const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin, int slot_end, float thr3d,
                            int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s);
int launch_score3d_ur_lane1(const FrameView& f, const HypGen* gen, const HypFast* fast, int slot_begin, int slot_end, float thr3d,
                            int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s);
// The 3-D / 3-D scorer fed straight from page-locked host arrays (fsrc.xw / fsrc.xc as the device sees them); the frame is
// written to dxw / dxc on the way (may be null). Returns the number of worklist segments (= CTAs).
int launch_score3d_stream(const FrameView& fsrc, float* dxw, float* dxc, const HypGen* gen, const HypFast* fast, int slot_begin,
                          int slot_end, float thr3d, int32_t* votes, FrameStats* st, Worklist wl, int num_sms, cudaStream_t s);

void launch_fixup(int method, const FrameView& f, const HypGen* gen, Thresh th, int32_t* votes, FrameStats* st,
                  Worklist wl, int nseg, int slot_begin, int slot_end, cudaStream_t s);
void launch_consume_worklist(FrameStats* st, cudaStream_t s);
// exact-order scoring of every (slot, correspondence); `only_if_overflow` makes it a no-op unless
// the fast pass overflowed its worklist.
void launch_score_exact(int method, const FrameView& f, const HypGen* gen, int slot_begin, int slot_end, Thresh th,
                        int32_t* votes, FrameStats* st, bool only_if_overflow, int num_sms, cudaStream_t s);

// -- replay / mask / refit --------------------------------------------------------------------------
// Replay of the sequential rule over the iterations [iter_base, iter_base+H) held in gen/votes; the state is
// carried in `rs`; begin_iter_max >= 0 (the frame's first pass) resets it to the loop's initial state with Iter =
// begin_iter_max. The kernel also clears the per-pass FrameStats counters.
void launch_replay(int method, const HypGen* gen, const int32_t* votes, int H, int iter_base, int n, float confidence,
                   FrameStats* st, ReplayState* rs, ReplayOut* out, bool finalize, int begin_iter_max, cudaStream_t s);

struct RefitBuffers {
  double* partials = nullptr;   // [blocks x kMomentCount]
  double* moments = nullptr;    // [kMomentCount] reduced
  double* suff = nullptr;       // [kMomentCount] sufficient statistics of the current inlier columns (suff_add_3d / suff_add_nl)
  int max_blocks = 0;
  int num_sms = 148;        // refit kernels use at most 2 CTAs per SM and stride over the correspondences
};
constexpr int kMomentCount = 48;  // 16 Kabsch moments / 29 LM entries / 40 nl_shinji_kneip_ls sums, padded

// pose_rw: the ReplayOut written by the replay kernel (read for the pose, per-column inlier counts are
// added to it); kabsch_out receives pose_rw with (q,t) replaced by the Kabsch refit over the 3-D inliers.
void launch_mask(int method, const FrameView& f, ReplayOut* pose_rw, Thresh th, int16_t* mask, ReplayOut* kabsch_out,
                 RefitBuffers rb, FrameStats* st, cudaStream_t s, uint32_t* bits = nullptr);
void launch_reset_corr_bound(FrameStats* st, cudaStream_t s);
// Kabsch from the moments left by launch_mask (or by launch_kabsch_moments); writes pose_out.
// returns the number of CTAs launched (= rows of rb.partials to reduce)
int launch_kabsch_moments(const FrameView& f, const int16_t* flags3d /*null: all points*/, RefitBuffers rb,
                          FrameStats* st, cudaStream_t s);
// statistics of explicit flag columns -> rb.partials; returns the CTAs launched
int launch_suffstats(const FrameView& f, const int16_t* flags3d, bool all3d, const int16_t* flagsN, RefitBuffers rb,
                     cudaStream_t s);
void launch_kabsch_solve(RefitBuffers rb, int blocks_used, ReplayOut* pose_inout, int32_t* refit_ok, cudaStream_t s);

struct GnState {  // device-resident LM state (mirrors oracle/refine.hpp refine_gn)
  double Rp[9], tp[3];  // proposal
  double Ra[9], ta[3];  // accepted
  double H[21], g[6];   // accepted normal equations
  double cost_acc;
  double mu;
  long long rows;
  int have, done, evals, accepted;
};
void launch_gn_init(const ReplayOut* pose, GnState* st, cudaStream_t s);
// One LM evaluation; its last CTA solves, updates the state and refreshes pose_out/cost_out/evals_out with
// the best pose accepted so far.
void launch_gn_iteration(const FrameView& f, const int16_t* mask, int mask_cols, float w2d, float w3d, float wnl,
                         RefitBuffers rb, GnState* gs, FrameStats* st, ReplayOut* pose_out, double* cost_out,
                         int32_t* evals_out, cudaStream_t s);

// The complete LM loop for 3-D / normal rows from the sufficient statistics (blocks_used > 0: reduce rb.partials
// first; 0: use rb.suff as left by the mask kernel). One launch, no pass over the correspondences.
void launch_gn_from_stats(RefitBuffers rb, int blocks_used, float w3d, float wnl, int max_iters, ReplayOut* pose_io,
                          GnState* gs, double* cost_out, int32_t* evals_out, cudaStream_t s);

// -- peer-memory vote exchange (peer.cu) --------------------------------------------------------------
// One block per rank, exported through CUDA IPC: 64 flag words (slot r = the epoch rank r has published here, slot
// kPeerErrSlot = latched time-out) followed by two vote tables of kPeerSlots int32.
constexpr int kMaxPeers = 16, kPeerErrSlot = 63, kPeerFlagWords = 64, kPeerSlots = 8192 * 3;
constexpr size_t kPeerBlockBytes = kPeerFlagWords * sizeof(unsigned int) + 2 * (size_t)kPeerSlots * sizeof(int32_t);
struct PeerTable {
  unsigned char* block[kMaxPeers];  // this process' mappings of every rank's block (own block included)
};
__host__ __device__ inline unsigned int* peer_flags(unsigned char* block) { return reinterpret_cast<unsigned int*>(block); }
__host__ __device__ inline int32_t* peer_table(unsigned char* block, int parity) {
  return reinterpret_cast<int32_t*>(block + kPeerFlagWords * sizeof(unsigned int)) + (size_t)parity * kPeerSlots;
}
void launch_exchange_votes(const PeerTable& peers, int rank, int world, unsigned int epoch, int slot_begin, int slot_end,
                           int n_slots, int32_t* votes, unsigned int* host_err, unsigned long long timeout_ns, cudaStream_t s);

// -- device-side Simulator (simulate.cu) ------------------------------------------------------------
struct SimParams {
  float R[9];  // R_cw row-major
  float t[3];
  int n;
  float noise2d, noise3d, noise_nl;
  int out2d, out3d, outnl;  // number of outliers per modality
  // affine permutations i -> (a*j + b) mod n ; membership: j = ainv*(i - b) mod n < out
  unsigned int ainv[3], b[3];
  float min_depth, max_depth, f;
  int gaussian;
  int kinect;  // camera points perturbed by the Kinect lateral / axial model instead of isotropic noise3d
  unsigned long long seed;
};
void launch_simulate(const SimParams& p, float* xw, float* xc, float* bv, float* nw, float* nc, int mode_3d3d, cudaStream_t s);

// -- binary64 RANSAC path (f64.cu) ------------------------------------------------------------------
struct FrameView64 {
  const double* bv;
  const double* xc;
  const double* nc;
  const double* xw;
  const double* nw;
  int n;
};
struct Thresh64 {
  double thr3d, cos_thr, cos_nl;
};
struct HypGen64 {
  double q[4];
  double t[3];
  int32_t valid;
  int32_t pad;
};
struct Pose64 {
  double q[4];
  double t[3];
};
struct ReplayState64 {
  int32_t best, iter, win, cur_iter, stop, slots_done, borderline, overflow;
  double q[4];
  double t[3];
};
void launch_f64_to_f32(const double* src, float* dst, size_t count, cudaStream_t s);
void launch_hypgen64(int method, const FrameView64& f, const int32_t* samples_dev, int H, HypGen64* gen, int32_t* votes,
                     cudaStream_t s, const int32_t* stale_eff = nullptr);
void launch_score64(int method, const FrameView64& f, const HypGen64* gen, int n_slots, Thresh64 th, int32_t* votes,
                    int num_sms, const FrameStats* only_if_overflow, cudaStream_t s);
void launch_derive_fast64(const HypGen64* g64, HypGen* gen, HypFast* fast, int n_slots, cudaStream_t s);
void launch_fixup64(int method, const FrameView64& f, const HypGen64* gen, Thresh64 th, int32_t* votes, FrameStats* st,
                    Worklist wl, int nseg, cudaStream_t s);
void launch_replay64_begin(ReplayState64* rs, int iter_max, cudaStream_t s);
void launch_replay64(int method, const HypGen64* gen, const int32_t* votes, int H, int iter_base, int n, double confidence,
                     FrameStats* st, ReplayState64* rs, ReplayOut* out, Pose64* out64, bool finalize, cudaStream_t s);
void launch_mask64(int method, const FrameView64& f, ReplayOut* pose_rw, const Pose64* pose64, Thresh64 th, int16_t* mask,
                   int num_sms, cudaStream_t s);

// nl_shinji_kneip_ls (AbsoluteOrientationNormal.hpp:447-552) as one pose-independent reduction pass plus three
// passes for the only sum that depends on the running camera centre (M23); the 3x3 SVDs, find_opt_cc's
// ray-intersection solve and the blending run in the last CTA of each pass, in binary64.
struct NlskState {
  double Cw[3], Cc[3], TV;
  double S33[9], sig, SNN[9], tl, tw;
  double M23[9], M33[9], MNN[9], TW, TL;
  double cp[3], c_opt[3], q_opt[4];
  double q0[4], t0[3];
  int N, k23, mnn, K, M, cp_ok, stopped, iter;
};
void launch_nlsk_prepass(const FrameView& f, const int16_t* mask3, const float* weights3, const ReplayOut* pose,
                         RefitBuffers rb, NlskState* ns, FrameStats* st, cudaStream_t s);
void launch_nlsk_iteration(const FrameView& f, const int16_t* mask3, const float* weights3, RefitBuffers rb, NlskState* ns,
                           FrameStats* st, ReplayOut* pose_out, cudaStream_t s);

// -- microbenchmark -----------------------------------------------------------------------------------
void launch_ffma_bench(float* sink, int iters, bool packed, int blocks, cudaStream_t s);

int grid_blocks_for(int n, int threads);

}  // namespace rpe

#endif  // RPE_KERNELS_CUH_
