// peer.cu — the one exchange step of the hypothesis-sharded single-frame mode, over peer memory.
//
// Every rank scores its slice of the vote table; the slices are then written straight into every peer's table through
// NVLink (peer mappings obtained with CUDA IPC, one process per GPU) by ONE single-CTA kernel per rank that
//   1. stores its slice into table[epoch & 1] of every rank (itself included),
//   2. __threadfence_system(), then publishes `epoch` in its slot of every rank's flag array,
//   3. waits (bounded) until every rank's flag in its OWN array has reached `epoch`,
//   4. copies the complete table into the context's vote table, where the replay kernel reads it.
// Two tables alternate by epoch: a fast rank may already write epoch e+1 while a slow one still reads epoch e, and it
// cannot reach e+2 before the slow rank has published e+1, i.e. has finished reading e (stream order).
// No NCCL call, no host round trip: generate -> score -> exchange -> replay -> mask is one asynchronous stream.
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace rpe {

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256)
exchange_votes_kernel(PeerTable peers, int rank, int world, unsigned int epoch, int slot_begin, int slot_end, int n_slots,
                      int32_t* __restrict__ votes, unsigned int* __restrict__ host_err, unsigned long long timeout_ns) {
  const int par = (int)(epoch & 1u);
  // 1. my slice -> every rank's table
  for (int r = 0; r < world; ++r) {
    int32_t* dst = peer_table(peers.block[r], par);
    for (int i = slot_begin + threadIdx.x; i < slot_end; i += blockDim.x) dst[i] = votes[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish
  if (threadIdx.x < world) {
    volatile unsigned int* f = peer_flags(peers.block[threadIdx.x]);
    f[rank] = epoch;
  }
  __threadfence_system();
  // 3. wait for everybody (bounded: a missing peer must not hang the GPU)
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    volatile unsigned int* mine = peer_flags(peers.block[rank]);
    const unsigned long long t0 = global_ns();
    while ((int)(mine[threadIdx.x] - epoch) < 0) {
      if (global_ns() - t0 > timeout_ns) {
        ok = 0;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
  // 4. the complete table -> the context's vote table (all empty on a timeout: no winner, and the error is latched)
  const int32_t* src = peer_table(peers.block[rank], par);
  for (int i = threadIdx.x; i < n_slots; i += blockDim.x) votes[i] = ok ? ((volatile const int32_t*)src)[i] : -1;
  if (!ok && threadIdx.x == 0) {
    peer_flags(peers.block[rank])[kPeerErrSlot] = 1u;
    if (host_err) {  // page-locked host word: what rpe_sync / the blocking calls look at (RPE_ERR_COMM)
      *(volatile unsigned int*)host_err = 1u;
      __threadfence_system();
    }
  }
}

void launch_exchange_votes(const PeerTable& peers, int rank, int world, unsigned int epoch, int slot_begin, int slot_end,
                           int n_slots, int32_t* votes, unsigned int* host_err, unsigned long long timeout_ns, cudaStream_t s) {
  exchange_votes_kernel<<<1, 256, 0, s>>>(peers, rank, world, epoch, slot_begin, slot_end, n_slots, votes, host_err,
                                          timeout_ns);
}

}  // namespace rpe
