// ltr_p2p.cuh -- all-reduce over NVLink peer memory for the ranks of one node (ltr_p2p_* of ltr_sm100.h).
//
// Every rank owns a mailbox in device memory, exported to the other ranks by CUDA IPC.  An exchange writes
// the rank's values straight into every peer's mailbox (64-bit stores carrying {sequence number, float
// bits}: data and flag arrive together, as in NCCL's LL protocol), waits until its own mailbox holds the
// current sequence number from every rank and sums the values in rank order (the same bits on every rank).
// The sequence numbers live on the device and are advanced by the kernels themselves, so a launch can be
// captured in a CUDA graph and replayed; two slot sets alternate, because a rank can run at most one
// exchange ahead of the slowest one.  A rank that never shows up trips a ~2 s timeout and raises the
// mailbox's error flag instead of hanging the GPU.
//   scalars: slot[2][rank][4], one single-CTA kernel (the sum of losses behind a global mean)
//   vectors: vslot[2][rank][capacity] behind the struct in the same allocation, any grid; used inline by the
//            kernel that produces the vector (mlp_reduce_kernel: the gradient of the MLP ranker leaves the
//            reduction already summed over the ranks) or by p2p_allreduce_vec_kernel on its own
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ltr {

constexpr int kP2PMaxRanks = 16;
constexpr int kP2PMaxValues = 4;
constexpr int kP2PVecCapacity = 1 << 16;       // floats per vector exchange (longer vectors go in pieces)
constexpr long long kP2PTimeoutCycles = 4000000000LL;

struct P2PMailbox {
  unsigned long long slot[2][kP2PMaxRanks][kP2PMaxValues];   // {sequence << 32 | float bits}
  P2PMailbox* peers[kP2PMaxRanks];                            // this rank's view of every rank's mailbox
  unsigned int seq;
  unsigned int error;
  unsigned int vseq;      // sequence number of the last completed vector exchange
  unsigned int vdone;     // CTAs of the running vector exchange that have finished
  unsigned int world;
  unsigned int pad[3];
  // followed by unsigned long long vslot[2][world][kP2PVecCapacity]
};

__host__ __device__ inline size_t p2p_mailbox_bytes(int world) {
  return sizeof(P2PMailbox) + 2u * static_cast<size_t>(world) * kP2PVecCapacity * sizeof(unsigned long long);
}

__device__ __forceinline__ volatile unsigned long long* p2p_vslot(P2PMailbox* box, int world, int par, int rank,
                                                                  int j) {
  unsigned long long* base = reinterpret_cast<unsigned long long*>(box + 1);
  return base + (static_cast<size_t>(par) * world + rank) * kP2PVecCapacity + j;
}

// sequence number of the vector exchange this launch performs (every CTA reads the same value: it only changes
// when the last CTA of the launch leaves through p2p_vec_finish)
__device__ __forceinline__ unsigned int p2p_vec_begin(P2PMailbox* mine) {
  return *reinterpret_cast<volatile unsigned int*>(&mine->vseq) + 1u;
}

// element j (< kP2PVecCapacity) of this rank's vector: value v out to every rank, the sum over the ranks back
__device__ __forceinline__ float p2p_vec_element(P2PMailbox* mine, int rank, int world, unsigned int seq, int j,
                                                 float v) {
  const int par = static_cast<int>(seq & 1u);
  const unsigned long long packed =
      (static_cast<unsigned long long>(seq) << 32) | static_cast<unsigned long long>(__float_as_uint(v));
  for (int peer = 0; peer < world; ++peer) *p2p_vslot(mine->peers[peer], world, par, rank, j) = packed;
  float s = 0.0f;
  const long long t0 = clock64();
  for (int r = 0; r < world; ++r) {
    const volatile unsigned long long* src = p2p_vslot(mine, world, par, r, j);
    unsigned long long w = *src;
    while (static_cast<unsigned int>(w >> 32) != seq) {
      if (clock64() - t0 > kP2PTimeoutCycles) {
        mine->error = 1u;
        break;
      }
      w = *src;
    }
    s += __uint_as_float(static_cast<unsigned int>(w & 0xffffffffull));   // rank order: the same bits on every rank
  }
  return s;
}

// every CTA calls this once, after all its threads have finished their elements (call from all threads)
__device__ __forceinline__ void p2p_vec_finish(P2PMailbox* mine, unsigned int seq) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&mine->vdone, 1u) == gridDim.x - 1) {
      mine->vdone = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned int*>(&mine->vseq) = seq;
    }
  }
}

}  // namespace ltr

// the handle behind ltr_p2p_create (host side)
struct ltr_p2p {
  int rank = 0, world = 1, device = 0;
  ltr::P2PMailbox* mine = nullptr;
  void* opened[ltr::kP2PMaxRanks] = {};
};
