// srb_kernels_peer.cuh -- multi-GPU peer path: the kernels that run on NVLink peer
// memory (CUDA-IPC mappings of the other ranks' buffers).  The reduce-scatter half is done by copy
// engines (srb_api.cu: srb_peer_scatter_dev); these kernels do the fixed-order sum of a rank's band,
// the all-gather stores and the flag barriers.
#pragma once
#include "srb_common.cuh"

namespace srb {

constexpr int SRB_MAX_PEERS = 8;

// Multi-GPU reduce + all-gather of one rank's band (after every rank has scattered its partial rows
// into this rank's slots): out_r[band] = sum_s slots[s][band] in fixed slot order (deterministic),
// written to the gradient buffer of EVERY rank (peer stores over NVLink).
struct GatherParams {
  int world, rank;
  long long band_begin, band_len, band_cap;
  const double* slots;           // this rank's slot array [world][band_cap]; slot [rank] is unused:
  const double* own;             // ... this rank's own contribution is its local partial gradient band
  double* out[SRB_MAX_PEERS];    // gradient buffers of all ranks (peer mappings)
};
// Multi-GPU reduce + all-gather of this rank's band: out_r[band] = sum_s partial_s[band] in fixed
// rank order (deterministic), stored into the gradient buffer of EVERY rank (peer stores over
// NVLink).  Every block first waits (bounded spin on the local phase-0 flags) until all ranks'
// contributions have arrived; the last block to finish publishes the total cost locally and raises
// this rank's phase-1 flag on every rank after a system fence.  A block whose wait times out records the
// failure in *err and neither sums nor stores nor raises the flag: nothing partial is ever published, the
// other ranks time out in turn, and the host reports SRB_ERR_STATE at the next synchronisation point
// (peer_status in srb_api.cu).
__global__ void __launch_bounds__(256)
k_sum_gather(GatherParams G, long long n, long long flag_base, unsigned long long epoch, unsigned int* done_counter,
             int* err) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if (threadIdx.x < G.world) {
    const volatile unsigned long long* f =
        reinterpret_cast<const volatile unsigned long long*>(G.out[G.rank] + flag_base) + threadIdx.x;
    unsigned long long spins = 0;
    while (*f < epoch)
      if (++spins > (1ull << 24)) {  // seconds: a rank is missing -- give up instead of hanging the GPU
        *err = 1;
        timed_out = 1;
        break;
      }
    __threadfence_system();
  }
  __syncthreads();
  if (timed_out || *reinterpret_cast<volatile int*>(err) != 0) return;  // block-uniform
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (((G.band_len | G.band_begin | G.band_cap) & 1) == 0) {
    // 16-byte accesses: 512 B per warp and store instruction on the link
    const long long len2 = G.band_len >> 1, cap2 = G.band_cap >> 1, first2 = G.band_begin >> 1;
    const double2* __restrict__ slots2 = reinterpret_cast<const double2*>(G.slots);
    const double2* __restrict__ own2 = reinterpret_cast<const double2*>(G.own);
    for (long long i = i0; i < len2; i += stride) {
      double2 acc = make_double2(0.0, 0.0);
      for (int s = 0; s < G.world; ++s) {
        const double2 v = s == G.rank ? own2[i] : slots2[(long long)s * cap2 + i];
        acc.x += v.x;
        acc.y += v.y;
      }
      for (int r = 0; r < G.world; ++r) reinterpret_cast<double2*>(G.out[r])[first2 + i] = acc;
    }
  } else {
    for (long long i = i0; i < G.band_len; i += stride) {
      double acc = 0.0;
      for (int s = 0; s < G.world; ++s) acc += s == G.rank ? G.own[i] : G.slots[(long long)s * G.band_cap + i];
      for (int r = 0; r < G.world; ++r) G.out[r][G.band_begin + i] = acc;
    }
  }
  // last block done: total cost + phase-1 flags
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done_counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int r = 0; r < G.world; ++r) acc += G.out[G.rank][n + 1 + r];
      G.out[G.rank][n] = acc;
      *done_counter = 0;
    }
    __syncthreads();  // the cost slots are read before any peer may start overwriting them for the next epoch
    if (threadIdx.x < G.world) {
      __threadfence_system();
      volatile unsigned long long* f =
          reinterpret_cast<volatile unsigned long long*>(G.out[threadIdx.x] + flag_base) + G.world + G.rank;
      *f = epoch;
    }
  }
}

// End of a rank's scatter phase, one block: cost = fixed-order sum of the per-CTA partials (as
// k_finish_partials), posted into slot [rank] of every rank's cost array, then -- after a system
// fence -- the phase-0 flag of this rank is raised on every rank.  Runs after k_tile on the same
// stream, i.e. after all of this rank's gradient rows have been stored to their owners.
__global__ void __launch_bounds__(1024)
k_peer_finish_scatter(const double* __restrict__ pd, size_t nd, const double* __restrict__ pr, size_t nr,
                      double* __restrict__ cost, GatherParams G, long long cost_slot_base, long long flag_base,
                      unsigned long long epoch) {
  double a = 0.0, b = 0.0;
  for (size_t i = threadIdx.x; i < nd; i += blockDim.x) a += pd[i];
  for (size_t i = threadIdx.x; i < nr; i += blockDim.x) b += pr[i];
  a = block_sum(a);
  b = block_sum(b);
  __shared__ double total;
  if (threadIdx.x == 0) {
    cost[0] = a;
    cost[1] = b;
    cost[2] = total = a + b;
  }
  __syncthreads();
  if (threadIdx.x < G.world) {
    G.out[threadIdx.x][cost_slot_base + G.rank] = total;
    __threadfence_system();
    volatile unsigned long long* f =
        reinterpret_cast<volatile unsigned long long*>(G.out[threadIdx.x] + flag_base) + G.rank;
    *f = epoch;
  }
}

// Device-side barrier between the ranks, on flags that live behind every rank's gradient buffer:
// k_peer_signal (after this rank's kernels of the phase, same stream) publishes `epoch` into slot
// [phase][rank] of every rank; k_peer_wait spins until all ranks have published it.  The spin is
// bounded: on timeout it records the failure in *err and returns instead of hanging the GPU.
__global__ void k_peer_wait(const double* out_local, long long flag_base, int phase, int world,
                            unsigned long long epoch, int* err) {
  if (threadIdx.x < world) {
    const volatile unsigned long long* f =
        reinterpret_cast<const volatile unsigned long long*>(out_local + flag_base) + phase * world + threadIdx.x;
    unsigned long long spins = 0;
    while (*f < epoch) {
      if (++spins > (1ull << 24)) {  // seconds: a rank is missing
        *err = 1;
        break;
      }
    }
    __threadfence_system();
  }
}

}  // namespace srb
