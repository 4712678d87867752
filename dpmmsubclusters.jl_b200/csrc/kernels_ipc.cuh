// aggregate_suff_stats across workers (src/priors/niw.jl:64-66; the worker -> leader -> master reduce of
// src/local_clusters_actions.jl:171-203, 229-251) as ONE kernel over NVLink peer memory.
//
// Every rank owns an exchange region [flags: 64 x u32 | pad to 1 KB | buffer 0 | buffer 1] that all peers
// map through CUDA IPC.  A statistics call with epoch e:
//   1. stats_finalize_kernel packs the rank's statistics straight into its buffer (e & 1);
//   2. ipc_allreduce_kernel: block 0 publishes "epoch e is complete on rank r" into slot r of EVERY peer's flag
//      array (system-scope fence + stores over NVLink); all blocks wait until the local flag array shows
//      epoch e for every rank; then every rank sums the peers' buffers (e & 1) in rank order -- the same
//      order everywhere, so all ranks hold bit-identical sums -- into its private result buffer.
// Two buffers are enough: a peer publishes epoch e + 1 only after its reduce of epoch e retired (stream
// order), and nobody writes buffer (e & 1) again before it has seen every peer's epoch e + 1.
// The loads of peer memory are volatile (no L1 caching of remote lines); the transfer is n * 8 * (world - 1)
// bytes per rank (C2: 3.5 MB at 8 ranks), i.e. a few microseconds at NVLink 5 rates, against ~18-37 us for
// the NCCL ring / tree at this size.
#pragma once
#include <cstdint>

#define IPC_FLAG_BYTES 1024
#define IPC_MAX_WORLD 16

struct IpcReduceArgs {
  int world, rank;
  uint32_t epoch;
  size_t n;                                  // doubles
  const double* src[IPC_MAX_WORLD];          // every rank's buffer (epoch & 1), own included
  uint32_t* peer_flags[IPC_MAX_WORLD];       // every rank's flag array
  const uint32_t* my_flags;
  double* out;
};

__global__ void __launch_bounds__(256) ipc_allreduce_kernel(const IpcReduceArgs a) {
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();   // the finalise kernel's writes (previous launch on this stream) before the flag
    volatile uint32_t* f = a.peer_flags[threadIdx.x] + a.rank;
    *f = a.epoch;
  }
  if (threadIdx.x < a.world) {
    const volatile uint32_t* f = a.my_flags + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(*f - a.epoch) < 0) {
      // a peer that never arrives (a failed launch on its side) must not hang this GPU: give up after ~4 s; the
      // sums are then incomplete and the statistics call's consistency checks (counts vs N) fail loudly
      if (clock64() - t0 > 8000000000LL) break;
    }
    __threadfence_system();
  }
  __syncthreads();
  // two doubles per thread and rank, every rank's load in flight before the first add (one NVLink round trip);
  // .cg: no L1 allocation of remote lines
  const size_t n2 = a.n >> 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 v[IPC_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < IPC_MAX_WORLD; ++r)
      if (r < a.world) v[r] = __ldcg(reinterpret_cast<const double2*>(a.src[r]) + i);
    double2 s = make_double2(0.0, 0.0);
#pragma unroll
    for (int r = 0; r < IPC_MAX_WORLD; ++r)
      if (r < a.world) {
        s.x += v[r].x;
        s.y += v[r].y;
      }
    reinterpret_cast<double2*>(a.out)[i] = s;
  }
  if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int r = 0; r < a.world; ++r) s += __ldcg(a.src[r] + a.n - 1);
    a.out[a.n - 1] = s;
  }
}
