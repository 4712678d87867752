// Stage 1+2 for the Dirichlet-multinomial prior.
//
//   log_likelihood!(r, x, ::multinomial_dist)   src/distributions/multinomial_dist.jl:13-15
//       r_j = sum_d alpha_d * x[d,j]   (alpha = log-probabilities, no multinomial coefficient)
//   sample_labels_worker! / create_subclusters_labels!  (as in kernels_gauss.cuh)
//
// The n x K product is a skinny GEMM (counts x log-probabilities) that is HBM-bound once X is
// streamed once: every thread owns one point, keeps KT running dot products in registers and reads
// the log-probability table from shared memory as float4 broadcasts.  D and K are runtime values.
#pragma once
#include "common.cuh"
#include "kernels_gauss.cuh"  // SubLabelArgs, sublabel_partition

struct MnmLabelArgs {
  const float* x;       // [n][D] counts
  int64_t n;
  int D;
  int DS;               // shared-memory row stride (odd)
  int K;
  int KP;               // K rounded up to a multiple of MNM_KT
  const float* logp_t;  // [D][KP] cluster log-probabilities, transposed + zero padded
  const float* logw;    // [K]
  int32_t* labels;
  int32_t* hist;
  const double* u_inj;
  uint64_t seed;
  uint32_t call;
  int64_t goff;
  int final_iter;
  int sampler;
  float* dump;
  int64_t ntiles;
};

#define MNM_KT 16

// K3: one thread = one point, tile = blockDim.x consecutive points.
//   xs [T][DS]  tile,   as [D][KP] log-probability table,   rs [K][T] slice of parr,   hs [K]
__global__ void mnm_label_kernel(const MnmLabelArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  float* as = smem;
  float* xs = as + (size_t)a.D * a.KP;
  float* rs = xs + (size_t)T * a.DS;
  int* hs = reinterpret_cast<int*>(rs + (size_t)a.K * T);
  for (int e = tid; e < a.D * a.KP / 4; e += T)
    reinterpret_cast<float4*>(as)[e] = __ldg(reinterpret_cast<const float4*>(a.logp_t) + e);
  for (int k = tid; k < a.K; k += T) hs[k] = 0;

  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t base = tile * T;
    const int npts = (int)min((int64_t)T, a.n - base);
    __syncthreads();
    {
      const float* src = a.x + base * a.D;
      const int nv = npts * a.D;
      for (int e = tid; e < T * a.D; e += T) {
        const int p = e / a.D, c = e - p * a.D;
        xs[(size_t)p * a.DS + c] = (e < nv) ? __ldg(src + e) : 0.f;
      }
    }
    __syncthreads();
    const float* xr = xs + (size_t)tid * a.DS;
    for (int k0 = 0; k0 < a.KP; k0 += MNM_KT) {
      float acc[MNM_KT];
#pragma unroll
      for (int j = 0; j < MNM_KT; ++j) acc[j] = 0.f;
      for (int d = 0; d < a.D; ++d) {
        const float xv = xr[d];
        const float4* ar = reinterpret_cast<const float4*>(as + (size_t)d * a.KP + k0);
#pragma unroll
        for (int j4 = 0; j4 < MNM_KT / 4; ++j4) {
          const float4 w = ar[j4];
          acc[4 * j4 + 0] = fmaf(w.x, xv, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(w.y, xv, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(w.z, xv, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(w.w, xv, acc[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < MNM_KT; ++j)
        if (k0 + j < a.K) rs[(size_t)(k0 + j) * T + tid] = __fadd_rn(acc[j], __ldg(a.logw + k0 + j));
    }
    if (tid < npts) {
      const int64_t i = base + tid;
      float* col = rs + tid;
      if (a.dump != nullptr)
        for (int k = 0; k < a.K; ++k) a.dump[(size_t)k * a.n + i] = col[(size_t)k * T];
      int lab;
      if (a.final_iter) {
        lab = dpmm_draw_argmax(col, T, a.K);
      } else if (a.sampler == 1) {
        lab = dpmm_draw_gumbel(col, T, a.K, a.seed, a.call, (uint64_t)(a.goff + i));
      } else {
        const double u = dpmm_uniform(a.u_inj, i, a.seed, DPMM_STREAM_LABEL, a.call, (uint64_t)(a.goff + i));
        lab = dpmm_draw_inverse_cdf_screened(col, T, a.K, u);
      }
      a.labels[i] = lab;
      atomicAdd(&hs[lab], 1);
    }
  }
  __syncthreads();
  for (int k = tid; k < a.K; k += T)
    if (hs[k] != 0) atomicAdd(&a.hist[k], hs[k]);
}

// Sub-label draw over the label-sorted permutation (+ left/right partition); a.recs = log_p [3K][D].
template <bool SAMPLE>
__global__ void mnm_sublabel_kernel(const SubLabelArgs a) {
  __shared__ int s_cnt[2 * 256];
  __shared__ int s_base[2 * 256];
  __shared__ int s_first[2];
  const int tid = threadIdx.x;
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + tid;
  const bool active = pos < a.n;
  int32_t idx = 0;
  int k = 0, side = 0;
  if (active) {
    idx = a.perm[pos];
    k = a.labels[idx];
    if constexpr (SAMPLE) {
      const float* xp = a.x + (size_t)idx * a.D;
      const float* al = a.recs + (size_t)(3 * k + 1) * a.D;
      const float* ar = a.recs + (size_t)(3 * k + 2) * a.D;
      float sl = 0.f, sr = 0.f;
      for (int d = 0; d < a.D; ++d) {
        const float xv = __ldg(xp + d);
        sl = fmaf(__ldg(al + d), xv, sl);
        sr = fmaf(__ldg(ar + d), xv, sr);
      }
      const float rl = __fadd_rn(sl, __ldg(a.loglr + 2 * k));
      const float rr = __fadd_rn(sr, __ldg(a.loglr + 2 * k + 1));
      if (a.dump != nullptr) {
        a.dump[idx] = rl;
        a.dump[a.n + idx] = rr;
      }
      const double u = dpmm_uniform(a.u_inj, idx, a.seed, DPMM_STREAM_SUBLABEL, a.call, (uint64_t)(a.goff + idx));
      side = dpmm_draw_two(rl, rr, u);
      a.sub[idx] = (uint8_t)side;
    } else {
      side = a.sub[idx];
    }
  }
  sublabel_partition(a, active, k, side, idx, s_cnt, s_base, s_first);
}

// Sub-label draw, D <= 128: a warp evaluates its 32 label-sorted positions one point at a time with
// lanes <-> features (each point is read as coalesced 128-byte segments; the l/r log-probability
// vectors of the current label sit in registers and are reloaded only when the label changes), then
// every lane draws and partitions its own point.
__global__ void __launch_bounds__(128) mnm_sublabel2_kernel(const SubLabelArgs a) {
  __shared__ int s_cnt[2 * 128];
  __shared__ int s_base[2 * 128];
  __shared__ int s_first[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int D = a.D;
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + tid;
  const bool active = pos < a.n;
  int32_t idx = -1;
  int k = 0, side = 0;
  if (active) {
    idx = a.perm[pos];
    k = a.labels[idx];
  }
  float my_sl = 0.f, my_sr = 0.f;
  float al[4], ar[4];
  int cur = -1;
#pragma unroll 1
  for (int j0 = 0; j0 < 32; j0 += 4) {
    int pi[4], pk[4];
    float xv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pi[j] = __shfl_sync(0xffffffffu, idx, j0 + j);
      pk[j] = __shfl_sync(0xffffffffu, k, j0 + j);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int d = lane + 32 * r;
        xv[j][r] = (pi[j] >= 0 && d < D) ? __ldg(a.x + (size_t)pi[j] * D + d) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (pi[j] < 0) continue;                       // warp-uniform
      if (pk[j] != cur) {
        cur = pk[j];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int d = lane + 32 * r;
          al[r] = d < D ? __ldg(a.recs + (size_t)(3 * cur + 1) * D + d) : 0.f;
          ar[r] = d < D ? __ldg(a.recs + (size_t)(3 * cur + 2) * D + d) : 0.f;
        }
      }
      float sl = 0.f, sr = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        sl = fmaf(al[r], xv[j][r], sl);
        sr = fmaf(ar[r], xv[j][r], sr);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
      }
      if (lane == j0 + j) {
        my_sl = sl;
        my_sr = sr;
      }
    }
  }
  if (active) {
    const float rl = __fadd_rn(my_sl, __ldg(a.loglr + 2 * k));
    const float rr = __fadd_rn(my_sr, __ldg(a.loglr + 2 * k + 1));
    if (a.dump != nullptr) {
      a.dump[idx] = rl;
      a.dump[a.n + idx] = rr;
    }
    const double u = dpmm_uniform(a.u_inj, idx, a.seed, DPMM_STREAM_SUBLABEL, a.call, (uint64_t)(a.goff + idx));
    side = dpmm_draw_two(rl, rr, u);
    a.sub[idx] = (uint8_t)side;
  }
  sublabel_partition(a, active, k, side, idx, s_cnt, s_base, s_first);
}
