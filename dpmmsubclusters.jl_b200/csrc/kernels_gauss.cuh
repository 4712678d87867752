// Stage 1+2 for the NIW-Gaussian prior on the FP32 FFMA pipe.
//
//   log_likelihood!(r, x, ::mv_gaussian)      src/distributions/mv_gaussian.jl:21-25
//   sample_labels_worker!                      src/local_clusters_actions.jl:112-134
//   create_subclusters_labels!                 src/local_clusters_actions.jl:83-95
//
// The reference evaluates  q = z' invSigma z  (z = x - mu) with an SGEMM + column dot; here the
// host-side factor invSigma = U'U (U upper triangular, the reference's own mv_gaussian.invChol)
// turns it into q = |U z|^2: D(D+1)/2 FMAs for the triangular product instead of D^2, all in
// registers, with z = x - mu formed in Float32 BEFORE the contraction exactly as the reference does
// (no expanded |Ux - U mu| form, which loses ~1.5 digits when |mu| >> sigma).
// The final Float32 operations are the reference's:  r = -c - q/2 ;  r += log(w)
// with c = (D^2 * Float32(log 2pi) + logdetSigma)/2 (the `length(Sigma)` = D^2 quirk).
#pragma once
#include "common.cuh"

template <int D>
struct GaussCfg {
  static constexpr int TRI = D * (D + 1) / 2;
  static constexpr int TRIP = (TRI + 3) & ~3;
  static constexpr int DP4 = (D + 3) & ~3;
  static constexpr int REC = TRIP + DP4;  // floats per distribution record: [U packed rows | mu]
  static constexpr bool VEC = (D % 4 == 0);
  // shared-memory row stride of a staged point: 16B-aligned rows whose float4 index is odd (no
  // bank conflicts for LDS.128 with lane <-> point), or an odd scalar stride.
  static constexpr int DS = VEC ? 4 * ((D / 4) | 1) : ((D & 1) ? D : D + 1);
};

struct GaussLabelArgs {
  const float* x;        // [n][D]
  int64_t n;
  int K;
  int KC;                // clusters staged in shared memory at a time
  const float* recs;     // [3K][REC]; the cluster distribution of k is record 3k
  const float* cst;      // [3K]  c = (D^2 log2pi + logdet)/2
  const float* logw;     // [K]
  int32_t* labels;       // out, 0-based
  int32_t* hist;         // [K] global histogram of the new labels (pre-zeroed)
  const double* u_inj;   // injected uniforms or nullptr
  uint64_t seed;
  uint32_t call;
  int64_t goff;
  int final_iter;
  int sampler;
  float* dump;           // optional [K][n] log-likelihood dump (parity hook)
  int64_t ntiles;
};

// z[pp][:] = x_pp - mu for points staged in shared memory (row pointers from `xrow`).
template <int D, int P, typename XF>
__device__ __forceinline__ void gauss_center_smem(const float* __restrict__ rec, XF xrow, float (&z)[P][D]) {
  using C = GaussCfg<D>;
  const float* mu = rec + C::TRIP;
  if constexpr (C::VEC) {
#pragma unroll
    for (int j4 = 0; j4 < D / 4; ++j4) {
      const float4 m = *reinterpret_cast<const float4*>(mu + 4 * j4);
#pragma unroll
      for (int pp = 0; pp < P; ++pp) {
        const float4 v = *reinterpret_cast<const float4*>(xrow(pp) + 4 * j4);
        z[pp][4 * j4 + 0] = v.x - m.x;
        z[pp][4 * j4 + 1] = v.y - m.y;
        z[pp][4 * j4 + 2] = v.z - m.z;
        z[pp][4 * j4 + 3] = v.w - m.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const float m = mu[j];
#pragma unroll
      for (int pp = 0; pp < P; ++pp) z[pp][j] = xrow(pp)[j] - m;
    }
  }
}

// q[pp] = |U z_pp|^2 for P centred points held in registers.  `rec` points to the distribution
// record [U packed upper-triangular rows | mu] in shared or global memory; the triangle is read as
// float4 words at compile-time offsets (one LDS.128 / LDG.128 per 4*P FMAs).
template <int D, int P>
__device__ __forceinline__ void gauss_quadform(const float* __restrict__ rec, const float (&z)[P][D], float (&q)[P]) {
#pragma unroll
  for (int pp = 0; pp < P; ++pp) q[pp] = 0.f;
  const float4* U4 = reinterpret_cast<const float4*>(rec);
  float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int e = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float acc[P];
#pragma unroll
    for (int pp = 0; pp < P; ++pp) acc[pp] = 0.f;
#pragma unroll
    for (int j = i; j < D; ++j) {
      if ((e & 3) == 0) u4 = U4[e >> 2];
      const float u = ((e & 3) == 0) ? u4.x : ((e & 3) == 1) ? u4.y : ((e & 3) == 2) ? u4.z : u4.w;
#pragma unroll
      for (int pp = 0; pp < P; ++pp) acc[pp] = fmaf(u, z[pp][j], acc[pp]);
      ++e;
    }
#pragma unroll
    for (int pp = 0; pp < P; ++pp) q[pp] = fmaf(acc[pp], acc[pp], q[pp]);
  }
}

// r = -c - q/2 (mv_gaussian.jl:24), then r += log(w) (local_clusters_actions.jl:126 / :92-93).
__device__ __forceinline__ float gauss_finish(float c, float q, float logw) {
  const float r = __fsub_rn(-c, __fmul_rn(q, 0.5f));
  return __fadd_rn(r, logw);
}

// ---------------------------------------------------------------------------------------------
// K1: fused log-likelihood + label draw.  One thread owns P points; a CTA of blockDim.x threads
// owns a tile of TP = P*blockDim.x consecutive points.  Shared memory:
//   xs [TP][DS]   the tile (coalesced copy of TP*D contiguous floats)
//   rs [K][TP]    the tile's slice of parr (never leaves the SM)
//   us [KC][REC]  the staged cluster distributions
//   hs [K]        histogram of the labels drawn by this CTA (feeds the label sort)
// ---------------------------------------------------------------------------------------------
template <int D, int P>
__global__ void gauss_label_kernel(const GaussLabelArgs a) {
  using C = GaussCfg<D>;
  extern __shared__ __align__(16) float smem[];
  const int T = blockDim.x;
  const int TP = T * P;
  float* xs = smem;
  float* rs = xs + (size_t)TP * C::DS;
  float* us = rs + (size_t)a.K * TP;
  int* hs = reinterpret_cast<int*>(us + (size_t)a.KC * C::REC);
  const int tid = threadIdx.x;

  for (int k = tid; k < a.K; k += T) hs[k] = 0;
  bool staged = false;

  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t base = tile * TP;
    const int npts = (int)min((int64_t)TP, a.n - base);
    __syncthreads();  // previous tile's readers of xs/rs are done
    // ---- stage the tile ----
    if constexpr (C::VEC) {
      const float4* src = reinterpret_cast<const float4*>(a.x + base * D);
      const int nv = npts * (D / 4);
      for (int e = tid; e < TP * (D / 4); e += T) {
        const int p = e / (D / 4), c = e - p * (D / 4);
        const float4 v = (e < nv) ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(xs + (size_t)p * C::DS + 4 * c) = v;
      }
    } else {
      const float* src = a.x + base * D;
      const int nv = npts * D;
      for (int e = tid; e < TP * D; e += T) {
        const int p = e / D, c = e - p * D;
        xs[(size_t)p * C::DS + c] = (e < nv) ? __ldg(src + e) : 0.f;
      }
    }
    // ---- log-likelihood of every point of the tile under every cluster ----
    for (int kc0 = 0; kc0 < a.K; kc0 += a.KC) {
      const int kcn = min(a.KC, a.K - kc0);
      if (!staged || a.KC < a.K) {
        __syncthreads();
        for (int e = tid; e < kcn * (C::REC / 4); e += T) {
          const int kk = e / (C::REC / 4), c = e - kk * (C::REC / 4);
          reinterpret_cast<float4*>(us)[e] =
              __ldg(reinterpret_cast<const float4*>(a.recs + (size_t)(3 * (kc0 + kk)) * C::REC) + c);
        }
        staged = true;
      }
      __syncthreads();
      for (int kk = 0; kk < kcn; ++kk) {
        const int k = kc0 + kk;
        float q[P];
        {
          float z[P][D];
          gauss_center_smem<D, P>(us + (size_t)kk * C::REC,
                                  [&](int pp) { return xs + (size_t)(tid + pp * T) * C::DS; }, z);
          gauss_quadform<D, P>(us + (size_t)kk * C::REC, z, q);
        }
        const float c = __ldg(a.cst + 3 * k), lw = __ldg(a.logw + k);
#pragma unroll
        for (int pp = 0; pp < P; ++pp) rs[(size_t)k * TP + tid + pp * T] = gauss_finish(c, q[pp], lw);
      }
    }
    // ---- draw ----
#pragma unroll
    for (int pp = 0; pp < P; ++pp) {
      const int p = tid + pp * T;
      if (p < npts) {
        const int64_t i = base + p;
        float* col = rs + p;
        if (a.dump != nullptr)
          for (int k = 0; k < a.K; ++k) a.dump[(size_t)k * a.n + i] = col[(size_t)k * TP];
        int lab;
        if (a.final_iter) {
          lab = dpmm_draw_argmax(col, TP, a.K);
        } else if (a.sampler == 1) {
          lab = dpmm_draw_gumbel(col, TP, a.K, a.seed, a.call, (uint64_t)(a.goff + i));
        } else {
          const double u = dpmm_uniform(a.u_inj, i, a.seed, DPMM_STREAM_LABEL, a.call, (uint64_t)(a.goff + i));
          lab = dpmm_draw_inverse_cdf(col, TP, a.K, u);
        }
        a.labels[i] = lab;
        atomicAdd(&hs[lab], 1);
      }
    }
  }
  __syncthreads();
  for (int k = tid; k < a.K; k += T)
    if (hs[k] != 0) atomicAdd(&a.hist[k], hs[k]);
}

// ---------------------------------------------------------------------------------------------
// K4: sub-label draw over the label-sorted permutation + partition of every label segment into
// its left / right halves (feeds the statistics kernel).  One thread = one sorted position; lanes
// of a warp almost always share the label, so the l/r records are read through the read-only path
// as warp-uniform broadcast loads.  With SAMPLE=false the kernel only partitions (used when the
// sub-labels were randomised or restored rather than sampled).
// ---------------------------------------------------------------------------------------------
struct SubLabelArgs {
  const float* x;
  int64_t n;
  int K;
  const float* recs;      // [3K][REC]    (Gaussian)   |  log_p [3K][D] (multinomial)
  const float* cst;       // [3K]
  const float* loglr;     // [K][2]
  const int32_t* labels;  // [n] 0-based
  uint8_t* sub;           // [n] 0 = left, 1 = right
  const int32_t* perm;    // [n] positions sorted by label
  int32_t* perm2;         // [n] out: label segments partitioned left | right
  int* cursor;            // [2K] cursor[2k] grows up from seg_off[k], cursor[2k+1] down from seg_off[k+1]
  const double* u_inj;
  uint64_t seed;
  uint32_t call;
  int64_t goff;
  float* dump;            // optional [2][n]
  int D;                  // runtime D (multinomial)
};

// Partition step shared by the Gaussian and multinomial kernels.  `k`/`side` are this thread's
// label and sub-label, `idx` its point, `active` whether the position exists.
__device__ __forceinline__ void sublabel_partition(const SubLabelArgs& a, bool active, int k, int side,
                                                   int32_t idx, int* s_cnt, int* s_base, int* s_first) {
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int64_t pos0 = (int64_t)blockIdx.x * T;
  const int nact = (int)min((int64_t)T, a.n - pos0);
  if (tid == 0) s_first[0] = k;
  if (tid == nact - 1) s_first[1] = k;
  for (int j = tid; j < 2 * T; j += T) s_cnt[j] = 0;
  __syncthreads();
  const int kfirst = s_first[0];
  const int span = s_first[1] - kfirst + 1;
  if (span <= T) {
    int rank = 0;
    const int local = active ? (k - kfirst) * 2 + side : 0;
    if (active) rank = atomicAdd(&s_cnt[local], 1);
    __syncthreads();
    for (int j = tid; j < 2 * span; j += T) {
      const int c = s_cnt[j];
      if (c > 0) {
        const int key = 2 * kfirst + j;
        s_base[j] = (j & 1) ? atomicSub(&a.cursor[key], c) - c : atomicAdd(&a.cursor[key], c);
      }
    }
    __syncthreads();
    if (active) a.perm2[s_base[local] + rank] = idx;
  } else if (active) {  // a tile spanning > T labels (many tiny clusters): direct reservation
    const int key = 2 * k + side;
    const int dst = side ? atomicSub(&a.cursor[key], 1) - 1 : atomicAdd(&a.cursor[key], 1);
    a.perm2[dst] = idx;
  }
}

template <int D, bool SAMPLE>
__global__ void gauss_sublabel_kernel(const SubLabelArgs a) {
  using C = GaussCfg<D>;
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
      float xr[C::DP4];
      const float* xp = a.x + (size_t)idx * D;
      if constexpr (C::VEC) {
#pragma unroll
        for (int j4 = 0; j4 < D / 4; ++j4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(xp) + j4);
          xr[4 * j4] = v.x; xr[4 * j4 + 1] = v.y; xr[4 * j4 + 2] = v.z; xr[4 * j4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < D; ++j) xr[j] = __ldg(xp + j);
      }
      float ql[1], qr[1];
      {
        const float* recl = a.recs + (size_t)(3 * k + 1) * C::REC;
        const float* recr = a.recs + (size_t)(3 * k + 2) * C::REC;
        float zl[1][D], zr[1][D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
          zl[0][j] = xr[j] - __ldg(recl + C::TRIP + j);
          zr[0][j] = xr[j] - __ldg(recr + C::TRIP + j);
        }
        gauss_quadform<D, 1>(recl, zl, ql);
        gauss_quadform<D, 1>(recr, zr, qr);
      }
      const float rl = gauss_finish(__ldg(a.cst + 3 * k + 1), ql[0], __ldg(a.loglr + 2 * k));
      const float rr = gauss_finish(__ldg(a.cst + 3 * k + 2), qr[0], __ldg(a.loglr + 2 * k + 1));
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
