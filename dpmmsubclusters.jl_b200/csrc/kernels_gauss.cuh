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

// Offset (floats) of column j of the packed factor: column j holds U[0..j][j] padded to a multiple
// of 4 floats so that every column starts 16-byte aligned.
__host__ __device__ constexpr int gauss_col_off(int j) {
  int s = 0;
  for (int c = 0; c < j; ++c) s += 4 * (c / 4 + 1);
  return s;
}

template <int D>
struct GaussCfg {
  static constexpr int TRIP = gauss_col_off(D);   // floats of the packed upper-triangular factor
  static constexpr int DP4 = (D + 3) & ~3;
  static constexpr int REC = TRIP + DP4;  // floats per distribution record: [U packed columns | mu]
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

// q[pp] = |U (x_pp - mu)|^2 for P points owned by this thread.
//   rec     : distribution record [U packed by columns | mu] in shared memory
//   load4   : load4(pp, j0) -> the 4 coordinates j0..j0+3 of point pp (zero beyond D)
// Column-major accumulation: all D partial sums y_i = sum_{j>=i} U_ij z_j of a point live in
// registers and column j adds U[0..j][j] * z_j to y[0..j], so a thread always has >= D independent
// FMA chains in flight and z_j = x_j - mu_j is formed once per column (never stored).  The factor is
// read as float4 words at compile-time offsets: one LDS.128 per 4*P FMAs.
template <int D, int P, typename LF>
__device__ __forceinline__ void gauss_quadform(const float* __restrict__ rec, LF load4, float (&q)[P]) {
  using C = GaussCfg<D>;
  constexpr int NP = C::DP4 / 2;  // packed row pairs (rows 2i, 2i+1)
  constexpr int NV = C::DP4 / 4;  // float4 words of the longest column
  f32x2_t y[P][NP];
#pragma unroll
  for (int pp = 0; pp < P; ++pp)
#pragma unroll
    for (int i = 0; i < NP; ++i) y[pp][i] = 0ull;
  const float* mu = rec + C::TRIP;
  // Software pipeline (everything is unrolled, so all of this lives in registers): the factor
  // words of column j+1 and the point/mean words of the next 4-column block are loaded while
  // column j is being accumulated, which keeps the shared-memory latency off the FMA chains.
  ulonglong2 ub[2][NV];
  float4 xb[2][P], mb[2];
  ub[0][0] = *reinterpret_cast<const ulonglong2*>(rec);
  mb[0] = *reinterpret_cast<const float4*>(mu);
#pragma unroll
  for (int pp = 0; pp < P; ++pp) xb[0][pp] = load4(pp, 0);
  int off = 0;
#pragma unroll
  for (int j0 = 0; j0 < D; j0 += 4) {
    const int cb = (j0 >> 2) & 1;
    if (j0 + 4 < D) {
      mb[cb ^ 1] = *reinterpret_cast<const float4*>(mu + j0 + 4);
#pragma unroll
      for (int pp = 0; pp < P; ++pp) xb[cb ^ 1][pp] = load4(pp, j0 + 4);
    }
    float zj[P][4];
#pragma unroll
    for (int pp = 0; pp < P; ++pp) {
      zj[pp][0] = xb[cb][pp].x - mb[cb].x;
      zj[pp][1] = xb[cb][pp].y - mb[cb].y;
      zj[pp][2] = xb[cb][pp].z - mb[cb].z;
      zj[pp][3] = xb[cb][pp].w - mb[cb].w;
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = j0 + jj;
      if (j < D) {
        const int ncur = 4 * (j / 4 + 1);
        if (j + 1 < D) {  // prefetch column j+1
#pragma unroll
          for (int i4 = 0; i4 <= (j + 1) / 4; ++i4)
            ub[(j + 1) & 1][i4] = *reinterpret_cast<const ulonglong2*>(rec + off + ncur + 4 * i4);
        }
        f32x2_t z2[P];
#pragma unroll
        for (int pp = 0; pp < P; ++pp) z2[pp] = f2_pack(zj[pp][jj], zj[pp][jj]);
#pragma unroll
        for (int i4 = 0; i4 <= j / 4; ++i4) {
          // rows 4*i4 .. 4*i4+3 of column j: two packed pairs (entries below the diagonal are 0)
          const ulonglong2 u = ub[j & 1][i4];
#pragma unroll
          for (int pp = 0; pp < P; ++pp) {
            y[pp][2 * i4] = f2_fma(u.x, z2[pp], y[pp][2 * i4]);
            if (4 * i4 + 2 <= j) y[pp][2 * i4 + 1] = f2_fma(u.y, z2[pp], y[pp][2 * i4 + 1]);
          }
        }
        off += ncur;
      }
    }
  }
#pragma unroll
  for (int pp = 0; pp < P; ++pp) {
    f32x2_t s = 0ull;
#pragma unroll
    for (int i = 0; i < NP; ++i) s = f2_fma(y[pp][i], y[pp][i], s);
    float lo, hi;
    f2_unpack(s, lo, hi);
    q[pp] = lo + hi;
  }
}

// load4 functor over a point row in shared or global memory (16B-aligned rows when D % 4 == 0)
template <int D>
__device__ __forceinline__ float4 gauss_row_load4(const float* row, int j0) {
  if constexpr (D % 4 == 0) {
    return *reinterpret_cast<const float4*>(row + j0);
  } else {
    float4 v;
    v.x = row[j0];
    v.y = (j0 + 1 < D) ? row[j0 + 1] : 0.f;
    v.z = (j0 + 2 < D) ? row[j0 + 2] : 0.f;
    v.w = (j0 + 3 < D) ? row[j0 + 3] : 0.f;
    return v;
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
__global__ void __launch_bounds__(128, 2) gauss_label_kernel(const GaussLabelArgs a) {
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
        gauss_quadform<D, P>(us + (size_t)kk * C::REC,
                             [&](int pp, int j0) {
                               return gauss_row_load4<D>(xs + (size_t)(tid + pp * T) * C::DS, j0);
                             },
                             q);
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
          lab = dpmm_draw_inverse_cdf_screened(col, TP, a.K, u);
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
// K1 (warp-autonomous form, used whenever all K cluster records fit in shared memory next to the
// per-warp buffers): one persistent CTA per SM; every warp owns a private tile of 32*P points and
// runs stage -> log-likelihood -> draw on its own, synchronising only with __syncwarp.  The warps of
// an SM drift out of phase, so the FMA-bound likelihood loop of one warp overlaps the latency-bound
// staging and draw phases of the others (the CTA-synchronous kernel above keeps all of its warps in
// the same phase).  Shared memory:
//   us [K][REC] cluster records (staged once) | hs [K] label histogram |
//   per warp: xs [32P][DS] tile, rs [K][32P] slice of parr
// ---------------------------------------------------------------------------------------------
template <int D, int P>
__global__ void __launch_bounds__(384, 1) gauss_label_warp_kernel(const GaussLabelArgs a) {
  using C = GaussCfg<D>;
  constexpr int TPW = 32 * P;  // points per warp tile
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
  float* us = smem;
  int* hs = reinterpret_cast<int*>(us + (size_t)a.K * C::REC);
  float* csm = reinterpret_cast<float*>(hs + ((a.K + 3) & ~3));   // [K] c_k, [K] log w_k
  float* wbase = csm + 2 * ((a.K + 3) & ~3);
  float* xs = wbase + (size_t)warp * (TPW * C::DS + a.K * TPW);
  float* rs = xs + TPW * C::DS;

  for (int e = tid; e < a.K * (C::REC / 4); e += blockDim.x) {
    const int kk = e / (C::REC / 4), c = e - kk * (C::REC / 4);
    reinterpret_cast<float4*>(us)[e] = __ldg(reinterpret_cast<const float4*>(a.recs + (size_t)(3 * kk) * C::REC) + c);
  }
  for (int k = tid; k < a.K; k += blockDim.x) {
    hs[k] = 0;
    csm[k] = __ldg(a.cst + 3 * k);
    csm[a.K + k] = __ldg(a.logw + k);
  }
  __syncthreads();

  for (int64_t tile = (int64_t)blockIdx.x * W + warp; tile < a.ntiles; tile += (int64_t)gridDim.x * W) {
    const int64_t base = tile * TPW;
    const int npts = (int)min((int64_t)TPW, a.n - base);
    __syncwarp();
    // ---- stage the warp's tile: all loads in flight before the first store ----
    if constexpr (C::VEC) {
      constexpr int NLD = P * (D / 4);  // float4 per lane
      const float4* src = reinterpret_cast<const float4*>(a.x + base * D);
      const int nv = npts * (D / 4);
      float4 v[NLD];
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int e = lane + 32 * i;
        v[i] = (e < nv) ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int e = lane + 32 * i;
        const int p = e / (D / 4), c = e - p * (D / 4);
        *reinterpret_cast<float4*>(xs + p * C::DS + 4 * c) = v[i];
      }
    } else {
      const float* src = a.x + base * D;
      const int nv = npts * D;
      for (int e = lane; e < TPW * D; e += 32) {
        const int p = e / D, c = e - p * D;
        xs[p * C::DS + c] = (e < nv) ? __ldg(src + e) : 0.f;
      }
    }
    __syncwarp();
    // ---- log-likelihood under every cluster ----
    if constexpr (D <= 8) {
      // low-dimensional points live in registers for the whole cluster loop (re-reading them from
      // shared memory per cluster would cost more than the D(D+1)/2 FMAs of the quadratic form)
      float4 xr[P][C::DP4 / 4];
#pragma unroll
      for (int pp = 0; pp < P; ++pp)
#pragma unroll
        for (int j4 = 0; j4 < C::DP4 / 4; ++j4) xr[pp][j4] = gauss_row_load4<D>(xs + (lane + pp * 32) * C::DS, 4 * j4);
      for (int k = 0; k < a.K; ++k) {
        float q[P];
        gauss_quadform<D, P>(us + (size_t)k * C::REC, [&](int pp, int j0) { return xr[pp][j0 >> 2]; }, q);
        const float c = csm[k], lw = csm[a.K + k];
#pragma unroll
        for (int pp = 0; pp < P; ++pp) rs[k * TPW + lane + pp * 32] = gauss_finish(c, q[pp], lw);
      }
    } else {
      for (int k = 0; k < a.K; ++k) {
        float q[P];
        gauss_quadform<D, P>(us + (size_t)k * C::REC,
                             [&](int pp, int j0) { return gauss_row_load4<D>(xs + (lane + pp * 32) * C::DS, j0); }, q);
        const float c = csm[k], lw = csm[a.K + k];
#pragma unroll
        for (int pp = 0; pp < P; ++pp) rs[k * TPW + lane + pp * 32] = gauss_finish(c, q[pp], lw);
      }
    }
    // ---- draw (each lane only touches its own columns of rs) ----
#pragma unroll
    for (int pp = 0; pp < P; ++pp) {
      const int p = lane + pp * 32;
      if (p < npts) {
        const int64_t i = base + p;
        float* col = rs + p;
        if (a.dump != nullptr)
          for (int k = 0; k < a.K; ++k) a.dump[(size_t)k * a.n + i] = col[k * TPW];
        int lab;
        if (a.final_iter) {
          lab = dpmm_draw_argmax(col, TPW, a.K);
        } else if (a.sampler == 1) {
          lab = dpmm_draw_gumbel(col, TPW, a.K, a.seed, a.call, (uint64_t)(a.goff + i));
        } else {
          const double u = dpmm_uniform(a.u_inj, i, a.seed, DPMM_STREAM_LABEL, a.call, (uint64_t)(a.goff + i));
          lab = dpmm_draw_inverse_cdf_screened(col, TPW, a.K, u);
        }
        a.labels[i] = lab;
        atomicAdd(&hs[lab], 1);
      }
    }
  }
  __syncthreads();
  for (int k = tid; k < a.K; k += blockDim.x)
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
  const int32_t* seg_off; // [K+1] label segment offsets in perm
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

#define SUBLABEL_SPAN 2  // clusters whose l/r records are staged per pass

template <int D, bool SAMPLE>
__global__ void gauss_sublabel_kernel(const SubLabelArgs a) {
  using C = GaussCfg<D>;
  __shared__ int s_cnt[2 * 256];
  __shared__ int s_base[2 * 256];
  __shared__ int s_first[2];
  __shared__ __align__(16) float us[SAMPLE ? SUBLABEL_SPAN * 2 * C::REC : 4];
  const int tid = threadIdx.x, T = blockDim.x;
  const int64_t pos0 = (int64_t)blockIdx.x * T;
  const int64_t pos = pos0 + tid;
  const bool active = pos < a.n;
  int32_t idx = 0;
  int k = 0, side = 0;
  if (active) {
    idx = a.perm[pos];
    k = a.labels[idx];
    if constexpr (!SAMPLE) side = a.sub[idx];
  }
  if constexpr (SAMPLE) {
    const int nact = (int)min((int64_t)T, a.n - pos0);
    if (tid == 0) s_first[0] = k;
    if (tid == nact - 1) s_first[1] = k;
    float xr[C::DP4];
#pragma unroll
    for (int j = 0; j < C::DP4; ++j) xr[j] = 0.f;
    if (active) {
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
    }
    __syncthreads();
    const int kfirst = s_first[0], klast = s_first[1];
    float rl = 0.f, rr = 0.f;
    // the tile is label-sorted: stage the l/r records of SUBLABEL_SPAN consecutive labels at a time
    // (almost always a single pass); lanes of a warp then read them as shared-memory broadcasts
    for (int kb = kfirst; kb <= klast; kb += SUBLABEL_SPAN) {
      const int nk = min(SUBLABEL_SPAN, klast - kb + 1);
      __syncthreads();
      for (int e = tid; e < nk * 2 * (C::REC / 4); e += T) {
        const int r = e / (C::REC / 4), c = e - r * (C::REC / 4);
        const int kk = kb + (r >> 1), s = 1 + (r & 1);
        reinterpret_cast<float4*>(us)[e] =
            __ldg(reinterpret_cast<const float4*>(a.recs + (size_t)(3 * kk + s) * C::REC) + c);
      }
      __syncthreads();
      if (active && k >= kb && k < kb + nk) {
        float ql[1], qr[1];
        auto ld = [&](int, int j0) { return make_float4(xr[j0], xr[j0 + 1], xr[j0 + 2], xr[j0 + 3]); };
        gauss_quadform<D, 1>(us + (size_t)((k - kb) * 2) * C::REC, ld, ql);
        gauss_quadform<D, 1>(us + (size_t)((k - kb) * 2 + 1) * C::REC, ld, qr);
        rl = gauss_finish(__ldg(a.cst + 3 * k + 1), ql[0], __ldg(a.loglr + 2 * k));
        rr = gauss_finish(__ldg(a.cst + 3 * k + 2), qr[0], __ldg(a.loglr + 2 * k + 1));
      }
    }
    if (active) {
      if (a.dump != nullptr) {
        a.dump[idx] = rl;
        a.dump[a.n + idx] = rr;
      }
      const double u = dpmm_uniform(a.u_inj, idx, a.seed, DPMM_STREAM_SUBLABEL, a.call, (uint64_t)(a.goff + idx));
      side = dpmm_draw_two(rl, rr, u);
      a.sub[idx] = (uint8_t)side;
    }
    __syncthreads();
  }
  sublabel_partition(a, active, k, side, idx, s_cnt, s_base, s_first);
}

// ---------------------------------------------------------------------------------------------
// K4, two points per thread: thread t owns the ADJACENT label-sorted positions 2t and 2t+1 of a
// 256-position tile, which share their label except at a segment boundary, so one pass of factor
// loads serves two points (the one-point form above is bound by the shared-memory pipe: 144
// LDS.128 per 288 FFMA2).  The tile's rows are gathered into shared memory with cp.async.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sl_cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}

template <int D>
__global__ void __launch_bounds__(128) gauss_sublabel2_kernel(const SubLabelArgs a) {
  using C = GaussCfg<D>;
  static_assert(C::VEC, "two-point sub-label kernel needs D % 4 == 0");
  constexpr int T = 128, TP = 256;
  extern __shared__ __align__(16) float sl_smem[];
  float* xs_all = sl_smem;                              // [2][TP][DS]  double-buffered gathered rows
  float* us = xs_all + 2 * TP * C::DS;                  // [SPAN][l, r][REC]
  int* segs = reinterpret_cast<int*>(us + SUBLABEL_SPAN * 2 * C::REC);   // [K+1] label segment offsets
  __shared__ int s_cnt[2 * TP];
  __shared__ int s_base[2 * TP];
  const int tid = threadIdx.x;
  const int K = a.K;
  for (int e = tid; e <= K; e += T) segs[e] = __ldg(a.seg_off + e);
  __syncthreads();
  // label of a sorted position = the segment that contains it (positions are sorted by label)
  auto label_of = [&](int64_t pos) {
    int lo = 0, hi = K;                                  // segs[lo] <= pos < segs[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (segs[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
  };
  const int64_t ntiles = (a.n + TP - 1) / TP;
  auto load_perm = [&](int64_t tile, int32_t (&idx)[2]) {
    const int64_t p = tile * TP + 2 * tid;
    idx[0] = (tile < ntiles && p < a.n) ? __ldg(a.perm + p) : -1;
    idx[1] = (tile < ntiles && p + 1 < a.n) ? __ldg(a.perm + p + 1) : -1;
  };
  auto gather = [&](int buf, const int32_t (&idx)[2]) {
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      float* dst = xs_all + (size_t)buf * TP * C::DS + (2 * tid + pp) * C::DS;
      const bool ok = idx[pp] >= 0;
      const float* src = a.x + (size_t)(ok ? idx[pp] : 0) * D;
#pragma unroll
      for (int c = 0; c < D / 4; ++c) sl_cp_async16(dst + 4 * c, src + 4 * c, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // software pipeline over this CTA's tiles: the permutation entries are loaded two tiles ahead and
  // the rows one tile ahead of the tile being evaluated
  int32_t idx[2], idx_n[2], idx_nn[2];
  int64_t tile = blockIdx.x;
  load_perm(tile, idx);
  load_perm(tile + gridDim.x, idx_n);
  gather(0, idx);
  int cached_first = -1, cached_last = -2;
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const float* xs = xs_all + (size_t)buf * TP * C::DS;
    load_perm(tile + 2 * (int64_t)gridDim.x, idx_nn);
    gather(buf ^ 1, idx_n);                               // rows of the next tile (a no-op group past the end)
    const int64_t pos0 = tile * TP;
    const int nact = (int)min((int64_t)TP, a.n - pos0);
    const int kfirst = label_of(pos0), klast = label_of(pos0 + nact - 1);
    int k[2], side[2] = {0, 0};
    bool act[2];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      act[pp] = idx[pp] >= 0;
      // most tiles lie inside one label segment: only boundary tiles search per point
      k[pp] = (act[pp] && kfirst != klast) ? label_of(pos0 + 2 * tid + pp) : kfirst;
    }
    for (int j = tid; j < 2 * TP; j += T) s_cnt[j] = 0;
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this tile's rows have landed
    __syncthreads();
    float rl[2] = {0.f, 0.f}, rr[2] = {0.f, 0.f};
    for (int kb = kfirst; kb <= klast; kb += SUBLABEL_SPAN) {
      const int nk = min(SUBLABEL_SPAN, klast - kb + 1);
      const bool reuse = (kb == kfirst) && (kfirst == cached_first) && (klast == cached_last) && (klast - kfirst < SUBLABEL_SPAN);
      if (!reuse) {
        if (kb != kfirst) __syncthreads();
        for (int e = tid; e < nk * 2 * (C::REC / 4); e += T) {
          const int r = e / (C::REC / 4), c = e - r * (C::REC / 4);
          const int kk = kb + (r >> 1), s = 1 + (r & 1);
          reinterpret_cast<float4*>(us)[e] = __ldg(reinterpret_cast<const float4*>(a.recs + (size_t)(3 * kk + s) * C::REC) + c);
        }
        __syncthreads();
      }
      // pass 0 evaluates both points under the first point's label, pass 1 (only at a segment
      // boundary) both under the second point's label
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const int kk = pass == 0 ? k[0] : k[1];
        const bool inchunk = kk >= kb && kk < kb + nk;
        const bool need = pass == 0 ? (act[0] && inchunk) || (act[1] && k[1] == k[0] && inchunk)
                                    : (act[1] && k[1] != k[0] && inchunk);
        if (need) {
          float ql[2], qr[2];
          auto ld = [&](int pp, int j0) { return *reinterpret_cast<const float4*>(xs + (2 * tid + pp) * C::DS + j0); };
          gauss_quadform<D, 2>(us + (size_t)((kk - kb) * 2) * C::REC, ld, ql);
          gauss_quadform<D, 2>(us + (size_t)((kk - kb) * 2 + 1) * C::REC, ld, qr);
          const float cl = __ldg(a.cst + 3 * kk + 1), cr = __ldg(a.cst + 3 * kk + 2);
          const float wl = __ldg(a.loglr + 2 * kk), wr = __ldg(a.loglr + 2 * kk + 1);
#pragma unroll
          for (int pp = 0; pp < 2; ++pp)
            if (k[pp] == kk) {
              rl[pp] = gauss_finish(cl, ql[pp], wl);
              rr[pp] = gauss_finish(cr, qr[pp], wr);
            }
        }
      }
    }
    cached_first = kfirst;
    cached_last = (klast - kfirst < SUBLABEL_SPAN) ? klast : -2;   // only a single-pass staging stays valid
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
      if (act[pp]) {
        if (a.dump != nullptr) {
          a.dump[idx[pp]] = rl[pp];
          a.dump[a.n + idx[pp]] = rr[pp];
        }
        const double u = dpmm_uniform(a.u_inj, idx[pp], a.seed, DPMM_STREAM_SUBLABEL, a.call, (uint64_t)(a.goff + idx[pp]));
        side[pp] = dpmm_draw_two(rl[pp], rr[pp], u);
        a.sub[idx[pp]] = (uint8_t)side[pp];
      }
    // ---- partition every label segment of the tile into left | right (as sublabel_partition) ----
    const int span = klast - kfirst + 1;
    if (span <= TP) {
      int rank[2] = {0, 0}, local[2] = {0, 0};
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
        if (act[pp]) {
          local[pp] = (k[pp] - kfirst) * 2 + side[pp];
          rank[pp] = atomicAdd(&s_cnt[local[pp]], 1);
        }
      __syncthreads();
      for (int j = tid; j < 2 * span; j += T) {
        const int c = s_cnt[j];
        if (c > 0) {
          const int key = 2 * kfirst + j;
          s_base[j] = (j & 1) ? atomicSub(&a.cursor[key], c) - c : atomicAdd(&a.cursor[key], c);
        }
      }
      __syncthreads();
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
        if (act[pp]) a.perm2[s_base[local[pp]] + rank[pp]] = idx[pp];
    } else {
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
        if (act[pp]) {
          const int key = 2 * k[pp] + side[pp];
          const int dst = side[pp] ? atomicSub(&a.cursor[key], 1) - 1 : atomicAdd(&a.cursor[key], 1);
          a.perm2[dst] = idx[pp];
        }
    }
    __syncthreads();   // s_cnt / s_base / xs[buf] are reused by the next tile
    idx[0] = idx_n[0]; idx[1] = idx_n[1];
    idx_n[0] = idx_nn[0]; idx_n[1] = idx_nn[1];
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
