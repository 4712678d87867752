// Parameter packing on the device: broadcast_cluster_params (src/local_clusters_actions.jl:518-549)
// ships mv_gaussian fields (mu, invSigma, logdetSigma; mv_gaussian.jl:12-18); this kernel turns them
// into what the sweep kernels read.  One CTA per distribution:
//   invSigma (Float32, symmetrised) -> Float64 Cholesky  invSigma = L L'  -> U = L' rounded to Float32,
//   stored by columns (FMA path records) and, for the cluster distributions of the tensor-core path,
//   by rows (K-major B operand) with b = U mu and |U|_F;  c = (D^2 * Float32(log 2pi) + logdetSigma)/2.
// The reference carries the same factor (mv_gaussian.invChol, niw.jl:38-39) but evaluates z' invSigma z
// directly (mv_gaussian.jl:21-24).  A Float32-rounded invSigma that lost positive definiteness by rounding
// alone is factored with a tiny diagonal jitter (see niw_pack_kernel); an indefinite or NaN input gives NaN
// factors, i.e. NaN log-likelihoods.
#pragma once
#include "common.cuh"
#include "kernels_gauss.cuh"   // gauss_col_off

struct NiwPackArgs {
  int D, K, rec_f, trip;
  int D_const;             // the D of the reference's D^2 log(2 pi) constant (= the caller's D when features are padded)
  const float* mu;         // [3K][D]
  const float* inv_sigma;  // [3K][D][D]
  const float* logdet;     // [3K]
  float* recs;             // [3K][rec_f]
  float* cst;              // [3K]
  float* tc_w;             // [K4][D][D] or nullptr
  float* tc_b;             // [K][D]
  float* tc_mu;            // [K][D]
  float* tc_fro;           // [K]
  // fused sub-label + statistics tensor-core path (kernels_substats_tc.cuh), or nullptr
  float* ss_w;             // [K][2][D][D] rows of U of the left / right distributions
  float* ss_b;             // [K][2][D]    U_s (mu_s - c_k)
  float* ss_c;             // [K][D]       c_k = cluster mean rounded to 12 significant bits
  // second-generation tensor-core label path (kernels_gauss_tc2.cuh), or t2_piv == nullptr:
  // operand images in the un-swizzled K-major core-matrix layout [k-step][row / 8][k half][row % 8][4]
  float* t2_piv;           // [K][D*D]  all D rows of U_k as D/8 k-step slabs
  float* t2_scr;           // [nch][(KS/8) * 1024]  8 screen rows of every cluster, 16 clusters per chunk
  float* t2_u;             // [K][D][D] rows of U_k (exact refinement, bias tables)
  float* t2_fro8;          // [K] Frobenius norm of the 8 screen rows
  int t2_KS;               // features of the screen: D (first 8 rows of U_k) or 8 (last 8 rows = last 8 features)
  int t2_n0;               // clusters in chunk 0 (behind the pivot's D columns)
  // optional: the factor itself, L lower with invSigma = L L' ([3K][D][D] Float64, from niw_draw_kernel);
  // inv_sigma is then ignored and no factorisation runs
  const double* lfac;
};

// The centre every point of cluster k is shifted by before it meets the tensor core: the cluster mean
// rounded to 12 significant bits, so that x - c is EXACT in Float32 for every point within a few
// widths of the cluster (the operands share their exponent range and c has 12 trailing zero bits).
// A component is NOT centred (c_i = 0) when one of the two sub-cluster means lies closer to the origin
// than half the cluster mean: the points of that sub-cluster would then be far from c_i relative to
// their own magnitude, and the statistics' rounding (relative to sum y_i^2) would not be small against
// the un-centred sum x_i^2 the result is measured by.  In that case the cluster is at least |c_i|/2
// wide in this component, so nothing is lost by not centring it.
__device__ __forceinline__ float niw_pack_center(float m, float ml, float mr) {
  if (!(fabsf(m) < CUDART_INF_F)) return 0.f;
  const float h = 0.5f * fabsf(m);
  if (!(ml * m > 0.f && mr * m > 0.f && fabsf(ml) >= h && fabsf(mr) >= h)) return 0.f;
  uint32_t u = __float_as_uint(m);
  u += 0x7FFu + ((u >> 12) & 1u);
  u &= 0xFFFFF000u;
  return __uint_as_float(u);
}

#define NIW_PACK_THREADS 256

// One CTA of 256 threads per distribution.  The right-looking Cholesky applies, per column j, the same
// operations in the same order as a one-thread loop would (every element L[i][k] receives its updates
// for j = 0, 1, ... in turn), spread over the CTA: 3 barriers per column instead of a D-long serial chain.
// Everything the sweep kernels read of ONE distribution t, from the Float64 factor L (invSigma = L L', lower,
// shared memory Ls[D][D+1]), the Float32 mean a.mu[t] and a.logdet[t].  Called by the whole CTA.
__device__ __forceinline__ void niw_pack_body(const NiwPackArgs& a, const int t, const double* Ls, const bool ok) {
  const int D = a.D, LD = D + 1;
  const int tid = threadIdx.x, NT = NIW_PACK_THREADS;
  const float nanv = __int_as_float(0x7fc00000);
  float* rec = a.recs + (size_t)t * a.rec_f;
  for (int e = tid; e < a.rec_f; e += NT) rec[e] = 0.f;
  __syncthreads();
  for (int e = tid; e < D * D; e += NT) {   // column j of U = row j of L
    const int j = e / D, i = e - j * D;
    if (i <= j) rec[gauss_col_off(j) + i] = ok ? (float)Ls[j * LD + i] : nanv;
  }
  for (int j = tid; j < D; j += NT) rec[a.trip + j] = a.mu[(size_t)t * D + j];
  if (tid == 0) {
    const float log2pi = 1.8378770664093453f;   // Float32(log(2pi)), mv_gaussian.jl:24
    a.cst[t] = __fmul_rn(__fadd_rn(__fmul_rn((float)(a.D_const * a.D_const), log2pi), a.logdet[t]), 0.5f);
  }
  if ((a.tc_w != nullptr || a.t2_piv != nullptr) && t % 3 == 0) {
    const int k = t / 3;
    if (a.tc_w != nullptr) for (int e = tid; e < D * D; e += NT) {
      const int i = e / D, j = e - i * D;
      a.tc_w[((size_t)k * D + i) * D + j] = (j >= i) ? (ok ? (float)Ls[j * LD + i] : nanv) : 0.f;   // U[i][j] = L[j][i]
    }
    double fro = 0.0;
    for (int i = tid; i < D; i += NT) {
      double bi = 0.0;
      for (int j = 0; j < D; ++j) {
        const float u = (j >= i) ? (ok ? (float)Ls[j * LD + i] : nanv) : 0.f;
        bi += (double)u * (double)a.mu[(size_t)t * D + j];
        fro += (double)u * (double)u;
      }
      a.tc_b[(size_t)k * D + i] = (float)bi;
      a.tc_mu[(size_t)k * D + i] = a.mu[(size_t)t * D + i];
    }
    // |U|_F: the per-row partial sums live in the first D threads (D <= 64: warps 0 and 1)
    __shared__ double fro_s[2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fro += __shfl_xor_sync(0xffffffffu, fro, o);
    if (tid < 64 && (tid & 31) == 0) fro_s[tid >> 5] = fro;
    __syncthreads();
    if (tid == 0) a.tc_fro[k] = (float)sqrt(fro_s[0] + fro_s[1]);
  }
  if (a.t2_piv != nullptr && t % 3 == 0) {
    const int k = t / 3;
    auto uval = [&](int i, int j) { return (j >= i) ? (ok ? (float)Ls[j * LD + i] : nanv) : 0.f; };   // U[i][j] = L[j][i]
    float* piv = a.t2_piv + (size_t)k * (D * D);
    float* urow = a.t2_u + (size_t)k * D * D;
    for (int e = tid; e < D * D; e += NT) {
      const int i = e / D, j = e - i * D;
      const float u = uval(i, j);
      urow[e] = u;
      piv[(j >> 3) * (D * 8) + (i >> 3) * 64 + ((j & 7) >> 2) * 32 + (i & 7) * 4 + (j & 3)] = u;
    }
    const int KS = a.t2_KS;
    const int row0 = (KS == D) ? 0 : D - 8, f0 = (KS == D) ? 0 : D - 8;
    // slot of the cluster in its chunk; the rows of the two clusters of a slot pair are INTERLEAVED (row of the
    // image = 16 (slot / 2) + 2 r + slot % 2) so that the label kernel squares two clusters per packed FFMA2
    const int ch = k < a.t2_n0 ? 0 : 1 + (k - a.t2_n0) / 16, slot = k < a.t2_n0 ? k : (k - a.t2_n0) % 16;
    float* scr = a.t2_scr + (size_t)ch * (KS / 8) * 1024;
    for (int e = tid; e < 8 * KS; e += NT) {
      const int r = e / KS, jj = e - r * KS;
      const int n = 16 * (slot >> 1) + 2 * r + (slot & 1);
      scr[(jj >> 3) * 1024 + (n >> 3) * 64 + ((jj & 7) >> 2) * 32 + (n & 7) * 4 + (jj & 3)] = uval(row0 + r, f0 + jj);
    }
    if (tid < 32) {   // |screen rows|_F
      double f8 = 0.0;
      for (int e = tid; e < 8 * D; e += 32) {
        const float u = uval(row0 + e / D, e % D);
        f8 += (double)u * (double)u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) f8 += __shfl_xor_sync(0xffffffffu, f8, o);
      if (tid == 0) a.t2_fro8[k] = (float)sqrt(f8);
    }
  }
  if (a.ss_w != nullptr) {
    const int k = t / 3, side = t % 3 - 1;
    const float* mu0 = a.mu + (size_t)(3 * k) * D;
    if (side < 0) {
      for (int j = tid; j < D; j += NT) a.ss_c[(size_t)k * D + j] = niw_pack_center(mu0[j], mu0[D + j], mu0[2 * D + j]);
    } else {
      float* W = a.ss_w + ((size_t)k * 2 + side) * D * D;
      for (int e = tid; e < D * D; e += NT) {
        const int i = e / D, j = e - i * D;
        W[e] = (j >= i) ? (ok ? (float)Ls[j * LD + i] : nanv) : 0.f;   // U[i][j] = L[j][i]
      }
      for (int i = tid; i < D; i += NT) {
        double bi = 0.0;
        for (int j = 0; j < D; ++j) {
          const float u = (j >= i) ? (ok ? (float)Ls[j * LD + i] : nanv) : 0.f;
          bi += (double)u * ((double)a.mu[(size_t)t * D + j] - (double)niw_pack_center(mu0[j], mu0[D + j], mu0[2 * D + j]));
        }
        a.ss_b[((size_t)k * 2 + side) * D + i] = (float)bi;
      }
    }
  }
}

__global__ void __launch_bounds__(NIW_PACK_THREADS) niw_pack_kernel(const NiwPackArgs a) {
  extern __shared__ double Ls[];   // [D][D+1]
  const int D = a.D, LD = D + 1;
  const int t = blockIdx.x, tid = threadIdx.x, NT = NIW_PACK_THREADS;
  bool ok = true;
  if (a.lfac != nullptr) {
    const double* Lf = a.lfac + (size_t)t * D * D;
    bool fine = true;
    for (int e = tid; e < D * D; e += NT) {
      const int i = e / D, j = e - i * D;
      const double v = Lf[e];
      Ls[i * LD + j] = v;
      if (j <= i && (!(fabs(v) < CUDART_INF) || (i == j && !(v > 0.0)))) fine = false;
    }
    ok = __syncthreads_and(fine ? 1 : 0) != 0;
  } else {
    // invSigma arrives rounded to Float32 (mv_gaussian.jl:15).  For a badly conditioned cluster (condition number
    // above ~1e7) the rounded matrix can lose positive definiteness by a hair although the reference's direct
    // z' invSigma z stays finite and meaningful; the factorisation is then retried with a relative diagonal
    // jitter of 1e-7 .. 1e-4 (far inside the 1e-4 log-likelihood tolerance).  A genuinely indefinite matrix still
    // fails and gives NaN log-likelihoods.
    const float* A = a.inv_sigma + (size_t)t * D * D;
    const double jit[5] = {0.0, 1e-7, 1e-6, 1e-5, 1e-4};
    for (int attempt = 0; attempt < 5; ++attempt) {
      __syncthreads();
      for (int e = tid; e < D * D; e += NT) {
        const int i = e / D, j = e - i * D;
        double v = 0.5 * ((double)A[(size_t)i * D + j] + (double)A[(size_t)j * D + i]);
        if (i == j) v += jit[attempt] * fabs(v);
        Ls[i * LD + j] = v;
      }
      __syncthreads();
      ok = true;
      for (int j = 0; j < D; ++j) {
        const double d = Ls[j * LD + j];
        ok = ok && (d > 0.0) && (d < CUDART_INF);
        const double ljj = sqrt(d);
        __syncthreads();
        if (tid == 0) Ls[j * LD + j] = ljj;
        for (int i = j + 1 + tid; i < D; i += NT) Ls[i * LD + j] /= ljj;
        __syncthreads();
        // trailing update of the lower triangle: element (i, k), j < k <= i < D
        const int m = D - j - 1;
        for (int e = tid; e < m * m; e += NT) {
          const int ii = e / m, kk = e - ii * m;
          if (kk <= ii) {
            const int i = j + 1 + ii, k = j + 1 + kk;
            Ls[i * LD + k] -= Ls[i * LD + j] * Ls[k * LD + j];
          }
        }
        __syncthreads();   // the next column's pivot is part of the trailing block
      }
      if (ok) break;   // (uniform over the CTA: every thread read the same pivots)
    }
  }
  niw_pack_body(a, t, Ls, ok);
}

// Bias tables of the second-generation label kernel.  Its tiles are centred by the pivot's mean,
// z = x - mu_p, so the screen rows of cluster k see  U_k (x - mu_k) = U_k z - U_k (mu_k - mu_p):
// one table per pivot p with b[k][r] = (U_k (mu_k - mu_p))_{row0 + r}, Float64 accumulation, stored as the
// compact B operand of the bias k-step: [chunk][128 rows][4] = (-b_hi, -b_lo, 0, 0) with b_hi TF32-exact.
// One CTA per pivot.
struct NiwT2BiasArgs {
  int D, K, KS, n0, nch;
  const float* u;      // [K][D][D]
  const float* mu;     // [K][D]
  float* bias;         // [K][nch * 512 + 32]
};
__global__ void __launch_bounds__(256) niw_t2_bias_kernel(const NiwT2BiasArgs a) {
  const int p = blockIdx.x, D = a.D, K = a.K;
  const int row0 = (a.KS == D) ? 0 : D - 8;
  const int stride = a.nch * 512 + 32;
  float* out = a.bias + (size_t)p * stride;
  for (int e = threadIdx.x; e < stride; e += blockDim.x) out[e] = 0.f;
  __syncthreads();
  const float* mup = a.mu + (size_t)p * D;
  for (int e = threadIdx.x; e < K * 8; e += blockDim.x) {
    const int k = e >> 3, r = e & 7, i = row0 + r;
    const float* urow = a.u + ((size_t)k * D + i) * D;
    const float* muk = a.mu + (size_t)k * D;
    double b = 0.0;
    for (int j = i; j < D; ++j) b += (double)urow[j] * ((double)muk[j] - (double)mup[j]);
    const float bf = (float)b;
    uint32_t ub = __float_as_uint(bf);
    ub += 0xFFFu + ((ub >> 13) & 1u);
    ub &= 0xFFFFE000u;                                // round to TF32 (10 explicit mantissa bits)
    const float hi = (bf == bf && fabsf(bf) < CUDART_INF_F) ? __uint_as_float(ub) : bf;
    const float lo = bf - hi;
    const int ch = k < a.n0 ? 0 : 1 + (k - a.n0) / 16, slot = k < a.n0 ? k : (k - a.n0) % 16;
    float* q = out + ch * 512 + (16 * (slot >> 1) + 2 * r + (slot & 1)) * 4;
    q[0] = -hi;
    q[1] = -lo;
  }
}
