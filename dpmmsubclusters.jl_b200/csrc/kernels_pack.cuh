// Parameter packing on the device: broadcast_cluster_params (src/local_clusters_actions.jl:518-549)
// ships mv_gaussian fields (mu, invSigma, logdetSigma; mv_gaussian.jl:12-18); this kernel turns them
// into what the sweep kernels read.  One CTA per distribution:
//   invSigma (Float32, symmetrised) -> Float64 Cholesky  invSigma = L L'  -> U = L' rounded to Float32,
//   stored by columns (FMA path records) and, for the cluster distributions of the tensor-core path,
//   by rows (K-major B operand) with b = U mu and |U|_F;  c = (D^2 * Float32(log 2pi) + logdetSigma)/2.
// The reference carries the same factor (mv_gaussian.invChol, niw.jl:38-39).  A non positive definite
// or NaN input gives NaN factors, i.e. NaN log-likelihoods, exactly like the reference's arithmetic.
#pragma once
#include "common.cuh"
#include "kernels_gauss.cuh"   // gauss_col_off

struct NiwPackArgs {
  int D, K, rec_f, trip;
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
__global__ void __launch_bounds__(NIW_PACK_THREADS) niw_pack_kernel(const NiwPackArgs a) {
  extern __shared__ double Ls[];   // [D][D+1]
  const int D = a.D, LD = D + 1;
  const int t = blockIdx.x, tid = threadIdx.x, NT = NIW_PACK_THREADS;
  const float* A = a.inv_sigma + (size_t)t * D * D;
  for (int e = tid; e < D * D; e += NT) {
    const int i = e / D, j = e - i * D;
    Ls[i * LD + j] = 0.5 * ((double)A[(size_t)i * D + j] + (double)A[(size_t)j * D + i]);
  }
  __syncthreads();
  bool ok = true;
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
    a.cst[t] = __fmul_rn(__fadd_rn(__fmul_rn((float)(D * D), log2pi), a.logdet[t]), 0.5f);
  }
  if (a.tc_w != nullptr && t % 3 == 0) {
    const int k = t / 3;
    for (int e = tid; e < D * D; e += NT) {
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
