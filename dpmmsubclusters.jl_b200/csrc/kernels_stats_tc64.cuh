// Stage 3 on the tensor cores for D = 64 (NIW, all clusters): the D = 32 scheme of kernels_stats_tc.cuh with one
// instruction shape twice as large.
//
//   create_suff_stats_dict_worker            src/local_clusters_actions.jl:149-169
//   create_sufficient_statistics (NIW)       src/priors/niw.jl:42-51      N, sum x, S = X X' (Float64)
//
// A tile = 128 consecutive positions of ONE key (2k + side) of the partitioned permutation perm2.  Its rows
// (256 bytes) are gathered with cp.async into FOUR [128 points][32 features] panels -- features 0-31 | 32-63 of
// h, then of l -- each in the canonical MN-major layout of 32-bit operands (rows of 128 bytes, SWIZZLE_128B with
// 32-byte atomicity), one panel (16 KB) apart, so that a single descriptor with LBO = 16 KB addresses
//        A = [h0 | h1 | l0 | l1]  (M = 128)          B = [h0 | h1]  (N = 64)
// and one tcgen05.mma (kind::tf32, M = 128, N = 64, K = 8 points) does the rank-8 update
//        D[128 x 64] += [h | l]' . h        rows 0-63: sum h h'      rows 64-127: sum l h'
// y = x - c (centre of the run, stats_centers_kernel) is split y = h + l with h = the TF32 bits of y (what the
// tensor core reads of the word) and l = y - h exactly; S_ij = (hh')_ij + (lh')_ij + (lh')_ji drops l l'
// (<= 2^-20 relative).  The accumulator of at most S64_FLUSH tiles (1024 points) is added to the key's Float64
// record; sum y is accumulated by the gather warps; stats_finalize_kernel shifts back by c in Float64.
//
// Warp roles (416 threads, one CTA per SM): warps 0-7 gather + centre + split (thread = one 16-byte chunk of 8
// rows), warp 8 issues the MMAs, warps 9-12 drain the accumulator (TMEM lane = accumulator row).
#pragma once
#include "kernels_stats_tc.cuh"

#define S64_D 64
#define S64_TILE 128
#define S64_STAGES 3
#define S64_FLUSH 8
#define S64_THREADS 416
#define S64_GATHER 256
#define S64_PANEL 16384                      // one [128][32] Float32 panel
#define S64_STAGE_BYTES (4 * S64_PANEL)      // h0 | h1 | l0 | l1
#define S64_TMEM_COLS 128                    // two accumulators of 64 columns
#define S64_TLD 65

static_assert(S64_TILE == STC_TILE, "the tile walk of kernels_stats_tc.cuh is shared");

struct StatsTc64Smem {
  size_t stages, tbuf, bnd, pre, bars, slot, total;
  __host__ __device__ explicit StatsTc64Smem(int K) {
    size_t o = 0;
    stages = o; o += (size_t)S64_STAGES * S64_STAGE_BYTES;
    tbuf = o;   o += 64 * S64_TLD * 4;       // sum l h' of the group being drained (read transposed)
    o = (o + 15) & ~(size_t)15;
    bnd = o;    o += (size_t)(2 * K + 1) * 4;
    pre = o;    o += (size_t)(2 * K + 1) * 4;
    o = (o + 15) & ~(size_t)15;
    bars = o;   o += 16 * 8;
    slot = o;   o += 16;
    total = o;
  }
};

// centre of every run = mean of its first <= 32 points (as stats_centers_kernel, for 256-byte rows)
__global__ void __launch_bounds__(512) stats_centers64_kernel(const float* __restrict__ x, const int32_t* __restrict__ perm2,
                                                              const int32_t* __restrict__ seg_off,
                                                              const int32_t* __restrict__ lr_cursor, float* __restrict__ centers) {
  __shared__ float4 sm[32][16];
  const int key = blockIdx.x, k = key >> 1;
  const int mid = lr_cursor[2 * k];
  const int beg = (key & 1) ? mid : seg_off[k];
  const int end = (key & 1) ? seg_off[k + 1] : mid;
  const int cnt = min(32, end - beg);
  const int p = threadIdx.x >> 4, c = threadIdx.x & 15;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p < cnt) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)perm2[beg + p] * S64_D) + c);
  sm[p][c] = v;
  __syncthreads();
  if (threadIdx.x < S64_D) {
    const float* col = reinterpret_cast<const float*>(&sm[0][0]) + threadIdx.x;
    float sacc = 0.f;
    for (int q = 0; q < 32; ++q) sacc += col[q * S64_D];
    centers[(size_t)key * S64_D + threadIdx.x] = cnt > 0 ? sacc / (float)cnt : 0.f;
  }
}

__global__ void __launch_bounds__(S64_THREADS, 1) niw_stats_tc64_kernel(const StatsTcArgs a) {
  extern __shared__ __align__(1024) uint8_t s64_smem[];
  const StatsTc64Smem L(a.K);
  uint8_t* stage0 = s64_smem + L.stages;
  float* T = reinterpret_cast<float*>(s64_smem + L.tbuf);
  int32_t* B = reinterpret_cast<int32_t*>(s64_smem + L.bnd);
  int32_t* P = reinterpret_cast<int32_t*>(s64_smem + L.pre);
  uint64_t* ready = reinterpret_cast<uint64_t*>(s64_smem + L.bars);   // [3] tile split and visible to the MMA
  uint64_t* empty = ready + S64_STAGES;                               // [3] MMAs of the stage retired
  uint64_t* accfull = empty + S64_STAGES;                             // [2]
  uint64_t* accempty = accfull + 2;                                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s64_smem + L.slot);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkeys = 2 * a.K;

  for (int j = tid; j <= nkeys; j += S64_THREADS)
    B[j] = (j & 1) ? __ldg(a.lr_cursor + (j - 1)) : __ldg(a.seg_off + (j >> 1));
  if (tid == 0) {
    for (int s = 0; s < S64_STAGES; ++s) {
      tc::mbar_init(&ready[s], S64_GATHER);
      tc::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&accfull[b], 1);
      tc::mbar_init(&accempty[b], 128);
    }
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {   // exclusive prefix of tiles per key
    int carry = 0;
    if (lane == 0) P[0] = 0;
    for (int base = 0; base < nkeys; base += 32) {
      const int j = base + lane;
      int v = j < nkeys ? (B[j + 1] - B[j] + S64_TILE - 1) / S64_TILE : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (j < nkeys) P[j + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, S64_TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ntot = P[nkeys];
  const int t0 = (int)(((int64_t)ntot * blockIdx.x) / gridDim.x);
  const int t1 = (int)(((int64_t)ntot * (blockIdx.x + 1)) / gridDim.x);
  const int nt = t1 - t0;

  if (nt > 0) {
    if (warp < 8) {
      // ======================= gather + centre + split =======================
      const int c16 = tid & 15, r0 = tid >> 4;              // 16-byte chunk of the 256-byte row; rows r0 + 16 j
      const int cc = c16 & 7;                               // chunk inside its 32-feature panel
      // panel (c16 >> 3); 32-byte chunk index XORed with (row & 3), (r0 + 16 j) & 3 == r0 & 3
      const uint32_t off0 = (uint32_t)((c16 >> 3) * S64_PANEL + r0 * 128 + (((((cc >> 1) ^ (r0 & 3)) << 1) | (cc & 1)) << 4));
      StcWalk wl, wc;
      stc_walk_init(wl, B, P, nkeys, t0, t1);
      wc = wl;
      int idx[8];
      auto load_idx = [&]() {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int p = wl.pos + r0 + 16 * j;
          idx[j] = p < wl.end ? __ldg(a.perm2 + p) : -1;
        }
      };
      auto issue = [&](int s) {
        uint8_t* h = stage0 + (size_t)s * S64_STAGE_BYTES + off0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = idx[j] >= 0;
          cp_async16(h + j * 2048, a.x + (size_t)(ok ? idx[j] : 0) * S64_D + 4 * c16, ok ? 16 : 0);
        }
      };
#pragma unroll
      for (int li = 0; li < S64_STAGES - 1; ++li) {
        if (li < nt) {
          load_idx();
          issue(li);
          stc_advance<S64_FLUSH>(wl, B);
        }
        cp_async_commit();
      }
      if (S64_STAGES - 1 < nt) load_idx();
      float sx[4] = {0.f, 0.f, 0.f, 0.f};
      auto load_center = [&](int key) { return __ldg(reinterpret_cast<const float4*>(a.centers + (size_t)key * S64_D) + c16); };
      int ckey = wc.key;
      float4 cen = load_center(ckey);
      for (int li = 0; li < nt; ++li) {
        const int s = li % S64_STAGES;
        if (wc.key != ckey) {
          ckey = wc.key;
          cen = load_center(ckey);
        }
        const int npts = wc.end - wc.pos;                    // rows >= npts are zero padding
        cp_async_wait_group<S64_STAGES - 2>();               // this thread's chunks of tile li have landed
        uint8_t* h = stage0 + (size_t)s * S64_STAGE_BYTES + off0;
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(h + (4 * hb + j) * 2048);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (r0 + 16 * (4 * hb + j) < npts) {
              v[j].x -= cen.x; v[j].y -= cen.y; v[j].z -= cen.z; v[j].w -= cen.w;
            }
            float4 lo;
            lo.x = v[j].x - tc::trunc_tf32(v[j].x); lo.y = v[j].y - tc::trunc_tf32(v[j].y);
            lo.z = v[j].z - tc::trunc_tf32(v[j].z); lo.w = v[j].w - tc::trunc_tf32(v[j].w);
            *reinterpret_cast<float4*>(h + (4 * hb + j) * 2048) = v[j];                    // y: the tensor core reads h = its TF32 bits
            *reinterpret_cast<float4*>(h + 2 * S64_PANEL + (4 * hb + j) * 2048) = lo;
            sx[0] += v[j].x; sx[1] += v[j].y; sx[2] += v[j].z; sx[3] += v[j].w;
          }
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&ready[s]);
        // next gather: tile li + 2 goes into the stage tile li - 1 used, once its MMAs have retired
        const int ln = li + S64_STAGES - 1;
        if (ln < nt) {
          const int sn = ln % S64_STAGES;
          tc::mbar_wait(&empty[sn], ((ln / S64_STAGES) & 1) ^ 1);
          issue(sn);
          stc_advance<S64_FLUSH>(wl, B);
          if (ln + 1 < nt) load_idx();
        }
        cp_async_commit();
        if (stc_is_last<S64_FLUSH>(wc)) {   // sum y of the flush group -> Float64 accumulator
#pragma unroll
          for (int q = 0; q < 4; ++q) sx[q] += __shfl_xor_sync(0xffffffffu, sx[q], 16);
          if (lane < 16) {
            double* dst = a.acc + (size_t)wc.key * a.rec + 1 + 4 * lane;   // lane == c16 for lanes 0-15
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (sx[q] != 0.f) atomicAdd(dst + q, (double)sx[q]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) sx[q] = 0.f;
        }
        stc_advance<S64_FLUSH>(wc, B);
      }
    } else if (warp == 8) {
      // ======================= MMA issuer (warp-uniform loop, one elected lane issues) =======================
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      StcWalk wm;
      stc_walk_init(wm, B, P, nkeys, t0, t1);
      const uint32_t idesc = tc::idesc_tf32_mn_m128(S64_D);
      const uint64_t desc0 = tc::smem_desc_mn128(tc::smem_u32(stage0), S64_PANEL);
      int buf = 0;
      uint32_t uses0 = 0u, uses1 = 0u;
      for (int li = 0; li < nt; ++li) {
        const int s = li % S64_STAGES;
        const bool first = wm.gcount == 0, last = stc_is_last<S64_FLUSH>(wm);
        if (first) {
          const uint32_t u = buf ? uses1 : uses0;
          tc::mbar_wait(&accempty[buf], (u & 1) ^ 1);   // the drain of this accumulator's previous group
          if (buf) ++uses1; else ++uses0;
        }
        tc::mbar_wait(&ready[s], (li / S64_STAGES) & 1);
        tc::tc_fence_after();
        // one k-step = 8 points = 1024 bytes of every panel
        const uint64_t desc = desc0 + (uint64_t)(s * (S64_STAGE_BYTES >> 4));
        const uint32_t tmem_d = tmem_u + buf * 64;
        const int npts = min(S64_TILE, wm.end - wm.pos);
        const int nks = (npts + 7) >> 3;
        if (first) tc::umma_tf32_first_w(tmem_d, desc, desc, idesc);
        else tc::umma_tf32_acc_w(tmem_d, desc, desc, idesc);
        if (nks == S64_TILE / 8) {
#pragma unroll
          for (int ks = 1; ks < S64_TILE / 8; ++ks) tc::umma_tf32_acc_w(tmem_d, desc + ks * 64, desc + ks * 64, idesc);
        } else {
          for (int ks = 1; ks < nks; ++ks) tc::umma_tf32_acc_w(tmem_d, desc + ks * 64, desc + ks * 64, idesc);
        }
        tc::umma_commit_w(&empty[s]);
        if (last) {
          tc::umma_commit_w(&accfull[buf]);
          buf ^= 1;
        }
        stc_advance<S64_FLUSH>(wm, B);
      }
    } else {
      // ======================= accumulator drain =======================
      const int sub = warp & 3;                 // TMEM sub-partition of this warp
      const int m = (sub << 5) | lane;          // accumulator row: 0-63 sum h h', 64-127 sum l h'
      StcWalk we;
      stc_walk_init(we, B, P, nkeys, t0, t1);
      int buf = 0;
      uint32_t uses0 = 0u, uses1 = 0u;
      for (int li = 0; li < nt; ++li) {
        if (stc_is_last<S64_FLUSH>(we)) {
          const uint32_t u = buf ? uses1 : uses0;
          tc::mbar_wait(&accfull[buf], u & 1);
          if (buf) ++uses1; else ++uses0;
          tc::tc_fence_after();
          uint32_t v0[32], v1[32];
          const uint32_t taddr = tmem_base + buf * 64 + ((uint32_t)(sub * 32) << 16);
          tc::tmem_ld32(taddr, v0);
          tc::tmem_ld32(taddr + 32, v1);
          tc::tmem_ld_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(&accempty[buf]);
          if (m >= 64) {
            float* trow = T + (m - 64) * S64_TLD;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              trow[j] = __uint_as_float(v0[j]);
              trow[32 + j] = __uint_as_float(v1[j]);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (m < 64) {
            double* dst = a.acc + (size_t)we.key * a.rec + 1 + S64_D + (size_t)m * S64_D;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              if (j >= m) {
                const float hh = __uint_as_float(j < 32 ? v0[j & 31] : v1[j & 31]);
                const float sv = (hh + T[m * S64_TLD + j]) + T[j * S64_TLD + m];
                if (sv != 0.f) atomicAdd(dst + j, (double)sv);
              }
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          buf ^= 1;
        }
        stc_advance<S64_FLUSH>(we, B);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_base, S64_TMEM_COLS);
}
