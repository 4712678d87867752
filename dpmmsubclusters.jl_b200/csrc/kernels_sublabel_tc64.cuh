// Stage 2b on the tensor cores for D = 64 (NIW): the sub-label draw of every point, one pass over the label-sorted
// points -- the GEMM1 half of kernels_substats_tc.cuh at twice the width (the statistics follow in
// kernels_stats_tc64.cuh: at D = 64 one CTA cannot hold the permuted panels of the fused scheme).
//
//   sample_sub_clusters_worker! / create_subclusters_labels!   src/local_clusters_actions.jl:70-95
//   log_likelihood!(mv_gaussian)                               src/distributions/mv_gaussian.jl:21-25
//   sample_log_cat_array! (C = 2)                              src/utils.jl:19-31
//
// A tile = 128 consecutive positions of ONE cluster k in `perm`.  Its 256-byte rows are gathered with cp.async into
// two K-major [128][32-feature] halves (128B swizzle), shifted by the cluster's centre c_k (exact in Float32, see
// niw_pack_center) in place -- the tensor core reads h = the TF32 bits of z -- and l = z - h goes to a second pair.
// The epilogue also partitions every label segment of perm2 into its left | right points for the statistics kernel.
//       Y[128 x 128] = h Wh' + l Wh' + h Wl' - b,   W = [U_left; U_right] = Wh + Wl,   b = U_s (mu_s - c_k)
// (25 tcgen05.mma kind::tf32, M = 128, N = 128: 8 k-steps per term + the bias k-step; only l Wl' is dropped).
// Accumulator row p holds U_l (x_p - mu_l) | U_r (x_p - mu_r); the epilogue warps read it with tcgen05.ld, form
// q = |y|^2, r = -c - q/2 + log w (the reference's Float32 final operations) and draw with the reference's
// inverse-CDF rule and the same Philox uniform as every other sub-label kernel.
//
// Warp roles (768 threads, one CTA per SM, contiguous range of the tile sequence per CTA):
//   warps 0-7 gather + centre + split      warp 8 GEMM issuer      warp 9 stages the next cluster's factors
//   warps 12-15 / 16-19 / 20-23 epilogue of the tiles li = 0 / 1 / 2 (mod 3)        (warps 10-11 idle)
// Shared memory: z ring 3 x 32 KB, l 32 KB, Wh | Wl 64 KB, bias operands 8 KB.  TMEM: 4 x 128 columns.
#pragma once
#include "kernels_substats_tc.cuh"

#define L64_D 64
#define L64_TILE 128
#define L64_ZR 3
#define L64_PF 2
#define L64_NG 3
#define L64_NTB 4
#define L64_THREADS 768
#define L64_GATHER 256
#define L64_HALF 16384                       // one [128][32] K-major half
#define L64_PANEL (2 * L64_HALF)             // a [128][64] operand
#define L64_TMEM_COLS 512

struct SubLabel64Args {
  const float* x;
  int64_t n;
  int K;
  const int32_t* perm;     // [n] point indices sorted by label
  const int32_t* seg_off;  // [K+1]
  const float* w;          // [K][2][64][64] rows of U_left, U_right
  const float* bias;       // [K][2][64]     U_s (mu_s - c_k)
  const float* cen;        // [K][64]        c_k
  const float* cst;        // [3K]
  const float* loglr;      // [2K]
  uint8_t* sub;            // [n] out
  int32_t* perm2;          // [n] out: every label segment partitioned into its left | right points
  int32_t* cursor;         // [2K] in/out: (first free left position, end of the free right positions) per cluster
  const double* u_inj;
  uint64_t seed;
  uint32_t call;
  int64_t goff;
  float* dump;             // optional [2][n]
};

struct SubLabel64Smem {
  size_t z, l, w, aaug, baug, bnd, pre, bars, slot, total;
  __host__ __device__ explicit SubLabel64Smem(int K) {
    size_t o = 0;
    z = o;     o += (size_t)L64_ZR * L64_PANEL;
    l = o;     o += L64_PANEL;
    w = o;     o += 2 * L64_PANEL;               // Wh | Wl, [128 rows (side, i)][64] each
    aaug = o;  o += 4096;
    baug = o;  o += 4096;
    bnd = o;   o += (size_t)(K + 1) * 4;
    pre = o;   o += (size_t)(K + 1) * 4;
    o = (o + 15) & ~(size_t)15;
    bars = o;  o += 32 * 8;
    slot = o;  o += 16;
    total = o;
  }
};

__global__ void __launch_bounds__(L64_THREADS, 1) niw_sublabel_tc64_kernel(const SubLabel64Args a) {
  extern __shared__ __align__(1024) uint8_t l64_smem[];
  const SubLabel64Smem L(a.K);
  uint8_t* z0 = l64_smem + L.z;
  uint8_t* lp = l64_smem + L.l;
  uint8_t* wsm = l64_smem + L.w;
  float* aaug = reinterpret_cast<float*>(l64_smem + L.aaug);
  float* baug = reinterpret_cast<float*>(l64_smem + L.baug);
  int32_t* B = reinterpret_cast<int32_t*>(l64_smem + L.bnd);
  int32_t* P = reinterpret_cast<int32_t*>(l64_smem + L.pre);
  uint64_t* bars = reinterpret_cast<uint64_t*>(l64_smem + L.bars);
  uint64_t* ready = bars;             // [1]  z (in place) and l of the tile written and visible
  uint64_t* sfree = bars + 1;         // [1]  l panel read by the GEMM
  uint64_t* zfree = bars + 2;         // [3]  z slot read by the GEMM
  uint64_t* d1full = bars + 8;        // [4]  accumulator complete
  uint64_t* d1empty = bars + 12;      // [4]  ... read by the epilogue
  uint64_t* wfull = bars + 16;        // factors of a cluster staged
  uint64_t* wempty = bars + 17;       // last GEMM of the cluster retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(l64_smem + L.slot);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkeys = a.K;

  for (int j = tid; j <= nkeys; j += L64_THREADS) B[j] = __ldg(a.seg_off + j);
  for (int e = tid; e < 1024; e += L64_THREADS) aaug[e] = baug[e] = 0.f;
  if (tid == 0) {
    tc::mbar_init(ready, L64_GATHER);
    tc::mbar_init(sfree, 1);
    for (int s = 0; s < L64_ZR; ++s) tc::mbar_init(&zfree[s], 1);
    for (int b = 0; b < L64_NTB; ++b) {
      tc::mbar_init(&d1full[b], 1);
      tc::mbar_init(&d1empty[b], 128);
    }
    tc::mbar_init(wfull, 32);
    tc::mbar_init(wempty, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  for (int r = tid; r < L64_TILE; r += L64_THREADS) {   // bias k-step A operand: (1, 1, 0, ...) per row
    float* p = aaug + (r >> 3) * 64 + (r & 7) * 4;
    p[0] = 1.f;
    p[1] = 1.f;
  }
  if (warp == 0) {   // exclusive prefix of tiles per cluster
    int carry = 0;
    if (lane == 0) P[0] = 0;
    for (int base = 0; base < nkeys; base += 32) {
      const int j = base + lane;
      int v = j < nkeys ? (B[j + 1] - B[j] + L64_TILE - 1) / L64_TILE : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (j < nkeys) P[j + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, L64_TMEM_COLS);
  tc::fence_proxy_async();
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
      const int c16 = tid & 15, r0 = tid >> 4;   // 16-byte chunk of the 256-byte row; rows r0 + 16 j
      const int cc = c16 & 7;
      // K-major halves with the 128B swizzle: 16-byte chunk index XORed with (row & 7); (r0 + 16 j) & 7 == r0 & 7
      const uint32_t off0 = (uint32_t)((c16 >> 3) * L64_HALF + r0 * 128 + ((cc ^ (r0 & 7)) << 4));
      StcWalk wl, wc;
      stc_walk_init(wl, B, P, nkeys, t0, t1);
      wc = wl;
      int idx[8];
      auto load_idx = [&]() {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int p = wl.pos + r0 + 16 * j;
          idx[j] = p < wl.end ? __ldg(a.perm + p) : -1;
        }
      };
      auto issue = [&](int s) {
        uint8_t* dst = z0 + (size_t)s * L64_PANEL + off0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = idx[j] >= 0;
          cp_async16(dst + j * 2048, a.x + (size_t)(ok ? idx[j] : 0) * L64_D + 4 * c16, ok ? 16 : 0);
        }
      };
#pragma unroll
      for (int li = 0; li < L64_PF; ++li) {
        if (li < nt) {
          load_idx();
          issue(li);
          stc_advance(wl, B);
        }
        cp_async_commit();
      }
      if (L64_PF < nt) load_idx();
      auto load_center = [&](int key) { return __ldg(reinterpret_cast<const float4*>(a.cen + (size_t)key * L64_D) + c16); };
      int ckey = wc.key;
      float4 cen = load_center(ckey);
      for (int li = 0; li < nt; ++li) {
        if (wc.key != ckey) {
          ckey = wc.key;
          cen = load_center(ckey);
        }
        const int npts = wc.end - wc.pos;   // rows >= npts are zero padding
        cp_async_wait_group<L64_PF - 1>();
        uint8_t* src = z0 + (size_t)(li % L64_ZR) * L64_PANEL + off0;
        float4 lo[8];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = *reinterpret_cast<const float4*>(src + (4 * hb + j) * 2048);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (r0 + 16 * (4 * hb + j) < npts) {
              v[j].x -= cen.x; v[j].y -= cen.y; v[j].z -= cen.z; v[j].w -= cen.w;
            }
            *reinterpret_cast<float4*>(src + (4 * hb + j) * 2048) = v[j];   // z, in place
            lo[4 * hb + j].x = v[j].x - tc::trunc_tf32(v[j].x); lo[4 * hb + j].y = v[j].y - tc::trunc_tf32(v[j].y);
            lo[4 * hb + j].z = v[j].z - tc::trunc_tf32(v[j].z); lo[4 * hb + j].w = v[j].w - tc::trunc_tf32(v[j].w);
          }
        }
        tc::mbar_wait(sfree, (li & 1) ^ 1);     // the GEMM of tile li - 1 has read the l panel
        uint8_t* lk = lp + off0;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(lk + j * 2048) = lo[j];
        tc::fence_proxy_async();
        tc::mbar_arrive(ready);
        // next gather: tile li + PF goes into the slot of tile li + PF - ZR once its GEMM has read it
        const int ln = li + L64_PF;
        if (ln < nt) {
          const int q = ln / L64_ZR;           // q-th use of the slot: wait for the GEMM of its (q - 1)-th tile
          if (q > 0) tc::mbar_wait(&zfree[ln % L64_ZR], (q - 1) & 1);
          issue(ln % L64_ZR);
          stc_advance(wl, B);
          if (ln + 1 < nt) load_idx();
        }
        cp_async_commit();
        stc_advance(wc, B);
      }
    } else if (warp == 8) {
      // ======================= GEMM issuer (warp-uniform loop, one elected lane issues) =======================
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      StcWalk wm;
      stc_walk_init(wm, B, P, nkeys, t0, t1);
      const uint32_t idesc = tc::idesc_tf32(128);
      const uint64_t aaug_desc = tc::smem_desc_k_noswz(tc::smem_u32(aaug));
      const uint64_t baug_desc = tc::smem_desc_k_noswz(tc::smem_u32(baug));
      const uint64_t z_desc = tc::smem_desc_k128(tc::smem_u32(z0)), l_desc = tc::smem_desc_k128(tc::smem_u32(lp));
      const uint64_t wh_desc = tc::smem_desc_k128(tc::smem_u32(wsm)), wl_desc = tc::smem_desc_k128(tc::smem_u32(wsm) + L64_PANEL);
      int kj = -1, prevkey = -1;
      for (int li = 0; li < nt; ++li) {
        if (wm.key != prevkey) {
          prevkey = wm.key;
          ++kj;
          tc::mbar_wait(wfull, kj & 1);
        }
        const int tb = li % L64_NTB;
        tc::mbar_wait(ready, li & 1);
        tc::mbar_wait(&d1empty[tb], ((li / L64_NTB) & 1) ^ 1);
        tc::tc_fence_after();
        const uint64_t zd = z_desc + (uint64_t)((li % L64_ZR) * (L64_PANEL >> 4));
        const uint32_t tmem_d = tmem_u + tb * 128;
        // k-step ks of half hf: + hf * 16 KB + ks * 32 B (descriptor units of 16 bytes)
        tc::umma_tf32_first_w(tmem_d, zd, wh_desc, idesc);
#pragma unroll
        for (int ks = 1; ks < 8; ++ks) {
          const uint64_t o = (uint64_t)((ks >> 2) * (L64_HALF >> 4) + (ks & 3) * 2);
          tc::umma_tf32_acc_w(tmem_d, zd + o, wh_desc + o, idesc);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t o = (uint64_t)((ks >> 2) * (L64_HALF >> 4) + (ks & 3) * 2);
          tc::umma_tf32_acc_w(tmem_d, zd + o, wl_desc + o, idesc);
        }
        tc::umma_commit_w(&zfree[li % L64_ZR]);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t o = (uint64_t)((ks >> 2) * (L64_HALF >> 4) + (ks & 3) * 2);
          tc::umma_tf32_acc_w(tmem_d, l_desc + o, wh_desc + o, idesc);
        }
        tc::umma_tf32_acc_w(tmem_d, aaug_desc, baug_desc, idesc);   // Y -= b
        tc::umma_commit_w(sfree);
        tc::umma_commit_w(&d1full[tb]);
        if (wm.pos + L64_TILE >= wm.end) tc::umma_commit_w(wempty);   // last tile of the cluster
        stc_advance(wm, B);
      }
    } else if (warp == 9) {
      // ======================= factor staging =======================
      StcWalk wp;
      stc_walk_init(wp, B, P, nkeys, t0, t1);
      int kj = 0, prevkey = -1;
      float* wh = reinterpret_cast<float*>(wsm);
      float* wlo = wh + L64_PANEL / 4;
      for (int li = 0; li < nt; ++li) {
        if (wp.key != prevkey) {
          prevkey = wp.key;
          tc::mbar_wait(wempty, (kj & 1) ^ 1);   // every GEMM of the previous cluster has retired
          const float4* src = reinterpret_cast<const float4*>(a.w + (size_t)wp.key * 2 * L64_D * L64_D);
          for (int e = lane; e < 2048; e += 32) {
            const int r = e >> 4, c16 = e & 15;   // row (side, i), 16-byte chunk of its 256 bytes
            const float4 v = __ldg(src + e);
            float4 hi, lo;
            hi.x = tc::to_tf32(v.x); hi.y = tc::to_tf32(v.y); hi.z = tc::to_tf32(v.z); hi.w = tc::to_tf32(v.w);
            lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
            const int o = (c16 >> 3) * (L64_HALF / 4) + r * 32 + (((c16 & 7) ^ (r & 7)) << 2);
            *reinterpret_cast<float4*>(wh + o) = hi;
            *reinterpret_cast<float4*>(wlo + o) = lo;
          }
          for (int r = lane; r < 128; r += 32) {
            const float bv = __ldg(a.bias + (size_t)wp.key * 128 + r);
            const float bhi = tc::to_tf32(bv), blo = bv - bhi;
            float* p = baug + (r >> 3) * 64 + (r & 7) * 4;
            p[0] = -bhi;
            p[1] = -blo;
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(wfull);
          ++kj;
        }
        stc_advance(wp, B);
      }
    } else if (warp >= 12) {
      // ======================= epilogue: q, r, draw =======================
      const int g = (warp - 12) >> 2;                 // group g owns the tiles li = g (mod 3)
      const int sub = warp & 3;                       // TMEM sub-partition of this warp
      const int row = (sub << 5) | lane;              // TMEM lane == row of the tile
      StcWalk we;
      stc_walk_init(we, B, P, nkeys, t0, t1);
      int ckey = -1;
      float cl = 0.f, cr = 0.f, lwl = 0.f, lwr = 0.f;
      for (int li = 0; li < nt; ++li) {
        if (li % L64_NG != g) {
          stc_advance(we, B);
          continue;
        }
        const int key = we.key;
        const int npts = min(L64_TILE, we.end - we.pos);
        const bool valid = row < npts;
        const int32_t idx = valid ? __ldg(a.perm + we.pos + row) : 0;
        if (key != ckey) {
          ckey = key;
          cl = __ldg(a.cst + 3 * key + 1); cr = __ldg(a.cst + 3 * key + 2);
          lwl = __ldg(a.loglr + 2 * key); lwr = __ldg(a.loglr + 2 * key + 1);
        }
        double u = 0.0;
        if (valid) u = dpmm_uniform(a.u_inj, idx, a.seed, DPMM_STREAM_SUBLABEL, a.call, (uint64_t)(a.goff + idx));
        const int tb = li % L64_NTB;
        tc::mbar_wait(&d1full[tb], (li / L64_NTB) & 1);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + tb * 128;
        float ql, qr;
        {   // 32 columns at a time: two register arrays in flight would not fit next to the partition bookkeeping
          uint32_t v[32];
          tc::tmem_ld32(taddr, v);
          tc::tmem_ld_wait();
          ql = gauss_tc_screen_q(v);
          tc::tmem_ld32(taddr + 32, v);
          tc::tmem_ld_wait();
          ql += gauss_tc_screen_q(v);
          tc::tmem_ld32(taddr + 64, v);
          tc::tmem_ld_wait();
          qr = gauss_tc_screen_q(v);
          tc::tmem_ld32(taddr + 96, v);
          tc::tmem_ld_wait();
          qr += gauss_tc_screen_q(v);
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&d1empty[tb]);
        int side = 2;
        if (valid) {
          const float rl = gauss_finish(cl, ql, lwl), rr = gauss_finish(cr, qr, lwr);
          if (a.dump != nullptr) {
            a.dump[idx] = rl;
            a.dump[a.n + idx] = rr;
          }
          side = dpmm_draw_two(rl, rr, u);
          a.sub[idx] = (uint8_t)side;
        }
        {   // left / right partition of the segment (what the FP32 sub-label kernel leaves in perm2): the left points
            // grow from the segment's start, the right ones from its end; one atomic per warp and side
          const uint32_t bl = __ballot_sync(0xffffffffu, side == 0), br = __ballot_sync(0xffffffffu, side == 1);
          int basel = 0, baser = 0;
          if (lane == 0) {
            if (bl) basel = atomicAdd(a.cursor + 2 * key, __popc(bl));
            if (br) baser = atomicSub(a.cursor + 2 * key + 1, __popc(br)) - __popc(br);
          }
          basel = __shfl_sync(0xffffffffu, basel, 0);
          baser = __shfl_sync(0xffffffffu, baser, 0);
          const uint32_t lt = (1u << lane) - 1u;
          if (side == 0) a.perm2[basel + __popc(bl & lt)] = idx;
          else if (side == 1) a.perm2[baser + __popc(br & lt)] = idx;
        }
        stc_advance(we, B);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem_base, L64_TMEM_COLS);
}
