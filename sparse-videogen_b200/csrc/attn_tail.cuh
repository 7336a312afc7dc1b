// Transposed attention for SHORT TAILS of variable-block plans (sm_100a).
//
// A k-means cluster of r query rows needs ceil(r / 128) M=128 tiles in attn_fwd_kernel; the last one holds only
// r mod 128 rows (297 rows = 256 + 41), and tcgen05 has no cheaper M: M=64 costs the cycles of M=128.  What IS
// proportional to the tile size is the MMA N.  So a tail of <= 64 rows is computed transposed -- the KEYS of a chunk
// sit on M (TMEM lanes), the tail's query rows on N:
//
//     S^T [128 keys x Nq]  =  K_tile [128 x D]  *  Q_tail^T [D x Nq]          (SS, both K-major)          8 x (Nq/2) cycles
//     O^T [D=128 x Nq]    +=  V_tile^T [D x 128]  *  P^T [128 keys x Nq]      (SS, both MN-major)         8 x (Nq/2) cycles
//
// i.e. 384 tensor cycles per 128-key chunk at Nq = 48 instead of 1024, and 48 instead of 128 exponentials per key.
// The price: the softmax reductions run ACROSS lanes (a TMEM lane is a key).  Column maxima come from
// redux.sync over the warp + a 4-warp combine through shared memory; they only steer the lazy rescale (P is formed
// against a per-column reference max that moves when the running max grew by more than 2^8), and row sums are kept as
// per-thread partial sums over all chunks and reduced once in the epilogue.  P^T goes to shared memory (it is the B
// operand of the second MMA; only A may live in TMEM) in the MN-major 128-byte-swizzled image a V tile has.
//
// Measured variants (one B200 box, QC=400 / KC=1000 rho 0.30, profiles/r02_tail_ab.txt): this version 809 TF/s (M=128
// tiles only: 707); replacing the 48 redux by a halving shuffle butterfly + serving the two tails' MMAs in arrival order
// + splitting the exp sweep: 765 -- the per-chunk step of a tail pair stayed at 4.1-4.7 k cycles (timeline in
// profiles/r02_attn_timelines_summary.txt: column maxima ~1.1 k, barriers + combine ~0.45 k, exp + P stores 1.0-1.4 k),
// i.e. ~4.4 cycles per instruction with one softmax warp per tail and sub-partition: the step is bound by dependent-issue
// latency, more tails (warps) per CTA is the lever, not fewer instructions.
//
// One CTA = two tails (T0 / T1, each with its own rows, chunk list and K/V tiles through the common ring, same ring
// order as the dual items of attn_fwd_kernel).  12 warps: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 softmax of T0, warps 8-11 softmax of T1.  bf16 / fp16, D = 128.
#pragma once
#include "attn_kernel.cuh"

namespace svgb {

constexpr int kTailRows = 64;  // max query rows of a transposed tail (N of both MMAs)

struct TailBars {
  uint64_t q_full;
  uint64_t o_final;
  uint64_t s_full[2][2];   // [tail][S buffer]
  uint64_t p_full[2];      // 128 arrivals
  uint64_t pv_done[2];
  uint64_t kv_full[8];
  uint64_t kv_empty[8];
  uint32_t tmem_base;
  uint32_t pad_[3];
  float wmax[2][4][kTailRows];   // per tail, per warp: column maxima of the current chunk
  float mc[2][kTailRows];        // per tail: reference max * c of each query column (log2 units)
  float m_used[2][kTailRows];    // per tail: reference max (raw score units)
  float alpha[2][kTailRows];     // per tail: rescale factor of the current chunk (1 = unchanged)
  float lsum[2][4][kTailRows];   // epilogue: per-warp partial row sums
  int flag[2][2];                // per tail, double-buffered by chunk parity: some column was rescaled
};

static_assert(sizeof(TailBars) <= 6144, "TailBars outgrew its shared-memory slot");

template <int DT>
struct TailCfg {
  static constexpr int D = 128;
  static constexpr int kTileBytes = 128 * 256;            // K / V tile: 128 keys x 256 B (two 64-column panels)
  static constexpr int kPanelBytes = 128 * 128;
  static constexpr int kQtBytes = kTailRows * 256;        // Q tail: 64 rows x 256 B (two panels of 64 x 128 B)
  static constexpr int kQtPanelBytes = kTailRows * 128;
  static constexpr int kPBytes = 128 * 128;               // P^T: 128 keys x 64 columns x 2 B
  static constexpr int kStages = 4;                       // 227 KB limit: 2 x 16 (Q) + 2 x 16 (P) + 4 x 32 (ring) + 6 KB
  static constexpr int kBarBytes = 6144;
  static constexpr int kSmemBytes = 2 * kQtBytes + 2 * kPBytes + kStages * kTileBytes + kBarBytes;
  static constexpr int kThreads = 384;
  // TMEM columns: tail t -> S buffers at t*192 + {0, 64}, O^T at t*192 + 128
  static constexpr uint32_t kTmemCols = 512;
};

__device__ __forceinline__ float redux_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

template <int DT>
__global__ void __launch_bounds__(384, 1)
attn_tail_kernel(const __grid_constant__ CUtensorMap qmap64, const __grid_constant__ CUtensorMap kmap,
                 const __grid_constant__ CUtensorMap vmap, const AttnArgs args) {
  using Cfg = TailCfg<DT>;
  constexpr bool BF16 = DT != DT_F16;
  constexpr int D = 128;
  const int bh = blockIdx.y;
  const int n_items = args.item_count[bh * args.counts_stride];
  if (static_cast<int>(blockIdx.x) >= n_items) return;
  const size_t iidx = static_cast<size_t>(bh) * args.items_stride + blockIdx.x;
  const int4 it0 = args.items[iidx];
  const int4 it1 = args.items2[iidx];  // .y == 0: no second tail
  const int ntails = it1.y > 0 ? 2 : 1;
  const int n0 = it0.w, n1 = ntails == 2 ? it1.w : 0;
  const int2* __restrict__ ch0 = args.chunks + it0.z;
  const int2* __restrict__ ch1 = args.chunks + it1.z;
  const int nmax = max(n0, n1);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sQt = smem_base;                         // 2 x kQtBytes
  const uint32_t sP = sQt + 2 * Cfg::kQtBytes;            // 2 x kPBytes
  const uint32_t sRing = sP + 2 * Cfg::kPBytes;
  TailBars* bars = reinterpret_cast<TailBars*>(smem_raw + 2 * Cfg::kQtBytes + 2 * Cfg::kPBytes + Cfg::kStages * Cfg::kTileBytes);
  constexpr int kStages = Cfg::kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef SVGB_ATTN_TRACE
  const bool trace_on = bh == 0 && blockIdx.x == SVGB_ATTN_TRACE && (warp == 1 || warp == 4 || warp == 8);
#endif
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&qmap64);
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&vmap);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(smem_u32(&bars->q_full), 1);
    mbar_init(smem_u32(&bars->o_final), 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->s_full[t][0]), 1);
      mbar_init(smem_u32(&bars->s_full[t][1]), 1);
      mbar_init(smem_u32(&bars->p_full[t]), 128);
      mbar_init(smem_u32(&bars->pv_done[t]), 1);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), 1);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(smem_u32(&bars->tmem_base));
  if (threadIdx.x >= 128 && threadIdx.x < 128 + 2 * kTailRows) {  // softmax state
    const int t = (threadIdx.x - 128) / kTailRows, qq = (threadIdx.x - 128) % kTailRows;
    bars->m_used[t][qq] = -INFINITY;
    bars->mc[t][qq] = 0.f;
    bars->alpha[t][qq] = 1.f;
    if (qq < 2) bars->flag[t][qq] = 0;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<kRegsLight>();
    if (lane == 0) {
      const uint32_t qbar = smem_u32(&bars->q_full);
      mbar_expect_tx(qbar, ntails * Cfg::kQtBytes);
      for (int t = 0; t < ntails; ++t)
        for (int h = 0; h < 2; ++h)
          tma_load_3d(sQt + t * Cfg::kQtBytes + h * Cfg::kQtPanelBytes, &qmap64, qbar, h * 64, t == 0 ? it0.x : it1.x, bh);
      int it = 0;
      auto load = [&](const CUtensorMap* map, int kv0) {
        const int slot = it % kStages;
        mbar_wait(smem_u32(&bars->kv_empty[slot]), ((it / kStages) & 1) ^ 1, 1);
        const uint32_t fb = smem_u32(&bars->kv_full[slot]);
        mbar_expect_tx(fb, Cfg::kTileBytes);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(sRing + slot * Cfg::kTileBytes + h * Cfg::kPanelBytes, map, fb, h * 64, kv0, bh);
        ++it;
      };
      // ring order (the MMA issuer consumes in exactly this order): K0(0) K1(0) | V0(j) K0(j+1) V1(j) K1(j+1) | ...
      if (n0 > 0) load(&kmap, __ldg(&ch0[0].x));
      if (n1 > 0) load(&kmap, __ldg(&ch1[0].x));
      for (int j = 0; j < nmax; ++j) {
        if (j < n0) {
          load(&vmap, __ldg(&ch0[j].x));
          if (j + 1 < n0) load(&kmap, __ldg(&ch0[j + 1].x));
        }
        if (j < n1) {
          load(&vmap, __ldg(&ch1[j].x));
          if (j + 1 < n1) load(&kmap, __ldg(&ch1[j + 1].x));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    setmaxnreg_dec<kRegsLight>();
    if (nmax > 0 && elect_one()) {
      const int nq0 = (it0.y + 15) & ~15, nq1 = (it1.y + 15) & ~15;
      auto chunk_n = [&](int t, int jj) -> int {  // MMA K extent of the second MMA: valid keys rounded to 16
        const int vld = chunk_valid(__ldg(t == 0 ? &ch0[jj].y : &ch1[jj].y));
        return (vld + 15) & ~15;
      };
      // S^T_t(buf) = K_tile * Q_t^T : A = K tile (K-major, 128 rows), B = Q tail (K-major, Nq rows)
      auto issue_qk = [&](int t, int buf, int slot) {
        const uint32_t idesc = make_idesc(128, t == 0 ? nq0 : nq1, DT == DT_BF16, false, false);
        const uint32_t d_tmem = tmem + t * 192 + buf * 64;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t a_off = (kk >> 2) * Cfg::kPanelBytes + (kk & 3) * 32;
          const uint32_t b_off = (kk >> 2) * Cfg::kQtPanelBytes + (kk & 3) * 32;
          mma_ss(d_tmem, desc_kmajor_sw128(sRing + slot * Cfg::kTileBytes + a_off),
                 desc_kmajor_sw128(sQt + t * Cfg::kQtBytes + b_off), idesc, kk > 0 ? 1u : 0u);
        }
      };
      // O^T_t += V_tile^T * P_t^T : A = V tile (MN-major: M = d, two 64-d panels 16 KB apart), B = P^T (MN-major, N = q)
      auto issue_pv = [&](int t, int slot, int nkeys, bool acc) {
        const uint32_t idesc = make_idesc(128, t == 0 ? nq0 : nq1, DT == DT_BF16, true, true);
        const uint32_t d_tmem = tmem + t * 192 + 128;
        const int nk = nkeys / 16;
        for (int kk = 0; kk < nk; ++kk) {
          mma_ss(d_tmem, desc_mnmajor_sw128(sRing + slot * Cfg::kTileBytes + kk * 16 * 128, Cfg::kPanelBytes),
                 desc_mnmajor_sw128(sP + t * Cfg::kPBytes + kk * 16 * 128, Cfg::kPBytes), idesc, (acc || kk > 0) ? 1u : 0u);
        }
      };
      mbar_wait(smem_u32(&bars->q_full), 0, 2);
      int ring = 0;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if ((t == 0 ? n0 : n1) > 0) {
          const int slot = ring % kStages;
          mbar_wait(smem_u32(&bars->kv_full[slot]), (ring / kStages) & 1, 3);
          ++ring;
          tc_fence_after();
          issue_qk(t, 0, slot);
          tc_commit(smem_u32(&bars->s_full[t][0]));
          tc_commit(smem_u32(&bars->kv_empty[slot]));
        }
      }
      for (int j = 0; j < nmax; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int nt = t == 0 ? n0 : n1;
          if (j >= nt) continue;
          const bool has_next = j + 1 < nt;
          const int vslot = ring % kStages;
          const uint32_t vph = (ring / kStages) & 1;
          ++ring;
          int kslot = 0;
          uint32_t kph = 0;
          if (has_next) {
            kslot = ring % kStages;
            kph = (ring / kStages) & 1;
            ++ring;
          }
          // S^T(j+1) first: its buffer was read out by the softmax of chunk j-1 (whose P arrival we already saw), so
          // the scores of the next chunk are ready long before the softmax of chunk j ends
          if (has_next) {
            mbar_wait(smem_u32(&bars->kv_full[kslot]), kph, 6);
            tc_fence_after();
            issue_qk(t, (j + 1) & 1, kslot);
            tc_commit(smem_u32(&bars->s_full[t][(j + 1) & 1]));
            tc_commit(smem_u32(&bars->kv_empty[kslot]));
          }
          mbar_wait(smem_u32(&bars->kv_full[vslot]), vph, 4);
          mbar_wait(smem_u32(&bars->p_full[t]), j & 1, 5 + 2 * t);
          tc_fence_after();
          issue_pv(t, vslot, chunk_n(t, j), j > 0);
          tc_commit(smem_u32(&bars->pv_done[t]));
          tc_commit(smem_u32(&bars->kv_empty[vslot]));
        }
      }
      tc_commit(smem_u32(&bars->o_final));
    }
  } else if (warp < 4) {
    setmaxnreg_dec<kRegsLight>();
  } else {
    // ------------------------------------------------------------------ softmax (keys on lanes) + epilogue
    setmaxnreg_inc<kRegsSoftmax>();
    const int t = (warp - 4) >> 2;
    if (t < ntails) {
      const int wq = warp & 3;
      const int key = wq * 32 + lane;  // key inside the chunk == TMEM lane of S^T; output dim d of O^T in the epilogue
      const int4 itm = t == 0 ? it0 : it1;
      const int nrows = itm.y, my_n = itm.w;
      const int nq = (nrows + 15) & ~15;
      const int2* __restrict__ my_chunks = t == 0 ? ch0 : ch1;
      const float c = args.scale_log2;
      const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16) + t * 192;
      const uint32_t o_addr = lane_addr + 128;
      const uint32_t barid = 1 + t;
      auto wg_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory"); };
      float* wmax = &bars->wmax[t][wq][0];
      float* mc_s = &bars->mc[t][0];
      float* mu_s = &bars->m_used[t][0];
      float* al_s = &bars->alpha[t][0];
      const uint32_t p_row = sP + t * Cfg::kPBytes + key * 128;
      float l_part[kTailRows];
#pragma unroll
      for (int i = 0; i < kTailRows; ++i) l_part[i] = 0.f;

      for (int j = 0; j < my_n; ++j) {
        const int valid = chunk_valid(__ldg(&my_chunks[j].y));
        const bool live = key < valid;  // keys past `valid` belong to unselected blocks (or lie past the sequence)
        SVGB_TRACE(t, j, 0);
        mbar_wait(smem_u32(&bars->s_full[t][j & 1]), (j >> 1) & 1, 8 + t);
        tc_fence_after();
        SVGB_TRACE(t, j, 1);
        uint32_t s0[32], s1[32];
        tmem_ld32(lane_addr + (j & 1) * 64, s0);
        if (nq > 32) tmem_ld32(lane_addr + (j & 1) * 64 + 32, s1);
        tc_wait_ld();
        SVGB_TRACE(t, j, 2);
        if (!live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            s0[i] = 0xff800000u;
            s1[i] = 0xff800000u;
          }
        }
        // ---- column maxima over the 128 keys: redux over the warp, 4-warp combine through shared memory
        {
          float keep0 = -INFINITY, keep1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float r = redux_max_f32(__uint_as_float(s0[i]));
            keep0 = lane == i ? r : keep0;
          }
          if (nq > 16) {
#pragma unroll
            for (int i = 16; i < 32; ++i) {
              const float r = redux_max_f32(__uint_as_float(s0[i]));
              keep0 = lane == i ? r : keep0;
            }
          }
          if (nq > 32) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float r = redux_max_f32(__uint_as_float(s1[i]));
              keep1 = lane == i ? r : keep1;
            }
          }
          if (nq > 48) {
#pragma unroll
            for (int i = 16; i < 32; ++i) {
              const float r = redux_max_f32(__uint_as_float(s1[i]));
              keep1 = lane == i ? r : keep1;
            }
          }
          wmax[lane] = keep0;
          wmax[32 + lane] = keep1;
        }
        SVGB_TRACE(t, j, 3);
        wg_sync();
        if (key < nq) {
          const float mx = fmaxf(fmaxf(bars->wmax[t][0][key], bars->wmax[t][1][key]),
                                 fmaxf(bars->wmax[t][2][key], bars->wmax[t][3][key]));
          const float mu = mu_s[key];
          const float mn = fmaxf(mu, mx);
          float a = 1.f;
          if ((mn - mu) * c > kRescaleTau) {  // false when both are -inf
            a = ex2_approx((mu - mn) * c);    // 0 when mu == -inf
            mu_s[key] = mn;
            mc_s[key] = mn * c;
            if (j > 0) bars->flag[t][j & 1] = 1;
          }
          al_s[key] = a;
        }
        if (key == 0) bars->flag[t][(j + 1) & 1] = 0;
        wg_sync();
        SVGB_TRACE(t, j, 4);
        if (j > 0) {
          // P^T(j-1) must have been consumed before it is overwritten (and O^T complete before a rescale)
          mbar_wait(smem_u32(&bars->pv_done[t]), (j - 1) & 1, 12);
          tc_fence_after();
          if (bars->flag[t][j & 1]) {
            // rescale: O^T[d = key][q] *= alpha[q], partial row sums likewise
            uint32_t o0[32], o1[32];
            tmem_ld32(o_addr, o0);
            if (nq > 32) tmem_ld32(o_addr + 32, o1);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              o0[i] = __float_as_uint(__uint_as_float(o0[i]) * al_s[i]);
              o1[i] = __float_as_uint(__uint_as_float(o1[i]) * al_s[32 + i]);
              l_part[i] *= al_s[i];
              l_part[32 + i] *= al_s[32 + i];
            }
            tmem_st32(o_addr, o0);
            if (nq > 32) tmem_st32(o_addr + 32, o1);
            tc_wait_st();
          }
        }
        // ---- P^T[key][q] = exp2(S*c - m_ref[q]*c) -> 16 bit -> shared memory, MN-major 128-byte-swizzled row `key`
        SVGB_TRACE(t, j, 5);
        // 16 columns (two 16-byte pieces of this key's row) at a time, only the nq / 16 live groups
        auto p_quarter = [&](const uint32_t(&ss)[32], int h, int qt) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = qt * 16 + 2 * i;  // column inside this 32-column half
            const float x0 = fmaf(__uint_as_float(ss[col]), c, -mc_s[h * 32 + col]);
            const float x1 = fmaf(__uint_as_float(ss[col + 1]), c, -mc_s[h * 32 + col + 1]);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            l_part[h * 32 + col] += p0;
            l_part[h * 32 + col + 1] += p1;
            pk[i] = pack2<BF16>(p0, p1);
          }
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int piece = h * 4 + qt * 2 + v;  // 16-byte piece (8 columns) of this key's 128-byte row
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((piece ^ (key & 7)) << 4)),
                         "r"(pk[4 * v]), "r"(pk[4 * v + 1]), "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                         : "memory");
          }
        };
        p_quarter(s0, 0, 0);
        if (nq > 16) p_quarter(s0, 0, 1);
        if (nq > 32) p_quarter(s1, 1, 0);
        if (nq > 48) p_quarter(s1, 1, 1);
        SVGB_TRACE(t, j, 6);
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->p_full[t]));
        SVGB_TRACE(t, j, 7);
      }

      // ---------------- epilogue: O^T / l -> 16 bit -> global; this thread owns output dim d = key of every row
      if (my_n > 0) {
        mbar_wait(smem_u32(&bars->o_final), 0, 10 + t);
        tc_fence_after();
      }
      // row sums: butterfly over the warp, then the 4 warps through shared memory
#pragma unroll
      for (int i = 0; i < kTailRows; ++i) {
        float v = l_part[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == (i & 31)) bars->lsum[t][wq][i] = v;
      }
      wg_sync();
      if (key < kTailRows) {
        const float l = bars->lsum[t][0][key] + bars->lsum[t][1][key] + bars->lsum[t][2][key] + bars->lsum[t][3][key];
        al_s[key] = l > 0.f ? 1.f / l : 0.f;
      }
      wg_sync();
      uint32_t o0[32], o1[32];
      if (my_n > 0) {
        tmem_ld32(o_addr, o0);
        if (nq > 32) tmem_ld32(o_addr + 32, o1);
        tc_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o0[i] = o1[i] = 0u;
      }
      uint16_t* obase = reinterpret_cast<uint16_t*>(args.o) + bh * args.o_head_stride + key;
      auto store_half = [&](const uint32_t(&oo)[32], int h) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int qi = h * 32 + i;
          if (qi < nrows) {
            long long out_row = itm.x + qi;
            if (args.o_rows) out_row = __ldg(&args.o_rows[static_cast<size_t>(bh) * args.S + itm.x + qi]);
            const float val = __uint_as_float(oo[i]) * al_s[qi];
            uint16_t bits;
            if constexpr (BF16) bits = __bfloat16_as_ushort(__float2bfloat16_rn(val));
            else bits = __half_as_ushort(__float2half_rn(val));
            obase[out_row * args.o_row_stride] = bits;
          }
        }
      };
      store_half(o0, 0);
      if (nq > 32) store_half(o1, 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem);
  }
}

}  // namespace svgb
