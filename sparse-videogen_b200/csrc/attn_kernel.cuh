// Block-sparse flash attention forward for sm_100a.
//
// One CTA per work item (<= 256 query rows of one head = two 128-row tiles T0,T1 that share every
// K/V tile).  Warp roles (12 warps, 1 CTA / SM):
//   warp 0      TMA producer : Q once, then K(0) V(0) K(1) V(1) ... through one smem ring
//   warp 1      MMA issuer   : S_t = Q_t K^T (SS, K-major, tcgen05) ; O_t += P_t V (TS: P from TMEM,
//                              V MN-major from smem).  Issue order  PV0(j) QK0(j+1) PV1(j) QK1(j+1)
//                              so the softmax of one tile overlaps the MMAs of the other.
//   warp 2      TMEM allocator (512 columns: S0 | S1 | O0 | O1), otherwise idle
//   warps 4-7   softmax + correction + epilogue for T0 (one thread per query row)
//   warps 8-11  same for T1
// P (bf16/fp16) overwrites the first half of its S tile in TMEM; O is rescaled lazily (only when the
// running max grows by more than 2^8) by the row's own softmax thread, which is safe because the
// S_t(j) commit also covers PV_t(j-1).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "attn_common.cuh"
#include "ptx.cuh"

namespace svgb {

// element type of Q/K/V (and of P): 0 = bf16, 1 = fp16, 2 = fp8 e4m3 (output bf16)
constexpr int DT_BF16 = 0, DT_F16 = 1, DT_E4M3 = 2;

template <int D, int DT = DT_BF16>
struct AttnCfg {
  static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
  static_assert(DT != DT_E4M3 || D == 128, "the e4m3 path is built for head_dim 128 (128-byte rows)");
  static constexpr int kElemBytes = DT == DT_E4M3 ? 1 : 2;
  static constexpr int kRowBytes = D * kElemBytes;
  static constexpr int kHalves = kRowBytes / 128;        // 128-byte swizzle panels per row
  static constexpr int kPanelElems = 128 / kElemBytes;   // elements per panel row (TMA box width)
  static constexpr int kMmaK = 32 / kElemBytes;          // K elements per tcgen05.mma (32 bytes)
  static constexpr int kPanelBytes = 128 * 128;          // 128 rows x 128 B
  static constexpr int kTileBytes = kPanelBytes * kHalves;
  static constexpr int kStages = (kRowBytes == 256) ? 5 : 8;
  // gather instantiation: one stage less; the freed tile holds the item's run table in shared memory
  static constexpr int kStagesGather = kStages;
  static constexpr int kMaxRunsSmem = kTileBytes / 8 - 1;
  static constexpr int kQBytes = 2 * kTileBytes;
  static constexpr int kRingBytes = kStages * kTileBytes;
  static constexpr int kBarBytes = 3072;
  // no alignment slack: the dynamic window starts 1 KB into the CTA's shared memory (driver-reserved), i.e.
  // 1024-byte aligned; the kernel traps if that ever stops being true
  static constexpr int kSmemBytes = kQBytes + kRingBytes + kBarBytes;
  static constexpr int kThreads = 384;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr int kSCol0 = 0, kSCol1 = 128, kOCol0 = 256, kOCol1 = 256 + D;
};

struct AttnBars {
  uint64_t q_full;
  uint64_t o_final;
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t kv_full[8];
  uint64_t kv_empty[8];
  uint64_t pv_done;  // single-tile (ping-pong) items: completes once per PV
  // sub-chunk pipeline (kSub): per (tile, 64-column half) S-ready / P-ready, per tile PV-complete
  uint64_t sub_s[2][2];
  uint64_t sub_p[2][2];
  uint64_t sub_pv[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  float xch[2 * 2 * 128];  // row-max / row-sum exchange between the two halves of a row (double-buffered)
};

#ifdef SVGB_ATTN_TRACE
// bring-up only: per-chunk clock64 timestamps of one CTA (item SVGB_ATTN_TRACE of head 0)
__device__ long long g_attn_trace[3 * 64 * 8];
#define SVGB_TRACE(role, j, ev)                                                                   \
  do {                                                                                            \
    if (trace_on && (j) < 64 && lane == 0) g_attn_trace[((role) * 64 + (j)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define SVGB_TRACE(role, j, ev) \
  do {                          \
  } while (0)
#endif

constexpr float kRescaleTau = 8.0f;  // log2 units
// which of the 16 column pairs of a 32-column group evaluate exp2 by polynomial on the FMA pipe instead of MUFU.EX2.
// Measured on a B200 (same box, profiles/r02_softmax_variants.jsonl): 1/4 -> 1003-1015 TF/s on the band plan,
// 3/8 -> 978, 1/2 -> 925-933 (and 12.5 % slower than 1/4 in round 1): every extra polynomial costs ~6 issue slots.
#ifndef SVGB_POLY_SEL
#define SVGB_POLY_SEL(i) (((i) & 3) == 3)
#endif
// register budget per role (launch: 384 threads x 168): warps 0-3 give registers to the 8 softmax warps
constexpr int kRegsLight = 56;   // 128 x 56 + 256 x 224 = 64512 = the launch allocation (384 x 168)
constexpr int kRegsSoftmax = 224;

// kGather selects the producer: false = TMA boxes over contiguous key ranges (chunk list), true = cp.async
// row gathers over a run list (separate instantiations keep each one's register footprint small).
template <int D, int DT, bool kGather, bool kSub = false>
__global__ void __launch_bounds__(384, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                const __grid_constant__ CUtensorMap vmap, const AttnArgs args) {
  using Cfg = AttnCfg<D, DT>;
  constexpr bool BF16 = DT != DT_F16;  // output / 16-bit P format (fp8 inputs produce bf16 output)
  constexpr bool FP8 = DT == DT_E4M3;
  static_assert(!(FP8 && kGather), "row-gather producers are 16-bit only");
  static_assert(!(kSub && (FP8 || kGather)), "the sub-chunk pipeline is built for 16-bit TMA plans");
  const int bh = blockIdx.y;
  const int n_items = args.item_count[bh * args.counts_stride];
  if (static_cast<int>(blockIdx.x) >= n_items) return;
  const int4 item = args.items[static_cast<size_t>(bh) * args.items_stride + blockIdx.x];
  int4 item2 = make_int4(0, 0, 0, 0);
  if constexpr (!kGather && !kSub) {
    if (args.items2) item2 = args.items2[static_cast<size_t>(bh) * args.items_stride + blockIdx.x];
  }
  // dual item: T0 and T1 are two independent single-tile streams (own rows, own chunk list, own K/V tiles)
  const bool dual = item2.y > 0;
  const int q_row0 = item.x, nrows = item.y, chunk0 = item.z;
  const int ntiles = (dual || nrows > kTileRows) ? 2 : 1;
  const bool per_tile_map = ntiles == 2 && !args.softmax_shared;  // softmax thread mapping (see below)
  const int2* __restrict__ chunks = args.chunks + chunk0;
  // per-tile view (a regular two-tile item is the special case "same list, rows 128 apart, K/V shared")
  const int q0_t0 = item.x, q0_t1 = dual ? item2.x : item.x + kTileRows;
  const int nr_t0 = dual ? item.y : min(item.y, kTileRows), nr_t1 = dual ? item2.y : item.y - kTileRows;
  const int2* __restrict__ chunks_t1 = dual ? args.chunks + item2.z : chunks;
  const int nch_t1 = dual ? item2.w : item.w;
  constexpr bool gather = kGather;
  // gather mode: item.w = number of runs, chunks are implicit (128 selected keys each, last one partial)
  const int nruns = gather ? item.w : 0;
  const int total_kv = gather ? args.item_total[static_cast<size_t>(bh) * args.items_stride + blockIdx.x] : 0;
  const int nchunks = gather ? (total_kv + kChunkCols - 1) / kChunkCols : item.w;
  const int nch_any = dual ? max(item.w, nch_t1) : nchunks;  // > 0 iff any MMA is issued for this item

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();  // 128-byte-swizzled tiles need a 1024-byte aligned base
  uint8_t* smem_al = smem_raw;
  const uint32_t sQ = smem_base;
  const uint32_t sRing = smem_base + Cfg::kQBytes;
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem_al + Cfg::kQBytes + Cfg::kRingBytes);
  constexpr int kStages = kGather ? Cfg::kStagesGather : Cfg::kStages;
  // gather: the run table {src_start, keys_before} (+ sentinel) stays in global memory; it is a few KB per
  // item, re-read by every producer lane, and therefore L1-resident after the first chunk
  const int2* __restrict__ s_runs = chunks;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef SVGB_ATTN_TRACE
  const bool trace_on = bh == 0 && blockIdx.x == SVGB_ATTN_TRACE && (warp == 1 || warp == 4 || warp == 8);
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&qmap);
    tma_prefetch_desc(&kmap);
    tma_prefetch_desc(&vmap);
  }
  if (warp == 1 && lane == 0) {
    const uint32_t n_prod = gather ? 3u : 1u;  // gather: one arrival per producer warp (0, 2, 3)
    mbar_init(smem_u32(&bars->q_full), n_prod);
    mbar_init(smem_u32(&bars->o_final), 1);
    mbar_init(smem_u32(&bars->pv_done), 1);
    if constexpr (kSub) {
      for (int t = 0; t < 2; ++t) {
        mbar_init(smem_u32(&bars->sub_pv[t]), 1);
        for (int h = 0; h < 2; ++h) {
          mbar_init(smem_u32(&bars->sub_s[t][h]), 1);
          mbar_init(smem_u32(&bars->sub_p[t][h]), 128);
        }
      }
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(smem_u32(&bars->s_full[t]), 1);
      mbar_init(smem_u32(&bars->p_full[t]), per_tile_map ? 128 : 256);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), n_prod);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(smem_u32(&bars->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (gather && (warp == 0 || warp == 2 || warp == 3)) {
    // ------------------------------------------------------------------ cp.async gather producers
    // 3 warps = 6 half-warps; a half-warp moves one 2*D-byte row per step (16 lanes x 16 B, coalesced),
    // writing the 128-byte-swizzled K-major / MN-major image the UMMA descriptors expect:
    //   piece p (16 B) of row r  ->  panel p/8, byte r*128 + ((p%8) ^ (r%8))*16.
    setmaxnreg_dec<kRegsLight>();
    const int pw = warp == 0 ? 0 : warp - 1;
    const int hw = pw * 2 + (lane >> 4);   // half-warp id 0..5
    const int pl = lane & 15;              // 16-byte piece within the row
    constexpr int kPieces = D / 8;         // pieces per row (8 or 16)
    const bool lane_live = pl < kPieces;
    const uint16_t* qp = static_cast<const uint16_t*>(args.q_ptr) + bh * args.in_head_stride;
    const uint16_t* kp = static_cast<const uint16_t*>(args.k_ptr) + bh * args.in_head_stride;
    const uint16_t* vp = static_cast<const uint16_t*>(args.v_ptr) + bh * args.in_head_stride;
    const int* q_rows = args.q_rows ? args.q_rows + static_cast<size_t>(bh) * args.S : nullptr;
    const int* kv_rows = args.kv_rows ? args.kv_rows + static_cast<size_t>(bh) * args.S : nullptr;
    auto dst_off = [&](int r) -> uint32_t {
      return static_cast<uint32_t>((pl >> 3) * Cfg::kPanelBytes + r * 128 + (((pl & 7) ^ (r & 7)) << 4));
    };
    // Each producer warp owns the rows r with (r mod 6) in {2*pw, 2*pw+1}.  Lane L first resolves the source
    // row of the warp's L-th row (all lanes in parallel: one window load of the run list + one optional
    // index load), then the half-warps stream the rows, fetching each row's source with a shuffle.
    auto my_row = [&](int L) -> int { return 6 * (L >> 1) + 2 * pw + (L & 1); };  // L-th row of this warp
    // ---- Q (rows past nrows are zero-filled)
    for (int base = 0; base < ntiles * kTileRows; base += 6 * 16) {  // 32 rows of this warp per pass
      const int r_mine = base + my_row(lane);
      long long src_mine = -1;
      if (r_mine < nrows) {
        const int qr = q_row0 + r_mine;
        src_mine = q_rows ? __ldg(&q_rows[qr]) : qr;
      }
#pragma unroll 4
      for (int k = 0; k < 16; ++k) {
        const int L = 2 * k + (lane >> 4);
        const int r = base + my_row(L);
        const long long src = __shfl_sync(0xffffffffu, src_mine, L);
        if (r < ntiles * kTileRows && lane_live)
          cp_async16(sQ + (r >> 7) * Cfg::kTileBytes + dst_off(r & 127),
                     qp + (src < 0 ? 0 : src) * args.in_row_stride + pl * 8, src >= 0 ? 16u : 0u);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars->q_full));
    // ---- K(j), V(j): two ring slots per chunk, software-pipelined one chunk ahead (measured faster than
    // hardware completion arrivals via cp.async.mbarrier.arrive.noinc, which let the producers run further
    // ahead but lowered the achieved copy rate)
    int pend_k = -1, pend_v = -1;
    for (int j = 0; j < nchunks; ++j) {
      const int it = 2 * j;
      const int kslot = it % kStages, vslot = (it + 1) % kStages;
      // resolve this warp's 42-44 rows of the chunk: position -> run (binary search in the shared-memory run
      // table) -> source row (optional index vector)
      long long src_a = -1, src_b = -1;  // lane L: rows my_row(L) and my_row(L + 32)
      {
        auto resolve = [&](int r) -> long long {
          const int pos = j * kChunkCols + r;
          if (r >= kChunkCols || pos >= total_kv) return -1;
          int lo = 0, hi = nruns - 1;  // largest run with keys_before <= pos
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(&s_runs[mid].y) <= pos) lo = mid;
            else hi = mid - 1;
          }
          const int2 rr = __ldg(&s_runs[lo]);
          const int prow = rr.x + (pos - rr.y);
          return kv_rows ? static_cast<long long>(__ldg(&kv_rows[prow])) : prow;
        };
        src_a = resolve(my_row(lane));
        src_b = resolve(my_row(lane + 32));
      }
      mbar_wait(smem_u32(&bars->kv_empty[kslot]), ((it / kStages) & 1) ^ 1, 1);
      mbar_wait(smem_u32(&bars->kv_empty[vslot]), (((it + 1) / kStages) & 1) ^ 1, 1);
      const uint32_t kd = sRing + kslot * Cfg::kTileBytes, vd = sRing + vslot * Cfg::kTileBytes;
#pragma unroll 4
      for (int k = 0; k < 22; ++k) {
        const int L = 2 * k + (lane >> 4);      // which of the warp's rows this half-warp copies now
        const int r = my_row(L);
        const long long sa = __shfl_sync(0xffffffffu, src_a, L & 31);
        const long long sb = __shfl_sync(0xffffffffu, src_b, L & 31);
        const long long src = L < 32 ? sa : sb;
        if (r < kChunkCols && lane_live) {
          const uint32_t off = dst_off(r);
          const long long so = (src < 0 ? 0 : src) * args.in_row_stride + pl * 8;
          cp_async16(kd + off, kp + so, src >= 0 ? 16u : 0u);
          cp_async16(vd + off, vp + so, src >= 0 ? 16u : 0u);
        }
      }
      cp_async_commit();
      if (pend_k >= 0) {  // the previous chunk's group has landed once at most one group is pending
        cp_async_wait<1>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&bars->kv_full[pend_k]));
          mbar_arrive(smem_u32(&bars->kv_full[pend_v]));
        }
      }
      pend_k = kslot;
      pend_v = vslot;
    }
    if (pend_k >= 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bars->kv_full[pend_k]));
        mbar_arrive(smem_u32(&bars->kv_full[pend_v]));
      }
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<kRegsLight>();
    if (lane == 0) {
      const uint32_t qbar = smem_u32(&bars->q_full);
      mbar_expect_tx(qbar, ntiles * Cfg::kTileBytes);
      for (int t = 0; t < ntiles; ++t)
        for (int h = 0; h < Cfg::kHalves; ++h)
          tma_load_3d(sQ + t * Cfg::kTileBytes + h * Cfg::kPanelBytes, &qmap, qbar, h * Cfg::kPanelElems,
                      t == 0 ? q0_t0 : q0_t1, bh);
      int it = 0;
      if (dual) {
        // ring order of a dual item (the MMA issuer consumes in exactly this order):
        //   K0(0) K1(0) | V0(0) K0(1) V1(0) K1(1) | V0(1) K0(2) V1(1) K1(2) | ...   (entries past a stream's end are skipped)
        auto load = [&](const CUtensorMap* map, int kv0) {
          const int slot = it % kStages;
          mbar_wait(smem_u32(&bars->kv_empty[slot]), ((it / kStages) & 1) ^ 1, 1);
          const uint32_t fb = smem_u32(&bars->kv_full[slot]);
          mbar_expect_tx(fb, Cfg::kTileBytes);
          for (int h = 0; h < Cfg::kHalves; ++h)
            tma_load_3d(sRing + slot * Cfg::kTileBytes + h * Cfg::kPanelBytes, map, fb, h * Cfg::kPanelElems, kv0, bh);
          ++it;
        };
        const int n0 = item.w, n1 = nch_t1, nmax = max(n0, n1);
        if (n0 > 0) load(&kmap, __ldg(&chunks[0].x));
        if (n1 > 0) load(&kmap, __ldg(&chunks_t1[0].x));
        for (int j = 0; j < nmax; ++j) {
          if (j < n0) {
            load(&vmap, __ldg(&chunks[j].x));
            if (j + 1 < n0) load(&kmap, __ldg(&chunks[j + 1].x));
          }
          if (j < n1) {
            load(&vmap, __ldg(&chunks_t1[j].x));
            if (j + 1 < n1) load(&kmap, __ldg(&chunks_t1[j + 1].x));
          }
        }
      } else
      for (int j = 0; j < nchunks; ++j) {
        const int kv0 = __ldg(&chunks[j].x);
#pragma unroll
        for (int kv = 0; kv < 2; ++kv, ++it) {
          const int slot = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(smem_u32(&bars->kv_empty[slot]), ph ^ 1, 1);
          const uint32_t fb = smem_u32(&bars->kv_full[slot]);
          mbar_expect_tx(fb, Cfg::kTileBytes);
          const CUtensorMap* map = kv == 0 ? &kmap : &vmap;
          for (int h = 0; h < Cfg::kHalves; ++h)
            tma_load_3d(sRing + slot * Cfg::kTileBytes + h * Cfg::kPanelBytes, map, fb, h * Cfg::kPanelElems, kv0,
                        bh);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    setmaxnreg_dec<kRegsLight>();
    if (nchunks > 0 && ntiles == 1) {
      // ---- single-tile item: the two S buffers ping-pong over the CHUNK stream (even chunks -> S0, odd -> S1), both
      // accumulate into O0.  QK(j+1) is in flight / done while the softmax of chunk j runs, PV(j) is issued as soon
      // as P(j) is written, QK(j+2) follows it into the buffer it just consumed.  Ring order is unchanged
      // (K(j) = entry 2j, V(j) = entry 2j+1).  pv_done completes once per PV: the lazy O rescale of chunk j waits
      // for PV(j-1) on it.  (+28 % on all-single-tile plans, sample_mse 1.33 -> 1.05 ms per 12 heads.)
      if (lane == 0) {
        auto chunk_n = [&](int jj) -> int {
          const int vld = gather ? min(kChunkCols, total_kv - jj * kChunkCols) : chunk_valid(__ldg(&chunks[jj].y));
          return (vld + Cfg::kMmaK - 1) & ~(Cfg::kMmaK - 1);
        };
        auto issue_qk = [&](int sbuf, int slot, int ncols) {  // S[sbuf] = Q_0 K^T
          const uint32_t idesc = FP8 ? make_idesc_e4m3(128, ncols, false, false)
                                     : make_idesc(128, ncols, DT == DT_BF16, false, false);
          const uint32_t d_tmem = tmem + (sbuf == 0 ? Cfg::kSCol0 : Cfg::kSCol1);
#pragma unroll
          for (int kk = 0; kk < Cfg::kRowBytes / 32; ++kk) {  // 32 bytes of the head dim per MMA
            const uint32_t off = (kk >> 2) * Cfg::kPanelBytes + (kk & 3) * 32;
            const uint64_t a = desc_kmajor_sw128(sQ + off);
            const uint64_t b = desc_kmajor_sw128(sRing + slot * Cfg::kTileBytes + off);
            if constexpr (FP8) mma_ss_f8(d_tmem, a, b, idesc, kk > 0 ? 1u : 0u);
            else mma_ss(d_tmem, a, b, idesc, kk > 0 ? 1u : 0u);
          }
        };
        auto issue_pv = [&](int sbuf, int slot, int ncols, bool acc) {  // O_0 += P[sbuf] V
          const uint32_t idesc = FP8 ? make_idesc_e4m3(128, D, false, true) : make_idesc(128, D, DT == DT_BF16, false, true);
          const uint32_t d_tmem = tmem + Cfg::kOCol0;
          const uint32_t p_tmem = tmem + (sbuf == 0 ? Cfg::kSCol0 : Cfg::kSCol1);
          const int nk = ncols / Cfg::kMmaK;
          for (int kk = 0; kk < nk; ++kk) {
            // kMmaK kv rows per MMA: kMmaK x 128 B down the V panel; P advances 8 TMEM columns (32 bytes).
            const uint64_t b = desc_mnmajor_sw128(sRing + slot * Cfg::kTileBytes + kk * Cfg::kMmaK * 128, Cfg::kPanelBytes);
            if constexpr (FP8) mma_ts_f8(d_tmem, p_tmem + kk * 8, b, idesc, (acc || kk > 0) ? 1u : 0u);
            else mma_ts(d_tmem, p_tmem + kk * 8, b, idesc, (acc || kk > 0) ? 1u : 0u);
          }
        };
        auto wait_full = [&](int idx) {
          mbar_wait(smem_u32(&bars->kv_full[idx % kStages]), (idx / kStages) & 1, 3);
          if constexpr (kGather) fence_proxy_async_smem();
          tc_fence_after();
        };
        auto release = [&](int idx) { tc_commit(smem_u32(&bars->kv_empty[idx % kStages])); };
        mbar_wait(smem_u32(&bars->q_full), 0, 2);
        if constexpr (kGather) fence_proxy_async_smem();
        for (int j = 0; j < 2 && j < nchunks; ++j) {
          wait_full(2 * j);
          issue_qk(j, (2 * j) % kStages, chunk_n(j));
          tc_commit(smem_u32(&bars->s_full[j]));
          release(2 * j);
        }
        for (int j = 0; j < nchunks; ++j) {
          const int b = j & 1;
          wait_full(2 * j + 1);
          SVGB_TRACE(2, j, 0);
          mbar_wait(smem_u32(&bars->p_full[b]), (j >> 1) & 1, 5);
          SVGB_TRACE(2, j, 1);
          tc_fence_after();
          issue_pv(b, (2 * j + 1) % kStages, chunk_n(j), j > 0);
          release(2 * j + 1);
          tc_commit(smem_u32(&bars->pv_done));
          SVGB_TRACE(2, j, 2);
          if (j + 2 < nchunks) {
            wait_full(2 * j + 4);
            issue_qk(b, (2 * j + 4) % kStages, chunk_n(j + 2));
            tc_commit(smem_u32(&bars->s_full[b]));
            release(2 * j + 4);
          }
          SVGB_TRACE(2, j, 3);
        }
        tc_commit(smem_u32(&bars->o_final));
      }
    } else
    // ---- two-tile items: streamed issue (elect.sync leader, running descriptors), so the tensor pipe runs at its
    // 64 cycles per M=128,N=128,K=16 MMA instead of the ~90 cycles a rebuilt-descriptor loop can issue at
    if (nch_any > 0 && elect_one()) {  // elect.sync: ptxas then knows a single lane runs the tcgen05 stream
      // MMA N of chunk jj: run-tail / band chunks carry it in the chunk list; gather chunks are all full
      // except the last
      auto chunk_n = [&](int jj) -> int {
        const int vld = gather ? min(kChunkCols, total_kv - jj * kChunkCols) : chunk_valid(__ldg(&chunks[jj].y));
        return (vld + Cfg::kMmaK - 1) & ~(Cfg::kMmaK - 1);
      };
      // The issuing thread's own instruction stream must not pace the tensor pipe (one M=128,N=128,K=16 MMA is 64
      // cycles of tensor time) nor sit between a barrier wait and the first MMA of a group.  Descriptors are
      // therefore split into a constant high word and a running low word (address field, units of 16 B; all tiles
      // live below 256 KB so the 14-bit field never carries): per MMA the stream is "add a constant, issue".  The
      // empty asm after each MMA pins that order -- without it ptxas hoists all eight address computations (and
      // their R2UR moves) in front of the first MMA, which puts ~60 scalar instructions on the softmax->MMA
      // critical path.
      auto lo32 = [](uint64_t d) { return static_cast<uint32_t>(d); };
      auto hi32 = [](uint64_t d) { return static_cast<uint32_t>(d >> 32); };
      auto mk64 = [](uint32_t lo, uint32_t hi) { return (static_cast<uint64_t>(hi) << 32) | lo; };
      const uint32_t k_hi = hi32(desc_kmajor_sw128(sRing)), v_hi = hi32(desc_mnmajor_sw128(sRing, Cfg::kPanelBytes));
      const uint32_t q_lo[2] = {lo32(desc_kmajor_sw128(sQ)), lo32(desc_kmajor_sw128(sQ + Cfg::kTileBytes))};
      const uint32_t k_lo0 = lo32(desc_kmajor_sw128(sRing)), v_lo0 = lo32(desc_mnmajor_sw128(sRing, Cfg::kPanelBytes));
      constexpr uint32_t kSlotStep = Cfg::kTileBytes >> 4;
      auto qk_idesc = [&](int ncols) -> uint32_t {
        return FP8 ? make_idesc_e4m3(128, ncols, false, false) : make_idesc(128, ncols, DT == DT_BF16, false, false);
      };
      // S_t = Q_t K^T; b_lo = low descriptor word of the K tile (k_lo0 + slot * kSlotStep)
      auto issue_qk = [&](int t, uint32_t b_lo, uint32_t idesc) {
        const uint32_t d_tmem = tmem + (t == 0 ? Cfg::kSCol0 : Cfg::kSCol1);
        uint32_t a_lo = q_lo[t];
#pragma unroll
        for (int kk = 0; kk < Cfg::kRowBytes / 32; ++kk) {  // 32 bytes of the head dim per MMA
          if constexpr (FP8) mma_ss_f8(d_tmem, mk64(a_lo, k_hi), mk64(b_lo, k_hi), idesc, kk > 0 ? 1u : 0u);
          else mma_ss(d_tmem, mk64(a_lo, k_hi), mk64(b_lo, k_hi), idesc, kk > 0 ? 1u : 0u);
          asm volatile("" : "+r"(a_lo), "+r"(b_lo));
          // next 32-byte K step: +32 B inside a 128-byte panel row, then on to the next 16 KB panel
          const uint32_t step = ((kk & 3) == 3) ? ((Cfg::kPanelBytes - 3 * 32) >> 4) : (32 >> 4);
          a_lo += step;
          b_lo += step;
        }
      };
      // O_t += P_t V; b_lo = low descriptor word of the V tile (v_lo0 + slot * kSlotStep)
      auto issue_pv = [&](int t, uint32_t b_lo, int ncols, bool acc) {
        const uint32_t idesc = FP8 ? make_idesc_e4m3(128, D, false, true) : make_idesc(128, D, DT == DT_BF16, false, true);
        const uint32_t d_tmem = tmem + (t == 0 ? Cfg::kOCol0 : Cfg::kOCol1);
        uint32_t p_tmem = tmem + (t == 0 ? Cfg::kSCol0 : Cfg::kSCol1);
        const int nk = ncols / Cfg::kMmaK;
#pragma unroll
        for (int kk = 0; kk < kChunkCols / Cfg::kMmaK; ++kk) {
          // kMmaK kv rows per MMA: kMmaK x 128 B down the V panel; P advances 8 TMEM columns (32 bytes).
          if (kk < nk) {
            if constexpr (FP8) mma_ts_f8(d_tmem, p_tmem, mk64(b_lo, v_hi), idesc, (acc || kk > 0) ? 1u : 0u);
            else mma_ts(d_tmem, p_tmem, mk64(b_lo, v_hi), idesc, (acc || kk > 0) ? 1u : 0u);
            asm volatile("" : "+r"(p_tmem), "+r"(b_lo));
            p_tmem += 8;
            b_lo += (Cfg::kMmaK * 128) >> 4;
          }
        }
      };
      if (dual) {
        // ---- dual item: two independent single-tile streams.  Same interleaving as the shared-K/V loop below
        // (PV_t(j), QK_t(j+1) per tile, tiles alternating) but every stream has its own ring entries, in the order
        // the producer loads them; a stream that ends early simply drops out.
        auto chunk_n_t = [&](int t, int jj) -> int {
          const int vld = chunk_valid(__ldg(t == 0 ? &chunks[jj].y : &chunks_t1[jj].y));
          return (vld + Cfg::kMmaK - 1) & ~(Cfg::kMmaK - 1);
        };
        const int n0 = item.w, n1 = nch_t1, nmax = max(n0, n1);
        mbar_wait(smem_u32(&bars->q_full), 0, 2);
        int ring = 0;
        int n_cur0 = n0 > 0 ? chunk_n_t(0, 0) : 0, n_cur1 = n1 > 0 ? chunk_n_t(1, 0) : 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if ((t == 0 ? n0 : n1) > 0) {
            const int slot = ring % kStages;
            mbar_wait(smem_u32(&bars->kv_full[slot]), (ring / kStages) & 1, 3);
            ++ring;
            tc_fence_after();
            issue_qk(t, k_lo0 + slot * kSlotStep, qk_idesc(t == 0 ? n_cur0 : n_cur1));
            tc_commit(smem_u32(&bars->s_full[t]));
            tc_commit(smem_u32(&bars->kv_empty[slot]));
          }
        }
        for (int j = 0; j < nmax; ++j) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int nt = t == 0 ? n0 : n1;
            if (j >= nt) continue;
            const bool has_next = (j + 1 < nt);
            const int vslot = ring % kStages;
            const uint32_t vph = (ring / kStages) & 1;
            ++ring;
            int kslot = 0, n_next = 0;
            uint32_t kph = 0;
            if (has_next) {
              kslot = ring % kStages;
              kph = (ring / kStages) & 1;
              ++ring;
              n_next = chunk_n_t(t, j + 1);
            }
            const uint32_t v_lo = v_lo0 + vslot * kSlotStep, k_lo = k_lo0 + kslot * kSlotStep;
            const uint32_t idesc_next = qk_idesc(n_next);
            mbar_wait(smem_u32(&bars->kv_full[vslot]), vph, 4);
            mbar_wait(smem_u32(&bars->p_full[t]), j & 1, 5 + 2 * t);
            tc_fence_after();
            issue_pv(t, v_lo, t == 0 ? n_cur0 : n_cur1, j > 0);
            tc_commit(smem_u32(&bars->kv_empty[vslot]));
            if (has_next) {
              mbar_wait(smem_u32(&bars->kv_full[kslot]), kph, 6);
              tc_fence_after();
              issue_qk(t, k_lo, idesc_next);
              tc_commit(smem_u32(&bars->s_full[t]));
              tc_commit(smem_u32(&bars->kv_empty[kslot]));
            }
            if (t == 0) n_cur0 = n_next;
            else n_cur1 = n_next;
          }
        }
        tc_commit(smem_u32(&bars->o_final));
      } else
      if constexpr (kSub) {
        // ---- sub-chunk pipeline (experimental, SVGB_ATTN_SUB=1): every 128-key chunk is two 64-key halves with
        // their own S sub-buffer (S_t columns [64h, 64h+64), P over the first 32 of them).  QK_t,h(j+1) follows
        // PV_t,h(j) immediately, i.e. while the softmax thread is still in the OTHER half of chunk j, so the
        // softmax never waits for an MMA round trip.  sub_pv[t] completes once per PV group (the lazy O rescale
        // of sub-step n waits for group n-1 on it).
        auto half_cols = [](int n, int h) { return min(max(n - 64 * h, 0), 64); };
        constexpr uint32_t kHalfRows = (64 * 128) >> 4;  // 64 key rows down a 128-byte-row panel, descriptor units
        auto issue_qk_h = [&](int t, int h, uint32_t k_lo_slot, int nh) {
          if (nh == 0) return;
          const uint32_t idesc = make_idesc(128, nh, DT == DT_BF16, false, false);
          const uint32_t d_tmem = tmem + (t == 0 ? Cfg::kSCol0 : Cfg::kSCol1) + 64 * h;
          uint32_t a_lo = q_lo[t], b_lo = k_lo_slot + h * kHalfRows;
#pragma unroll
          for (int kk = 0; kk < Cfg::kRowBytes / 32; ++kk) {
            mma_ss(d_tmem, mk64(a_lo, k_hi), mk64(b_lo, k_hi), idesc, kk > 0 ? 1u : 0u);
            asm volatile("" : "+r"(a_lo), "+r"(b_lo));
            const uint32_t step = ((kk & 3) == 3) ? ((Cfg::kPanelBytes - 3 * 32) >> 4) : (32 >> 4);
            a_lo += step;
            b_lo += step;
          }
        };
        auto issue_pv_h = [&](int t, int h, uint32_t v_lo_slot, int nh, bool acc) {
          const uint32_t idesc = make_idesc(128, D, DT == DT_BF16, false, true);
          const uint32_t d_tmem = tmem + (t == 0 ? Cfg::kOCol0 : Cfg::kOCol1);
          uint32_t p_tmem = tmem + (t == 0 ? Cfg::kSCol0 : Cfg::kSCol1) + 64 * h;
          uint32_t b_lo = v_lo_slot + h * kHalfRows;
          const int nk = nh / Cfg::kMmaK;
#pragma unroll
          for (int kk = 0; kk < 64 / Cfg::kMmaK; ++kk) {
            if (kk < nk) {
              mma_ts(d_tmem, p_tmem, mk64(b_lo, v_hi), idesc, (acc || kk > 0) ? 1u : 0u);
              asm volatile("" : "+r"(p_tmem), "+r"(b_lo));
              p_tmem += 8;
              b_lo += (Cfg::kMmaK * 128) >> 4;
            }
          }
        };
        mbar_wait(smem_u32(&bars->q_full), 0, 2);
        mbar_wait(smem_u32(&bars->kv_full[0]), 0, 3);
        tc_fence_after();
        {
          const int n0c = chunk_n(0);
          for (int h = 0; h < 2; ++h)
            for (int t = 0; t < 2; ++t) {
              issue_qk_h(t, h, k_lo0, half_cols(n0c, h));
              tc_commit(smem_u32(&bars->sub_s[t][h]));
            }
        }
        tc_commit(smem_u32(&bars->kv_empty[0]));
        // ---- ARRIVAL-ORDER service.  Every (tile, half) sub-step of chunk j is served as soon as its P half is
        // written: PV_t,h(j) then QK_t,h(j+1).  The two tiles are served in whatever order their softmax warpgroups
        // arrive (a fixed T0,T1 order makes the issuer sit on the slower tile's barrier while the other one's P is
        // ready, which locks the two warpgroups into the same phase: both in their MUFU-bound exp section at once,
        // the tensor pipe fed in bursts).  Ring entries: K(0) = 0, V(j) = 2j+1, K(j+1) = 2j+2; V(j) and K(j+1) are
        // released when BOTH tiles have served (j, half 1).
        int jt0 = 0, jt1 = 0, ht0 = 0, ht1 = 0;  // next sub-step of each tile
        int last_full0 = -1, last_full1 = -1;    // last chunk whose half 1 the tile has served
        auto bar_ready = [](uint32_t bar, uint32_t parity) -> bool {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
          return ok != 0;
        };
        while (jt0 < nchunks || jt1 < nchunks) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int j = t == 0 ? jt0 : jt1, h = t == 0 ? ht0 : ht1;
            if (j >= nchunks) continue;
            const bool has_next = j + 1 < nchunks;
            const int ve = 2 * j + 1, ke = 2 * j + 2;
            const int vslot = ve % kStages, kslot = ke % kStages;
            if (!bar_ready(smem_u32(&bars->sub_p[t][h]), j & 1)) continue;
            if (!bar_ready(smem_u32(&bars->kv_full[vslot]), (ve / kStages) & 1)) continue;
            if (has_next && !bar_ready(smem_u32(&bars->kv_full[kslot]), (ke / kStages) & 1)) continue;
            tc_fence_after();
#ifdef SVGB_ATTN_TRACE
            const int svc = (jt0 * 2 + ht0) + (jt1 * 2 + ht1);  // services done so far
            if (trace_on && svc < 64) g_attn_trace[(2 * 64 + svc) * 8 + 0] = clock64(), g_attn_trace[(2 * 64 + svc) * 8 + 3] = t * 2 + h;
#endif
            const int n_c = chunk_n(j);
            issue_pv_h(t, h, v_lo0 + vslot * kSlotStep, half_cols(n_c, h), j > 0 || h > 0);
            tc_commit(smem_u32(&bars->sub_pv[t]));
            if (has_next) {
              issue_qk_h(t, h, k_lo0 + kslot * kSlotStep, half_cols(chunk_n(j + 1), h));
              tc_commit(smem_u32(&bars->sub_s[t][h]));
            }
#ifdef SVGB_ATTN_TRACE
            if (trace_on && svc < 64) g_attn_trace[(2 * 64 + svc) * 8 + 1] = clock64();
#endif
            if (h == 1) {
              const int other = t == 0 ? last_full1 : last_full0;
              if (other >= j) {  // the other tile is already past this chunk: its K/V tiles are free
                tc_commit(smem_u32(&bars->kv_empty[vslot]));
                if (has_next) tc_commit(smem_u32(&bars->kv_empty[kslot]));
              }
              if (t == 0) { last_full0 = j; ++jt0; ht0 = 0; }
              else { last_full1 = j; ++jt1; ht1 = 0; }
            } else {
              if (t == 0) ht0 = 1;
              else ht1 = 1;
            }
          }
        }
        tc_commit(smem_u32(&bars->o_final));
      } else if (!kGather && ntiles == 2 && args.arrival_order) {
        // ---- two-tile item, ARRIVAL-ORDER service (SVGB_ATTN_ORDER=1): PV_t(j) + QK_t(j+1) are issued for whichever
        // tile's P arrives first instead of always T0 then T1, so a tile never waits behind the other one's barrier.
        // Ring entries: K(0) = 0, V(j) = 2j+1, K(j+1) = 2j+2; V(j) / K(j+1) are released by the second tile to pass j.
        auto bar_ready = [](uint32_t bar, uint32_t parity) -> bool {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
          return ok != 0;
        };
        mbar_wait(smem_u32(&bars->q_full), 0, 2);
        mbar_wait(smem_u32(&bars->kv_full[0]), 0, 3);
        tc_fence_after();
        {
          const uint32_t id0 = qk_idesc(chunk_n(0));
          issue_qk(0, k_lo0, id0);
          tc_commit(smem_u32(&bars->s_full[0]));
          issue_qk(1, k_lo0, id0);
          tc_commit(smem_u32(&bars->s_full[1]));
        }
        tc_commit(smem_u32(&bars->kv_empty[0]));
        int jt0 = 0, jt1 = 0;
        while (jt0 < nchunks || jt1 < nchunks) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int j = t == 0 ? jt0 : jt1;
            if (j >= nchunks) continue;
            const bool has_next = j + 1 < nchunks;
            const int ve = 2 * j + 1, ke = 2 * j + 2;
            const int vslot = ve % kStages, kslot = ke % kStages;
            if (!bar_ready(smem_u32(&bars->p_full[t]), j & 1)) continue;
            if (!bar_ready(smem_u32(&bars->kv_full[vslot]), (ve / kStages) & 1)) continue;
            if (has_next && !bar_ready(smem_u32(&bars->kv_full[kslot]), (ke / kStages) & 1)) continue;
            tc_fence_after();
            issue_pv(t, v_lo0 + vslot * kSlotStep, chunk_n(j), j > 0);
            if (has_next) {
              issue_qk(t, k_lo0 + kslot * kSlotStep, qk_idesc(chunk_n(j + 1)));
              tc_commit(smem_u32(&bars->s_full[t]));
            }
            if ((t == 0 ? jt1 : jt0) > j) {  // the other tile is already past chunk j: its K/V tiles are free
              tc_commit(smem_u32(&bars->kv_empty[vslot]));
              if (has_next) tc_commit(smem_u32(&bars->kv_empty[kslot]));
            }
            if (t == 0) ++jt0;
            else ++jt1;
          }
        }
        tc_commit(smem_u32(&bars->o_final));
      } else {
      mbar_wait(smem_u32(&bars->q_full), 0, 2);
      if constexpr (kGather) fence_proxy_async_smem();
      int ring = 0;
      int n_cur = chunk_n(0);
      {
        const int slot = 0;
        mbar_wait(smem_u32(&bars->kv_full[slot]), 0, 3);
        if constexpr (kGather) fence_proxy_async_smem();  // cp.async wrote through the generic proxy
        tc_fence_after();
        issue_qk(0, k_lo0, qk_idesc(n_cur));
        tc_commit(smem_u32(&bars->s_full[0]));
        if (ntiles > 1) {
          issue_qk(1, k_lo0, qk_idesc(n_cur));
          tc_commit(smem_u32(&bars->s_full[1]));
        }
        tc_commit(smem_u32(&bars->kv_empty[slot]));
        ring = 1;
      }
      for (int j = 0; j < nchunks; ++j) {
        const bool has_next = (j + 1 < nchunks);
        const int vslot = ring % kStages;
        const uint32_t vph = (ring / kStages) & 1;
        ++ring;
        int kslot = 0, n_next = 0;
        uint32_t kph = 0;
        if (has_next) {
          kslot = ring % kStages;
          kph = (ring / kStages) & 1;
          ++ring;
          n_next = chunk_n(j + 1);
        }
        // everything the MMAs need is computed before the waits
        const uint32_t v_lo = v_lo0 + vslot * kSlotStep, k_lo = k_lo0 + kslot * kSlotStep;
        const uint32_t idesc_next = qk_idesc(n_next);
        mbar_wait(smem_u32(&bars->kv_full[vslot]), vph, 4);
        if constexpr (kGather) fence_proxy_async_smem();
        SVGB_TRACE(2, j, 0);
        mbar_wait(smem_u32(&bars->p_full[0]), j & 1, 5);
        SVGB_TRACE(2, j, 1);
        tc_fence_after();
        issue_pv(0, v_lo, n_cur, j > 0);
        if (has_next) {
          mbar_wait(smem_u32(&bars->kv_full[kslot]), kph, 6);
          if constexpr (kGather) fence_proxy_async_smem();
          tc_fence_after();
          SVGB_TRACE(2, j, 2);
          issue_qk(0, k_lo, idesc_next);
          tc_commit(smem_u32(&bars->s_full[0]));
          SVGB_TRACE(2, j, 3);
        }
        if (ntiles > 1) {
          mbar_wait(smem_u32(&bars->p_full[1]), j & 1, 7);
          SVGB_TRACE(2, j, 4);
          tc_fence_after();
          issue_pv(1, v_lo, n_cur, j > 0);
        }
        tc_commit(smem_u32(&bars->kv_empty[vslot]));
        if (has_next) {
          if (ntiles > 1) {
            SVGB_TRACE(2, j, 5);
            issue_qk(1, k_lo, idesc_next);
            tc_commit(smem_u32(&bars->s_full[1]));
            SVGB_TRACE(2, j, 6);
          }
          tc_commit(smem_u32(&bars->kv_empty[kslot]));
        }
        n_cur = n_next;
      }
      tc_commit(smem_u32(&bars->o_final));
      }
    }
  } else if (warp < 4) {
    setmaxnreg_dec<kRegsLight>();
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    // Both softmax warpgroups work on the SAME tile at a time, alternating T0, T1, T0, ...: thread (half, row)
    // owns 64 of the 128 key columns of one query row (half 0 = warps 4-7, half 1 = warps 8-11; both map to
    // the same TMEM lanes).  While they process T0's chunk the tensor core runs T1's PV / QK, so the softmax
    // units stay busy and the per-tile latency on the MMA critical chain is halved.  The two halves of a row
    // agree on the row maximum through shared memory (one named barrier per tile-chunk); row sums are only
    // combined in the epilogue.
    setmaxnreg_inc<kRegsSoftmax>();
    if (per_tile_map) {
    // ---- two-tile items: one warpgroup per tile, one thread per query row (the two tiles' softmaxes run
    // concurrently on different warps, so their latency bubbles overlap)
    const int t = (warp - 4) >> 2;
    if (t < ntiles) {
      const int wq = warp & 3;
      const int row = wq * 32 + lane;  // row inside the 128-row tile == TMEM lane
      const int q = (t == 0 ? q0_t0 : q0_t1) + row;
      // this tile's chunk stream (dual items: its own list; otherwise the item's)
      const int2* __restrict__ my_chunks = t == 0 ? chunks : chunks_t1;
      const int my_n = dual ? (t == 0 ? item.w : nch_t1) : nchunks;
      const int qm = (args.q_index && q < args.S) ? __ldg(&args.q_index[q]) : q;  // position seen by the mask
      const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
      const uint32_t s_addr = lane_addr + (t == 0 ? Cfg::kSCol0 : Cfg::kSCol1);
      const uint32_t o_addr = lane_addr + (t == 0 ? Cfg::kOCol0 : Cfg::kOCol1);
      // fp8: logits carry s_q * s_k; P = 2^(x - m + 4) with the lazy-rescale slack at 4 keeps P in (0, 2^8]
      // (e4m3 max 448) while the fresh-max case still has 13 binades below it
      const float c = args.scale_log2 * (args.q_scale ? __ldg(&args.q_scale[bh]) * __ldg(&args.k_scale[bh]) : 1.f);
      constexpr float kTau = FP8 ? 4.f : kRescaleTau;
      constexpr float kPOff = FP8 ? 4.f : 0.f;
      const int mode = args.mask_mode, m0 = args.m0, m1 = args.m1, m2 = args.m2;
      const uint32_t sbar = smem_u32(&bars->s_full[t]);
      const uint32_t pbar = smem_u32(&bars->p_full[t]);

      float m_used = -INFINITY;  // reference max the stored P / O are scaled against
      float l_run = 0.f;
      int2 ch = (my_n > 0 && !gather) ? __ldg(&my_chunks[0]) : make_int2(0, 0);
      MaskRow mrow;
      mrow.init(mode, qm, m0, m1, m2);

      if constexpr (kSub) {
      // ---- sub-chunk pipeline: two 64-column online-softmax steps per chunk (see the MMA issuer)
      int nsub = 0;  // PV groups issued into O_t so far == sub-steps finished
      for (int j = 0; j < nchunks; ++j) {
        const int kv0 = ch.x;
        const int valid = chunk_valid(ch.y);
        const bool elem = (ch.y & kChunkElem) != 0;
        const int ncols = (valid + Cfg::kMmaK - 1) & ~(Cfg::kMmaK - 1);
        if (j + 1 < nchunks) ch = __ldg(&chunks[j + 1]);
#pragma unroll 1
        for (int h = 0; h < 2; ++h, ++nsub) {
          const int nh = min(max(ncols - 64 * h, 0), 64);  // MMA columns of this half
          const int vh = min(max(valid - 64 * h, 0), 64);  // valid key columns of this half
          const uint32_t sh_addr = s_addr + 64 * h;
          SVGB_TRACE(t, nsub, 0);
          mbar_wait(smem_u32(&bars->sub_s[t][h]), j & 1, 8 + t);
          tc_fence_after();
          SVGB_TRACE(t, nsub, 1);
          if (nh > 0) {
            float rs;
            auto half_body = [&](auto plain_tag) {
              constexpr bool kPlain = decltype(plain_tag)::value;
              uint32_t r0[32], r1[32];
              tmem_ld32(sh_addr, r0);
              if (kPlain || nh > 32) tmem_ld32(sh_addr + 32, r1);
              tc_wait_ld();
              SVGB_TRACE(t, nsub, 2);
              if constexpr (!kPlain) {
                auto sanitize = [&](uint32_t(&rr)[32], int gl) {
                  const int left = vh - gl * 32;
                  if (gl * 32 >= nh || !(elem || left < 32)) return;  // warp-uniform
                  uint32_t bits = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
                  if (elem && left > 0) bits &= mrow.bits32(kv0 + 64 * h + gl * 32);
#pragma unroll
                  for (int i = 0; i < 32; ++i) rr[i] = (bits >> i) & 1u ? rr[i] : 0xff800000u;  // -inf
                };
                sanitize(r0, 0);
                sanitize(r1, 1);
              }
              float mx = -INFINITY;
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r0[i]));
              if (kPlain || nh > 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r1[i]));
              }
              const float m_new = fmaxf(m_used, mx);
              float alpha = 1.f;
              if ((m_new - m_used) * c > kTau) {
                alpha = ex2_approx((m_used - m_new) * c);
                m_used = m_new;
              }
              SVGB_TRACE(t, nsub, 3);
              if (nsub > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
                // every PV group issued so far must have landed in O_t before it is rescaled
                mbar_wait(smem_u32(&bars->sub_pv[t]), (nsub - 1) & 1, 12);
                tc_fence_after();
#pragma unroll 1
                for (int g = 0; g < D / 32; ++g) {
                  uint32_t o[32];
                  tmem_ld32(o_addr + g * 32, o);
                  tc_wait_ld();
#pragma unroll
                  for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                  tmem_st32(o_addr + g * 32, o);
                }
              }
              l_run *= alpha;
              const float mc = (m_used == -INFINITY) ? 0.f : m_used * c - kPOff;
              const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
              uint64_t sum2 = pack_f32x2(0.f, 0.f);
              auto group_p = [&](const uint32_t(&rr)[32], int gl) {
                if constexpr (!kPlain) {
                  if (gl * 32 >= nh) return;
                }
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint64_t x2 =
                      ffma2(pack_f32x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1])), c2, nmc2);
                  float p0, p1;
                  if (SVGB_POLY_SEL(i)) {
                    ex2_poly2(x2, p0, p1);
                  } else {
                    float x0, x1;
                    unpack_f32x2(x2, x0, x1);
                    p0 = ex2_approx(x0);
                    p1 = ex2_approx(x1);
                  }
                  sum2 = fadd2(sum2, pack_f32x2(p0, p1));
                  pk[i] = pack2_p<BF16>(p0, p1);
                }
                tmem_st16(sh_addr + gl * 16, pk);
              };
              group_p(r0, 0);
              group_p(r1, 1);
              SVGB_TRACE(t, nsub, 4);
              float s0, s1;
              unpack_f32x2(sum2, s0, s1);
              rs = s0 + s1;
            };
            if (!elem && vh == 64) half_body(std::true_type{});
            else half_body(std::false_type{});
            l_run += rs;
            tc_wait_st();
            SVGB_TRACE(t, nsub, 5);
          }
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->sub_p[t][h]));
          SVGB_TRACE(t, nsub, 6);
        }
      }
      } else
      for (int j = 0; j < my_n; ++j) {
        const int kv0 = ch.x;
        const int valid = gather ? min(kChunkCols, total_kv - j * kChunkCols) : chunk_valid(ch.y);
        const bool elem = !gather && (ch.y & kChunkElem) != 0;
        const int ncols = (valid + Cfg::kMmaK - 1) & ~(Cfg::kMmaK - 1);
        const int ngroups = (ncols + 31) >> 5;
        if (!gather && j + 1 < my_n) ch = __ldg(&my_chunks[j + 1]);

        SVGB_TRACE(t, j, 0);
        mbar_wait(sbar, j & 1, 8 + t);
        tc_fence_after();
        SVGB_TRACE(t, j, 1);

        // ---------------- single pass: the whole score row (<=128 columns) lives in registers.
        // Compiled twice: kPlain = full unmasked 128-column chunk (the common case, no guards at all) and
        // the general form, which first overwrites disallowed scores with -inf in place (run tails,
        // band edges, profiling masks) -- exp2(-inf) = 0 then makes the rest identical to the plain path.
        float rs;
        // Three straight-line specialisations (no per-group run-time guards, static register allocation):
        //   plain   : full unmasked 128-column chunk
        //   masked<4>: any other chunk wider than 64 columns -- all four 32-column groups are processed; columns past
        //             `valid` (stale TMEM contents when the MMA N was < 128) and element-masked ones become -inf
        //   masked<2>: chunks of at most 64 columns -- two groups
        auto chunk_body = [&](auto plain_tag, auto groups_tag) {
          constexpr bool kPlain = decltype(plain_tag)::value;
          constexpr int kG = decltype(groups_tag)::value;
          uint32_t r0[32], r1[32], r2[kG == 4 ? 32 : 1], r3[kG == 4 ? 32 : 1];
          tmem_ld32(s_addr, r0);
          tmem_ld32(s_addr + 32, r1);
          if constexpr (kG == 4) {
            tmem_ld32(s_addr + 64, r2);
            tmem_ld32(s_addr + 96, r3);
          }
          tc_wait_ld();
          SVGB_TRACE(t, j, 2);
          if constexpr (!kPlain) {
            auto sanitize = [&](uint32_t(&rr)[32], int g) {
              const int left = valid - g * 32;
              if (!(elem || left < 32)) return;  // warp-uniform
              uint32_t bits = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
              if (elem && left > 0) bits &= mrow.bits32(kv0 + g * 32);
#pragma unroll
              for (int i = 0; i < 32; ++i) rr[i] = (bits >> i) & 1u ? rr[i] : 0xff800000u;  // -inf
            };
            sanitize(r0, 0);
            sanitize(r1, 1);
            if constexpr (kG == 4) {
              sanitize(r2, 2);
              sanitize(r3, 3);
            }
          }
          auto group_max = [&](const uint32_t(&rr)[32]) -> float {
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(rr[i]));
            return m;
          };
          float mx = fmaxf(group_max(r0), group_max(r1));
          if constexpr (kG == 4) mx = fmaxf(mx, fmaxf(group_max(r2), group_max(r3)));
          const float m_new = fmaxf(m_used, mx);
          // lazy rescale: keep the stale reference max unless it grew by more than tau (log2 units)
          float alpha = 1.f;
          if ((m_new - m_used) * c > kTau) {  // false when both are -inf (NaN compare)
            alpha = ex2_approx((m_used - m_new) * c);  // 0 when m_used == -inf
            m_used = m_new;
          }
          if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
            // correction: O_row *= alpha (PV_t(j-1) is complete: the S_t(j) commit covered it)
#pragma unroll 1
            for (int g = 0; g < D / 32; ++g) {
              uint32_t o[32];
              tmem_ld32(o_addr + g * 32, o);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(o_addr + g * 32, o);
            }
          }
          SVGB_TRACE(t, j, 3);
          l_run *= alpha;
          const float mc = (m_used == -INFINITY) ? 0.f : m_used * c - kPOff;
          const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
          uint64_t sum2 = pack_f32x2(0.f, 0.f);
          // P = exp2(S*c - m*c) -> 16-bit, packed two per TMEM column over the first half of the S tile.
          // Every 4th pair is evaluated on the FMA pipe (polynomial) to unload the MUFU.
          auto group_p = [&](const uint32_t(&rr)[32], int g) {
            uint32_t pk[16];
            float rs_hi[FP8 ? 16 : 1];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t x2 =
                  ffma2(pack_f32x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1])), c2, nmc2);
              float p0, p1;
              if (SVGB_POLY_SEL(i)) {
                ex2_poly2(x2, p0, p1);
              } else {
                float x0, x1;
                unpack_f32x2(x2, x0, x1);
                p0 = ex2_approx(x0);
                p1 = ex2_approx(x1);
              }
              sum2 = fadd2(sum2, pack_f32x2(p0, p1));
              if constexpr (FP8) {  // keep the fp32 pair; four of them make one e4m3x4 word below
                pk[i] = __float_as_uint(p0);
                rs_hi[i] = p1;
              } else {
                pk[i] = pack2_p<BF16>(p0, p1);
              }
            }
            if constexpr (FP8) {
              uint32_t p8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                p8[i] = pack4_e4m3(__uint_as_float(pk[2 * i]), rs_hi[2 * i], __uint_as_float(pk[2 * i + 1]), rs_hi[2 * i + 1]);
              tmem_st8(s_addr + g * 8, p8);
            } else {
              tmem_st16(s_addr + g * 16, pk);
            }
          };
          group_p(r0, 0);
          group_p(r1, 1);
          if constexpr (kG == 4) {
            group_p(r2, 2);
            group_p(r3, 3);
          }
          SVGB_TRACE(t, j, 4);
          float s0, s1;
          unpack_f32x2(sum2, s0, s1);
          rs = s0 + s1;
        };
        if (!elem && valid == kChunkCols) chunk_body(std::true_type{}, std::integral_constant<int, 4>{});
        else if (ngroups <= 2) chunk_body(std::false_type{}, std::integral_constant<int, 2>{});
        else chunk_body(std::false_type{}, std::integral_constant<int, 4>{});
        l_run += rs;
        tc_wait_st();
        SVGB_TRACE(t, j, 5);
        tc_fence_before();
        mbar_arrive(pbar);
        SVGB_TRACE(t, j, 6);
      }

      // ---------------- epilogue: O / l -> 16-bit -> global (optionally scattered rows)
      const bool row_ok = row < (t == 0 ? nr_t0 : nr_t1);
      if (nch_any > 0) {
        mbar_wait(smem_u32(&bars->o_final), 0, 10 + t);
        tc_fence_after();
      }
      const float inv_l = (l_run > 0.f ? 1.f / l_run : 0.f) * (args.v_scale ? __ldg(&args.v_scale[bh]) : 1.f);
      long long out_row = q;
      if (row_ok && args.o_rows) out_row = __ldg(&args.o_rows[static_cast<size_t>(bh) * args.S + q]);
      uint16_t* optr = reinterpret_cast<uint16_t*>(args.o) + bh * args.o_head_stride +
                       out_row * args.o_row_stride;
      float* optr32 = reinterpret_cast<float*>(args.o) + bh * args.o_head_stride + out_row * args.o_row_stride;
#pragma unroll 1
      for (int g = 0; g < D / 32; ++g) {
        uint32_t o[32];
        if (my_n > 0) {
          tmem_ld32(o_addr + g * 32, o);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0u;
        }
        if (row_ok && args.out_f32) {
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            float4 w;
            w.x = __uint_as_float(o[4 * v + 0]) * inv_l;
            w.y = __uint_as_float(o[4 * v + 1]) * inv_l;
            w.z = __uint_as_float(o[4 * v + 2]) * inv_l;
            w.w = __uint_as_float(o[4 * v + 3]) * inv_l;
            *reinterpret_cast<float4*>(optr32 + g * 32 + v * 4) = w;
          }
        } else if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint4 w;
            w.x = pack2<BF16>(__uint_as_float(o[8 * v + 0]) * inv_l, __uint_as_float(o[8 * v + 1]) * inv_l);
            w.y = pack2<BF16>(__uint_as_float(o[8 * v + 2]) * inv_l, __uint_as_float(o[8 * v + 3]) * inv_l);
            w.z = pack2<BF16>(__uint_as_float(o[8 * v + 4]) * inv_l, __uint_as_float(o[8 * v + 5]) * inv_l);
            w.w = pack2<BF16>(__uint_as_float(o[8 * v + 6]) * inv_l, __uint_as_float(o[8 * v + 7]) * inv_l);
            *reinterpret_cast<uint4*>(optr + g * 32 + v * 8) = w;
          }
        }
      }
      if (row_ok && args.lse) {
        // natural-log LSE of the scaled scores; -inf for rows that saw no key
        const float lse = l_run > 0.f ? (m_used * c - kPOff + log2f(l_run)) * 0.6931471805599453f : -INFINITY;
        args.lse[static_cast<size_t>(bh) * args.S + out_row] = lse;
      }
    }
    } else
    // ---- shared mapping: both warpgroups work on the same tile (each thread owns half of a row), alternating
    // T0, T1.  Used for single-tile items (cluster tails, split-KV profiling passes: halves the lone tile's
    // softmax latency) and for plans made of narrow chunks (small k-means clusters), where the step is
    // latency- rather than throughput-bound and interleaving the two tiles on the same threads hides it.
    {
      const int half = (warp - 4) >> 2;
      const int wq = warp & 3;
      const int row = wq * 32 + lane;  // row inside the 128-row tile == TMEM lane
      const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16);
      // fp8: logits carry s_q * s_k; P = 2^(x - m + 4) with the lazy-rescale slack at 4 keeps P in (0, 2^8]
      // (e4m3 max 448) while the fresh-max case still has 13 binades below it
      const float c = args.scale_log2 * (args.q_scale ? __ldg(&args.q_scale[bh]) * __ldg(&args.k_scale[bh]) : 1.f);
      constexpr float kTau = FP8 ? 4.f : kRescaleTau;
      constexpr float kPOff = FP8 ? 4.f : 0.f;
      const int mode = args.mask_mode, m0 = args.m0, m1 = args.m1, m2 = args.m2;
      float* xch = bars->xch;  // [2 parity][2 half][128 rows]
      int xstep = 0;
      // exchange one float with the thread that owns the other half of this row
      auto exchange = [&](float v) -> float {
        float* buf = xch + (xstep & 1) * 256;
        buf[half * 128 + row] = v;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float other = buf[(half ^ 1) * 128 + row];
        ++xstep;
        return other;
      };

      float m_used[2] = {-INFINITY, -INFINITY};  // per tile: reference max the stored P / O are scaled against
      float l_run[2] = {0.f, 0.f};               // per tile: this thread's half of the row sum
      const int n_t0 = dual ? item.w : nchunks, n_t1 = dual ? nch_t1 : nchunks;  // chunk count of each tile's stream

      for (int j = 0; j < nch_any; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {  // unrolled: m_used / l_run stay in registers
          if (t >= ntiles || j >= (t == 0 ? n_t0 : n_t1)) continue;  // dual items: the shorter stream has ended
          // chunk j of this tile's stream (L1-resident list; both tiles of a regular item read the same entry)
          const int2 ch = gather ? make_int2(0, 0) : __ldg(t == 0 ? &chunks[j] : &chunks_t1[j]);
          const int kv0 = ch.x;
          const int valid = gather ? min(kChunkCols, total_kv - j * kChunkCols) : chunk_valid(ch.y);
          const bool elem = !gather && (ch.y & kChunkElem) != 0;
          const int ncols = (valid + Cfg::kMmaK - 1) & ~(Cfg::kMmaK - 1);
          const int c0 = half * 64;            // first key column of this thread's half
          const int mycols = ncols - c0;       // <= 0: nothing in this half (narrow chunk)
          // single-tile items ping-pong the two S buffers over the chunk stream (see the MMA issuer)
          const int sb = ntiles == 1 ? (j & 1) : t;
          const uint32_t s_addr = lane_addr + (sb == 0 ? Cfg::kSCol0 : Cfg::kSCol1);
          const uint32_t o_addr = lane_addr + (t == 0 ? Cfg::kOCol0 : Cfg::kOCol1);
          const int q = (t == 0 ? q0_t0 : q0_t1) + row;
#ifdef SVGB_ATTN_TRACE
          const int tj = ntiles == 1 ? j : 2 * j + t;  // trace slot: (chunk, tile) steps in execution order
#endif
          SVGB_TRACE(half, tj, 0);
          mbar_wait(smem_u32(&bars->s_full[sb]), ntiles == 1 ? ((j >> 1) & 1) : (j & 1), 8 + t);
          tc_fence_after();
          SVGB_TRACE(half, tj, 1);

          float rs;
          auto chunk_body = [&](auto plain_tag) {
            constexpr bool kPlain = decltype(plain_tag)::value;
            uint32_t ra[32], rb[32];  // columns [c0, c0+32) and [c0+32, c0+64)
            const bool live = kPlain || mycols > 0;  // general form: a half with columns processes both groups
            if (live) {
              tmem_ld32(s_addr + c0, ra);
              tmem_ld32(s_addr + c0 + 32, rb);
            }
            tc_wait_ld();
            SVGB_TRACE(half, tj, 2);
            if constexpr (!kPlain) {
              auto sanitize = [&](uint32_t(&rr)[32], int g) {  // g: global 32-column group index (0..3)
                const int left = valid - g * 32;
                if (!live || !(elem || left < 32)) return;  // warp-uniform
                uint32_t bits = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
                if (elem && left > 0) {
                  const int qm = (args.q_index && q < args.S) ? __ldg(&args.q_index[q]) : q;
                  MaskRow mrow;
                  mrow.init(mode, qm, m0, m1, m2);
                  bits &= mrow.bits32(kv0 + g * 32);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) rr[i] = (bits >> i) & 1u ? rr[i] : 0xff800000u;  // -inf
              };
              sanitize(ra, 2 * half);
              sanitize(rb, 2 * half + 1);
            }
            float mx = -INFINITY;
            if (live) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
            }
            const float m_new = fmaxf(m_used[t], fmaxf(mx, exchange(mx)));
            SVGB_TRACE(half, tj, 3);
            // lazy rescale: keep the stale reference max unless it grew by more than tau (log2 units)
            float alpha = 1.f;
            if ((m_new - m_used[t]) * c > kTau) {  // false when both are -inf (NaN compare)
              alpha = ex2_approx((m_used[t] - m_new) * c);  // 0 when m_used == -inf
              m_used[t] = m_new;
            }
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
              // correction of this thread's half of the O row.  Two-tile items: PV_t(j-1) is complete (the S_t(j)
              // commit covered it).  Ping-pong items: PV(j-1) may still be accumulating -> wait for its commit.
              if (ntiles == 1) {
                mbar_wait(smem_u32(&bars->pv_done), (j - 1) & 1, 12);
                tc_fence_after();
              }
#pragma unroll 1
              for (int g = 0; g < D / 64; ++g) {
                uint32_t o[32];
                tmem_ld32(o_addr + half * (D / 2) + g * 32, o);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st32(o_addr + half * (D / 2) + g * 32, o);
              }
            }
            l_run[t] *= alpha;
            const float mc = (m_used[t] == -INFINITY) ? 0.f : m_used[t] * c - kPOff;
            const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
            uint64_t sum2 = pack_f32x2(0.f, 0.f);
            // P = exp2(S*c - m*c) -> 16-bit (or e4m3), packed over the first half of the S tile.  Every 4th
            // pair is evaluated on the FMA pipe (polynomial) to unload the MUFU.
            auto group_p = [&](const uint32_t(&rr)[32], int g) {  // g: global group index
              if (!live) return;
              uint32_t pk[16];
              float rs_hi[FP8 ? 16 : 1];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint64_t x2 =
                    ffma2(pack_f32x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1])), c2, nmc2);
                float p0, p1;
                if (SVGB_POLY_SEL(i)) {
                  ex2_poly2(x2, p0, p1);
                } else {
                  float x0, x1;
                  unpack_f32x2(x2, x0, x1);
                  p0 = ex2_approx(x0);
                  p1 = ex2_approx(x1);
                }
                sum2 = fadd2(sum2, pack_f32x2(p0, p1));
                if constexpr (FP8) {  // keep the fp32 pair; four of them make one e4m3x4 word below
                  pk[i] = __float_as_uint(p0);
                  rs_hi[i] = p1;
                } else {
                  pk[i] = pack2_p<BF16>(p0, p1);
                }
              }
              if constexpr (FP8) {
                uint32_t p8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  p8[i] = pack4_e4m3(__uint_as_float(pk[2 * i]), rs_hi[2 * i], __uint_as_float(pk[2 * i + 1]),
                                     rs_hi[2 * i + 1]);
                tmem_st8(s_addr + g * 8, p8);
              } else {
                tmem_st16(s_addr + g * 16, pk);
              }
            };
            group_p(ra, 2 * half);
            group_p(rb, 2 * half + 1);
            SVGB_TRACE(half, tj, 4);
            float s0, s1;
            unpack_f32x2(sum2, s0, s1);
            rs = s0 + s1;
          };
          if (!elem && valid == kChunkCols) chunk_body(std::true_type{});
          else chunk_body(std::false_type{});

          l_run[t] += rs;
          tc_wait_st();
          SVGB_TRACE(half, tj, 5);
          tc_fence_before();
          mbar_arrive(smem_u32(&bars->p_full[sb]));
          SVGB_TRACE(half, tj, 6);
        }
      }

      // ---------------- epilogue: O / l -> 16-bit -> global (optionally scattered rows); each thread stores its
      // half of the D output columns of both tiles' rows
      if (nch_any > 0) {
        mbar_wait(smem_u32(&bars->o_final), 0, 10);
        tc_fence_after();
      }
#pragma unroll 1
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t o_addr = lane_addr + (t == 0 ? Cfg::kOCol0 : Cfg::kOCol1);
        const int q = (t == 0 ? q0_t0 : q0_t1) + row;
        const bool row_ok = row < (t == 0 ? nr_t0 : nr_t1);
        const bool tile_has_o = (t == 0 ? n_t0 : n_t1) > 0;
        const float l_tot = l_run[t] + exchange(l_run[t]);
        const float inv_l = (l_tot > 0.f ? 1.f / l_tot : 0.f) * (args.v_scale ? __ldg(&args.v_scale[bh]) : 1.f);
        long long out_row = q;
        if (row_ok && args.o_rows) out_row = __ldg(&args.o_rows[static_cast<size_t>(bh) * args.S + q]);
        const int col0 = half * (D / 2);
        uint16_t* optr = reinterpret_cast<uint16_t*>(args.o) + bh * args.o_head_stride +
                         out_row * args.o_row_stride + col0;
        float* optr32 = reinterpret_cast<float*>(args.o) + bh * args.o_head_stride + out_row * args.o_row_stride + col0;
#pragma unroll 1
        for (int g = 0; g < D / 64; ++g) {
          uint32_t o[32];
          if (tile_has_o) {
            tmem_ld32(o_addr + col0 + g * 32, o);
            tc_wait_ld();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0u;
          }
          if (row_ok && args.out_f32) {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 w;
              w.x = __uint_as_float(o[4 * v + 0]) * inv_l;
              w.y = __uint_as_float(o[4 * v + 1]) * inv_l;
              w.z = __uint_as_float(o[4 * v + 2]) * inv_l;
              w.w = __uint_as_float(o[4 * v + 3]) * inv_l;
              *reinterpret_cast<float4*>(optr32 + g * 32 + v * 4) = w;
            }
          } else if (row_ok) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint4 w;
              w.x = pack2<BF16>(__uint_as_float(o[8 * v + 0]) * inv_l, __uint_as_float(o[8 * v + 1]) * inv_l);
              w.y = pack2<BF16>(__uint_as_float(o[8 * v + 2]) * inv_l, __uint_as_float(o[8 * v + 3]) * inv_l);
              w.z = pack2<BF16>(__uint_as_float(o[8 * v + 4]) * inv_l, __uint_as_float(o[8 * v + 5]) * inv_l);
              w.w = pack2<BF16>(__uint_as_float(o[8 * v + 6]) * inv_l, __uint_as_float(o[8 * v + 7]) * inv_l);
              *reinterpret_cast<uint4*>(optr + g * 32 + v * 8) = w;
            }
          }
        }
        if (row_ok && args.lse && half == 0) {
          // natural-log LSE of the scaled scores; -inf for rows that saw no key
          const float lse =
              l_tot > 0.f ? (m_used[t] * c - kPOff + log2f(l_tot)) * 0.6931471805599453f : -INFINITY;
          args.lse[static_cast<size_t>(bh) * args.S + out_row] = lse;
        }
      }
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
