// Attention plans (device-built work lists), the attention launcher, density, and the tile
// self-test.  See attn_kernel.cuh for the kernel itself.
#include <stdlib.h>

#include "../../include/svgb200.h"
#include "attn_kernel.cuh"
#include "attn_tail.cuh"
#include "host_common.h"

namespace svgb {

// =============================================================================================
// Plan: variable blocks (SVG2 dynamic map / BSR / dense)
//   one CTA per head; thread per q-block walks its map row, merges selected k-blocks that are
//   adjacent in token space into runs and cuts runs into <=128-column chunks.
// =============================================================================================
// Two kernels.  (1) plan_lists_kernel -- grid (ceil(QC/8), BH), one WARP per q-block: the map row is read with
// coalesced byte loads 32 columns at a time, a ballot gives the selected non-empty k-blocks, the (warp-uniform) run
// walk visits only the set bits and the chunks of a finished run are written by the lanes in parallel;
// (2) plan_items_kernel -- one CTA per head: work items from the row sizes and list lengths.
// (The first version walked each row byte by byte with one thread: 0.37 ms at 24 x 400 x 1000; this one ~0.02 ms.)
constexpr int kPlanWarps = 8;
__global__ void __launch_bounds__(kPlanWarps * 32)
plan_lists_kernel(const uint8_t* __restrict__ map, const int* __restrict__ row_sz, const int* __restrict__ col_sz,
                  int QC, int KC, int chunk_cap, int gather, int2* __restrict__ chunks, int* __restrict__ nch_of,
                  int* __restrict__ tot_of) {
  // gather form: the list holds RUNS {start, keys before this run} plus a sentinel {0, total}; the kernel then
  // gathers exactly-full 128-key chunks across run boundaries.
  extern __shared__ int sm[];
  int* coloff = sm;  // KC + 1
  __shared__ int wsum[kPlanWarps];
  const int bh = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  row_sz += static_cast<size_t>(bh) * QC;
  col_sz += static_cast<size_t>(bh) * KC;
  // exclusive prefix sum of the key-block sizes: a contiguous slice per thread, warp scan, serial pass over 8 warp sums
  {
    const int nthr = kPlanWarps * 32;
    const int per = (KC + nthr - 1) / nthr;
    const int j0 = min(KC, static_cast<int>(threadIdx.x) * per), j1 = min(KC, j0 + per);
    int acc = 0;
    for (int j = j0; j < j1; ++j) acc += col_sz[j];
    int incl = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = incl - acc;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    for (int j = j0; j < j1; ++j) {
      coloff[j] = base;
      base += col_sz[j];
    }
    if (threadIdx.x == nthr - 1) coloff[KC] = base;  // the last thread's slice ends at KC (possibly empty)
  }
  __syncthreads();
  const int qb = blockIdx.x * kPlanWarps + warp;
  if (qb >= QC) return;
  const size_t list = (static_cast<size_t>(bh) * QC + qb) * chunk_cap;
  int2* out = chunks + list;
  const uint8_t* mrow = map + (static_cast<size_t>(bh) * QC + qb) * KC;
  int n = 0, total = 0;  // warp-uniform
  if (row_sz[qb] > 0) {
    int run_s = -1, run_e = -1;
    auto flush = [&]() {
      if (gather) {
        if (n < chunk_cap - 1) {
          if (lane == 0) out[n] = make_int2(run_s, total);
          ++n;
        }
        total += run_e - run_s;
      } else {
        const int cnt = min((run_e - run_s + kChunkCols - 1) / kChunkCols, chunk_cap - n);
        for (int c = lane; c < cnt; c += 32) {
          const int p = run_s + c * kChunkCols;
          out[n + c] = make_int2(p, chunk_meta(min(kChunkCols, run_e - p), false));
        }
        n += cnt;
      }
    };
    for (int j0 = 0; j0 < KC; j0 += 32) {
      const int j = j0 + lane;
      const bool sel = j < KC && mrow[j] != 0 && coloff[j + 1] > coloff[j];
      unsigned m = __ballot_sync(0xffffffffu, sel);
      while (m) {
        const int jj = j0 + __ffs(m) - 1;
        m &= m - 1;
        const int s = coloff[jj], e = coloff[jj + 1];
        if (run_s >= 0 && s == run_e) {
          run_e = e;  // contiguous in token space (only empty k-blocks in between): extend
        } else {
          if (run_s >= 0) flush();
          run_s = s;
          run_e = e;
        }
      }
    }
    if (run_s >= 0) flush();
    if (gather && lane == 0) out[n] = make_int2(0, total);  // sentinel: ends the last run
  }
  if (lane == 0) {
    nch_of[static_cast<size_t>(bh) * QC + qb] = n;
    tot_of[static_cast<size_t>(bh) * QC + qb] = total;
  }
}

__global__ void plan_items_kernel(const int* __restrict__ row_sz, const int* __restrict__ nch_all,
                                  const int* __restrict__ tot_all, int QC, int max_items, int chunk_cap,
                                  int pair_tails, int short_rows, int tmax, int* __restrict__ counts,
                                  int4* __restrict__ items, int4* __restrict__ items2, int* __restrict__ item_total,
                                  int4* __restrict__ titems, int4* __restrict__ titems2, int* __restrict__ tcount) {
  // TMA form (items2 != nullptr): a q-block of r rows becomes r/256 two-tile items (+ one more when r % 256 > 128,
  // its second tile partial) and, when 0 < r % 256 <= 128, a single-tile TAIL.  Tails are where k-means clusters
  // lose tensor-core rows (a 297-row cluster = 256 + 41) and a lone single-tile CTA leaves half of the softmax
  // warps idle, so the tails of a head are sorted by chunk count and packed two per CTA as DUAL items
  // (items[i] = stream of T0, items2[i] = stream of T1; see AttnArgs::items2).  Launch order: two-tile items in
  // q-block order, then the dual items from long to short.  Tails of at most `short_rows` rows go to a separate list
  // (titems / titems2 / tcount, same format) that the transposed tail kernel runs (attn_tail.cuh).
  const bool gather = item_total != nullptr;
  extern __shared__ int sm[];
  int* rowoff = sm;                 // QC + 1
  int* itembase = rowoff + QC + 1;  // QC + 1 : first two-tile item of the q-block (gather: first item)
  int* long_of = itembase + QC + 1; // QC     : rows of the q-block's long tail (65..128; 0 = none) -> dual item
  int* short_of = long_of + QC;     // QC     : rows of its short tail (1..short_rows; 0 = none) -> transposed kernel
  int* nch_of = short_of + QC;      // QC
  __shared__ int s_total2, s_nlong, s_nshort;
  const int bh = blockIdx.x;
  row_sz += static_cast<size_t>(bh) * QC;
  const bool pair = !gather && pair_tails;
  for (int i = threadIdx.x; i < QC; i += blockDim.x) {
    nch_of[i] = nch_all[static_cast<size_t>(bh) * QC + i];
    rowoff[i] = row_sz[i];  // staged: the serial pass below turns the sizes into offsets in place
  }
  __syncthreads();
  // split of the last r % 256 rows of a q-block (pair == true):
  //   rem <= short                : short tail
  //   short < rem <= 128          : long tail
  //   128 < rem <= 128 + short    : long tail of 128 rows + short tail of rem - 128
  //   rem > 128 + short           : one more two-tile item (second tile partial)
  if (threadIdx.x == 32) {
    int acc = 0, it = 0, nl = 0, ns = 0;
    for (int i = 0; i < QC; ++i) {
      const int r = rowoff[i];
      rowoff[i] = acc;
      itembase[i] = it;
      acc += r;
      const int rem = r % kItemRows;
      int lg = 0, sh = 0, n2 = (r + kItemRows - 1) / kItemRows;
      if (pair && rem > 0) {
        if (rem <= short_rows) sh = rem;
        else if (rem <= kTileRows) lg = rem;
        else if (rem <= kTileRows + short_rows) { lg = kTileRows; sh = rem - kTileRows; }
        if (lg > 0 || sh > 0) n2 = r / kItemRows;
      }
      long_of[i] = lg;
      short_of[i] = sh;
      nl += lg > 0;
      ns += sh > 0;
      it += n2;
    }
    rowoff[QC] = acc;
    itembase[QC] = it;
    s_total2 = it;
    s_nlong = nl;
    s_nshort = ns;
    const int total = it + (nl + 1) / 2;
    counts[bh] = total < max_items ? total : max_items;
    if (tcount) tcount[bh] = min((ns + 1) / 2, tmax);
  }
  __syncthreads();
  for (int qb = threadIdx.x; qb < QC; qb += blockDim.x) {
    const int r = rowoff[qb + 1] - rowoff[qb];
    if (r == 0) continue;
    const size_t list = (static_cast<size_t>(bh) * QC + qb) * chunk_cap;
    const int n = nch_of[qb];
    const int nit = itembase[qb + 1] - itembase[qb];
    for (int t = 0; t < nit; ++t) {
      const int idx = itembase[qb] + t;
      if (idx < max_items) {
        items[static_cast<size_t>(bh) * max_items + idx] =
            make_int4(rowoff[qb] + t * kItemRows, min(kItemRows, r - t * kItemRows), static_cast<int>(list), n);
        if (items2) items2[static_cast<size_t>(bh) * max_items + idx] = make_int4(0, 0, 0, 0);
        if (gather) item_total[static_cast<size_t>(bh) * max_items + idx] = tot_all[static_cast<size_t>(bh) * QC + qb];
      }
    }
  }
  if (!pair) return;
  // tails of each class (long -> dual items of the main kernel, short -> transposed kernel): rank by (chunk count
  // descending, q-block ascending); ranks 2i and 2i+1 share CTA i
  const int base = s_total2;
  for (int cls = 0; cls < 2; ++cls) {
    const int* rows_of = cls == 0 ? long_of : short_of;
    const int n_class = cls == 0 ? s_nlong : s_nshort;
    int4* a = cls == 0 ? items + static_cast<size_t>(bh) * max_items + base : titems + static_cast<size_t>(bh) * tmax;
    int4* b = cls == 0 ? items2 + static_cast<size_t>(bh) * max_items + base : titems2 + static_cast<size_t>(bh) * tmax;
    const int cap = cls == 0 ? max_items - base : tmax;
    for (int qb = threadIdx.x; qb < QC; qb += blockDim.x) {
      const int rows = rows_of[qb];
      if (rows == 0) continue;
      const int mine = nch_of[qb];
      int rank = 0;
      for (int o = 0; o < QC; ++o)
        if (rows_of[o] > 0 && (nch_of[o] > mine || (nch_of[o] == mine && o < qb))) ++rank;
      const int r = rowoff[qb + 1] - rowoff[qb];
      const int off = (r / kItemRows) * kItemRows + ((cls == 1 && long_of[qb] > 0) ? kTileRows : 0);
      const int4 it = make_int4(rowoff[qb] + off, rows, static_cast<int>((static_cast<size_t>(bh) * QC + qb) * chunk_cap), mine);
      const int idx = rank >> 1;
      if (idx < cap) {
        if ((rank & 1) == 0) {
          a[idx] = it;
          if (rank == n_class - 1) b[idx] = make_int4(0, 0, 0, 0);  // odd one out
        } else {
          b[idx] = it;
        }
      }
    }
  }
}

// =============================================================================================
// Plan: element-exact band masks (SVG1).  One CTA per work item; thread = query row; for every
// 128-column chunk the block votes any/all over the item's rows.  Exact by construction.
// Items never straddle a row-region boundary of the mask (video | prompt | padding for HunyuanVideo,
// text | video for CogVideoX), so text rows -- which see every column -- do not drag 256-row items of
// video rows onto the per-element path; the heavy text items are issued first (LPT).
// =============================================================================================
struct BandSegs {
  int n;             // number of row segments
  int bound[6];      // bound[s] .. bound[s+1]
  int item_base[6];  // first item index of segment s (launch order)
  int item_cnt[6];
};

__global__ void plan_band_kernel(int mode, int m0, int m1, int m2, int S, int n_chunks_total, BandSegs segs,
                                 int* __restrict__ counts, int4* __restrict__ items,
                                 int2* __restrict__ chunks) {
  const int item = blockIdx.x;
  int seg = 0;
  for (int s = 0; s < segs.n; ++s)
    if (item >= segs.item_base[s] && item < segs.item_base[s] + segs.item_cnt[s]) seg = s;
  const int q_row0 = segs.bound[seg] + (item - segs.item_base[seg]) * kItemRows;
  const int nrows = min(kItemRows, segs.bound[seg + 1] - q_row0);
  const int q = q_row0 + threadIdx.x;
  const bool row_live = static_cast<int>(threadIdx.x) < nrows;
  int2* out = chunks + static_cast<size_t>(item) * n_chunks_total;
  int n = 0;
  for (int c = 0; c < n_chunks_total; ++c) {
    const int kv0 = c * kChunkCols;
    const int valid = min(kChunkCols, S - kv0);
    bool any = false, all = true;
    if (row_live) {
      for (int i = 0; i < valid; ++i) {
        const bool a = mask_allowed(mode, q, kv0 + i, m0, m1, m2);
        any |= a;
        all &= a;
      }
    }
    const int block_any = __syncthreads_or(any ? 1 : 0);
    const int block_all = __syncthreads_and((all || !row_live) ? 1 : 0);
    if (block_any) {
      if (threadIdx.x == 0) out[n] = make_int2(kv0, chunk_meta(valid, !block_all));
      ++n;
    }
  }
  if (threadIdx.x == 0) {
    items[item] = make_int4(q_row0, nrows, static_cast<int>(static_cast<size_t>(item) * n_chunks_total), n);
    if (item == 0) counts[0] = gridDim.x;
  }
}

// row-region boundaries of each mask family, and the launch order of the segments
static BandSegs band_segments(int mode, int m0, int m1, int S) {
  int cuts[4], nc = 0;
  auto add = [&](int c) {
    if (c > 0 && c < S && (nc == 0 || cuts[nc - 1] < c)) cuts[nc++] = c;
  };
  if (mode == MASK_HY) {
    add(m0);  // end of video
    add(m1);  // end of real prompt
  } else if (mode == MASK_COG) {
    add(m1);  // end of text prefix
  }
  BandSegs sg{};
  sg.n = nc + 1;
  sg.bound[0] = 0;
  for (int i = 0; i < nc; ++i) sg.bound[i + 1] = cuts[i];
  sg.bound[sg.n] = S;
  for (int s = 0; s < sg.n; ++s) sg.item_cnt[s] = (sg.bound[s + 1] - sg.bound[s] + kItemRows - 1) / kItemRows;
  int base = 0;
  if (mode == MASK_HY) {  // text segments (heavy: they see every column) first
    for (int s = sg.n - 1; s >= 0; --s) {
      sg.item_base[s] = base;
      base += sg.item_cnt[s];
    }
  } else {
    for (int s = 0; s < sg.n; ++s) {
      sg.item_base[s] = base;
      base += sg.item_cnt[s];
    }
  }
  return sg;
}
static int band_total_items(const BandSegs& sg) {
  int t = 0;
  for (int s = 0; s < sg.n; ++s) t += sg.item_cnt[s];
  return t;
}

// =============================================================================================
// density_calculation (svg/kmeans_utils.py:13-31)
// =============================================================================================
__global__ void density_kernel(const uint8_t* __restrict__ map, const int* __restrict__ row_sz,
                               const int* __restrict__ col_sz, int QC, int KC,
                               float* __restrict__ density) {
  const int bh = blockIdx.x;
  __shared__ unsigned long long s_num[32], s_den[32];
  unsigned long long num = 0, den = 0;
  for (int idx = threadIdx.x; idx < QC * KC; idx += blockDim.x) {
    const int i = idx / KC, j = idx - i * KC;
    const unsigned long long a =
        static_cast<unsigned long long>(row_sz[static_cast<size_t>(bh) * QC + i]) *
        static_cast<unsigned long long>(col_sz[static_cast<size_t>(bh) * KC + j]);
    den += a;
    if (map[(static_cast<size_t>(bh) * QC + i) * KC + j]) num += a;
  }
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_num[threadIdx.x >> 5] = num;
    s_den[threadIdx.x >> 5] = den;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    num = den = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) {
      num += s_num[w];
      den += s_den[w];
    }
    density[bh] = static_cast<float>(static_cast<double>(num) / static_cast<double>(den));
  }
}

// =============================================================================================
// Tile self-test: the exact descriptor paths of the attention kernel on one 128x128xD tile.
// =============================================================================================
template <int D, bool BF16>
__global__ void __launch_bounds__(128, 1)
selftest_tile_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                     const __grid_constant__ CUtensorMap vmap, float* __restrict__ s_out,
                     float* __restrict__ o_out, float p_scale) {
  using Cfg = AttnCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base, sK = sQ + Cfg::kTileBytes, sV = sK + Cfg::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + 3 * Cfg::kTileBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ld_bar = smem_u32(&bars[0]), mma_bar = smem_u32(&bars[1]), mma2_bar = smem_u32(&bars[2]);
  if (threadIdx.x == 0) {
    mbar_init(ld_bar, 1);
    mbar_init(mma_bar, 1);
    mbar_init(mma2_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(ld_bar, 3 * Cfg::kTileBytes);
    for (int h = 0; h < Cfg::kHalves; ++h) {
      tma_load_3d(sQ + h * Cfg::kPanelBytes, &qmap, ld_bar, h * 64, 0, 0);
      tma_load_3d(sK + h * Cfg::kPanelBytes, &kmap, ld_bar, h * 64, 0, 0);
      tma_load_3d(sV + h * Cfg::kPanelBytes, &vmap, ld_bar, h * 64, 0, 0);
    }
    mbar_wait(ld_bar, 0, 20);
    const uint32_t idesc = make_idesc(128, 128, BF16, false, false);
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
      const uint32_t off = (kk >> 2) * Cfg::kPanelBytes + (kk & 3) * 32;
      mma_ss(tmem + 0, desc_kmajor_sw128(sQ + off), desc_kmajor_sw128(sK + off), idesc, kk > 0);
    }
    tc_commit(mma_bar);
  }
  // S -> global, P = 16-bit(S * p_scale) -> TMEM columns [0,64)
  mbar_wait(mma_bar, 0, 21);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int g = 0; g < 4; ++g) {
    uint32_t r[32];
    tmem_ld32(lane_addr + g * 32, r);
    tc_wait_ld();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; ++i) s_out[row * 128 + g * 32 + i] = __uint_as_float(r[i]);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      pk[i] = pack2<BF16>(__uint_as_float(r[2 * i]) * p_scale, __uint_as_float(r[2 * i + 1]) * p_scale);
    tmem_st16(lane_addr + g * 16, pk);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, D, BF16, false, true);
    for (int kk = 0; kk < 8; ++kk)
      mma_ts(tmem + 128, tmem + kk * 8, desc_mnmajor_sw128(sV + kk * 2048, Cfg::kPanelBytes), idesc,
             kk > 0);
    tc_commit(mma2_bar);
  }
  mbar_wait(mma2_bar, 0, 22);
  tc_fence_after();
  for (int g = 0; g < D / 32; ++g) {
    uint32_t r[32];
    tmem_ld32(lane_addr + 128 + g * 32, r);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) o_out[row * D + g * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

template <int D, int DT>
static int launch_attn(const CUtensorMap& qm, const CUtensorMap& km, const CUtensorMap& vm,
                       const AttnArgs& args, dim3 grid, cudaStream_t stream) {
  using Cfg = AttnCfg<D, DT>;
  if constexpr (DT != DT_E4M3) {
    if (args.gather) {
      auto kern = attn_fwd_kernel<D, DT, true>;
      SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(qm, km, vm, args);
      SVGB_LAUNCH_OK();
      return 0;
    }
  }
  if constexpr (DT != DT_E4M3 && D == 128) {
    // experimental sub-chunk pipeline (attn_kernel.cuh, kSub): opt-in until it is validated on the GPU
    static const int sub = [] { const char* e = getenv("SVGB_ATTN_SUB"); return (e && e[0] == '1') ? 1 : 0; }();
    if (sub && !args.softmax_shared && !args.items2) {
      auto kern = attn_fwd_kernel<D, DT, false, true>;
      SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      AttnArgs sub_args = args;
      sub_args.sub_mode = sub;
      kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(qm, km, vm, sub_args);
      SVGB_LAUNCH_OK();
      return 0;
    }
  }
  {
    auto kern = attn_fwd_kernel<D, DT, false>;
    SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(qm, km, vm, args);
  }
  SVGB_LAUNCH_OK();
  return 0;
}

template <int D, bool BF16>
static int launch_selftest(const CUtensorMap& qm, const CUtensorMap& km, const CUtensorMap& vm,
                           float* s_out, float* o_out, float p_scale, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  auto kern = selftest_tile_kernel<D, BF16>;
  const int smem = 1024 + 3 * Cfg::kTileBytes + 256;
  SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<1, 128, smem, stream>>>(qm, km, vm, s_out, o_out, p_scale);
  SVGB_LAUNCH_OK();
  return 0;
}

int attn_fwd_impl(const void* q, int Sq, long long q_rs, long long q_hs, const void* k, const void* v, int Skv,
                  long long kv_rs, long long kv_hs, int dtype, int BH, int D, const AttnArgs& a, int grid_x,
                  cudaStream_t st) {
  SVGB_REQUIRE(D == 64 || D == 128, "head_dim %d unsupported (64 or 128)", D);
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16 || dtype == SVGB_E4M3, "dtype %d unsupported", dtype);
  CUtensorMap qm, km, vm;
  if (encode_tmap_hsd(&qm, q, dtype, BH, Sq, D, q_rs, q_hs)) return -1;
  if (encode_tmap_hsd(&km, k, dtype, BH, Skv, D, kv_rs, kv_hs)) return -1;
  if (encode_tmap_hsd(&vm, v, dtype, BH, Skv, D, kv_rs, kv_hs)) return -1;
  dim3 grid(grid_x, BH);
  if (dtype == SVGB_E4M3) {
    SVGB_REQUIRE(D == 128 && !a.gather, "the e4m3 path needs head_dim 128 and a non-gather plan");
    return launch_attn<128, DT_E4M3>(qm, km, vm, a, grid, st);
  }
  if (D == 128) {
    return dtype == SVGB_BF16 ? launch_attn<128, DT_BF16>(qm, km, vm, a, grid, st)
                              : launch_attn<128, DT_F16>(qm, km, vm, a, grid, st);
  }
  return dtype == SVGB_BF16 ? launch_attn<64, DT_BF16>(qm, km, vm, a, grid, st)
                            : launch_attn<64, DT_F16>(qm, km, vm, a, grid, st);
}

// transposed tail kernel launch (attn_tail.cuh): bf16 / fp16, head_dim 128, 16-bit output without LSE
int attn_tail_impl(const void* q, int Sq, long long q_rs, long long q_hs, const void* k, const void* v, int Skv,
                   long long kv_rs, long long kv_hs, int dtype, int BH, const AttnArgs& a, int grid_x, cudaStream_t st) {
  CUtensorMap qm, km, vm;
  if (encode_tmap_hsd(&qm, q, dtype, BH, Sq, 128, q_rs, q_hs, kTailRows)) return -1;
  if (encode_tmap_hsd(&km, k, dtype, BH, Skv, 128, kv_rs, kv_hs)) return -1;
  if (encode_tmap_hsd(&vm, v, dtype, BH, Skv, 128, kv_rs, kv_hs)) return -1;
  dim3 grid(grid_x, BH);
  if (dtype == SVGB_BF16) {
    auto kern = attn_tail_kernel<DT_BF16>;
    SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TailCfg<DT_BF16>::kSmemBytes));
    kern<<<grid, 384, TailCfg<DT_BF16>::kSmemBytes, st>>>(qm, km, vm, a);
  } else {
    auto kern = attn_tail_kernel<DT_F16>;
    SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TailCfg<DT_F16>::kSmemBytes));
    kern<<<grid, 384, TailCfg<DT_F16>::kSmemBytes, st>>>(qm, km, vm, a);
  }
  SVGB_LAUNCH_OK();
  return 0;
}

static int varblock_chunk_cap(int S, int KC) { return S / kChunkCols + (KC + 1) / 2 + 2; }

// scratch + transposed-tail lists that follow the aux region of a variable-block plan workspace
struct VarTailLayout {
  size_t nch_off, tot_off, t1_off, t2_off, tc_off;
  int tmax;
};
static VarTailLayout var_tail_layout(long long aux_off, int BH, int max_items, int QC) {
  VarTailLayout L;
  const size_t aux_bytes = align_up(sizeof(int4) * BH * max_items, 256);
  const size_t len_bytes = align_up(sizeof(int) * static_cast<size_t>(BH) * QC, 256);
  L.tmax = (QC + 1) / 2;
  const size_t t_bytes = align_up(sizeof(int4) * static_cast<size_t>(BH) * L.tmax, 256);
  L.nch_off = aux_off + aux_bytes;
  L.tot_off = L.nch_off + len_bytes;
  L.t1_off = L.tot_off + len_bytes;
  L.t2_off = L.t1_off + t_bytes;
  L.tc_off = L.t2_off + t_bytes;
  return L;
}
static int varblock_max_items(int S, int QC) { return S / kItemRows + QC + 1; }

}  // namespace svgb

using namespace svgb;

extern "C" {

int svgb_attn_plan_varblock_bytes(int BH, int S, int QC, int KC, size_t* bytes) {
  SVGB_REQUIRE(BH > 0 && S > 0 && QC > 0 && KC > 0 && bytes, "bad arguments");
  const size_t counts = align_up(sizeof(int) * BH, 256);
  const size_t items = align_up(sizeof(int4) * BH * varblock_max_items(S, QC), 256);
  const size_t chunks = align_up(sizeof(int2) * static_cast<size_t>(BH) * QC * varblock_chunk_cap(S, KC), 256);
  // aux: TMA plans = second stream of dual items (int4 per item); gather plans = selected keys per item (int)
  const size_t aux = align_up(sizeof(int4) * BH * varblock_max_items(S, QC), 256);
  const size_t lens = 2 * align_up(sizeof(int) * static_cast<size_t>(BH) * QC, 256);  // per q-block list length / key total
  // transposed-tail items: two int4 arrays of (QC+1)/2 pairs per head + a count per head
  const size_t tails = 2 * align_up(sizeof(int4) * static_cast<size_t>(BH) * ((QC + 1) / 2), 256) + align_up(sizeof(int) * BH, 256);
  *bytes = counts + items + chunks + aux + lens + tails;
  return 0;
}

static int plan_varblock_impl(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH,
                              int S, int QC, int KC, void* plan_ws, size_t ws_bytes, svgb_plan* plan,
                              void* stream, bool gather) {
  size_t need = 0;
  if (svgb_attn_plan_varblock_bytes(BH, S, QC, KC, &need)) return -1;
  SVGB_REQUIRE(map && row_sz && col_sz && plan_ws && plan, "null pointer");
  SVGB_REQUIRE(ws_bytes >= need, "plan workspace too small: %zu < %zu", ws_bytes, need);
  SVGB_REQUIRE(static_cast<size_t>(BH) * QC * varblock_chunk_cap(S, KC) < (1ull << 31),
               "plan too large for 32-bit chunk offsets");
  const int max_items = varblock_max_items(S, QC);
  const int cap = varblock_chunk_cap(S, KC);
  plan->kind = gather ? 3 : 1;
  plan->BH = BH;
  plan->S = S;
  plan->max_items = max_items;
  plan->items_stride = max_items;
  plan->counts_stride = 1;
  plan->mask_mode = MASK_NONE;
  plan->m0 = gather ? (KC + 1) / 2 + 1 : 0;  // gather: upper bound of runs per q-block (kernel smem table)
  plan->m1 = S / KC;                          // average key-cluster size (selects the softmax thread mapping)
  plan->m2 = 0;
  plan->counts_off = 0;
  plan->items_off = align_up(sizeof(int) * BH, 256);
  plan->chunks_off = plan->items_off + align_up(sizeof(int4) * BH * max_items, 256);
  plan->aux_off = plan->chunks_off + align_up(sizeof(int2) * static_cast<size_t>(BH) * QC * cap, 256);
  plan->bytes = need;
  char* ws = static_cast<char*>(plan_ws);
  const VarTailLayout tl = var_tail_layout(plan->aux_off, BH, max_items, QC);
  int* nch_of = reinterpret_cast<int*>(ws + tl.nch_off);
  int* tot_of = reinterpret_cast<int*>(ws + tl.tot_off);
  plan->m2 = gather ? 0 : QC;  // lets the launcher find the transposed-tail lists (var_tail_layout)
  const size_t smem1 = sizeof(int) * (KC + 1);
  const size_t smem2 = sizeof(int) * (2 * (QC + 1) + 3 * QC);
  SVGB_REQUIRE(smem1 <= 48 * 1024 && smem2 <= 48 * 1024, "QC/KC too large for the plan kernels (%zu / %zu B smem)", smem1, smem2);
  // SVGB_ATTN_PAIR=0 keeps every tail a single-tile item (A/B switch for bring-up)
  static const int pair_tails = [] { const char* e = getenv("SVGB_ATTN_PAIR"); return (e && e[0] == '0') ? 0 : 1; }();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  plan_lists_kernel<<<dim3((QC + kPlanWarps - 1) / kPlanWarps, BH), kPlanWarps * 32, smem1, st>>>(map, row_sz, col_sz, QC, KC, cap, gather ? 1 : 0,
                                                                 reinterpret_cast<int2*>(ws + plan->chunks_off), nch_of, tot_of);
  SVGB_LAUNCH_OK();
  // SVGB_ATTN_TAIL=0 keeps short tails in the main kernel's dual items (A/B switch for bring-up)
  static const int short_rows = [] { const char* e = getenv("SVGB_ATTN_TAIL"); return (e && e[0] == '0') ? 0 : kTailRows; }();
  plan_items_kernel<<<BH, 256, smem2, st>>>(
      row_sz, nch_of, tot_of, QC, max_items, cap, pair_tails, gather ? 0 : short_rows, tl.tmax,
      reinterpret_cast<int*>(ws + plan->counts_off), reinterpret_cast<int4*>(ws + plan->items_off),
      gather ? nullptr : reinterpret_cast<int4*>(ws + plan->aux_off),
      gather ? reinterpret_cast<int*>(ws + plan->aux_off) : nullptr, reinterpret_cast<int4*>(ws + tl.t1_off),
      reinterpret_cast<int4*>(ws + tl.t2_off), reinterpret_cast<int*>(ws + tl.tc_off));
  SVGB_LAUNCH_OK();
  return 0;
}

int svgb_attn_plan_varblock(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH,
                            int S, int QC, int KC, void* plan_ws, size_t ws_bytes, svgb_plan* plan,
                            void* stream) {
  return plan_varblock_impl(map, row_sz, col_sz, BH, S, QC, KC, plan_ws, ws_bytes, plan, stream, false);
}

int svgb_attn_plan_varblock_gather(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH,
                                   int S, int QC, int KC, void* plan_ws, size_t ws_bytes, svgb_plan* plan,
                                   void* stream) {
  return plan_varblock_impl(map, row_sz, col_sz, BH, S, QC, KC, plan_ws, ws_bytes, plan, stream, true);
}

int svgb_attn_plan_band_bytes(int S, size_t* bytes) {
  SVGB_REQUIRE(S > 0 && bytes, "bad arguments");
  const size_t n_items = (S + kItemRows - 1) / kItemRows + 4;  // + one partial item per extra row segment
  const size_t n_chunks = (S + kChunkCols - 1) / kChunkCols;
  *bytes = 256 + align_up(sizeof(int4) * n_items, 256) + align_up(sizeof(int2) * n_items * n_chunks, 256);
  return 0;
}

int svgb_attn_plan_band(int mask_mode, int m0, int m1, int m2, int BH, int S, void* plan_ws,
                        size_t ws_bytes, svgb_plan* plan, void* stream) {
  size_t need = 0;
  if (svgb_attn_plan_band_bytes(S, &need)) return -1;
  SVGB_REQUIRE(plan_ws && plan, "null pointer");
  SVGB_REQUIRE(ws_bytes >= need, "plan workspace too small: %zu < %zu", ws_bytes, need);
  SVGB_REQUIRE(mask_mode >= MASK_NONE && mask_mode <= MASK_COG, "unknown mask mode %d", mask_mode);
  const BandSegs sg = band_segments(mask_mode, m0, m1, S);
  const int n_items = band_total_items(sg);
  const int n_items_cap = (S + kItemRows - 1) / kItemRows + 4;
  const int n_chunks = (S + kChunkCols - 1) / kChunkCols;
  SVGB_REQUIRE(n_items <= n_items_cap, "internal: item count");
  SVGB_REQUIRE(static_cast<size_t>(n_items_cap) * n_chunks < (1ull << 31), "plan too large");
  plan->kind = 2;
  plan->BH = BH;
  plan->S = S;
  plan->max_items = n_items;
  plan->items_stride = 0;
  plan->counts_stride = 0;
  plan->mask_mode = mask_mode;
  plan->m0 = m0;
  plan->m1 = m1;
  plan->m2 = m2;
  plan->counts_off = 0;
  plan->items_off = 256;
  plan->chunks_off = 256 + align_up(sizeof(int4) * n_items_cap, 256);
  plan->aux_off = 0;
  plan->bytes = need;
  char* ws = static_cast<char*>(plan_ws);
  plan_band_kernel<<<n_items, kItemRows, 0, static_cast<cudaStream_t>(stream)>>>(
      mask_mode, m0, m1, m2, S, n_chunks, sg, reinterpret_cast<int*>(ws + plan->counts_off),
      reinterpret_cast<int4*>(ws + plan->items_off), reinterpret_cast<int2*>(ws + plan->chunks_off));
  SVGB_LAUNCH_OK();
  return 0;
}

static int attn_fwd_entry(const void* q, const void* k, const void* v, const float* q_scale, const float* k_scale,
                          const float* v_scale, void* o, float* lse,
                          const int32_t* q_rows, const int32_t* kv_rows, const int32_t* o_rows, int dtype, int BH,
                          int S, int D, long long row_stride, long long head_stride, long long o_row_stride,
                          long long o_head_stride, float sm_scale, const svgb_plan* plan, const void* plan_ws,
                          void* stream) {
  SVGB_REQUIRE(q && k && v && o && plan && plan_ws, "null pointer");
  SVGB_REQUIRE(D == 64 || D == 128, "head_dim %d unsupported (64 or 128)", D);
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16 || dtype == SVGB_E4M3, "dtype %d unsupported", dtype);
  SVGB_REQUIRE(plan->S == S && (plan->items_stride == 0 || plan->BH == BH),
               "plan was built for BH=%d S=%d, called with BH=%d S=%d", plan->BH, plan->S, BH, S);
  SVGB_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0 && o_row_stride % 8 == 0 && o_head_stride % 8 == 0,
               "output must be 16-byte aligned with strides multiple of 8 elements");
  SVGB_REQUIRE(plan->kind == 3 || (!q_rows && !kv_rows), "row gathers need a plan from svgb_attn_plan_varblock_gather");
  SVGB_REQUIRE(plan->kind != 3 || plan->m0 <= AttnCfg<64>::kMaxRunsSmem, "gather plan has too many runs per q-block (%d)", plan->m0);
  const char* ws = static_cast<const char*>(plan_ws);
  AttnArgs a;
  a.items = reinterpret_cast<const int4*>(ws + plan->items_off);
  a.item_count = reinterpret_cast<const int*>(ws + plan->counts_off);
  a.chunks = reinterpret_cast<const int2*>(ws + plan->chunks_off);
  a.items2 = plan->kind == 1 ? reinterpret_cast<const int4*>(ws + plan->aux_off) : nullptr;
  a.items_stride = plan->items_stride;
  a.counts_stride = plan->counts_stride;
  a.o = o;
  a.o_row_stride = o_row_stride;
  a.o_head_stride = o_head_stride;
  a.o_rows = o_rows;
  a.lse = lse;
  a.scale_log2 = sm_scale * 1.4426950408889634f;
  a.S = S;
  a.mask_mode = plan->mask_mode;
  a.m0 = plan->m0;
  a.m1 = plan->m1;
  a.m2 = plan->m2;
  a.q_index = nullptr;
  a.out_f32 = 0;
  // Softmax thread mapping of two-tile (and dual) items: 0 = one warpgroup per tile, 1 = both warpgroups share a tile.
  // Round 1 chose the shared mapping for plans made of narrow chunks (small key clusters); with the straight-line
  // masked specialisations (attn_kernel.cuh) the per-tile mapping is as fast there (705 vs 704 TF/s at QC=400/KC=1000)
  // and 4-5 % faster on aligned plans (1177-1188 vs 1126), so it is used everywhere.  Single-tile items always use the
  // shared mapping.
  a.softmax_shared = 0;
  {  // SVGB_ATTN_MAP=0 / 1 forces the per-tile / shared softmax mapping (A/B switch for bring-up)
    static const int force = [] { const char* e = getenv("SVGB_ATTN_MAP"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1; }();
    if (force >= 0) a.softmax_shared = force;
  }
  a.sub_mode = 0;
  {
    static const int order = [] { const char* e = getenv("SVGB_ATTN_ORDER"); return (e && e[0] == '1') ? 1 : 0; }();
    a.arrival_order = order;
  }
  a.q_scale = q_scale;
  a.k_scale = k_scale;
  a.v_scale = v_scale;
  a.gather = plan->kind == 3 ? 1 : 0;
  a.item_total = plan->kind == 3 ? reinterpret_cast<const int*>(ws + plan->aux_off) : nullptr;
  a.q_rows = q_rows;
  a.kv_rows = kv_rows;
  a.q_ptr = q;
  a.k_ptr = k;
  a.v_ptr = v;
  a.in_row_stride = row_stride;
  a.in_head_stride = head_stride;
  a.titems = a.titems2 = nullptr;
  a.tcount = nullptr;
  if (attn_fwd_impl(q, S, row_stride, head_stride, k, v, S, row_stride, head_stride, dtype, BH, D, a,
                    plan->max_items, static_cast<cudaStream_t>(stream)))
    return -1;
  if (plan->kind == 1 && plan->m2 > 0) {
    // short tails (<= 64 rows) of the variable-block plan: transposed kernel when it applies, otherwise the same
    // lists go through the main kernel as dual items
    const VarTailLayout tl = var_tail_layout(plan->aux_off, plan->BH, plan->max_items, plan->m2);
    AttnArgs t = a;
    t.items = reinterpret_cast<const int4*>(ws + tl.t1_off);
    t.items2 = reinterpret_cast<const int4*>(ws + tl.t2_off);
    t.item_count = reinterpret_cast<const int*>(ws + tl.tc_off);
    t.items_stride = plan->items_stride ? tl.tmax : 0;
    const bool transposed = D == 128 && (dtype == SVGB_BF16 || dtype == SVGB_F16) && !lse;
    if (transposed)
      return attn_tail_impl(q, S, row_stride, head_stride, k, v, S, row_stride, head_stride, dtype, BH, t, tl.tmax,
                            static_cast<cudaStream_t>(stream));
    return attn_fwd_impl(q, S, row_stride, head_stride, k, v, S, row_stride, head_stride, dtype, BH, D, t, tl.tmax,
                         static_cast<cudaStream_t>(stream));
  }
  return 0;
}

int svgb_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                  const int32_t* o_rows, int dtype, int BH, int S, int D, long long row_stride,
                  long long head_stride, long long o_row_stride, long long o_head_stride,
                  float sm_scale, const svgb_plan* plan, const void* plan_ws, void* stream) {
  return attn_fwd_entry(q, k, v, nullptr, nullptr, nullptr, o, lse, nullptr, nullptr, o_rows, dtype, BH, S, D,
                        row_stride, head_stride, o_row_stride, o_head_stride, sm_scale, plan, plan_ws, stream);
}

int svgb_attn_fwd_fp8(const void* q8, const void* k8, const void* v8, const float* q_scale, const float* k_scale,
                      const float* v_scale, void* o, float* lse, const int32_t* o_rows, int BH, int S, int D,
                      long long row_stride, long long head_stride, long long o_row_stride,
                      long long o_head_stride, float sm_scale, const svgb_plan* plan, const void* plan_ws,
                      void* stream) {
  SVGB_REQUIRE(q_scale && k_scale && v_scale, "fp8 attention needs the three per-head scale vectors");
  return attn_fwd_entry(q8, k8, v8, q_scale, k_scale, v_scale, o, lse, nullptr, nullptr, o_rows, SVGB_E4M3, BH, S,
                        D, row_stride, head_stride, o_row_stride, o_head_stride, sm_scale, plan, plan_ws, stream);
}

int svgb_attn_fwd_gather(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* q_rows,
                         const int32_t* kv_rows, const int32_t* o_rows, int dtype, int BH, int S, int D,
                         long long row_stride, long long head_stride, long long o_row_stride,
                         long long o_head_stride, float sm_scale, const svgb_plan* plan, const void* plan_ws,
                         void* stream) {
  return attn_fwd_entry(q, k, v, nullptr, nullptr, nullptr, o, lse, q_rows, kv_rows, o_rows, dtype, BH, S, D,
                        row_stride, head_stride, o_row_stride, o_head_stride, sm_scale, plan, plan_ws, stream);
}

int svgb_density(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH, int QC,
                 int KC, float* density, void* stream) {
  SVGB_REQUIRE(map && row_sz && col_sz && density && BH > 0 && QC > 0 && KC > 0, "bad arguments");
  density_kernel<<<BH, 256, 0, static_cast<cudaStream_t>(stream)>>>(map, row_sz, col_sz, QC, KC, density);
  SVGB_LAUNCH_OK();
  return 0;
}

int svgb_selftest_tile(const void* q, const void* k, const void* v, float* s_out, float* o_out,
                       int D, int dtype, float p_scale, void* stream) {
  SVGB_REQUIRE(q && k && v && s_out && o_out, "null pointer");
  SVGB_REQUIRE(D == 64 || D == 128, "head_dim %d unsupported", D);
  CUtensorMap qm, km, vm;
  if (encode_tmap_hsd(&qm, q, dtype, 1, 128, D, D, 128LL * D)) return -1;
  if (encode_tmap_hsd(&km, k, dtype, 1, 128, D, D, 128LL * D)) return -1;
  if (encode_tmap_hsd(&vm, v, dtype, 1, 128, D, D, 128LL * D)) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D == 128)
    return dtype == SVGB_BF16 ? launch_selftest<128, true>(qm, km, vm, s_out, o_out, p_scale, st)
                              : launch_selftest<128, false>(qm, km, vm, s_out, o_out, p_scale, st);
  return dtype == SVGB_BF16 ? launch_selftest<64, true>(qm, km, vm, s_out, o_out, p_scale, st)
                            : launch_selftest<64, false>(qm, km, vm, s_out, o_out, p_scale, st);
}

}  // extern "C"

#ifdef SVGB_ATTN_TRACE
extern "C" int svgb_debug_attn_trace(long long* host, int n) {
  SVGB_CUDA(cudaDeviceSynchronize());
  SVGB_CUDA(cudaMemcpyFromSymbol(host, svgb::g_attn_trace, sizeof(long long) * (n < 1536 ? n : 1536)));
  return 0;
}
#endif
