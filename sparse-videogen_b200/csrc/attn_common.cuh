// Shared definitions for the block-sparse attention path: the device-resident "plan" (work items +
// KV chunk lists) that every mask family (SVG2 variable blocks, SVG1 band masks, BSR, dense) is
// lowered to, and the element-mask predicates used on partial tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svgb {

// One work item = one CTA: up to 256 consecutive query rows of one head (two 128-row MMA tiles that
// share every K/V tile) and a list of KV chunks.
//   item  = {q_row0, nrows (1..256), chunk_off, nchunks}
//   chunk = {kv_start, meta}; meta bits [0,8) = valid columns - 1 (0..127), bit 8 = ELEM (evaluate
//           the element mask predicate on this chunk), bit 9 = reserved.
constexpr int kItemRows = 256;
constexpr int kTileRows = 128;
constexpr int kChunkCols = 128;
constexpr int kChunkElem = 1 << 8;

__host__ __device__ inline int chunk_meta(int valid, bool elem) {
  return ((valid - 1) & 0xff) | (elem ? kChunkElem : 0);
}
__host__ __device__ inline int chunk_valid(int meta) { return (meta & 0xff) + 1; }

// Element mask families: parameters m0,m1,m2.
enum MaskMode : int {
  MASK_NONE = 0,
  // HunyuanVideo (text last), reference svg/models/hyvideo/utils.py:20-44
  //   m0 = V = F*P, m1 = R = V + prompt_len, m2 = W
  MASK_HY = 1,
  // Wan / Cosmos (no text, first-frame sink), reference svg/models/wan/utils.py:25-41
  //   m0 = P (tokens per frame), m2 = W   : kv < P | |q-kv| <= W
  MASK_WAN = 2,
  // CogVideoX (text first), reference svg/models/cog/utils.py:30-46
  //   m0 = first-column limit (prompt_len or prompt_len + P), m1 = prompt_len, m2 = W
  MASK_COG = 3,
  // SVG1 profiling masks of get_attention_mask (hyvideo/utils.py:47-93, wan/utils.py:63-110),
  // evaluated analytically: 128-token-block band |bq - bk| < thres in frame-major (spatial) or
  // token-major (temporal) order.  m0 = F, m1 = P, m2 = thres (blocks).  HY: text rows / columns
  // (index >= F*P) always attend; WAN: first-frame sink painted before the token-major permutation.
  MASK_PROF_HY_S = 4,
  MASK_PROF_HY_T = 5,
  MASK_PROF_WAN_S = 6,
  MASK_PROF_WAN_T = 7,
  // CogVideoX (text FIRST) profiling masks, svg/models/cog/utils.py:61-88.  m2 = thres | ctx << 12.
  //   spatial : text rows / columns all-ones; the 128-token block band is painted in ABSOLUTE sequence
  //             coordinates (blocks 0 .. ceil(F*P/128)-1 from row / column 0, cog/utils.py:68-74)
  //   temporal: band in token-major order on the video x video part only; text rows and columns stay zero
  //             (cog/utils.py:76-86) -- a sampled text row has no allowed key at all
  MASK_PROF_COG_S = 8,
  MASK_PROF_COG_T = 9,
};

__host__ __device__ inline bool mask_allowed(int mode, int q, int kv, int m0, int m1, int m2) {
  // reference semantics, written branch-free (bitwise) so warps do not diverge on it
  int d = q - kv;
  d = d < 0 ? -d : d;
  switch (mode) {
    case MASK_HY: {
      const bool real = (q < m1) & (kv < m1);
      const bool fake = (q >= m1) & (kv >= m1);
      const bool vid = (d < m2) | (kv >= m0) | (q >= m0);
      return (real & vid) | fake;
    }
    case MASK_WAN:
      return (kv < m0) | (d <= m2);
    case MASK_COG:
      return (kv < m0) | (q < m1) | (d < m2);
    case MASK_PROF_COG_S: {
      const int thres = m2 & 0xfff, ctx = m2 >> 12;
      const int lim = ((m0 * m1 + 127) / 128) * 128;
      int bd = q / 128 - kv / 128;
      bd = bd < 0 ? -bd : bd;
      return (q < ctx) | (kv < ctx) | ((q < lim) & (kv < lim) & (bd < thres));
    }
    case MASK_PROF_COG_T: {
      const int thres = m2 & 0xfff, ctx = m2 >> 12;
      if (q < ctx || kv < ctx) return false;
      const int F = m0, P = m1;
      const int qv = q - ctx, kvv = kv - ctx;
      int bd = ((qv % P) * F + qv / P) / 128 - ((kvv % P) * F + kvv / P) / 128;
      bd = bd < 0 ? -bd : bd;
      return bd < thres;
    }
    case MASK_PROF_HY_S:
    case MASK_PROF_HY_T:
    case MASK_PROF_WAN_S:
    case MASK_PROF_WAN_T: {
      const int F = m0, P = m1, V = m0 * m1;
      const bool hy = mode <= MASK_PROF_HY_T;
      const bool temporal = (mode == MASK_PROF_HY_T) || (mode == MASK_PROF_WAN_T);
      if (q >= V || kv >= V) return hy;  // text rows / columns (HY); cannot occur for WAN (S == V)
      const int qi = temporal ? (q % P) * F + q / P : q;
      const int ki = temporal ? (kv % P) * F + kv / P : kv;
      int bd = qi / 128 - ki / 128;
      bd = bd < 0 ? -bd : bd;
      return (bd < m2) | (!hy & (ki < P));
    }
    default:
      return true;
  }
}

// Device-side evaluator of the same predicate for 32 consecutive key columns of one query row:
// row-constant terms are hoisted, the token-major index of the profiling masks is advanced
// incrementally (one division per 32 columns instead of two per element), and the result is a bitmask.
struct MaskRow {
  int mode, q, m0, m1, m2;
  int qi_blk;     // profiling masks: 128-token block of the (possibly token-major) query index
  bool q_text;    // profiling masks: query row is a text row
  __device__ __forceinline__ void init(int mode_, int q_, int m0_, int m1_, int m2_) {
    mode = mode_; q = q_; m0 = m0_; m1 = m1_; m2 = m2_;
    qi_blk = 0; q_text = false;
    if (mode >= MASK_PROF_HY_S && mode <= MASK_PROF_WAN_T) {
      const int F = m0, P = m1, V = F * P;
      const bool temporal = (mode == MASK_PROF_HY_T) | (mode == MASK_PROF_WAN_T);
      q_text = q >= V;
      const int qq = q_text ? 0 : q;
      qi_blk = (temporal ? (qq % P) * F + qq / P : qq) >> 7;
    }
  }
  // bit i set <=> (q, kv0 + i) allowed
  __device__ __forceinline__ uint32_t bits32(int kv0) const {
    uint32_t out = 0;
    if (mode >= MASK_PROF_HY_S && mode <= MASK_PROF_WAN_T) {
      const int F = m0, P = m1, V = F * P;
      const bool hy = mode <= MASK_PROF_HY_T;
      const bool temporal = (mode == MASK_PROF_HY_T) | (mode == MASK_PROF_WAN_T);
      const int kvc = kv0 < V ? kv0 : 0;
      int f = kvc / P, p = kvc - f * P;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kv = kv0 + i;
        const int ki = temporal ? p * F + f : kv;
        int bd = qi_blk - (ki >> 7);
        bd = bd < 0 ? -bd : bd;
        const bool text = q_text | (kv >= V);
        const bool a = text ? hy : ((bd < m2) | (!hy & (ki < P)));
        out |= (a ? 1u : 0u) << i;
        ++p;
        const bool wrap = p == P;
        p = wrap ? 0 : p;
        f += wrap ? 1 : 0;
      }
      return out;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) out |= (mask_allowed(mode, q, kv0 + i, m0, m1, m2) ? 1u : 0u) << i;
    return out;
  }
};

struct AttnArgs {
  const int4* items;
  // variable-block plans: optional second stream per work item.  items2[i].y > 0 makes item i a DUAL item: two
  // single-tile streams (<= 128 rows each, usually the tails of two different q-blocks) share one CTA -- T0 runs
  // items[i], T1 runs items2[i], each with its own chunk list and its own K/V tiles through the common ring.
  const int4* items2;
  // variable-block plans: transposed-tail items (attn_tail.cuh): tails of <= 64 rows, two per CTA.  Same int4 format
  // as items / items2, so they can also be run by attn_fwd_kernel as dual items when the tail kernel does not apply
  // (head_dim 64, fp8, LSE / fp32 output).
  const int4* titems;
  const int4* titems2;
  const int* tcount;
  const int* item_count;
  const int2* chunks;
  int items_stride;   // 0: one plan shared by all heads (band masks); else items per head
  int counts_stride;  // 0 or 1
  void* o;
  long long o_row_stride;   // elements
  long long o_head_stride;  // elements
  const int* o_rows;        // optional [BH, S]: output row for query row q (fused inverse permutation)
  float* lse;               // optional [BH, S] (natural log), indexed like o rows
  float scale_log2;         // softmax scale * log2(e)
  int S;                    // query rows per head (indexing of o_rows / lse / q_index)
  int mask_mode, m0, m1, m2;
  const int* q_index;       // optional [S]: query position used by the element mask (sampled rows)
  int out_f32;              // 1: o is fp32 (used for split-KV partials)
  int softmax_shared;       // 1: both softmax warpgroups share one tile at a time (plans with narrow chunks)
  int sub_mode;             // opt-in sub-chunk pipeline (SVGB_ATTN_SUB=1): 0 off, 1 on
  int arrival_order;        // two-tile items: serve the tiles' MMAs in P-arrival order (SVGB_ATTN_ORDER=1)
  // FP8 (e4m3) inputs: per-head dequantisation scales [BH]; s_q * s_k multiplies the logits, s_v the
  // output.  NULL for 16-bit inputs.
  const float* q_scale;
  const float* k_scale;
  const float* v_scale;
  // ---- gather mode (SVG2): `chunks` holds RUNS {src_start, cum_before} (+ a sentinel {0, total}) of the
  // selected key ranges; K/V rows are gathered with cp.async into exactly-full 128-token chunks, optionally
  // through row-index vectors so the cluster permutation of Q, K, V never materialises in HBM.
  int gather;
  const int* item_total;    // [BH * items_stride | items]: selected keys per item
  const int* q_rows;        // optional [BH, S]: source row of (permuted) query row
  const int* kv_rows;       // optional [BH, S]: source row of (permuted) key/value row
  const void* q_ptr;        // raw tensors for the gather path ([BH,S,D] with the strides below)
  const void* k_ptr;
  const void* v_ptr;
  long long in_row_stride, in_head_stride;  // elements
};

}  // namespace svgb
