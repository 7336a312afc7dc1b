// flash-k-means for sm_100a (reference: svg/kmeans_utils.py:258-322, 375-421, 464-733).
//
//   assign : labels[n] = argmin_k max(0, |x_n|^2 + |c_k|^2 - 2 x_n.c_k).  X.C^T on tcgen05, persistent (one CTA per
//            SM; work item = 256 points = two 128-row tiles sharing every centroid tile, double-buffered TMEM
//            accumulators and X buffers), fused running-argmin epilogue (one thread per point).  Tensor-bound.
//   update : deterministic segmented mean.  The reference sorts labels and uses fp32 atomics
//            (non-deterministic order); we reuse the stable counting sort (layout_ops.cu) and sum each
//            cluster's members in index order.  HBM-bound (x read once).
//   run    : the whole Lloyd loop on the device, no host sync: a `done` flag turns later iterations
//            into no-ops and a device-side buffer index replaces the reference's Python swap.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/svgb200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace svgb {

// ---------------------------------------------------------------------------------------------
// squared row norms
//   round_result=1: (x**2).sum(-1) in the 16-bit dtype (squares rounded, fp32 sum, result rounded)
//                   -- batch_kmeans_Euclid:704
//   round_result=0: fp32 sum of 16-bit-rounded squares -- _euclid_assign_kernel:531
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ float to_f32(uint16_t h) {
  if constexpr (BF16) return __uint_as_float(static_cast<uint32_t>(h) << 16);
  else return __half2float(__ushort_as_half(h));
}
template <bool BF16>
__device__ __forceinline__ float round16(float f) {
  if constexpr (BF16) return __bfloat162float(__float2bfloat16_rn(f));
  else return __half2float(__float2half_rn(f));
}

template <bool BF16>
__global__ void row_sqnorm_kernel(const uint16_t* __restrict__ x, float* __restrict__ out, long long rows,
                                  int D, int round_result) {
  // one warp per row; lanes stride the row, fixed-order tree reduction (deterministic)
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint16_t* p = x + row * D;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = to_f32<BF16>(p[d]);
    acc += round16<BF16>(v * v);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = round_result ? round16<BF16>(acc) : acc;
}

// ---------------------------------------------------------------------------------------------
// assign
// ---------------------------------------------------------------------------------------------
struct KmState {  // device-resident loop state
  int done;
  int cur;     // which centroid buffer the next assignment reads
  int n_iter;
  float shift;
};

// experiment switches (tools/kmeans_ab.sh builds variants; the defaults are the product configuration)
#ifndef SVGB_KM_STAGES
#define SVGB_KM_STAGES 3
#endif
#ifndef SVGB_KM_PREFETCH
#define SVGB_KM_PREFETCH 1
#endif
#ifndef SVGB_KM_EPI
#define SVGB_KM_EPI 0  // 1: accumulators are loaded but not reduced; 2: not even loaded (timing decomposition only)
#endif

template <int D>
struct AssignCfg {
  static constexpr int kHalves = D / 64;
  static constexpr int kPanelBytes = 128 * 128;
  static constexpr int kTileBytes = kPanelBytes * kHalves;   // 128 rows x D
  static constexpr int kStages = SVGB_KM_STAGES;              // centroid ring (128 centroids per stage)
  static constexpr int kXBytes = 2 * kTileBytes;             // one work item = 256 points
  static constexpr int kRingBytes = kStages * kTileBytes;
  static constexpr int kBarBytes = 256;
  // 2 X buffers (the next item's points land while this item's MMAs run) + ring + barriers + alignment slack
  static constexpr int kSmemBytes = 1024 + 2 * kXBytes + kRingBytes + kBarBytes;
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
};

struct AssignBars {
  uint64_t x_full[2], x_empty[2];
  uint64_t c_full[4], c_empty[4];
  uint64_t s_full[2][2], s_empty[2][2];
  uint32_t tmem_base;
};
static_assert(sizeof(AssignBars) <= 256, "barrier block");

// (bits(d) & mask) | idx in ONE LOP3 (mask in a register, idx an immediate; nvcc emits two for the C expression)
__device__ __forceinline__ uint32_t pack_idx(float d, int idx, uint32_t mask) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEC;" : "=r"(r) : "r"(__float_as_uint(d)), "r"(idx), "r"(mask));
  return r;
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}

// |x_n|^2 of one point from the swizzled X tile in shared memory, with the reference's rounding points
// ((x**2).sum(-1) in the 16-bit dtype, batch_kmeans_Euclid:704): 16-bit products, fp32 sum, 16-bit result.
template <int D, bool BF16>
__device__ __forceinline__ float row_sqnorm_smem(const uint8_t* tile, int r) {
  float acc = 0.f;
#pragma unroll
  for (int h = 0; h < D / 64; ++h) {
    const uint8_t* row = tile + h * AssignCfg<D>::kPanelBytes + r * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // 128-byte swizzle: logical 16-byte chunk c of row r sits at chunk position c ^ (r & 7)
      const uint4 v = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (BF16) {
          const __nv_bfloat162 x2 = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
          const float2 f = __bfloat1622float2(__hmul2(x2, x2));
          acc += f.x;
          acc += f.y;
        } else {
          const __half2 x2 = *reinterpret_cast<const __half2*>(&w[i]);
          const float2 f = __half22float2(__hmul2(x2, x2));
          acc += f.x;
          acc += f.y;
        }
      }
    }
  }
  return round16<BF16>(acc);
}

// Persistent: one CTA per SM walks the (head, 256-point block) work items; warp 0 = TMA producer, warp 1 = MMA
// issuer, warps 4-11 = two epilogue warpgroups (one per 128-point tile).  Nothing is torn down between items: the
// centroid ring, the TMEM double buffers and their barrier phases run on one global chunk counter, and the next
// item's points are loaded into the second X buffer while the current item's MMAs run.
//
// Epilogue arithmetic: d_k = fl(fl(|x|^2 + |c_k|^2) - 2 x.c_k) as the reference's kernel forms it (:540-545), one packed
// FADD2 + FFMA2 per two centroids; the running argmin is branch-free (in-group index in the low 5 mantissa bits, one
// FMNMX3 per two centroids, see the epilogue).  Distances that differ by less than 32 ulp (4e-6 relative) resolve to the
// lower index -- far inside the noise of the reference's own 16-bit |c|^2.
// x_sq == nullptr: |x|^2 is computed from the X tile in shared memory (the Lloyd loop does this: no separate pass
// over x); otherwise the caller's values are used (svgb_kmeans_assign, where the reference passes its own x_sq).
template <int D, bool BF16>
__global__ void __launch_bounds__(384, 1)
kmeans_assign_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap cmap0,
                     const __grid_constant__ CUtensorMap cmap1, const float* __restrict__ x_sq,
                     const float* __restrict__ c_sq, int* __restrict__ labels, int N, int K, int BH,
                     const KmState* __restrict__ state) {
  using Cfg = AssignCfg<D>;
  int cur = 0;
  if (state) {
    if (state->done) return;
    cur = state->cur;
  }
  const CUtensorMap* cmap = cur == 0 ? &cmap0 : &cmap1;
  const int nchunks = (K + 127) / 128;
  const int blocks_per_head = (N + 255) / 256;
  const int total = BH * blocks_per_head;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sX = smem_base, sRing = smem_base + 2 * Cfg::kXBytes;
  AssignBars* bars = reinterpret_cast<AssignBars*>(smem_al + 2 * Cfg::kXBytes + Cfg::kRingBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&xmap);
    tma_prefetch_desc(cmap);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->x_full[b]), 1);
      mbar_init(smem_u32(&bars->x_empty[b]), 1 + 8);  // MMA commit + the 8 epilogue warps (they read |x|^2 from it)
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(smem_u32(&bars->c_full[s]), 1);
      mbar_init(smem_u32(&bars->c_empty[s]), 1);
    }
    for (int t = 0; t < 2; ++t)
      for (int b = 0; b < 2; ++b) {
        mbar_init(smem_u32(&bars->s_full[t][b]), 1);
        mbar_init(smem_u32(&bars->s_empty[t][b]), 128);
      }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(&bars->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      auto load_x = [&](int i, int item) {
        const int xb = i & 1, bh = item / blocks_per_head, row0 = (item % blocks_per_head) * 256;
        mbar_wait(smem_u32(&bars->x_empty[xb]), ((i >> 1) & 1) ^ 1, 30);
        const uint32_t fb = smem_u32(&bars->x_full[xb]);
        mbar_expect_tx(fb, Cfg::kXBytes);
        for (int t = 0; t < 2; ++t) {
          // a second tile that would start past the end is loaded from the last row instead (its labels are not
          // stored); partially out-of-range tiles are zero-filled by TMA
          const int r = min(row0 + t * 128, N - 1);
          for (int h = 0; h < Cfg::kHalves; ++h)
            tma_load_3d(sX + xb * Cfg::kXBytes + t * Cfg::kTileBytes + h * Cfg::kPanelBytes, &xmap, fb, h * 64, r, bh);
        }
      };
      int g = 0, i = 0;
      if (static_cast<int>(blockIdx.x) < total) load_x(0, blockIdx.x);
      // the next item's points are requested once the ring has wrapped inside this item: by then the MMAs of the
      // previous item (which release that X buffer) are known to have completed, so that wait cannot hold up this
      // item's centroid loads
      const int x_at = min(static_cast<int>(Cfg::kStages), nchunks);
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++i) {
        const int bh = item / blocks_per_head;
        const int next = item + gridDim.x;
        for (int j = 0; j <= nchunks; ++j) {
          const int slot = g % Cfg::kStages;
          if (j < nchunks) mbar_wait(smem_u32(&bars->c_empty[slot]), ((g / Cfg::kStages) & 1) ^ 1, 31);
          if (j == x_at && next < total) load_x(i + 1, next);
          if (j < nchunks) {
            const uint32_t fb = smem_u32(&bars->c_full[slot]);
            mbar_expect_tx(fb, Cfg::kTileBytes);
            for (int h = 0; h < Cfg::kHalves; ++h)
              tma_load_3d(sRing + slot * Cfg::kTileBytes + h * Cfg::kPanelBytes, cmap, fb, h * 64, j * 128, bh);
            ++g;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync: ptxas emits the tcgen05 stream without per-instruction election loops
      const uint32_t idesc = make_idesc(128, 128, BF16, false, false);
      // descriptors = constant high word + running low word (address field in units of 16 B), see attn_kernel.cuh
      const uint64_t d0 = desc_kmajor_sw128(sX);
      const uint32_t hi = static_cast<uint32_t>(d0 >> 32), x_lo0 = static_cast<uint32_t>(d0);
      const uint32_t c_lo0 = static_cast<uint32_t>(desc_kmajor_sw128(sRing));
      int g = 0, i = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++i) {
        const int xb = i & 1;
        mbar_wait(smem_u32(&bars->x_full[xb]), (i >> 1) & 1, 32);
        for (int j = 0; j < nchunks; ++j, ++g) {
          const int slot = g % Cfg::kStages, buf = g & 1;
          const uint32_t c_lo = c_lo0 + slot * (Cfg::kTileBytes >> 4);
          mbar_wait(smem_u32(&bars->c_full[slot]), (g / Cfg::kStages) & 1, 33);
          for (int t = 0; t < 2; ++t) {
            mbar_wait(smem_u32(&bars->s_empty[t][buf]), ((g >> 1) & 1) ^ 1, 34);
            tc_fence_after();
            const uint32_t d_tmem = tmem + t * 256 + buf * 128;
            uint32_t a_lo = x_lo0 + ((xb * Cfg::kXBytes + t * Cfg::kTileBytes) >> 4), b_lo = c_lo;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
              mma_ss(d_tmem, (static_cast<uint64_t>(hi) << 32) | a_lo, (static_cast<uint64_t>(hi) << 32) | b_lo, idesc, kk > 0);
              asm volatile("" : "+r"(a_lo), "+r"(b_lo));
              const uint32_t step = ((kk & 3) == 3) ? ((Cfg::kPanelBytes - 3 * 32) >> 4) : (32 >> 4);
              a_lo += step;
              b_lo += step;
            }
            tc_commit(smem_u32(&bars->s_full[t][buf]));
          }
          tc_commit(smem_u32(&bars->c_empty[slot]));
        }
        tc_commit(smem_u32(&bars->x_empty[xb]));
      }
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2, wq = warp & 3;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(wq * 32) << 16) + t * 256;
    const uint64_t neg2 = pack_f32x2(-2.f, -2.f);
    uint32_t idx_mask;  // ~31, kept opaque so that it lives in a register (two immediates would cost two LOP3 per value)
    asm volatile("mov.u32 %0, 0xFFFFFFE0;" : "=r"(idx_mask));
    // Timing decomposition on a B200 (tools/kmeans_ab.sh, K = 1000): with the accumulators loaded but not reduced the
    // kernel runs at 1555 TFLOP/s (MMA + TMA + tcgen05.ld are not the limit; a 2-stage ring still gives 1475), the
    // reduction arithmetic of the epilogue warps is what the rest of the time goes to.  SVGB_KM_PREFETCH keeps one
    // tcgen05.ld in flight across chunk and item boundaries (the first group of the next chunk is requested before the
    // last group of the current one is reduced).
    int g = 0, i = 0;
    bool primed = false;
    uint32_t ra[32], rb[32];
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++i) {
      const int xb = i & 1, bh = item / blocks_per_head, row0 = (item % blocks_per_head) * 256;
      const int r = wq * 32 + lane;
      const int n = row0 + t * 128 + r;
      if (!primed) {  // first item of this CTA
        mbar_wait(smem_u32(&bars->s_full[t][g & 1]), (g >> 1) & 1, 36);
        tc_fence_after();
#if SVGB_KM_EPI != 2
        tmem_ld32(lane_addr + (g & 1) * 128, ra);
#endif
      }
      mbar_wait(smem_u32(&bars->x_full[xb]), (i >> 1) & 1, 35);  // complete: this item's first MMA has been committed
      float xs;
      if (x_sq) xs = (n < N) ? x_sq[static_cast<size_t>(bh) * N + n] : 0.f;
      else xs = row_sqnorm_smem<D, BF16>(smem_al + xb * Cfg::kXBytes + t * Cfg::kTileBytes, r);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->x_empty[xb]));
      const float4* cs_head = reinterpret_cast<const float4*>(c_sq + static_cast<size_t>(bh) * nchunks * 128);
      const uint64_t xs2 = pack_f32x2(xs, xs);
      float best = 3.4e38f;
      int best_k = 0;
      const bool next_item = item + static_cast<int>(gridDim.x) < total;
      for (int j = 0; j < nchunks; ++j, ++g) {
        const int buf = g & 1;
        const float4* cs = cs_head + j * 32;
        const int kbase = j * 128;
        auto group = [&](const uint32_t (&rr)[32], int gi) {
          // d_k = (|x|^2 + |c_k|^2) - 2 x.c_k with the reference's two roundings (the sum, then the fma); |c|^2 comes
          // through L1 (warp-uniform 16-byte loads, padded with +3e38 past K so that the columns TMA zero-filled can
          // never win).  The argmin inside the group is branch-free: the in-group index replaces the low 5 mantissa
          // bits of every (non-negative) distance, so one FMNMX3 tree returns value and index together and equal
          // distances resolve to the lower index.  (A compare-and-search on improvement was tried first: a WARP takes
          // that branch whenever any of its 32 rows improves, i.e. in almost every group, and it doubled the kernel time.)
          uint32_t v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 c4 = __ldg(cs + gi * 8 + q);
            const uint64_t d01 = ffma2(pack_f32x2(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1])), neg2,
                                       fadd2(pack_f32x2(c4.x, c4.y), xs2));
            const uint64_t d23 = ffma2(pack_f32x2(__uint_as_float(rr[4 * q + 2]), __uint_as_float(rr[4 * q + 3])), neg2,
                                       fadd2(pack_f32x2(c4.z, c4.w), xs2));
            float a0, a1, a2, a3;
            unpack_f32x2(d01, a0, a1);
            unpack_f32x2(d23, a2, a3);
            v[4 * q] = pack_idx(a0, 4 * q, idx_mask);
            v[4 * q + 1] = pack_idx(a1, 4 * q + 1, idx_mask);
            v[4 * q + 2] = pack_idx(a2, 4 * q + 2, idx_mask);
            v[4 * q + 3] = pack_idx(a3, 4 * q + 3, idx_mask);
          }
          float gmin = fminf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
          for (int q = 1; q < 16; ++q) gmin = fmin3(gmin, __uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
          uint32_t gbits = __float_as_uint(gmin);
          if (__any_sync(0xffffffffu, gmin < 0.f)) {  // warp-uniform vote: a real branch, not predicated code
            // a distance that rounded below zero (a point sitting on a centroid): the reference clamps to 0 and takes
            // the FIRST such centroid; negative floats order the packed index the wrong way round, so search (rare)
            if (gmin < 0.f) {
              int first = 31;
#pragma unroll
              for (int q = 31; q >= 0; --q)
                if (__uint_as_float(v[q] & ~31u) <= 0.f) first = q;
              gbits = static_cast<uint32_t>(first);  // value 0, index `first`
            }
          }
          const float gval = __uint_as_float(gbits & ~31u);
          if (gval < best) {  // strict: an earlier group keeps ties (selects, no branch)
            best = gval;
            best_k = kbase + gi * 32 + static_cast<int>(gbits & 31u);
          }
        };
#if SVGB_KM_EPI == 2
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->s_empty[t][buf]));
        primed = (j + 1 < nchunks) || next_item;
        if (primed) {
          const int nb = (g + 1) & 1;
          mbar_wait(smem_u32(&bars->s_full[t][nb]), ((g + 1) >> 1) & 1, 37);
          tc_fence_after();
        }
#else
#if SVGB_KM_EPI == 1
#define SVGB_KM_GROUP(arr, gi) best += __uint_as_float(arr[gi])
#else
#define SVGB_KM_GROUP(arr, gi) group(arr, gi)
#endif
#if !SVGB_KM_PREFETCH
        if (j > 0) {
          mbar_wait(smem_u32(&bars->s_full[t][buf]), (g >> 1) & 1, 37);
          tc_fence_after();
          tmem_ld32(lane_addr + buf * 128, ra);
        }
#endif
        tc_wait_ld();
        tmem_ld32(lane_addr + buf * 128 + 32, rb);
        SVGB_KM_GROUP(ra, 0);
        tc_wait_ld();
        tmem_ld32(lane_addr + buf * 128 + 64, ra);
        SVGB_KM_GROUP(rb, 1);
        tc_wait_ld();
        tmem_ld32(lane_addr + buf * 128 + 96, rb);
        SVGB_KM_GROUP(ra, 2);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->s_empty[t][buf]));  // the accumulator is in registers: hand the buffer back early
#if SVGB_KM_PREFETCH
        primed = (j + 1 < nchunks) || next_item;
        if (primed) {
          const int nb = (g + 1) & 1;
          mbar_wait(smem_u32(&bars->s_full[t][nb]), ((g + 1) >> 1) & 1, 37);
          tc_fence_after();
          tmem_ld32(lane_addr + nb * 128, ra);
        }
#endif
        SVGB_KM_GROUP(rb, 3);
#endif
      }
      if (n < N) labels[static_cast<size_t>(bh) * N + n] = best_k;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------
// update: one CTA per (head, cluster); members = perm[off .. off+cnt) (ascending token index).
// 256 threads = 16 row-groups x (D/8) lanes, each lane owns 8 consecutive dims (one 16-byte load per
// member row); group g sums members g, g+16, ... in order, then the 16 partials are combined in a fixed
// order -> run-to-run deterministic (the reference's fp32 atomics are not).
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256)
kmeans_update_kernel(const uint16_t* __restrict__ x, const int* __restrict__ perm,
                     const int* __restrict__ counts, const int* __restrict__ chunk0_base,
                     const uint16_t* __restrict__ c_old0, const uint16_t* __restrict__ c_old1,
                     uint16_t* __restrict__ c_new0, uint16_t* __restrict__ c_new1, float* __restrict__ shift_max,
                     int N, int K, int D, long long x_head_stride, const KmState* __restrict__ state) {
  int cur = 0;
  if (state) {
    if (state->done) return;
    cur = state->cur;
  }
  const uint16_t* c_old = cur == 0 ? c_old0 : c_old1;
  uint16_t* c_new = cur == 0 ? c_new0 : c_new1;
  const int k = blockIdx.x, bh = blockIdx.y;
  const int cnt = counts[static_cast<size_t>(bh) * K + k];
  const int off = chunk0_base[static_cast<size_t>(bh) * K + k];  // exclusive prefix of counts
  const int lanes = D / 8;                                        // 16-byte vectors per row
  const int groups = blockDim.x / lanes;
  const int grp = threadIdx.x / lanes, w = threadIdx.x % lanes;
  __shared__ float part[32][129];
  __shared__ float s_norm[8];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add_row = [&](const uint4& v) {
    const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] += to_f32<BF16>(static_cast<uint16_t>(ws[i] & 0xffff));
      acc[2 * i + 1] += to_f32<BF16>(static_cast<uint16_t>(ws[i] >> 16));
    }
  };
  if (grp < groups) {
    const int* pp = perm + static_cast<size_t>(bh) * N + off;
    const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<size_t>(bh) * x_head_stride);
    // 8 member rows in flight per thread, fully predicated (clusters are short -- ~7 rows per group at K=1000 -- so a
    // remainder loop of single dependent index->row loads used to dominate the CTA's lifetime); rows are still added
    // in ascending member order, so the result does not depend on the batching
    for (int i = grp; i < cnt; i += 8 * groups) {
      int idx[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) idx[u] = (i + u * groups < cnt) ? __ldg(pp + i + u * groups) : -1;
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = idx[u] >= 0 ? __ldg(xb + static_cast<size_t>(idx[u]) * lanes + w) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (idx[u] >= 0) add_row(v[u]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) part[grp][w * 8 + e] = acc[e];
  }
  __syncthreads();
  // D <= 128 threads finish the mean; the squared shift is reduced by warp shuffles (fixed order -> deterministic)
  float sq = 0.f;
  if (static_cast<int>(threadIdx.x) < D) {
    const int d = threadIdx.x;
    float s = 0.f;
    for (int gI = 0; gI < groups; ++gI) s += part[gI][d];
    const float o = to_f32<BF16>(c_old[(static_cast<size_t>(bh) * K + k) * D + d]);
    const float n = cnt > 0 ? round16<BF16>(s / static_cast<float>(cnt)) : o;
    uint16_t bits;
    if constexpr (BF16) bits = static_cast<uint16_t>(__float_as_uint(n) >> 16);
    else bits = __half_as_ushort(__float2half_rn(n));
    c_new[(static_cast<size_t>(bh) * K + k) * D + d] = bits;
    // shift = |round16(new - old)|_2, then rounded to 16 bit like the reference's bf16 tensor ops
    const float dd = round16<BF16>(n - o);
    sq = dd * dd;
  }
  if (shift_max) {
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_norm[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (D + 31) / 32; ++i) t += s_norm[i];
      const float nrm = round16<BF16>(sqrtf(t));
      atomicMax(reinterpret_cast<int*>(shift_max), __float_as_int(nrm));  // non-negative floats order as ints
    }
  }
}

// |c|^2 of the centroid buffer selected by the loop state, written in the padded [BH, Kpad] layout
// the assign kernel bulk-copies.  The reference computes it as `tl.sum(c_tile * c_tile, axis=0)` on 16-bit tiles
// (_euclid_assign_kernel, svg/kmeans_utils.py:531): products AND the reduction live in the 16-bit type, in a
// layout-dependent order (labels recorded from the reference's Triton kernel on a B200 are explained to the last
// point by per-centroid offsets of a few bf16 ulps of |c|^2 -- tests/golden/kmeans_golden.npz).  We take the
// correctly rounded value of that quantity: fp32 sum of the 16-bit-rounded squares, rounded once to the 16-bit type.
template <bool BF16>
__global__ void csq_kernel(const uint16_t* __restrict__ c0, const uint16_t* __restrict__ c1,
                           float* __restrict__ out, int K, int Kpad, int D, const KmState* __restrict__ state) {
  int cur = 0;
  if (state) {
    if (state->done) return;
    cur = state->cur;
  }
  const uint16_t* c = cur == 0 ? c0 : c1;
  const int bh = blockIdx.y;
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= Kpad) return;
  float acc = 0.f;
  if (k < K) {
    const uint16_t* p = c + (static_cast<size_t>(bh) * K + k) * D;
    for (int d = lane; d < D; d += 32) {
      const float v = to_f32<BF16>(p[d]);
      acc += round16<BF16>(v * v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  // columns past K (zero-filled by TMA in the assign kernel) get a norm no point can prefer
  if (lane == 0) out[static_cast<size_t>(bh) * Kpad + k] = k < K ? round16<BF16>(acc) : 3.0e38f;
}

__global__ void km_commit_kernel(KmState* st, float* shift_max, float tol, int it) {
  if (st->done) return;
  const float sh = *shift_max;
  st->shift = sh;
  st->n_iter = it + 1;
  if (sh < tol) st->done = 1;  // break BEFORE `centroids = centroids_new` (kmeans_utils.py:723-726)
  else st->cur ^= 1;
  *shift_max = 0.f;
}

__global__ void km_init_state_kernel(KmState* st, float* shift_max) {
  st->done = 0;
  st->cur = 0;
  st->n_iter = 0;
  st->shift = 0.f;
  *shift_max = 0.f;
}

// copy the final centroids (buffer `cur`) and n_iter out
__global__ void km_finalize_kernel(const KmState* st, const uint4* __restrict__ c0, const uint4* __restrict__ c1,
                                   uint4* __restrict__ out, long long n_vec, int* __restrict__ n_iter_out) {
  const uint4* src = st->cur == 0 ? c0 : c1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = src[i];
  if (n_iter_out && blockIdx.x == 0 && threadIdx.x == 0) *n_iter_out = st->n_iter;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int D, bool BF16>
static int launch_assign(const CUtensorMap& xm, const CUtensorMap& cm0, const CUtensorMap& cm1,
                         const float* x_sq, const float* c_sq_padded, int* labels, int BH, int N, int K,
                         const KmState* st, cudaStream_t stream) {
  using Cfg = AssignCfg<D>;
  auto kern = kmeans_assign_kernel<D, BF16>;
  SVGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int items = BH * ((N + 255) / 256);
  const int grid = items < sm_count() ? items : sm_count();  // persistent: one CTA per SM
  kern<<<grid, 384, Cfg::kSmemBytes, stream>>>(xm, cm0, cm1, x_sq, c_sq_padded, labels, N, K, BH, st);
  SVGB_LAUNCH_OK();
  return 0;
}

static int dispatch_assign(const void* x, long long x_head_stride, const void* c0, const void* c1, const float* x_sq,
                           const float* c_sq_padded, int* labels, int BH, int N, int K, int D, int dtype,
                           const KmState* st, cudaStream_t stream) {
  CUtensorMap xm, cm0, cm1;
  if (encode_tmap_hsd(&xm, x, dtype, BH, N, D, D, x_head_stride)) return -1;
  if (encode_tmap_hsd(&cm0, c0, dtype, BH, K, D, D, static_cast<long long>(K) * D)) return -1;
  if (encode_tmap_hsd(&cm1, c1, dtype, BH, K, D, D, static_cast<long long>(K) * D)) return -1;
  if (D == 128)
    return dtype == SVGB_BF16 ? launch_assign<128, true>(xm, cm0, cm1, x_sq, c_sq_padded, labels, BH, N, K, st, stream)
                              : launch_assign<128, false>(xm, cm0, cm1, x_sq, c_sq_padded, labels, BH, N, K, st, stream);
  return dtype == SVGB_BF16 ? launch_assign<64, true>(xm, cm0, cm1, x_sq, c_sq_padded, labels, BH, N, K, st, stream)
                            : launch_assign<64, false>(xm, cm0, cm1, x_sq, c_sq_padded, labels, BH, N, K, st, stream);
}

static int launch_sqnorm(const void* x, float* out, long long rows, int D, int dtype, int round_result,
                         cudaStream_t stream) {
  const int block = 256;
  const long long blocks = (rows * 32 + block - 1) / block;
  if (dtype == SVGB_BF16)
    row_sqnorm_kernel<true><<<static_cast<unsigned>(blocks), block, 0, stream>>>(
        static_cast<const uint16_t*>(x), out, rows, D, round_result);
  else
    row_sqnorm_kernel<false><<<static_cast<unsigned>(blocks), block, 0, stream>>>(
        static_cast<const uint16_t*>(x), out, rows, D, round_result);
  SVGB_LAUNCH_OK();
  return 0;
}

static int launch_csq(const void* c0, const void* c1, float* out, int BH, int K, int D, int dtype,
                      const KmState* st, cudaStream_t stream) {
  const int Kpad = (K + 127) / 128 * 128;
  dim3 grid((Kpad * 32 + 255) / 256, BH);
  if (dtype == SVGB_BF16)
    csq_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(c0), static_cast<const uint16_t*>(c1),
                                               out, K, Kpad, D, st);
  else
    csq_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(c0), static_cast<const uint16_t*>(c1),
                                                out, K, Kpad, D, st);
  SVGB_LAUNCH_OK();
  return 0;
}

static int launch_update(const void* x, long long x_head_stride, const int* perm, const int* counts, const int* offs, const void* c_old0,
                         const void* c_old1, void* c_new0, void* c_new1, float* shift_max, int BH, int N, int K,
                         int D, int dtype, const KmState* st, cudaStream_t stream) {
  dim3 grid(K, BH);
  const int threads = 256;  // 16 (D=128) or 32 (D=64) row-groups of D/8 lanes
  auto X = static_cast<const uint16_t*>(x);
  auto O0 = static_cast<const uint16_t*>(c_old0);
  auto O1 = static_cast<const uint16_t*>(c_old1);
  auto N0 = static_cast<uint16_t*>(c_new0);
  auto N1 = static_cast<uint16_t*>(c_new1);
  if (dtype == SVGB_BF16)
    kmeans_update_kernel<true><<<grid, threads, 0, stream>>>(X, perm, counts, offs, O0, O1, N0, N1, shift_max, N, K, D, x_head_stride, st);
  else
    kmeans_update_kernel<false><<<grid, threads, 0, stream>>>(X, perm, counts, offs, O0, O1, N0, N1, shift_max, N, K, D, x_head_stride, st);
  SVGB_LAUNCH_OK();
  return 0;
}

struct KmLayout {
  size_t csq_pad, perm, offs, sort, sort_bytes, cbuf0, cbuf1, state, total;
};
static KmLayout km_layout(int BH, int N, int K, int D) {
  KmLayout L;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t at = o;
    o += align_up(bytes, 256);
    return at;
  };
  const int Kpad = (K + 127) / 128 * 128;
  L.csq_pad = take(sizeof(float) * BH * Kpad);
  L.perm = take(sizeof(int) * static_cast<size_t>(BH) * N);
  L.offs = take(sizeof(int) * BH * K);
  L.sort_bytes = 0;
  svgb_argsort_labels_bytes(BH, N, K, &L.sort_bytes);
  L.sort = take(L.sort_bytes);
  L.cbuf0 = take(2ull * BH * K * D);
  L.cbuf1 = take(2ull * BH * K * D);
  L.state = take(256);
  L.total = o;
  return L;
}

}  // namespace svgb

using namespace svgb;

extern "C" {

int svgb_row_sqnorm(const void* x, float* x_sq, int BH, int N, int D, int dtype, int round_result,
                    void* stream) {
  SVGB_REQUIRE(x && x_sq && BH > 0 && N > 0 && D > 0, "bad arguments");
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16, "dtype %d unsupported", dtype);
  return launch_sqnorm(x, x_sq, static_cast<long long>(BH) * N, D, dtype, round_result,
                       static_cast<cudaStream_t>(stream));
}

int svgb_kmeans_bytes(int BH, int N, int K, int D, size_t* bytes) {
  SVGB_REQUIRE(BH > 0 && N > 0 && K > 0 && K <= 4096 && (D == 64 || D == 128) && bytes,
               "bad arguments (need K <= 4096, D in {64,128})");
  *bytes = km_layout(BH, N, K, D).total;
  return 0;
}

int svgb_kmeans_assign(const void* x, const void* c, const float* x_sq, int32_t* labels, int BH, int N,
                       int K, int D, int dtype, void* ws, size_t ws_bytes, void* stream) {
  SVGB_REQUIRE(x && c && x_sq && labels && ws, "null pointer");
  size_t need = 0;
  if (svgb_kmeans_bytes(BH, N, K, D, &need)) return -1;
  SVGB_REQUIRE(ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  const KmLayout L = km_layout(BH, N, K, D);
  char* w = static_cast<char*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* csqp = reinterpret_cast<float*>(w + L.csq_pad);
  if (launch_csq(c, c, csqp, BH, K, D, dtype, nullptr, st)) return -1;
  return dispatch_assign(x, static_cast<long long>(N) * D, c, c, x_sq, csqp, labels, BH, N, K, D, dtype, nullptr, st);
}

int svgb_kmeans_update(const void* x, const int32_t* labels, const void* c_old, void* c_new,
                       int32_t* counts, float* shift_max, int BH, int N, int K, int D, int dtype,
                       void* ws, size_t ws_bytes, void* stream) {
  SVGB_REQUIRE(x && labels && c_old && c_new && counts && ws, "null pointer");
  size_t need = 0;
  if (svgb_kmeans_bytes(BH, N, K, D, &need)) return -1;
  SVGB_REQUIRE(ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  const KmLayout L = km_layout(BH, N, K, D);
  char* w = static_cast<char*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* perm = reinterpret_cast<int*>(w + L.perm);
  int* offs = reinterpret_cast<int*>(w + L.offs);
  if (argsort_labels_impl(labels, BH, N, K, perm, counts, offs, w + L.sort, nullptr, st)) return -1;
  if (shift_max) SVGB_CUDA(cudaMemsetAsync(shift_max, 0, sizeof(float), st));
  return launch_update(x, static_cast<long long>(N) * D, perm, counts, offs, c_old, c_old, c_new, c_new, shift_max, BH, N, K, D, dtype, nullptr, st);
}

int svgb_kmeans_run(const void* x, const void* init_centroids, int BH, int N, int K, int D, int dtype,
                    int max_iters, float tol, int32_t* labels, void* centroids_out, int32_t* counts,
                    int32_t* n_iter_out, void* ws, size_t ws_bytes, void* stream) {
  return svgb_kmeans_run_sorted(x, 0, init_centroids, BH, N, K, D, dtype, max_iters, tol, labels, centroids_out, counts,
                                n_iter_out, nullptr, ws, ws_bytes, stream);
}

int svgb_kmeans_run_sorted(const void* x, long long x_head_stride, const void* init_centroids, int BH, int N, int K,
                           int D, int dtype, int max_iters, float tol, int32_t* labels, void* centroids_out,
                           int32_t* counts, int32_t* n_iter_out, int32_t* perm_out, void* ws, size_t ws_bytes,
                           void* stream) {
  SVGB_REQUIRE(x && init_centroids && labels && centroids_out && counts && ws, "null pointer");
  SVGB_REQUIRE(max_iters >= 1, "max_iters must be >= 1");
  const long long xhs = x_head_stride ? x_head_stride : static_cast<long long>(N) * D;
  SVGB_REQUIRE(xhs >= static_cast<long long>(N) * D && xhs % 8 == 0, "x_head_stride %lld invalid (>= N*D, multiple of 8)", xhs);
  size_t need = 0;
  if (svgb_kmeans_bytes(BH, N, K, D, &need)) return -1;
  SVGB_REQUIRE(ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  const KmLayout L = km_layout(BH, N, K, D);
  char* w = static_cast<char*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* csqp = reinterpret_cast<float*>(w + L.csq_pad);
  int* perm = reinterpret_cast<int*>(w + L.perm);
  int* offs = reinterpret_cast<int*>(w + L.offs);
  void* cb0 = w + L.cbuf0;
  void* cb1 = w + L.cbuf1;
  KmState* state = reinterpret_cast<KmState*>(w + L.state);
  float* shift = reinterpret_cast<float*>(w + L.state + 64);

  km_init_state_kernel<<<1, 1, 0, st>>>(state, shift);
  SVGB_LAUNCH_OK();
  SVGB_CUDA(cudaMemcpyAsync(cb0, init_centroids, 2ull * BH * K * D, cudaMemcpyDeviceToDevice, st));
  // |x|^2 (batch_kmeans_Euclid:704) is computed inside the assign kernel from the X tile it already holds
  for (int it = 0; it < max_iters; ++it) {
    // every kernel of an iteration is a no-op once state->done is set (device-side `break`)
    if (launch_csq(cb0, cb1, csqp, BH, K, D, dtype, state, st)) return -1;
    if (dispatch_assign(x, xhs, cb0, cb1, nullptr, csqp, labels, BH, N, K, D, dtype, state, st)) return -1;
    if (argsort_labels_impl(labels, BH, N, K, perm, counts, offs, w + L.sort, &state->done, st)) return -1;
    // update reads buffer `cur`, writes the other one
    if (launch_update(x, xhs, perm, counts, offs, cb0, cb1, cb1, cb0, shift, BH, N, K, D, dtype, state, st)) return -1;
    km_commit_kernel<<<1, 1, 0, st>>>(state, shift, tol, it);
    SVGB_LAUNCH_OK();
  }
  const long long n_vec = 2ll * BH * K * D / 16;
  km_finalize_kernel<<<64, 256, 0, st>>>(state, static_cast<const uint4*>(cb0), static_cast<const uint4*>(cb1),
                                         static_cast<uint4*>(centroids_out), n_vec, n_iter_out);
  SVGB_LAUNCH_OK();
  // the member lists the last executed update summed over are the stable argsort of the returned labels
  if (perm_out)
    SVGB_CUDA(cudaMemcpyAsync(perm_out, perm, sizeof(int) * static_cast<size_t>(BH) * N, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
