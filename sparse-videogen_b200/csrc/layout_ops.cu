// HBM-bound layout transforms on the hot path: token permutation by cluster order (gather /
// inverse scatter), stable argsort of labels (counting sort), SVG1 head placement and its inverse.
// All 16-byte vectorised, rows are D*2 bytes (128 or 256 B) so every access is full-sector.
#include "../../include/svgb200.h"
#include "host_common.h"

namespace svgb {

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

constexpr int kUnroll = 4;

// kScatter == false: out[h, s] = in[h, perm[h, s]]     (gather)
// kScatter == true : out[h, perm[h, s]] = in[h, s]     (inverse scatter)
template <bool kScatter>
__global__ void __launch_bounds__(256)
permute_rows_kernel(const uint4* __restrict__ in, const int* __restrict__ perm, uint4* __restrict__ out,
                    long long n_rows_total, int S, int vec_per_row) {
  const long long total = n_rows_total * vec_per_row;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i < total; i += stride * kUnroll) {
    uint4 val[kUnroll];
    long long dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long idx = i + u * stride;
      dst[u] = -1;
      if (idx < total) {
        const long long row = idx / vec_per_row;
        const int vcol = static_cast<int>(idx - row * vec_per_row);
        const long long h = row / S;
        const long long other = h * S + __ldg(&perm[row]);
        const long long src = kScatter ? row : other;
        dst[u] = (kScatter ? other : row) * vec_per_row + vcol;
        val[u] = ld_stream(in + src * vec_per_row + vcol);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u] >= 0) st_stream(out + dst[u], val[u]);
  }
}

// ---------------------------------------------------------------------------------------------
// Head placement: gather formulation (contiguous writes) for both directions.
//   forward  (inverse=0): video row d = p*F + f  <-  f*P + p
//   inverse  (inverse=1): video row d = f*P + p  <-  p*F + f
// ---------------------------------------------------------------------------------------------
struct PlacementPtrs {
  const uint4* in[3];
  uint4* out[3];
};

__global__ void __launch_bounds__(256)
head_placement_kernel(PlacementPtrs ptrs, const int* __restrict__ best_mask_idx, int S, int vec_per_row,
                      int video0, int F, int P, int inverse) {
  const int h = blockIdx.y;
  const uint4* __restrict__ in = ptrs.in[blockIdx.z] + static_cast<long long>(h) * S * vec_per_row;
  uint4* __restrict__ out = ptrs.out[blockIdx.z] + static_cast<long long>(h) * S * vec_per_row;
  const bool temporal = best_mask_idx[h] == 1;
  const int video1 = video0 + F * P;
  const long long total = static_cast<long long>(S) * vec_per_row;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i < total; i += stride * kUnroll) {
    uint4 val[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long idx = i + u * stride;
      if (idx < total) {
        int row = static_cast<int>(idx / vec_per_row);
        const int vcol = static_cast<int>(idx - static_cast<long long>(row) * vec_per_row);
        if (temporal && row >= video0 && row < video1) {
          const int d = row - video0;
          int src;
          if (!inverse) {
            const int p = d / F, f = d - p * F;
            src = f * P + p;
          } else {
            const int f = d / P, p = d - f * P;
            src = p * F + f;
          }
          row = video0 + src;
        }
        val[u] = ld_stream(in + static_cast<long long>(row) * vec_per_row + vcol);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long idx = i + u * stride;
      if (idx < total) st_stream(out + idx, val[u]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stable counting-sort argsort of labels in [0,K).
//   pass 1: per (head, 1024-token chunk) histogram
//   pass 2: per head, column scan over chunks + exclusive scan over clusters -> chunk bases, counts
//   pass 3: one warp per chunk places tokens in index order (match_any ranks) -> stable
// ---------------------------------------------------------------------------------------------
constexpr int kSortChunk = 1024;

__global__ void sort_hist_kernel(const int* __restrict__ labels, int S, int K, int n_chunks,
                                 int* __restrict__ hist, const int* __restrict__ skip) {
  extern __shared__ int sh[];
  if (skip && *skip) return;
  const int h = blockIdx.y, c = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  const int s0 = c * kSortChunk, s1 = min(S, s0 + kSortChunk);
  for (int s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
    const int l = labels[static_cast<long long>(h) * S + s];
    if (l >= 0 && l < K) atomicAdd(&sh[l], 1);
  }
  __syncthreads();
  int* dst = hist + (static_cast<long long>(h) * n_chunks + c) * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) dst[k] = sh[k];
}

// one CTA per head, one thread per cluster (blockDim.x >= K is not required: clusters are strided): the column scan over
// the chunks keeps 8 independent loads in flight per thread, then a block-wide exclusive scan over the K totals gives
// the cluster offsets.  hist is left holding, per (chunk, cluster), the number of that cluster's tokens in EARLIER
// chunks; the place kernel adds the cluster offset itself (one pass over the table instead of two).
__global__ void __launch_bounds__(1024)
sort_scan_kernel(int K, int n_chunks, int* __restrict__ hist, int* __restrict__ counts,
                 int* __restrict__ offs, const int* __restrict__ skip) {
  extern __shared__ int sh[];  // K totals -> exclusive offsets
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (skip && *skip) return;
  const int h = blockIdx.x;
  int* base = hist + static_cast<long long>(h) * n_chunks * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int acc = 0;
    int c = 0;
    for (; c + 8 <= n_chunks; c += 8) {
      int v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = base[static_cast<long long>(c + u) * K + k];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        base[static_cast<long long>(c + u) * K + k] = acc;  // tokens of cluster k in earlier chunks
        acc += v[u];
      }
    }
    for (; c < n_chunks; ++c) {
      const int v = base[static_cast<long long>(c) * K + k];
      base[static_cast<long long>(c) * K + k] = acc;
      acc += v;
    }
    sh[k] = acc;
    if (counts) counts[static_cast<long long>(h) * K + k] = acc;
  }
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  // exclusive scan of sh[0..K) in tiles of blockDim.x
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k0 = 0; k0 < K; k0 += blockDim.x) {
    const int k = k0 + threadIdx.x;
    const int v = k < K ? sh[k] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int t = lane < nwarps ? warp_tot[lane] : 0;
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += n;
      }
      warp_tot[lane] = ti - t;  // exclusive over warps
    }
    __syncthreads();
    const int excl = carry_s + warp_tot[warp] + incl - v;
    if (k < K) {
      sh[k] = excl;
      if (offs) offs[static_cast<long long>(h) * K + k] = excl;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
    __syncthreads();
  }
  // the place kernel needs the offsets even when the caller passed offs == nullptr: keep them behind the table
  int* offs_ws = hist + static_cast<long long>(gridDim.x) * n_chunks * K + static_cast<long long>(h) * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) offs_ws[k] = sh[k];
}

__global__ void __launch_bounds__(32)
sort_place_kernel(const int* __restrict__ labels, int S, int K, int n_chunks,
                  const int* __restrict__ hist, int* __restrict__ perm, const int* __restrict__ skip) {
  extern __shared__ int cnt[];  // running position per cluster for this chunk
  if (skip && *skip) return;
  const int h = blockIdx.y, c = blockIdx.x, lane = threadIdx.x;
  const int* base = hist + (static_cast<long long>(h) * n_chunks + c) * K;
  const int* offs_ws = hist + static_cast<long long>(gridDim.y) * n_chunks * K + static_cast<long long>(h) * K;
  for (int k = lane; k < K; k += 32) cnt[k] = base[k] + offs_ws[k];
  __syncwarp();
  const int s0 = c * kSortChunk, s1 = min(S, s0 + kSortChunk);
  for (int s = s0; s < s1; s += 32) {
    const int idx = s + lane;
    const bool live = idx < s1;
    int l = live ? labels[static_cast<long long>(h) * S + idx] : -1 - lane;
    if (live && (l < 0 || l >= K)) l = -1 - lane;  // out-of-range labels are dropped
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    const int rank = __popc(peers & ((1u << lane) - 1));
    int pos = 0;
    if (l >= 0) pos = cnt[l] + rank;
    __syncwarp();
    if (l >= 0 && rank == 0) cnt[l] += __popc(peers);
    __syncwarp();
    if (l >= 0) perm[static_cast<long long>(h) * S + pos] = idx;
  }
}

static int grid_for(long long work_items, int block) {
  long long g = (work_items + static_cast<long long>(block) * kUnroll - 1) / (static_cast<long long>(block) * kUnroll);
  const long long cap = 148LL * 16;  // a few waves of persistent-ish blocks
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int argsort_labels_impl(const int* labels, int BH, int S, int K, int* perm, int* counts, int* offs,
                        void* ws, const int* skip_flag, cudaStream_t st) {
  SVGB_REQUIRE(K * sizeof(int) <= 48 * 1024, "K=%d too large", K);
  const int n_chunks = (S + kSortChunk - 1) / kSortChunk;
  int* hist = static_cast<int*>(ws);
  sort_hist_kernel<<<dim3(n_chunks, BH), 256, K * sizeof(int), st>>>(labels, S, K, n_chunks, hist, skip_flag);
  SVGB_LAUNCH_OK();
  const int scan_threads = K >= 1024 ? 1024 : ((K + 31) / 32) * 32;
  sort_scan_kernel<<<BH, scan_threads, K * sizeof(int), st>>>(K, n_chunks, hist, counts, offs, skip_flag);
  SVGB_LAUNCH_OK();
  sort_place_kernel<<<dim3(n_chunks, BH), 32, K * sizeof(int), st>>>(labels, S, K, n_chunks, hist, perm, skip_flag);
  SVGB_LAUNCH_OK();
  return 0;
}

}  // namespace svgb

using namespace svgb;

extern "C" {

int svgb_permute_gather(const void* in, const int32_t* perm, void* out, int BH, int S, int D,
                        void* stream) {
  SVGB_REQUIRE(in && perm && out && BH > 0 && S > 0 && D > 0 && D % 8 == 0, "bad arguments");
  const int vpr = D * 2 / 16;
  const long long rows = static_cast<long long>(BH) * S;
  permute_rows_kernel<false><<<grid_for(rows * vpr, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), perm, static_cast<uint4*>(out), rows, S, vpr);
  SVGB_LAUNCH_OK();
  return 0;
}

int svgb_permute_scatter(const void* in, const int32_t* perm, void* out, int BH, int S, int D,
                         void* stream) {
  SVGB_REQUIRE(in && perm && out && BH > 0 && S > 0 && D > 0 && D % 8 == 0, "bad arguments");
  const int vpr = D * 2 / 16;
  const long long rows = static_cast<long long>(BH) * S;
  permute_rows_kernel<true><<<grid_for(rows * vpr, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(in), perm, static_cast<uint4*>(out), rows, S, vpr);
  SVGB_LAUNCH_OK();
  return 0;
}

int svgb_head_placement(const void* const* in, void* const* out, int n_tensors,
                        const int32_t* best_mask_idx, int BH, int S, int D, int ctx, int F, int P,
                        int text_first, int inverse, void* stream) {
  SVGB_REQUIRE(in && out && best_mask_idx, "null pointer");
  SVGB_REQUIRE(n_tensors >= 1 && n_tensors <= 3, "n_tensors must be 1..3");
  SVGB_REQUIRE(S == ctx + F * P, "seq_len %d != ctx %d + F %d * P %d", S, ctx, F, P);
  SVGB_REQUIRE(D % 8 == 0, "head_dim must be a multiple of 8");
  PlacementPtrs p{};
  for (int i = 0; i < n_tensors; ++i) {
    SVGB_REQUIRE(in[i] && out[i], "null tensor %d", i);
    p.in[i] = static_cast<const uint4*>(in[i]);
    p.out[i] = static_cast<uint4*>(out[i]);
  }
  const int vpr = D * 2 / 16;
  int gx = grid_for(static_cast<long long>(S) * vpr, 256);
  if (gx > 64) gx = 64;
  dim3 grid(gx, BH, n_tensors);
  head_placement_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, best_mask_idx, S, vpr, text_first ? ctx : 0, F, P, inverse);
  SVGB_LAUNCH_OK();
  return 0;
}

int svgb_argsort_labels_bytes(int BH, int S, int K, size_t* bytes) {
  SVGB_REQUIRE(BH > 0 && S > 0 && K > 0 && bytes, "bad arguments");
  const size_t n_chunks = (S + kSortChunk - 1) / kSortChunk;
  *bytes = align_up(sizeof(int) * (BH * n_chunks * K + static_cast<size_t>(BH) * K), 256);  // table + cluster offsets
  return 0;
}

int svgb_argsort_labels(const int32_t* labels, int BH, int S, int K, int32_t* perm, int32_t* counts,
                        void* ws, size_t ws_bytes, void* stream) {
  size_t need = 0;
  if (svgb_argsort_labels_bytes(BH, S, K, &need)) return -1;
  SVGB_REQUIRE(labels && perm && ws, "null pointer");
  SVGB_REQUIRE(ws_bytes >= need, "workspace too small: %zu < %zu", ws_bytes, need);
  return argsort_labels_impl(labels, BH, S, K, perm, counts, nullptr, ws, nullptr,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
