// Host-side helpers shared by every translation unit of libsvgb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace svgb {

// thread-local last-error message (api.cu)
void set_error(const char* fmt, ...);
const char* get_error();

#define SVGB_REQUIRE(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      svgb::set_error(__VA_ARGS__); \
      return -1;                  \
    }                             \
  } while (0)

#define SVGB_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      svgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                      __LINE__);                                                           \
      return -2;                                                                           \
    }                                                                                      \
  } while (0)

// check the launch that just happened (no sync)
#define SVGB_LAUNCH_OK()                                                                   \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      svgb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                           \
      return -3;                                                                           \
    }                                                                                      \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 3-D TMA descriptor over a [BH, S, D] 16-bit tensor addressed as (d, s, h) with byte strides.
// box = 64 x 128 x 1, 128-byte swizzle.  Returns 0 on success.
int encode_tmap_hsd(CUtensorMap* map, const void* base, int dtype, int BH, int S, int D,
                    long long row_stride_elems, long long head_stride_elems, int box_rows = 128);

struct AttnArgs;
// attention launch with independent query / key-value geometry (attn_fwd.cu); used by svgb_attn_fwd
// and by sample_mse's split-KV passes.
int attn_fwd_impl(const void* q, int Sq, long long q_rs, long long q_hs, const void* k, const void* v, int Skv,
                  long long kv_rs, long long kv_hs, int dtype, int BH, int D, const AttnArgs& a, int grid_x,
                  cudaStream_t st);

// stable counting-sort argsort (layout_ops.cu); `offs` (optional) receives the exclusive prefix of
// counts per head; `skip_flag` (optional, device) turns the three kernels into no-ops when non-zero.
int argsort_labels_impl(const int* labels, int BH, int S, int K, int* perm, int* counts, int* offs,
                        void* ws, const int* skip_flag, cudaStream_t stream);

}  // namespace svgb
