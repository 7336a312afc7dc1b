// Transformer-block glue either side of attention in Wan: the reference's Triton kernels
//   triton_layernorm_forward  (svg/kernels/triton/layernorm.py: with / without affine, fp32 result)
//   triton_modulate_shift_forward / triton_modulate_gate_residual_forward (svg/kernels/triton/modulate.py)
//   triton_rmsnorm_forward    (svg/kernels/triton/rmsnorm.py: RMS over the full hidden row)
// as used by WanTransformerBlock_Sparse.forward (svg/models/wan/custom_models.py:37-111).  On B200 the
// LayerNorm and the modulation that always follows it are ONE pass (the reference writes and re-reads an
// fp32 copy of the hidden state between them).  HBM-bound: one CTA per hidden row, the row lives in registers,
// 16-byte vectors, two block reductions for LayerNorm (mean, then centred variance, like the reference).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/svgb200.h"
#include "host_common.h"

namespace svgb {

constexpr int kGlueThreads = 256;
constexpr int kMaxVec = 4;  // 8-element groups per thread: N <= 256 * 4 * 8 = 8192

// A row segment of 8 elements as it sits in HBM: one 16-byte vector (16-bit storage) or two (fp32).  Rows are
// kept in this packed form in registers and unpacked in every pass, so that the NEXT row can be prefetched into
// a second set of registers while the current one is reduced (bytes in flight, not ALU, bound these kernels).
template <bool F32>
struct Raw8 {
  uint4 v[F32 ? 2 : 1];
};
template <bool F32>
__device__ __forceinline__ Raw8<F32> load_raw(const void* base, long long idx) {
  Raw8<F32> r;
  if constexpr (F32) {
    const uint4* p = reinterpret_cast<const uint4*>(static_cast<const float*>(base) + idx);
    r.v[0] = p[0];
    r.v[1] = p[1];
  } else {
    r.v[0] = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(base) + idx);
  }
  return r;
}
template <bool F32>
__device__ __forceinline__ void unpack_raw(const Raw8<F32>& r, int dtype, float (&f)[8]) {
  if constexpr (F32) {
    f[0] = __uint_as_float(r.v[0].x); f[1] = __uint_as_float(r.v[0].y);
    f[2] = __uint_as_float(r.v[0].z); f[3] = __uint_as_float(r.v[0].w);
    f[4] = __uint_as_float(r.v[1].x); f[5] = __uint_as_float(r.v[1].y);
    f[6] = __uint_as_float(r.v[1].z); f[7] = __uint_as_float(r.v[1].w);
  } else {
    const uint32_t w[4] = {r.v[0].x, r.v[0].y, r.v[0].z, r.v[0].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == SVGB_BF16) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      } else {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        f[2 * i] = __low2float(h);
        f[2 * i + 1] = __high2float(h);
      }
    }
  }
}
// small, cache-resident vectors (weights, modulation): any storage type, decided at run time
__device__ __forceinline__ void load8(const void* base, long long idx, int dtype, float (&f)[8]) {
  if (dtype == SVGB_F32) unpack_raw<true>(load_raw<true>(base, idx), dtype, f);
  else unpack_raw<false>(load_raw<false>(base, idx), dtype, f);
}
__device__ __forceinline__ void store8(void* base, long long idx, int dtype, const float (&f)[8]) {
  if (dtype == SVGB_F32) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + idx);
    p[0] = make_float4(f[0], f[1], f[2], f[3]);
    p[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == SVGB_BF16) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&b);
      } else {
        const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    *reinterpret_cast<uint4*>(static_cast<uint16_t*>(base) + idx) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // red[] may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kGlueThreads / 32; ++w) r += red[w];
  return r;
}

struct GlueArgs {
  const void* x;
  const void* res;    // residual (gate_residual); same storage width as x
  void* y;
  const void* w;      // norm weight (optional)
  const void* b;      // norm bias (optional)
  const float* scale; // [B or 1, N]
  const float* shift;
  const float* gate;
  long long rows;
  long long rows_per_batch;  // modulation vectors are indexed by row / rows_per_batch (0: one vector for all)
  int N, x_dtype, y_dtype, w_dtype, res_dtype;
  float eps;
};

enum GlueOp { kOpLayerNorm = 0, kOpRmsNorm = 1, kOpModulate = 2, kOpGateResidual = 3 };

// kOpLayerNorm: y = LN(x)[*w + b][*(1+scale) + shift];  kOpRmsNorm: y = x * rstd * w
// kOpModulate : y = x * (1+scale) + shift;              kOpGateResidual: y = res + x * gate
template <int OP, bool F32, bool RF32>
__global__ void __launch_bounds__(kGlueThreads)
glue_rows_kernel(const GlueArgs a) {
  __shared__ float red[kGlueThreads / 32];
  constexpr bool kRes = OP == kOpGateResidual;
  const int nvec = a.N / 8;
  Raw8<F32> cur[kMaxVec], nxt[kMaxVec];
  Raw8<RF32> rcur[kRes ? kMaxVec : 1], rnxt[kRes ? kMaxVec : 1];
  auto fetch = [&](long long row, Raw8<F32>(&d)[kMaxVec], Raw8<RF32>(&r)[kRes ? kMaxVec : 1]) {
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = threadIdx.x + j * kGlueThreads;
      if (vi < nvec) {
        d[j] = load_raw<F32>(a.x, row * a.N + vi * 8);
        if constexpr (kRes) r[j] = load_raw<RF32>(a.res, row * a.N + vi * 8);
      }
    }
  };
  long long row = blockIdx.x;
  if (row < a.rows) fetch(row, cur, rcur);
  for (; row < a.rows; row += gridDim.x) {
    if (row + gridDim.x < a.rows) fetch(row + gridDim.x, nxt, rnxt);
    const long long off = row * a.N;
    const long long mod_off = a.rows_per_batch > 0 ? (row / a.rows_per_batch) * a.N : 0;
    float mean = 0.f, rstd = 1.f;
    if constexpr (OP == kOpLayerNorm) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        if (threadIdx.x + j * kGlueThreads < nvec) {
          float f[8];
          unpack_raw<F32>(cur[j], a.x_dtype, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) s += f[i];
        }
      }
      mean = block_sum(s, red) / a.N;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        if (threadIdx.x + j * kGlueThreads < nvec) {
          float f[8];
          unpack_raw<F32>(cur[j], a.x_dtype, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) ss += (f[i] - mean) * (f[i] - mean);
        }
      }
      rstd = 1.0f / sqrtf(block_sum(ss, red) / a.N + a.eps);
    } else if constexpr (OP == kOpRmsNorm) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) {
        if (threadIdx.x + j * kGlueThreads < nvec) {
          float f[8];
          unpack_raw<F32>(cur[j], a.x_dtype, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
        }
      }
      rstd = 1.0f / sqrtf(block_sum(ss, red) / a.N + a.eps);
    }
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = threadIdx.x + j * kGlueThreads;
      if (vi < nvec) {
        float o[8];
        unpack_raw<F32>(cur[j], a.x_dtype, o);
        if constexpr (OP == kOpLayerNorm || OP == kOpRmsNorm) {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = (o[i] - mean) * rstd;
          if (a.w) {
            float w[8];
            load8(a.w, vi * 8, a.w_dtype, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] *= w[i];
          }
          if (a.b) {
            float b[8];
            load8(a.b, vi * 8, a.w_dtype, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += b[i];
          }
        }
        if constexpr (OP == kOpLayerNorm || OP == kOpModulate) {
          if (a.scale) {
            float sc[8], sh[8];
            load8(a.scale, mod_off + vi * 8, SVGB_F32, sc);
            load8(a.shift, mod_off + vi * 8, SVGB_F32, sh);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = o[i] * (1.0f + sc[i]) + sh[i];
          }
        }
        if constexpr (kRes) {
          float r[8], g[8];
          unpack_raw<RF32>(rcur[j], a.res_dtype, r);
          load8(a.gate, mod_off + vi * 8, SVGB_F32, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = r[i] + o[i] * g[i];
        }
        store8(a.y, off + vi * 8, a.y_dtype, o);
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      cur[j] = nxt[j];
      if constexpr (kRes) rcur[j] = rnxt[j];
    }
  }
}

static bool dtype_ok(int d) { return d == SVGB_BF16 || d == SVGB_F16 || d == SVGB_F32; }
static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int OP>
static int launch_glue(const GlueArgs& a, cudaStream_t st) {
  SVGB_REQUIRE(a.x && a.y, "null pointer");
  SVGB_REQUIRE(a.rows >= 0 && a.N > 0 && a.N % 8 == 0 && a.N <= kGlueThreads * kMaxVec * 8,
               "hidden size must be a multiple of 8 and <= %d", kGlueThreads * kMaxVec * 8);
  SVGB_REQUIRE(dtype_ok(a.x_dtype) && dtype_ok(a.y_dtype), "dtype unsupported");
  SVGB_REQUIRE(al16(a.x) && al16(a.y) && al16(a.w) && al16(a.b) && al16(a.scale) && al16(a.shift) && al16(a.gate) &&
                   al16(a.res),
               "pointers must be 16-byte aligned");
  if (a.rows == 0) return 0;
  // a few resident waves; every CTA streams rows with the next one prefetched
  long long blocks = a.rows < 148LL * 8 ? a.rows : 148LL * 8;
  const unsigned gb = static_cast<unsigned>(blocks);
  const bool xf = a.x_dtype == SVGB_F32;
  if constexpr (OP == kOpGateResidual) {
    const bool rf = a.res_dtype == SVGB_F32;
    if (xf && rf) glue_rows_kernel<OP, true, true><<<gb, kGlueThreads, 0, st>>>(a);
    else if (xf) glue_rows_kernel<OP, true, false><<<gb, kGlueThreads, 0, st>>>(a);
    else if (rf) glue_rows_kernel<OP, false, true><<<gb, kGlueThreads, 0, st>>>(a);
    else glue_rows_kernel<OP, false, false><<<gb, kGlueThreads, 0, st>>>(a);
  } else {
    if (xf) glue_rows_kernel<OP, true, false><<<gb, kGlueThreads, 0, st>>>(a);
    else glue_rows_kernel<OP, false, false><<<gb, kGlueThreads, 0, st>>>(a);
  }
  SVGB_LAUNCH_OK();
  return 0;
}

}  // namespace svgb

using namespace svgb;

extern "C" {

int svgb_layernorm_modulate(const void* x, int x_dtype, const void* w, const void* b, int w_dtype, float eps,
                            const float* scale, const float* shift, long long rows_per_batch, void* y, int y_dtype,
                            long long rows, int N, void* stream) {
  SVGB_REQUIRE((w == nullptr) == (b == nullptr), "weight and bias come together (elementwise_affine)");
  SVGB_REQUIRE(!w || dtype_ok(w_dtype), "weight dtype unsupported");
  SVGB_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift come together");
  GlueArgs a{};
  a.x = x; a.y = y; a.w = w; a.b = b; a.scale = scale; a.shift = shift;
  a.rows = rows; a.rows_per_batch = rows_per_batch; a.N = N;
  a.x_dtype = x_dtype; a.y_dtype = y_dtype; a.w_dtype = w_dtype; a.eps = eps;
  return launch_glue<kOpLayerNorm>(a, static_cast<cudaStream_t>(stream));
}

int svgb_rmsnorm_hidden(const void* x, int x_dtype, const void* w, int w_dtype, float eps, void* y, int y_dtype,
                        long long rows, int N, void* stream) {
  SVGB_REQUIRE(w && dtype_ok(w_dtype), "weight required");
  GlueArgs a{};
  a.x = x; a.y = y; a.w = w; a.rows = rows; a.N = N;
  a.x_dtype = x_dtype; a.y_dtype = y_dtype; a.w_dtype = w_dtype; a.eps = eps;
  return launch_glue<kOpRmsNorm>(a, static_cast<cudaStream_t>(stream));
}

int svgb_modulate_shift(const void* x, int x_dtype, const float* scale, const float* shift,
                        long long rows_per_batch, void* y, int y_dtype, long long rows, int N, void* stream) {
  SVGB_REQUIRE(scale && shift, "null pointer");
  GlueArgs a{};
  a.x = x; a.y = y; a.scale = scale; a.shift = shift; a.rows = rows; a.rows_per_batch = rows_per_batch; a.N = N;
  a.x_dtype = x_dtype; a.y_dtype = y_dtype;
  return launch_glue<kOpModulate>(a, static_cast<cudaStream_t>(stream));
}

int svgb_gate_residual(const void* residual, int res_dtype, const void* x, int x_dtype, const float* gate,
                       long long rows_per_batch, void* y, int y_dtype, long long rows, int N, void* stream) {
  SVGB_REQUIRE(residual && gate && dtype_ok(res_dtype), "null pointer / dtype");
  GlueArgs a{};
  a.x = x; a.res = residual; a.y = y; a.gate = gate; a.rows = rows; a.rows_per_batch = rows_per_batch; a.N = N;
  a.x_dtype = x_dtype; a.y_dtype = y_dtype; a.res_dtype = res_dtype;
  return launch_glue<kOpGateResidual>(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
