// Pre-attention elementwise chain: the reference's native `_kernels` ops
//   rms_norm_forward / layer_norm_forward            (svg/kernels/csrc/ops.h:20-75, include/norm/narrow_*.cuh)
//   apply_qk_rope_inplace_cossin{,_txtlast,_complex} (ops.h:77-260, include/rope/rope_enc*.cuh)
// as in-place drop-ins, plus ONE fused pass that replaces the whole reference chain
//   [B,S,H*D] -> unflatten/transpose/contiguous (hyvideo/attention.py:260-266) -> QK norm -> RoPE -> torch.cat
// reading each projection output once and writing Q/K/V once in the [B,H,S,D] layout the attention kernel's
// TMA descriptors read.  All kernels are HBM-bound: 16-byte vectors, a row segment of 8 elements per thread,
// D/8 lanes cooperate on one head row (sub-warp butterfly, same order as the reference's reduction).
//
// Rounding points reproduced from the reference chain: the norm result is rounded to the 16-bit dtype
// (in-place store) before RoPE reads it; RoPE computes in fp32 (fp64 for the `complex` variant,
// rope_enc_complex.cuh:29-43) and rounds once.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../../include/svgb200.h"
#include "host_common.h"

namespace svgb {

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (BF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      f[2 * i] = __low2float(h);
      f[2 * i + 1] = __high2float(h);
    }
  }
}
template <bool BF16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (BF16) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&b);
    } else {
      const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <bool BF16>
__device__ __forceinline__ float round16f(float x) {
  if constexpr (BF16) return __bfloat162float(__float2bfloat16_rn(x));
  else return __half2float(__float2half_rn(x));
}
template <bool BF16>
__device__ __forceinline__ float round16d(double x) {  // static_cast<T>(double), T the 16-bit type
  if constexpr (BF16) return __bfloat162float(__double2bfloat16(x));
  else return __half2float(__double2half(x));
}

// butterfly over the LPR lanes that share a row (LPR = D/8 is a power of two <= 32)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- the three per-row transforms, shared by the in-place kernels and the fused pass -----------
template <int D>
__device__ __forceinline__ void rms_apply(float (&f)[8], const float (&g)[8], float eps) {
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
  ss = group_sum<D / 8>(ss);
  const float inv = rsqrtf(ss / D + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = f[i] * inv * g[i];
}
template <int D>
__device__ __forceinline__ void ln_apply(float (&f)[8], const float (&g)[8], const float (&b)[8]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  const float mean = group_sum<D / 8>(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) ss += (f[i] - mean) * (f[i] - mean);
  const float inv = rsqrtf(group_sum<D / 8>(ss) / D + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * inv * g[i] + b[i];
}
// interleaved pairs (2i, 2i+1): out[e] = x[e]*cos[e] + (e even ? -x[e^1] : x[e^1]) * sin[e]
__device__ __forceinline__ void rope_apply(float (&f)[8], const float (&c)[8], const float (&s)[8]) {
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = f[i] * c[i] + ((i & 1) ? f[i ^ 1] : -f[i ^ 1]) * s[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = o[i];
}
template <bool BF16>
__device__ __forceinline__ void rope_apply_complex(float (&f)[8], const float (&c)[4], const float (&s)[4]) {
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double x = f[i], y = f[i ^ 1], cc = c[i >> 1], sn = s[i >> 1];
    o[i] = round16d<BF16>(x * cc + ((i & 1) ? y : -y) * sn);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = o[i];
}

__device__ __forceinline__ void load_f8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load_f4(const float* p, float (&f)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}

constexpr int kThreads = 256;
#ifndef SVGB_PREP_UNROLL
#define SVGB_PREP_UNROLL 4
#endif
constexpr int kUnrollRows = SVGB_PREP_UNROLL;  // independent 16-byte loads in flight per thread

// ---- in-place row norms on [m, N] ---------------------------------------------------------------
template <bool BF16, int N, bool kLayer>
__global__ void __launch_bounds__(kThreads)
row_norm_kernel(uint4* __restrict__ x, const uint4* __restrict__ gamma, const uint4* __restrict__ beta, float eps,
                long long m) {
  constexpr int LPR = N / 8, RPB = kThreads / LPR;
  const int seg = threadIdx.x % LPR, sub = threadIdx.x / LPR;
  float g[8], b[8];
  unpack8<BF16>(__ldg(gamma + seg), g);
  if constexpr (kLayer) unpack8<BF16>(__ldg(beta + seg), b);
  const long long pass = static_cast<long long>(gridDim.x) * RPB;
  for (long long base = static_cast<long long>(blockIdx.x) * RPB; base < m; base += pass * kUnrollRows) {
    uint4 v[kUnrollRows];
#pragma unroll
    for (int u = 0; u < kUnrollRows; ++u) {
      const long long r = base + u * pass + sub;
      v[u] = r < m ? x[r * LPR + seg] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kUnrollRows; ++u) {
      const long long r = base + u * pass + sub;
      float f[8];
      unpack8<BF16>(v[u], f);
      if constexpr (kLayer) ln_apply<N>(f, g, b);
      else rms_apply<N>(f, g, eps);
      if (r < m) x[r * LPR + seg] = pack8<BF16>(f);
    }
  }
}

// ---- in-place RoPE on q [B,Hq,S,D], k [B,Hk,S,D]; rows [row0, row0+valid) of every head ------------
template <bool BF16, int D, bool kComplex>
__global__ void __launch_bounds__(kThreads)
qk_rope_kernel(uint4* __restrict__ q, uint4* __restrict__ k, const float* __restrict__ cos_t,
               const float* __restrict__ sin_t, int Hq, int Hk, int S, int row0, int valid) {
  constexpr int LPR = D / 8, RPB = kThreads / LPR;
  const int seg = threadIdx.x % LPR;
  const int t = blockIdx.x * RPB + threadIdx.x / LPR;
  if (t >= valid) return;
  float c8[8], s8[8], c4[4], s4[4];
  if constexpr (kComplex) {
    load_f4(cos_t + static_cast<long long>(t) * (D / 2) + seg * 4, c4);
    load_f4(sin_t + static_cast<long long>(t) * (D / 2) + seg * 4, s4);
  } else {
    load_f8(cos_t + static_cast<long long>(t) * D + seg * 8, c8);
    load_f8(sin_t + static_cast<long long>(t) * D + seg * 8, s8);
  }
  const long long hs = static_cast<long long>(S) * LPR;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const int H = which ? Hk : Hq;
    uint4* p = (which ? k : q) + (static_cast<long long>(blockIdx.y) * H * S + row0 + t) * LPR + seg;
#pragma unroll 1
    for (int h0 = 0; h0 < H; h0 += kUnrollRows) {
      uint4 v[kUnrollRows];
#pragma unroll
      for (int u = 0; u < kUnrollRows; ++u)
        if (h0 + u < H) v[u] = p[(h0 + u) * hs];
#pragma unroll
      for (int u = 0; u < kUnrollRows; ++u) {
        if (h0 + u < H) {
          float f[8];
          unpack8<BF16>(v[u], f);
          if constexpr (kComplex) rope_apply_complex<BF16>(f, c4, s4);
          else rope_apply(f, c8, s8);
          p[(h0 + u) * hs] = pack8<BF16>(f);
        }
      }
    }
  }
}

// ---- fused transpose + norm + RoPE ----------------------------------------------------------------
struct PrepArgs {
  const uint16_t* in[3];   // q, k, v projection outputs [B, S_in, H*D] (token stride / batch stride below)
  uint16_t* out[3];        // [B, H, S_out, D]; this call fills rows [out_row0, out_row0 + S_in)
  const uint16_t* gamma[2];
  const uint16_t* beta[2];
  const float* cos_t;
  const float* sin_t;
  long long in_token_stride, in_batch_stride, out_head_stride, out_batch_stride;
  int S_in, H, out_row0, norm, rope, rope_lo, rope_n;
  float eps;
};

template <bool BF16, int D>
__global__ void __launch_bounds__(kThreads)
qkv_prep_kernel(const PrepArgs a) {
  constexpr int LPR = D / 8, RPB = kThreads / LPR;
  const int seg = threadIdx.x % LPR;
  const int t = blockIdx.x * RPB + threadIdx.x / LPR;
  const int which = blockIdx.z;  // 0 q, 1 k, 2 v
  const bool live = t < a.S_in;
  const int tt = live ? t : a.S_in - 1;  // keep the whole warp in the shuffles; dead lanes never store
  const uint16_t* in = a.in[which] + blockIdx.y * a.in_batch_stride + tt * a.in_token_stride + seg * 8;
  uint16_t* out = a.out[which] + blockIdx.y * a.out_batch_stride + static_cast<long long>(a.out_row0 + tt) * D + seg * 8;
  const int norm = which < 2 ? a.norm : 0;
  const int rp = a.rope;
  const bool rot = which < 2 && rp != 0 && tt >= a.rope_lo && tt < a.rope_lo + a.rope_n;
  float g[8], b[8], c8[8], s8[8], c4[4], s4[4];
  if (norm == 1 || norm == 2) unpack8<BF16>(__ldg(reinterpret_cast<const uint4*>(a.gamma[which]) + seg), g);
  if (norm == 2) unpack8<BF16>(__ldg(reinterpret_cast<const uint4*>(a.beta[which]) + seg), b);
  if (rot) {
    const long long tr = tt - a.rope_lo;
    if (rp == 2) {
      load_f4(a.cos_t + tr * (D / 2) + seg * 4, c4);
      load_f4(a.sin_t + tr * (D / 2) + seg * 4, s4);
    } else {
      load_f8(a.cos_t + tr * D + seg * 8, c8);
      load_f8(a.sin_t + tr * D + seg * 8, s8);
    }
  }
  float inv_full = 0.f;
  if (norm == 3) {  // RMS over the full hidden row H*D (Wan QK-norm before the head split, wan/attention.py:107-120)
    float ss = 0.f;
#pragma unroll 1
    for (int h0 = 0; h0 < a.H; h0 += kUnrollRows) {
      uint4 v[kUnrollRows];
#pragma unroll
      for (int u = 0; u < kUnrollRows; ++u)
        v[u] = h0 + u < a.H ? *reinterpret_cast<const uint4*>(in + (h0 + u) * D) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < kUnrollRows; ++u) {
        float f[8];
        unpack8<BF16>(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
      }
    }
    ss = group_sum<LPR>(ss);
    inv_full = 1.0f / sqrtf(ss / (a.H * D) + a.eps);
  }
#pragma unroll 1
  for (int h0 = 0; h0 < a.H; h0 += kUnrollRows) {
    uint4 v[kUnrollRows];
#pragma unroll
    for (int u = 0; u < kUnrollRows; ++u)
      if (h0 + u < a.H) v[u] = *reinterpret_cast<const uint4*>(in + (h0 + u) * D);
#pragma unroll
    for (int u = 0; u < kUnrollRows; ++u) {
      if (h0 + u < a.H) {  // block-uniform
        if (norm != 0 || rot) {
          float f[8];
          unpack8<BF16>(v[u], f);
          if (norm == 1) rms_apply<D>(f, g, a.eps);
          else if (norm == 2) ln_apply<D>(f, g, b);
          else if (norm == 3) {
            float gw[8];
            unpack8<BF16>(__ldg(reinterpret_cast<const uint4*>(a.gamma[which] + (h0 + u) * D) + seg), gw);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = f[i] * inv_full * gw[i];
          }
          if (rot) {
            if (norm != 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = round16f<BF16>(f[i]);
            }
            if (rp == 2) rope_apply_complex<BF16>(f, c4, s4);
            else rope_apply(f, c8, s8);
          }
          v[u] = pack8<BF16>(f);
        }
        if (live) *reinterpret_cast<uint4*>(out + (h0 + u) * a.out_head_stride) = v[u];
      }
    }
  }
}

template <bool BF16, bool kLayer>
static int launch_row_norm(void* x, const void* gamma, const void* beta, float eps, long long m, int n,
                           cudaStream_t st) {
  const long long rpb = kThreads / (n / 8);
  long long blocks = (m + rpb * kUnrollRows - 1) / (rpb * kUnrollRows);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  auto X = static_cast<uint4*>(x);
  auto G = static_cast<const uint4*>(gamma);
  auto B = static_cast<const uint4*>(beta);
  const unsigned gb = static_cast<unsigned>(blocks);
  switch (n) {
    case 32: row_norm_kernel<BF16, 32, kLayer><<<gb, kThreads, 0, st>>>(X, G, B, eps, m); break;
    case 64: row_norm_kernel<BF16, 64, kLayer><<<gb, kThreads, 0, st>>>(X, G, B, eps, m); break;
    case 128: row_norm_kernel<BF16, 128, kLayer><<<gb, kThreads, 0, st>>>(X, G, B, eps, m); break;
    default: row_norm_kernel<BF16, 256, kLayer><<<gb, kThreads, 0, st>>>(X, G, B, eps, m); break;
  }
  SVGB_LAUNCH_OK();
  return 0;
}

template <bool BF16, bool kComplex>
static int launch_rope(void* q, void* k, const float* c, const float* s, int B, int Hq, int Hk, int S, int D, int row0,
                       int valid, cudaStream_t st) {
  const int rpb = kThreads / (D / 8);
  dim3 grid((valid + rpb - 1) / rpb, B);
  auto Q = static_cast<uint4*>(q);
  auto K = static_cast<uint4*>(k);
  switch (D) {
    case 64: qk_rope_kernel<BF16, 64, kComplex><<<grid, kThreads, 0, st>>>(Q, K, c, s, Hq, Hk, S, row0, valid); break;
    case 128: qk_rope_kernel<BF16, 128, kComplex><<<grid, kThreads, 0, st>>>(Q, K, c, s, Hq, Hk, S, row0, valid); break;
    default: qk_rope_kernel<BF16, 256, kComplex><<<grid, kThreads, 0, st>>>(Q, K, c, s, Hq, Hk, S, row0, valid); break;
  }
  SVGB_LAUNCH_OK();
  return 0;
}

}  // namespace svgb

using namespace svgb;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" {

int svgb_rms_norm(void* x, const void* gamma, long long m, int n, float eps, int dtype, void* stream) {
  SVGB_REQUIRE(x && gamma && m >= 0, "null pointer");
  SVGB_REQUIRE(n == 32 || n == 64 || n == 128 || n == 256, "Unsupported head_dim: %d", n);
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16, "dtype %d unsupported", dtype);
  SVGB_REQUIRE(aligned16(x) && aligned16(gamma), "pointers must be 16-byte aligned");
  if (m == 0) return 0;
  auto st = static_cast<cudaStream_t>(stream);
  return dtype == SVGB_BF16 ? launch_row_norm<true, false>(x, gamma, nullptr, eps, m, n, st)
                            : launch_row_norm<false, false>(x, gamma, nullptr, eps, m, n, st);
}

int svgb_layer_norm(void* x, const void* gamma, const void* beta, long long m, int n, int dtype, void* stream) {
  SVGB_REQUIRE(x && gamma && beta && m >= 0, "null pointer");
  SVGB_REQUIRE(n == 32 || n == 64 || n == 128 || n == 256, "Unsupported head_dim: %d", n);
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16, "dtype %d unsupported", dtype);
  SVGB_REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta), "pointers must be 16-byte aligned");
  if (m == 0) return 0;
  auto st = static_cast<cudaStream_t>(stream);
  return dtype == SVGB_BF16 ? launch_row_norm<true, true>(x, gamma, beta, 0.f, m, n, st)
                            : launch_row_norm<false, true>(x, gamma, beta, 0.f, m, n, st);
}

int svgb_qk_rope(void* q, void* k, const float* cos_t, const float* sin_t, int B, int Hq, int Hk, int S, int D,
                 int len_text, int mode, int dtype, void* stream) {
  SVGB_REQUIRE(q && k && cos_t && sin_t, "null pointer");
  SVGB_REQUIRE(B > 0 && Hq > 0 && Hk > 0 && S > 0, "bad sizes");
  SVGB_REQUIRE(D == 64 || D == 128 || D == 256, "Unsupported head_dim: %d", D);
  SVGB_REQUIRE(len_text >= 0 && len_text < S, "len_text_prompt must leave at least one rotated row");
  SVGB_REQUIRE(mode >= SVGB_ROPE_TXT_FIRST && mode <= SVGB_ROPE_COMPLEX_TXT_FIRST, "mode %d unknown", mode);
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16, "dtype %d unsupported", dtype);
  SVGB_REQUIRE(aligned16(q) && aligned16(k) && aligned16(cos_t) && aligned16(sin_t), "pointers must be 16-byte aligned");
  const int valid = S - len_text;
  const int row0 = mode == SVGB_ROPE_TXT_LAST ? 0 : len_text;
  auto st = static_cast<cudaStream_t>(stream);
  if (mode == SVGB_ROPE_COMPLEX_TXT_FIRST)
    return dtype == SVGB_BF16 ? launch_rope<true, true>(q, k, cos_t, sin_t, B, Hq, Hk, S, D, row0, valid, st)
                              : launch_rope<false, true>(q, k, cos_t, sin_t, B, Hq, Hk, S, D, row0, valid, st);
  return dtype == SVGB_BF16 ? launch_rope<true, false>(q, k, cos_t, sin_t, B, Hq, Hk, S, D, row0, valid, st)
                            : launch_rope<false, false>(q, k, cos_t, sin_t, B, Hq, Hk, S, D, row0, valid, st);
}

int svgb_qkv_prep(const void* q_in, const void* k_in, const void* v_in, long long in_token_stride,
                  long long in_batch_stride, void* q_out, void* k_out, void* v_out, long long out_head_stride,
                  long long out_batch_stride, int B, int S_in, int H, int D, int out_row0, int norm,
                  const void* gamma_q, const void* gamma_k, const void* beta_q, const void* beta_k, float eps,
                  int rope, const float* cos_t, const float* sin_t, int rope_lo, int rope_n, int dtype,
                  void* stream) {
  SVGB_REQUIRE(q_in && k_in && v_in && q_out && k_out && v_out, "null pointer");
  SVGB_REQUIRE(B > 0 && S_in > 0 && H > 0 && out_row0 >= 0, "bad sizes");
  SVGB_REQUIRE(D == 64 || D == 128 || D == 256, "Unsupported head_dim: %d", D);
  SVGB_REQUIRE(dtype == SVGB_BF16 || dtype == SVGB_F16, "dtype %d unsupported", dtype);
  SVGB_REQUIRE(norm >= SVGB_NORM_NONE && norm <= SVGB_NORM_RMS_HIDDEN, "norm %d unknown", norm);
  SVGB_REQUIRE(norm == SVGB_NORM_NONE || (gamma_q && gamma_k), "norm needs gamma_q and gamma_k");
  SVGB_REQUIRE(norm != SVGB_NORM_LAYER || (beta_q && beta_k), "layer norm needs beta_q and beta_k");
  SVGB_REQUIRE(rope >= 0 && rope <= 2, "rope %d unknown (0 none, 1 cos/sin [n,D], 2 complex [n,D/2])", rope);
  SVGB_REQUIRE(rope == 0 || (cos_t && sin_t && rope_lo >= 0 && rope_n >= 0 && rope_lo + rope_n <= S_in),
               "rope needs tables and a token range inside [0, S_in)");
  SVGB_REQUIRE(in_token_stride % 8 == 0 && in_batch_stride % 8 == 0 && out_head_stride % 8 == 0 &&
                   out_batch_stride % 8 == 0,
               "strides must be multiples of 8 elements (16 bytes)");
  SVGB_REQUIRE(aligned16(q_in) && aligned16(k_in) && aligned16(v_in) && aligned16(q_out) && aligned16(k_out) &&
                   aligned16(v_out),
               "pointers must be 16-byte aligned");
  PrepArgs a;
  a.in[0] = static_cast<const uint16_t*>(q_in);
  a.in[1] = static_cast<const uint16_t*>(k_in);
  a.in[2] = static_cast<const uint16_t*>(v_in);
  a.out[0] = static_cast<uint16_t*>(q_out);
  a.out[1] = static_cast<uint16_t*>(k_out);
  a.out[2] = static_cast<uint16_t*>(v_out);
  a.gamma[0] = static_cast<const uint16_t*>(gamma_q);
  a.gamma[1] = static_cast<const uint16_t*>(gamma_k);
  a.beta[0] = static_cast<const uint16_t*>(beta_q);
  a.beta[1] = static_cast<const uint16_t*>(beta_k);
  a.cos_t = cos_t;
  a.sin_t = sin_t;
  a.in_token_stride = in_token_stride;
  a.in_batch_stride = in_batch_stride;
  a.out_head_stride = out_head_stride;
  a.out_batch_stride = out_batch_stride;
  a.S_in = S_in;
  a.H = H;
  a.out_row0 = out_row0;
  a.norm = norm;
  a.rope = rope;
  a.rope_lo = rope_lo;
  a.rope_n = rope_n;
  a.eps = eps;
  const int rpb = kThreads / (D / 8);
  dim3 grid((S_in + rpb - 1) / rpb, B, 3);
  auto st = static_cast<cudaStream_t>(stream);
  const bool bf = dtype == SVGB_BF16;
  switch (D) {
    case 64:
      if (bf) qkv_prep_kernel<true, 64><<<grid, kThreads, 0, st>>>(a);
      else qkv_prep_kernel<false, 64><<<grid, kThreads, 0, st>>>(a);
      break;
    case 128:
      if (bf) qkv_prep_kernel<true, 128><<<grid, kThreads, 0, st>>>(a);
      else qkv_prep_kernel<false, 128><<<grid, kThreads, 0, st>>>(a);
      break;
    default:
      if (bf) qkv_prep_kernel<true, 256><<<grid, kThreads, 0, st>>>(a);
      else qkv_prep_kernel<false, 256><<<grid, kThreads, 0, st>>>(a);
      break;
  }
  SVGB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
