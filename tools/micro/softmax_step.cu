// Microbenchmark (bring-up tool): cycles of ONE softmax step of the attention kernel (thread per row, 128 key
// columns: tcgen05.ld S -> row max -> P = exp2(S*c - m*c) -> pack -> tcgen05.st P -> wait) for several ways of
// doing the exponentials, with 4 warps (one tile alone) and 8 warps (both tiles' softmax overlapping), no MMAs:
//   0  as in attn_kernel.cuh: fp32 MUFU for 3 of 4 pairs, degree-3 polynomial for the 4th, bf16 pack
//   1  all MUFU                      2  polynomial for every second pair
//   3  MUFU.EX2 on packed f16x2 (two exponentials per MUFU op), fp32 row sum, bf16 pack
//   4  as 3 with an fp16 P (the f16x2 result is stored as is)
//   5  as 0 but in two 64-column sub-steps (load, max, exp, store, wait twice)
//   6  as 0 without the row sum (cost of the 64 FADD2)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o softmax_step softmax_step.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sparse-videogen_b200/csrc/ptx.cuh"

using namespace svgb;

__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ void f16x2_to_f32(uint32_t h, float& lo, float& hi) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(h));
}

template <int V>
__device__ __forceinline__ void group_p(const uint32_t (&rr)[32], uint32_t s_addr, int g, uint64_t c2, uint64_t nmc2,
                                        uint64_t& sum2) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint64_t x2 = ffma2(pack_f32x2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1])), c2, nmc2);
    float p0, p1;
    if constexpr (V == 3 || V == 4) {
      float x0, x1;
      unpack_f32x2(x2, x0, x1);
      const uint32_t e = ex2_f16x2(cvt_f16x2(x0, x1));
      f16x2_to_f32(e, p0, p1);
      sum2 = fadd2(sum2, pack_f32x2(p0, p1));
      pk[i] = V == 4 ? e : pack2<true>(p0, p1);
    } else {
      const bool poly = V == 1 ? false : (V == 2 ? (i & 1) == 1 : (i & 3) == 3);
      if (poly) {
        ex2_poly2(x2, p0, p1);
      } else {
        float x0, x1;
        unpack_f32x2(x2, x0, x1);
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
      }
      if constexpr (V != 6) sum2 = fadd2(sum2, pack_f32x2(p0, p1));
      pk[i] = pack2<true>(p0, p1);
    }
  }
  tmem_st16(s_addr + g * 16, pk);
}

__device__ __forceinline__ float max32(const uint32_t (&rr)[32]) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(rr[i]));
  return m;
}

template <int V>
__global__ void __launch_bounds__(256, 1) step_kernel(int nwarps, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  float l_run = 0.f, m_used = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t s_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    const uint32_t keep = s_addr + 256;  // a second copy of S that the step re-reads (P overwrites the first)
    {
      uint32_t r[32];
      for (int g = 0; g < 4; ++g) {
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(-0.01f * ((lane * 7 + i * 3 + g * 11) % 97));
        tmem_st32(keep + g * 32, r);
      }
      tc_wait_st();
    }
    const float c = 0.1275f;
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
      if constexpr (V == 5) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[32], r1[32];
          tmem_ld32(keep + h * 64, r0);
          tmem_ld32(keep + h * 64 + 32, r1);
          tc_wait_ld();
          const float mx = fmaxf(max32(r0), max32(r1));
          if ((mx - m_used) * c > 8.f) m_used = mx;
          const float mc = m_used * c;
          const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
          group_p<0>(r0, s_addr, 2 * h, c2, nmc2, sum2);
          group_p<0>(r1, s_addr, 2 * h + 1, c2, nmc2, sum2);
          tc_wait_st();
        }
      } else {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32(keep, r0);
        tmem_ld32(keep + 32, r1);
        tmem_ld32(keep + 64, r2);
        tmem_ld32(keep + 96, r3);
        tc_wait_ld();
        const float mx = fmaxf(fmaxf(max32(r0), max32(r1)), fmaxf(max32(r2), max32(r3)));
        if ((mx - m_used) * c > 8.f) m_used = mx;
        const float mc = m_used * c;
        const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
        group_p<V>(r0, s_addr, 0, c2, nmc2, sum2);
        group_p<V>(r1, s_addr, 1, c2, nmc2, sum2);
        group_p<V>(r2, s_addr, 2, c2, nmc2, sum2);
        group_p<V>(r3, s_addr, 3, c2, nmc2, sum2);
        tc_wait_st();
      }
      float s0, s1;
      unpack_f32x2(sum2, s0, s1);
      l_run += s0 + s1;
    }
    t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
  }
  sink[blockIdx.x * 256 + threadIdx.x] = l_run + m_used;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int V>
void run(const char* name, long long* cyc, float* sink) {
  const int iters = 400;
  for (int nw : {4, 8}) {
    cudaMemset(cyc, 0, 148 * 8 * sizeof(long long));
    step_kernel<V><<<148, 256>>>(nw, iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("{\"variant\": \"%s\", \"softmax_warps\": %d, \"cycles_per_128x128_tile_step\": %.1f}\n", name, nw, double(mx) / iters);
  }
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  cudaMalloc(&sink, 148 * 384 * sizeof(float));
  run<0>("fp32 MUFU 75% + poly 25% (current)", cyc, sink);
  run<1>("fp32 MUFU 100%", cyc, sink);
  run<2>("fp32 MUFU 50% + poly 50%", cyc, sink);
  run<3>("f16x2 MUFU, fp32 sum, bf16 P", cyc, sink);
  run<4>("f16x2 MUFU, fp32 sum, fp16 P", cyc, sink);
  run<5>("current, two 64-column sub-steps", cyc, sink);
  run<6>("current, no row sum", cyc, sink);
  return 0;
}
