// Microbenchmark (bring-up tool): TMEM read / write throughput and MUFU.EX2 throughput per SM as a function of
// the number of warps issuing.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sparse-videogen_b200/csrc/ptx.cuh"

using namespace svgb;

// mode 0: tcgen05.ld 32x32b.x32 x4 (128 columns) + wait per iteration
// mode 1: same but wait after every x32 load
// mode 2: tcgen05.st x32 x4 + wait
// mode 3: 128 ex2.approx per thread per iteration (MUFU)
// mode 4: ld x4 + 96 MUFU + 32 poly per thread (no dependency between load and math except through registers)
__global__ void __launch_bounds__(384, 1) bw_kernel(int mode, int nwarps, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
    uint32_t r0[32], r1[32], r2[32], r3[32];
    for (int i = 0; i < 32; ++i) r0[i] = r1[i] = r2[i] = r3[i] = __float_as_uint(0.001f * (lane + i));
    tmem_st32(lane_addr, r0); tmem_st32(lane_addr + 32, r1); tmem_st32(lane_addr + 64, r2); tmem_st32(lane_addr + 96, r3);
    tc_wait_st();
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0 || mode == 4) {
        tmem_ld32(lane_addr, r0); tmem_ld32(lane_addr + 32, r1); tmem_ld32(lane_addr + 64, r2); tmem_ld32(lane_addr + 96, r3);
        tc_wait_ld();
      } else if (mode == 1) {
        tmem_ld32(lane_addr, r0); tc_wait_ld();
        tmem_ld32(lane_addr + 32, r1); tc_wait_ld();
        tmem_ld32(lane_addr + 64, r2); tc_wait_ld();
        tmem_ld32(lane_addr + 96, r3); tc_wait_ld();
      } else if (mode == 2) {
        tmem_st32(lane_addr, r0); tmem_st32(lane_addr + 32, r1); tmem_st32(lane_addr + 64, r2); tmem_st32(lane_addr + 96, r3);
        tc_wait_st();
      }
      if (mode == 3 || mode == 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          a0 += ex2_approx(__uint_as_float(r0[i]));
          a1 += ex2_approx(__uint_as_float(r1[i]));
          a2 += ex2_approx(__uint_as_float(r2[i]));
          if (mode == 3) a3 += ex2_approx(__uint_as_float(r3[i]));
          else a3 = fmaf(__uint_as_float(r3[i]), a3, 0.5f);
        }
        acc += a0 + a1 + a2 + a3;
        if (mode == 3) r0[0] = __float_as_uint(acc * 1e-9f);  // keep the loop body live across iterations
      } else {
        acc += __uint_as_float(r0[0]) + __uint_as_float(r1[7]) + __uint_as_float(r2[13]) + __uint_as_float(r3[31]);
      }
    }
    t1 = clock64();
  }
  if (lane == 0 && warp < nwarps) cycles[blockIdx.x * 12 + warp] = t1 - t0;
  sink[blockIdx.x * 384 + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 12 * sizeof(long long));
  cudaMalloc(&sink, 148 * 384 * sizeof(float));
  const int iters = 2000;
  const char* names[] = {"ld128+wait", "ld32+wait x4", "st128+wait", "mufu128", "ld128+96mufu+32fma"};
  for (int mode = 0; mode < 5; ++mode)
    for (int nw : {1, 4, 8}) {
      cudaMemset(cyc, 0, 148 * 12 * sizeof(long long));
      bw_kernel<<<148, 384>>>(mode, nw, iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[12];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
      const double per_iter = double(mx) / iters;
      // bytes moved per iteration per SM: nw warps x 32 lanes x 128 cols x 4 B
      printf("{\"mode\": \"%s\", \"warps\": %d, \"cycles_per_iter\": %.1f, \"tmem_bytes_per_cycle_per_sm\": %.1f, \"elems_per_cycle_per_sm\": %.2f}\n",
             names[mode], nw, per_iter, nw * 32 * 128 * 4 / per_iter, nw * 32 * 128 / per_iter);
    }
  return 0;
}
