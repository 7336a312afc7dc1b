// Microbenchmark (bring-up tool): execution rate of tcgen05.mma (M=128, K=16, bf16) with a lean issue stream
// (elect.sync leader, descriptors prebuilt, 8 MMAs unrolled) per operand form / N / number of independent
// accumulators, with and without 8 warps reading TMEM + running MUFU at the same time.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sparse-videogen_b200/csrc/ptx.cuh"

using namespace svgb;

template <int FORM, int CHAINS, int N>  // FORM 0: SS (QK-like), 1: TS + MN-major B (PV-like), 2: alternate groups of 8 SS / TS
__global__ void __launch_bounds__(384, 1) mma_kernel(int noise, int n_groups, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); stop = 0; }
  if (warp == 2) tmem_alloc<512>(smem_u32(&tmem_base_s));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    if (elect_one()) {
      const uint64_t a0 = desc_kmajor_sw128(sbase), bk0 = desc_kmajor_sw128(sbase + 32768);
      const uint64_t bv0 = desc_mnmajor_sw128(sbase + 32768, 16384);
      const uint32_t id_ss = make_idesc(128, N, true, false, false), id_ts = make_idesc(128, N, true, false, true);
      long long t0 = clock64();
      for (int g = 0; g < n_groups; ++g) {
        // CHAINS == 1: the 8 MMAs of a group accumulate into one tile (a K loop); == 2: they alternate between two
        const bool ts = FORM == 1 || (FORM == 2 && (g & 1));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int acc = CHAINS == 1 ? (g & 1) : (kk & 1);
          const uint32_t d = tmem + 256 + acc * 128;
          const uint64_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if (ts) mma_ts(d, tmem + acc * 128 + kk * 8, bv0 + kk * 128, id_ts, 1u);
          else mma_ss(d, a0 + off, bk0 + off, id_ss, 1u);
        }
      }
      long long t1 = clock64();
      tc_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, 1);
      long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
      stop = 1;
    }
  } else if (warp >= 4 && noise) {
    const int lane = threadIdx.x & 31;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 3) & 1) * 128;
    uint32_t r0[32], r1[32], r2[32], r3[32];
    float acc = 0.f;
    while (!stop) {
      tmem_ld32(lane_addr, r0); tmem_ld32(lane_addr + 32, r1); tmem_ld32(lane_addr + 64, r2); tmem_ld32(lane_addr + 96, r3);
      tc_wait_ld();
      if (noise > 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += ex2_approx(__uint_as_float(r0[i])) + ex2_approx(__uint_as_float(r1[i])) + ex2_approx(__uint_as_float(r2[i]));
      }
      acc += __uint_as_float(r0[0]) + __uint_as_float(r3[31]);
    }
    if (acc == 123.f) out[1000 + lane] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int FORM, int CHAINS, int N>
void run(const char* name, long long* out) {
  cudaFuncSetAttribute(mma_kernel<FORM, CHAINS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int groups = 512;
  for (int noise : {0, 2}) {
    mma_kernel<FORM, CHAINS, N><<<148, 384, 100 * 1024>>>(noise, groups, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[2];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"form\": \"%s\", \"N\": %d, \"chains\": %d, \"softmax_like_noise\": %d, \"cycles_per_mma_issue\": %.1f, \"cycles_per_mma\": %.1f, \"nominal\": %d}\n",
           name, N, CHAINS, noise, double(h[0]) / (groups * 8), double(h[1]) / (groups * 8), N / 2);
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 4096 * sizeof(long long));
  run<0, 1, 128>("SS K-major x K-major (QK)", out);
  run<1, 1, 128>("TS TMEM x MN-major (PV)", out);
  run<2, 1, 128>("alternating groups QK / PV", out);
  run<0, 2, 128>("SS K-major x K-major (QK)", out);
  run<1, 2, 128>("TS TMEM x MN-major (PV)", out);
  run<0, 1, 64>("SS K-major x K-major (QK)", out);
  run<1, 1, 64>("TS TMEM x MN-major (PV)", out);
  return 0;
}
