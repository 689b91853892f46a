// Microbenchmark: tcgen05.mma (kind::f16, M=128, K=16) issue/execute rate vs N, operand layout and accumulator count.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ayolov2_b200/csrc -o /tmp/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "ay2_ptx.cuh"
using namespace ay2;

__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi),
      "r"(idesc), "r"(accumulate) : "memory");
}

// mode 0: A SW128 K-major; mode 1: A no-swizzle, SBO=128 (dense groups), LBO=2048; mode 2: A no-swizzle SBO=288, LBO=5184 (chain T)
__global__ void k(int N, int mode, int nacc, int nmma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&tptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tptr;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(128, N);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
    uint32_t a_lo, a_hi, kstep;
    if (mode == 0) { a_lo = (a_addr >> 4) | (1u << 16); a_hi = (1024u >> 4) | (1u << 14) | (2u << 29); kstep = 2; }
    else if (mode == 1) { a_lo = (a_addr >> 4) | ((2048u >> 4) << 16); a_hi = (128u >> 4) | (1u << 14); kstep = 256; }
    else { a_lo = (a_addr >> 4) | ((5184u >> 4) << 16); a_hi = (288u >> 4) | (1u << 14); kstep = 648; }
    const uint32_t b_lo = (b_addr >> 4) | (1u << 16), b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < nmma; ++i) {
        const int kk = i & 3;
        umma_ss(tm + (i % nacc) * N, a_lo + kk * kstep, a_hi, b_lo + kk * 2, b_hi, idesc, i >= nacc ? 1u : 0u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int nmma = 256;
  for (int mode = 0; mode < 3; ++mode)
    for (int N : {32, 64, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * N > 512) continue;
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          k<<<148, 64, 200 * 1024>>>(N, mode, nacc, nmma, d);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %d N %3d nacc %d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (math floor %d)\n", mode, N, nacc,
               (double)h[0] / nmma, (double)h[1] / nmma, N / 2);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
