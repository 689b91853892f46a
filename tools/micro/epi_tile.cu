// Microbenchmark: the conv epilogue alone (TMEM -> +bias -> SiLU -> bf16 -> swizzled smem), clocks per 128 x N tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ayolov2_b200/csrc -o tools/micro/epi_tile.bin tools/micro/epi_tile.cu
#include <cstdio>
#include <cstdlib>
#include "ay2_ptx.cuh"
using namespace ay2;

__device__ __forceinline__ void nbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// MODE bits: 1 = explicit ld.shared bias (else pointer-derived loads, generic if provenance is lost), 2 = no SiLU,
//            4 = LDTM only (no math / stores), 8 = software-pipelined LDTM (next block in flight during the math),
//            16 = no LDTM (math on registers only)
template <int MODE, int BLOCK_N, int EPI_WARPS, int SPIN_WARPS = 2>
__global__ void __launch_bounds__(128 + EPI_WARPS * 32) k(int iters, long long* clk, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem;
  float* bias_s = reinterpret_cast<float*>(staging + 128 * BLOCK_N * 2);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bias_s + BLOCK_N);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x) bias_s[i] = 0.01f * i;
  if (warp == 2) tmem_alloc(tptr, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tptr;
  uint64_t* never = reinterpret_cast<uint64_t*>(tptr + 2);
  volatile uint32_t* done = tptr + 4;
  if (threadIdx.x == 0) { mbar_init(never, 1); *done = 0; fence_barrier_init(); }
  __syncthreads();
  if ((MODE & 96) && warp < SPIN_WARPS) {  // spinning role warps, like a producer / MMA warp waiting on its mbarrier
    while (*done == 0) {
      if (MODE & 32) mbar_try_wait(never, 0);
      else {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(never)), "r"(0), "r"(1000000) : "memory");
      }
    }
  }
  constexpr int GROUPS = EPI_WARPS / 4;
  constexpr int COLS = BLOCK_N / GROUPS;
  if (warp >= 4) {
    const int eall = threadIdx.x - 128;
    const int et = eall & 127, egrp = eall >> 7, ewarp = warp & 3;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16);
    float acc = 0.f;
    nbar(1, EPI_WARPS * 32);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t v[32], w[32];
      if (MODE & 8) tmem_ld_32x32b_x32(taddr + egrp * COLS, v);
#pragma unroll 1
      for (int c0 = egrp * COLS; c0 < (egrp + 1) * COLS; c0 += 32) {
        if (MODE & 8) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = v[i];
          if (c0 + 32 < (egrp + 1) * COLS) tmem_ld_32x32b_x32(taddr + c0 + 32, v);
        } else if (MODE & 16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(acc + i);
        } else {
          tmem_ld_32x32b_x32(taddr + c0, w);
          tmem_ld_wait();
        }
        if (MODE & 4) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += __uint_as_float(w[i]);
          continue;
        }
        uint8_t* slab = staging + (c0 / 64) * (128 * 128);
        const int chunk0 = (c0 % 64) / 8;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8], bb[8];
          if (MODE & 1) {
            const uint32_t ba = smem_u32(bias_s) + (c0 + g * 8) * 4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[0]), "=f"(bb[1]), "=f"(bb[2]), "=f"(bb[3]) : "r"(ba));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[4]), "=f"(bb[5]), "=f"(bb[6]), "=f"(bb[7]) : "r"(ba + 16));
          } else {
            const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[c0 + g * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[c0 + g * 8 + 4]);
            bb[0] = b0.x, bb[1] = b0.y, bb[2] = b0.z, bb[3] = b0.w, bb[4] = b1.x, bb[5] = b1.y, bb[6] = b1.z, bb[7] = b1.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = __uint_as_float(w[g * 8 + i]) + bb[i];
            if (!(MODE & 2)) x = silu_f(x);
            f[i] = x;
          }
          const uint32_t dst = smem_u32(slab) + swizzled_offset<128>(et, chunk0 + g);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16x2(f[0], f[1])),
                       "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7])) : "memory");
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      nbar(1, EPI_WARPS * 32);
    }
    const long long t1 = clock64();
    if (eall == 0) { clk[blockIdx.x] = t1 - t0; *done = 1; }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int MODE, int BLOCK_N, int EPI_WARPS>
void run(const char* name, int ctas_per_sm) {
  long long* clk;
  float* sink;
  const int grid = 148 * ctas_per_sm;
  cudaMalloc(&clk, grid * 8);
  cudaMalloc(&sink, 4096);
  const int smem = 128 * BLOCK_N * 2 + BLOCK_N * 4 + 128 + 1024;
  cudaFuncSetAttribute(k<MODE, BLOCK_N, EPI_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 200;
  for (int r = 0; r < 2; ++r) k<MODE, BLOCK_N, EPI_WARPS><<<grid, 128 + EPI_WARPS * 32, smem>>>(iters, clk, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[148 * 4];
  cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < grid; ++i) c += h[i];
  c /= grid;
  printf("N=%3d epi warps %d CTAs/SM %d %-34s: %7.0f clk per tile per CTA -> %6.0f clk per 128x128 outputs per SM\n", BLOCK_N, EPI_WARPS,
         ctas_per_sm, name, c / iters, c / iters / ctas_per_sm * 128.0 / BLOCK_N);
  cudaFree(clk);
  cudaFree(sink);
}

int main() {
  run<1 | 32, 128, 4>("ld.shared bias + 2 spinning warps", 2);
  run<1 | 64, 128, 4>("same, try_wait with 1 ms hint", 2);
  run<1 | 32, 256, 8>("ld.shared bias + 2 spinning warps", 1);
  run<1 | 64, 256, 8>("same, try_wait with 1 ms hint", 1);
  for (int c : {1, 2}) {
    run<0, 128, 4>("as-is (pointer bias)", c);
    run<1, 128, 4>("ld.shared bias", c);
    run<3, 128, 4>("ld.shared bias, no SiLU", c);
    run<4, 128, 4>("LDTM only", c);
    run<1 | 16, 128, 4>("no LDTM, math + STS", c);
    run<1 | 8, 128, 4>("ld.shared bias, pipelined LDTM", c);
    run<1, 128, 8>("ld.shared bias, 8 warps", c);
    run<1 | 8, 128, 8>("ld.shared, pipelined, 8 warps", c);
  }
  run<0, 256, 8>("as-is (pointer bias)", 1);
  run<1, 256, 8>("ld.shared bias", 1);
  run<1 | 8, 256, 8>("ld.shared bias, pipelined LDTM", 1);
  run<0, 64, 4>("as-is (pointer bias)", 3);
  run<1, 64, 4>("ld.shared bias", 3);
  return 0;
}
