// Microbenchmark: per-SM throughput of the activation math candidates for the conv epilogue (elements / clock / SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/epi_rate.bin tools/micro/epi_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

template <int MODE>
__device__ __forceinline__ float op(float x) {
  float t;
  if (MODE == 0) {  // tanh.approx.f32 SiLU
    float h = 0.5f * x;
    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  } else if (MODE == 1) {  // ex2 + rcp SiLU
    float e;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(1.0f + e));
    return x * t;
  } else if (MODE == 2) {  // ex2 only
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
  } else if (MODE == 3) {  // rcp only
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
  } else if (MODE == 4) {  // tanh only
    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
  } else if (MODE == 5) {  // pure FMA: 8 fmas
    t = x;
#pragma unroll
    for (int i = 0; i < 8; ++i) t = fmaf(t, 0.999f, 0.001f);
    return t;
  }
  return x;
}

template <int MODE>
__global__ void k(int iters, float seed, float* out, long long* clk) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + 0.01f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<MODE>(v[i]);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// packed tanh on two bf16 / f16 values per instruction
template <int MODE>
__global__ void k2(int iters, float seed, float* out, long long* clk) {
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 b = __floats2bfloat162_rn(seed + 0.01f * i, seed + 0.02f * threadIdx.x);
    v[i] = *reinterpret_cast<uint32_t*>(&b);
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(v[i]) : "r"(v[i]));
      else if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(v[i]) : "r"(v[i]));
      else asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(v[i]) : "r"(v[i]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <class F>
void run(const char* name, F launch, int threads, int per_iter_elems) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8 * 8);
  const int iters = 2000;
  launch(iters, out, clk);
  launch(iters, out, clk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  printf("%-28s threads/SM %4d : %.2f elements/clk/SM  (%.1f clk per warp-instruction group)\n", name, threads,
         (double)threads * per_iter_elems * iters / c, c / iters / 8);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  for (int threads : {128, 256, 512}) {
#define R(MODE, NAME) run(NAME, [&](int it, float* o, long long* c) { k<MODE><<<148, threads>>>(it, 0.3f, o, c); }, threads, 8)
    R(0, "silu tanh.approx.f32");
    R(1, "silu ex2+rcp");
    R(2, "ex2.approx");
    R(3, "rcp.approx");
    R(4, "tanh.approx.f32");
    R(5, "8 x fma chain");
#define R2(MODE, NAME) run(NAME, [&](int it, float* o, long long* c) { k2<MODE><<<148, threads>>>(it, 0.3f, o, c); }, threads, 16)
    R2(0, "tanh.approx.bf16x2 (2 el)");
    R2(1, "tanh.approx.f16x2 (2 el)");
    R2(2, "ex2.approx.bf16x2 (2 el)");
  }
  return 0;
}
