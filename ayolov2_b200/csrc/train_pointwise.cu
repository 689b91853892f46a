// Training-mode pointwise / reduction kernels over NHWC bf16 activations:
//   * BatchNorm2d with batch statistics (+ SiLU): statistics, normalise+activate, and the backward pass
//     (kindle Conv in train mode = conv -> nn.BatchNorm2d(eps 1e-3, momentum 0.03) -> SiLU; reference call site
//     scripts/train/yolo_trainer.py:322-329: `pred = model(imgs)` ... `scaler.scale(loss).backward()`)
//   * backward of the data-movement operators (nearest-2x upsample, SPPF/SPP max pools, residual / concat adds)
//   * gradient layout conversion for the YOLOHead and the fused SGD-nesterov + EMA update
// Every kernel is HBM-bound: one 16-byte access per 8 channels, fp32 math, per-channel sums reduced in registers,
// then shared memory, then one double-precision atomic per (block, channel).
#include "ay2_common.h"
#include "ay2_ptx.cuh"

namespace ay2 {

__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
constexpr int kUnroll = 4;  // pixels per thread per loop trip of the streaming kernels below
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
// sigmoid(x) = 0.5 + 0.5 * tanh(x/2): ONE MUFU op (tanh.approx, |rel err| <= 2^-11, below the bf16 storage of every
// tensor it feeds). exp2 + rcp would be two, and these kernels are MUFU-bound: 2 passes x 5e9 elements per step.
__device__ __forceinline__ float sigmoid_acc(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
// forward activations keep the 2-MUFU form (~2 ulp): a random-init train-mode network amplifies forward perturbations
// chaotically (DESIGN.md §training parity), the forward kernel is not MUFU-bound, the two backward passes are
__device__ __forceinline__ float sigmoid_fwd(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Reduce `nval` per-thread partials (8 channels each) over the pixel lanes of a block and add them to `dst`
// (double, [C]) with one atomic per channel per block. smem: [lanes][tpp*8] floats.
template <int NVAL>
__device__ __forceinline__ void block_channel_reduce(float (*acc)[8], int tpp, int lanes, int cg, int pl, bool active,
                                                     float* sm, double* const* dst, int C) {
  for (int v = 0; v < NVAL; ++v) {
    __syncthreads();
    if (active)
      for (int j = 0; j < 8; ++j) sm[(pl * tpp + cg) * 8 + j] = acc[v][j];
    __syncthreads();
    for (int ch = threadIdx.x; ch < tpp * 8; ch += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += sm[l * tpp * 8 + ch];
      if (ch < C) atomicAdd(&dst[v][ch], (double)s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Packed fp32 math (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per instruction). The streaming BatchNorm kernels below
// were bound by instruction issue, not HBM: ~1e12 elements/s whatever the byte count (tools/prof_trainops.py: the backward
// reduce pass ran at 3.5 TB/s next to an apply pass at 6.0 TB/s with the same math and 1.5x the bytes). A thread's 8
// channels are 4 float2 pairs; every per-element multiply/add below is one packed instruction per PAIR.
// ------------------------------------------------------------------------------------------------
struct F8 {
  float2 v[4];
};
__device__ __forceinline__ F8 unpack8p(const uint4& q) {
  F8 r;
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
  return r;
}
__device__ __forceinline__ uint4 pack8p(const F8& f) {
  uint4 o;
  o.x = pack_bf16x2(f.v[0].x, f.v[0].y);
  o.y = pack_bf16x2(f.v[1].x, f.v[1].y);
  o.z = pack_bf16x2(f.v[2].x, f.v[2].y);
  o.w = pack_bf16x2(f.v[3].x, f.v[3].y);
  return o;
}
// per-channel constants of a thread's 8 channels as 4 pairs (two 16-byte loads)
__device__ __forceinline__ F8 load8p(const float* __restrict__ p) {
  F8 r;
  if (reinterpret_cast<uintptr_t>(p) & 15) {  // a parameter that is a view at an odd offset of a flat buffer
#pragma unroll
    for (int j = 0; j < 4; ++j) r.v[j] = make_float2(p[2 * j], p[2 * j + 1]);
    return r;
  }
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = make_float2(a.x, a.y);
  r.v[1] = make_float2(a.z, a.w);
  r.v[2] = make_float2(b.x, b.y);
  r.v[3] = make_float2(b.z, b.w);
  return r;
}
__device__ __forceinline__ float exp2f_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// d * silu'(u) for a pair, from h = u / 2: t = tanh(h), sigmoid = 0.5 + 0.5 t, sigmoid (1 - sigmoid) = (1 - t^2) / 4, so
// silu'(u) = sigmoid + u sigmoid (1 - sigmoid) = sigmoid + 0.5 h (1 - t^2). One MUFU per element, 5 packed instructions per pair.
__device__ __forceinline__ float2 silu_grad_times(float2 d, float2 h) {
  const float2 t = make_float2(tanh_approx(h.x), tanh_approx(h.y));
  const float2 half2 = make_float2(0.5f, 0.5f), one2 = make_float2(1.0f, 1.0f);
  const float2 sg = __ffma2_rn(half2, t, half2);
  const float2 q = __ffma2_rn(make_float2(-t.x, -t.y), t, one2);
  const float2 w = __fmul2_rn(h, q);
  return __fmul2_rn(d, __ffma2_rn(half2, w, sg));
}

// ------------------------------------------------------------------------------------------------
// per-channel sum / sum of squares of z [npix][cstride] (first C channels)
// ------------------------------------------------------------------------------------------------
__global__ void bn_stats_kernel(const __nv_bfloat16* __restrict__ z, long long npix, int C, int cs, double* sum,
                                double* sumsq) {
  extern __shared__ float sm[];
  const int tpp = C >> 3;
  const int lanes = blockDim.x / tpp;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const bool active = pl < lanes;
  float2 a0[4], a1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) a0[j] = a1[j] = make_float2(0.f, 0.f);
  if (active) {
    // kUnroll independent 16-byte loads per thread in flight (bandwidth x latency / 148 SMs ~ 45 KB per SM)
    const long long step = (long long)gridDim.x * lanes;
    long long p = (long long)blockIdx.x * lanes + pl;
    const __nv_bfloat16* zp = z + cg * 8;
    auto one = [&](const uint4& q) {
      const F8 f = unpack8p(q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a0[j] = __fadd2_rn(a0[j], f.v[j]);
        a1[j] = __ffma2_rn(f.v[j], f.v[j], a1[j]);
      }
    };
    for (; p + (kUnroll - 1) * step < npix; p += kUnroll * step) {
      uint4 q[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) q[u] = ld_stream(zp + (p + u * step) * cs);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) one(q[u]);
    }
    for (; p < npix; p += step) one(ld_stream(zp + p * cs));
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[0][2 * j] = a0[j].x, acc[0][2 * j + 1] = a0[j].y;
    acc[1][2 * j] = a1[j].x, acc[1][2 * j + 1] = a1[j].y;
  }
  double* dst[2] = {sum, sumsq};
  block_channel_reduce<2>(acc, tpp, lanes, cg, pl, active, sm, dst, C);
}

// mean / invstd from the sums (biased variance for normalisation, unbiased for the running estimate,
// exactly nn.BatchNorm2d in training mode) + running-statistics update with `momentum`.
__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, long long n, int C, float eps, float momentum,
                                   float* running_mean, float* running_var, float* mean, float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sum[c] / (double)n;
  double var = sumsq[c] / (double)n - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// y = act(gamma * (z - mean) * invstd + beta) (+ residual)
// Thread (cg, pl) owns channel group cg for the pixels pl, pl + lanes*grid, ...: u = z*A + B with A, B in registers.
__global__ void bn_act_fwd_kernel(const __nv_bfloat16* __restrict__ z, long long npix, int C, int zcs,
                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                  __nv_bfloat16* __restrict__ y, int ycs, const __nv_bfloat16* __restrict__ res, int rcs) {
  const int tpp = C >> 3;
  const int lanes = blockDim.x / tpp;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  if (pl >= lanes) return;
  F8 A, Bc;
  {
    const F8 ga = load8p(gamma + cg * 8), be = load8p(beta + cg * 8), mu = load8p(mean + cg * 8), is = load8p(invstd + cg * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      A.v[j] = __fmul2_rn(ga.v[j], is.v[j]);
      Bc.v[j] = make_float2(be.v[j].x - A.v[j].x * mu.v[j].x, be.v[j].y - A.v[j].y * mu.v[j].y);
    }
  }
  const bool silu = act == AY2_ACT_SILU;
  const long long step = (long long)gridDim.x * lanes;
  long long p = (long long)blockIdx.x * lanes + pl;
  auto one = [&](const uint4& qz, const uint4& qr, long long pp) {
    F8 f = unpack8p(qz);
    const F8 r = unpack8p(qr);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 u = __ffma2_rn(f.v[j], A.v[j], Bc.v[j]);
      if (silu) {
        // u * sigmoid(u), sigmoid = 1 / (1 + 2^(-u log2 e)): the 2-MUFU form (~2 ulp), see sigmoid_fwd
        const float2 e2 = __fmul2_rn(u, make_float2(-1.4426950408889634f, -1.4426950408889634f));
        const float2 den = __fadd2_rn(make_float2(exp2f_approx(e2.x), exp2f_approx(e2.y)), make_float2(1.0f, 1.0f));
        u = __fmul2_rn(u, make_float2(rcp_approx(den.x), rcp_approx(den.y)));
      }
      if (res) u = __fadd2_rn(u, r.v[j]);
      f.v[j] = u;
    }
    *reinterpret_cast<uint4*>(y + pp * ycs + cg * 8) = pack8p(f);
  };
  for (; p + (kUnroll - 1) * step < npix; p += kUnroll * step) {
    uint4 qz[kUnroll], qr[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      qz[u] = ld_stream(z + (p + u * step) * zcs + cg * 8);
      qr[u] = res ? ld_stream(res + (p + u * step) * rcs + cg * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) one(qz[u], qr[u], p + u * step);
  }
  for (; p < npix; p += step)
    one(ld_stream(z + p * zcs + cg * 8), res ? ld_stream(res + p * rcs + cg * 8) : make_uint4(0, 0, 0, 0), p);
}

// backward, pass 1: s1[c] = sum dyh, s2[c] = sum dyh * xhat, with dyh = dy * act'(u), u = gamma*xhat + beta
__global__ void bn_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, int dcs, const __nv_bfloat16* __restrict__ z,
                                         int zcs, long long npix, int C, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, int act, double* s1, double* s2) {
  extern __shared__ float sm[];
  const int tpp = C >> 3;
  const int lanes = blockDim.x / tpp;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const bool active = pl < lanes;
  float2 a0[4], a1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) a0[j] = a1[j] = make_float2(0.f, 0.f);
  if (active) {
    F8 is, nms, gah, beh;  // xhat = z*is + nms ; h = u/2 = gah*xhat + beh
    {
      const F8 ga = load8p(gamma + cg * 8), be = load8p(beta + cg * 8), mu = load8p(mean + cg * 8);
      is = load8p(invstd + cg * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        nms.v[j] = make_float2(-mu.v[j].x * is.v[j].x, -mu.v[j].y * is.v[j].y);
        gah.v[j] = make_float2(0.5f * ga.v[j].x, 0.5f * ga.v[j].y);
        beh.v[j] = make_float2(0.5f * be.v[j].x, 0.5f * be.v[j].y);
      }
    }
    const bool silu = act == AY2_ACT_SILU;
    auto one = [&](const uint4& qg, const uint4& qz) {
      const F8 g = unpack8p(qg), f = unpack8p(qz);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xh = __ffma2_rn(f.v[j], is.v[j], nms.v[j]);
        float2 d = g.v[j];
        if (silu) d = silu_grad_times(d, __ffma2_rn(gah.v[j], xh, beh.v[j]));
        a0[j] = __fadd2_rn(a0[j], d);
        a1[j] = __ffma2_rn(d, xh, a1[j]);
      }
    };
    const long long step = (long long)gridDim.x * lanes;
    long long p = (long long)blockIdx.x * lanes + pl;
    for (; p + (kUnroll - 1) * step < npix; p += kUnroll * step) {
      uint4 qg[kUnroll], qz[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        qg[u] = ld_stream(dy + (p + u * step) * dcs + cg * 8);
        qz[u] = ld_stream(z + (p + u * step) * zcs + cg * 8);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) one(qg[u], qz[u]);
    }
    for (; p < npix; p += step) one(ld_stream(dy + p * dcs + cg * 8), ld_stream(z + p * zcs + cg * 8));
  }
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[0][2 * j] = a0[j].x, acc[0][2 * j + 1] = a0[j].y;
    acc[1][2 * j] = a1[j].x, acc[1][2 * j + 1] = a1[j].y;
  }
  double* dst[2] = {s1, s2};
  block_channel_reduce<2>(acc, tpp, lanes, cg, pl, active, sm, dst, C);
}

// backward, pass 2: dz = gamma * invstd * (dyh - s1/N - xhat * s2/N)
__global__ void bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dcs, const __nv_bfloat16* __restrict__ z,
                                        int zcs, long long npix, long long npix_norm, int C, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, int act, const double* __restrict__ s1,
                                        const double* __restrict__ s2, __nv_bfloat16* __restrict__ dz, int zdcs,
                                        float* __restrict__ dbeta_acc, float* __restrict__ dgamma_acc) {
  const int tpp = C >> 3;
  const int lanes = blockDim.x / tpp;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  if (pl >= lanes) return;
  if (dbeta_acc && blockIdx.x == 0 && pl == 0) {
    // the parameter gradients are the two sums themselves: d beta = s1, d gamma = s2 -- accumulated into the caller's fp32
    // (flat) gradient here instead of two conversion + two add launches per layer on the host side
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      dbeta_acc[c] += (float)s1[c];
      dgamma_acc[c] += (float)s2[c];
    }
  }
  const float invn = 1.0f / (float)npix_norm;  // pixels of the WHOLE (possibly cross-rank, SyncBatchNorm) batch
  F8 is, nms, gah, beh, k0, nk1, nk2;  // dz = k0*dyh + nk1 + xhat*nk2
  {
    const F8 ga = load8p(gamma + cg * 8), be = load8p(beta + cg * 8), mu = load8p(mean + cg * 8);
    is = load8p(invstd + cg * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = cg * 8 + 2 * j;
      nms.v[j] = make_float2(-mu.v[j].x * is.v[j].x, -mu.v[j].y * is.v[j].y);
      gah.v[j] = make_float2(0.5f * ga.v[j].x, 0.5f * ga.v[j].y);
      beh.v[j] = make_float2(0.5f * be.v[j].x, 0.5f * be.v[j].y);
      k0.v[j] = __fmul2_rn(ga.v[j], is.v[j]);
      nk1.v[j] = make_float2(-k0.v[j].x * (float)s1[c] * invn, -k0.v[j].y * (float)s1[c + 1] * invn);
      nk2.v[j] = make_float2(-k0.v[j].x * (float)s2[c] * invn, -k0.v[j].y * (float)s2[c + 1] * invn);
    }
  }
  const bool silu = act == AY2_ACT_SILU;
  auto one = [&](const uint4& qg, const uint4& qz, long long pp) {
    F8 g = unpack8p(qg);
    const F8 f = unpack8p(qz);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 xh = __ffma2_rn(f.v[j], is.v[j], nms.v[j]);
      float2 d = g.v[j];
      if (silu) d = silu_grad_times(d, __ffma2_rn(gah.v[j], xh, beh.v[j]));
      g.v[j] = __ffma2_rn(d, k0.v[j], __ffma2_rn(xh, nk2.v[j], nk1.v[j]));
    }
    *reinterpret_cast<uint4*>(dz + pp * zdcs + cg * 8) = pack8p(g);
  };
  const long long step = (long long)gridDim.x * lanes;
  long long p = (long long)blockIdx.x * lanes + pl;
  for (; p + (kUnroll - 1) * step < npix; p += kUnroll * step) {
    uint4 qg[kUnroll], qz[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      qg[u] = ld_stream(dy + (p + u * step) * dcs + cg * 8);
      qz[u] = ld_stream(z + (p + u * step) * zcs + cg * 8);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) one(qg[u], qz[u], p + u * step);
  }
  for (; p < npix; p += step) one(ld_stream(dy + p * dcs + cg * 8), ld_stream(z + p * zcs + cg * 8), p);
}

// ------------------------------------------------------------------------------------------------
// dst (+)= src over a channel slice (residual / concat / fan-out gradient accumulation)
// ------------------------------------------------------------------------------------------------
__global__ void add_slices_kernel(const __nv_bfloat16* __restrict__ src, int scs, __nv_bfloat16* __restrict__ dst, int dcs,
                                  long long npix, int C, int accumulate) {
  const int tpp = C >> 3;
  const long long total = npix * tpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % tpp);
    const long long p = i / tpp;
    float a[8];
    unpack8(*reinterpret_cast<const uint4*>(src + p * scs + cg * 8), a);
    if (accumulate) {
      float b[8];
      unpack8(*reinterpret_cast<const uint4*>(dst + p * dcs + cg * 8), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
    *reinterpret_cast<uint4*>(dst + p * dcs + cg * 8) = pack8(a);
  }
}

// nearest-2x upsample backward: dx[y][x] (+)= sum of the 2x2 block of dy
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dcs, int B, int H, int W, int C,
                                      __nv_bfloat16* __restrict__ dx, int xcs, int accumulate) {
  const int tpp = C >> 3;
  const long long total = (long long)B * H * W * tpp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % tpp);
    long long p = i / tpp;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
    for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
      for (int dxx = 0; dxx < 2; ++dxx) {
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(dy + ((((long long)b * 2 * H + 2 * y + dyy) * 2 * W) + 2 * x + dxx) * dcs + cg * 8), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += g[j];
      }
    if (accumulate) {
      float o[8];
      unpack8(*reinterpret_cast<const uint4*>(dx + p * xcs + cg * 8), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += o[j];
    }
    *reinterpret_cast<uint4*>(dx + p * xcs + cg * 8) = pack8(a);
  }
}

// MaxPool2d(k, stride 1, pad k/2) backward in gather form: dx[p] (+)= sum over the windows q that contain p of
// dy[q] * [argmax(q) == p], with nn.MaxPool2d's tie rule (first maximum in row-major window order).
// One thread per (pixel, channel); channels-last so neighbouring threads touch neighbouring bf16.
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, int xcs, const __nv_bfloat16* __restrict__ dy, int dcs,
                                   int B, int H, int W, int C, int k, __nv_bfloat16* __restrict__ dx, int gcs, int accumulate) {
  const long long total = (long long)B * H * W * C;
  const int r = k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int px = (int)(p % W), py = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
    const __nv_bfloat16* xb = x + (long long)b * H * W * xcs + c;
    float acc = 0.f;
    for (int qy = max(py - r, 0); qy <= min(py + r, H - 1); ++qy)
      for (int qx = max(px - r, 0); qx <= min(px + r, W - 1); ++qx) {
        // argmax of the window centred at (qy, qx)
        float best = -INFINITY;
        int by = -1, bx = -1;
        for (int wy = max(qy - r, 0); wy <= min(qy + r, H - 1); ++wy)
          for (int wx = max(qx - r, 0); wx <= min(qx + r, W - 1); ++wx) {
            const float v = __bfloat162float(xb[((long long)wy * W + wx) * xcs]);
            if (v > best) {
              best = v;
              by = wy;
              bx = wx;
            }
          }
        if (by == py && bx == px) acc += __bfloat162float(dy[(((long long)b * H + qy) * W + qx) * dcs + c]);
      }
    __nv_bfloat16* o = dx + p * gcs + c;
    if (accumulate) acc += __bfloat162float(*o);
    *o = __float2bfloat16_rn(acc);
  }
}

// The same, for planes that fit shared memory (the SPP/SPPF maps: 20x20 at 640 px): one block per (image, group of 8
// channels) stages x and dy, finds every window's argmax ONCE and separably -- a row pass (first maximum of the 2r+1 row
// neighbours) then a column pass over the row maxima (first row that holds the maximum): 2k reads instead of k*k, and the
// same element nn.MaxPool2d picks (first maximum in row-major window order) -- and scatters dy to the winner with shared-
// memory fp32 adds (sums of <= k*k bf16 values are exact in fp32, so the order of the adds does not matter).
__global__ void maxpool_bwd_tile_kernel(const __nv_bfloat16* __restrict__ x, int xcs, const __nv_bfloat16* __restrict__ dy,
                                        int dcs, int H, int W, int C, int k, __nv_bfloat16* __restrict__ dx, int gcs,
                                        int accumulate) {
  extern __shared__ __align__(16) uint8_t tile_sm[];
  const int HW = H * W;
  uint4* xs = reinterpret_cast<uint4*>(tile_sm);                        // [HW] 8 channels of x
  uint4* ds = xs + HW;                                                  // [HW] 8 channels of dy
  uint4* rm = ds + HW;                                                  // [HW] 8 row maxima (bf16)
  float* acc = reinterpret_cast<float*>(rm + HW);                       // [HW][8] gathered gradient
  unsigned short* ra = reinterpret_cast<unsigned short*>(acc + HW * 8); // [HW][8] column of the row maximum
  const int groups = C >> 3;
  const int b = blockIdx.x / groups, cg = blockIdx.x - b * groups;
  const long long pix0 = (long long)b * HW;
  for (int q = threadIdx.x; q < HW; q += blockDim.x) {
    xs[q] = *reinterpret_cast<const uint4*>(x + (pix0 + q) * xcs + cg * 8);
    ds[q] = *reinterpret_cast<const uint4*>(dy + (pix0 + q) * dcs + cg * 8);
  }
  for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int r = k / 2;
  const __nv_bfloat16* xe = reinterpret_cast<const __nv_bfloat16*>(xs);
  const __nv_bfloat16* de = reinterpret_cast<const __nv_bfloat16*>(ds);
  __nv_bfloat16* re = reinterpret_cast<__nv_bfloat16*>(rm);
  for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) {  // row pass
    const int j = i & 7, q = i >> 3;
    const int qy = q / W, qx = q - qy * W;
    float best = -INFINITY;
    int bx = 0xffff;
    for (int wx = max(qx - r, 0); wx <= min(qx + r, W - 1); ++wx) {
      const float v = __bfloat162float(xe[(qy * W + wx) * 8 + j]);
      if (v > best) {
        best = v;
        bx = wx;
      }
    }
    re[i] = __float2bfloat16_rn(best);  // exact: the maximum IS one of the bf16 inputs (-inf if the row is all NaN)
    ra[i] = static_cast<unsigned short>(bx);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) {  // column pass + scatter
    const int j = i & 7, q = i >> 3;
    const int qy = q / W, qx = q - qy * W;
    float best = -INFINITY;
    int bi = -1;
    for (int wy = max(qy - r, 0); wy <= min(qy + r, H - 1); ++wy) {
      const int w = (wy * W + qx) * 8 + j;
      const float v = __bfloat162float(re[w]);
      if (v > best) {
        best = v;
        bi = wy * W + ra[w];
      }
    }
    if (bi >= 0) atomicAdd(&acc[bi * 8 + j], __bfloat162float(de[i]));
  }
  __syncthreads();
  for (int q = threadIdx.x; q < HW; q += blockDim.x) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = acc[q * 8 + j];
    __nv_bfloat16* o = dx + (pix0 + q) * gcs + cg * 8;
    if (accumulate) {
      float old[8];
      unpack8(*reinterpret_cast<const uint4*>(o), old);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += old[j];
    }
    *reinterpret_cast<uint4*>(o) = pack8(a);
  }
}

// One warp per pixel: lane l handles the channel pairs (2l + 64j, 2l + 64j + 1), so that the fp32 side is touched in
// runs of consecutive floats (one anchor's `no` values are contiguous there) and the bf16 NHWC side in 128-byte lines.
__global__ void head_grad_to_nhwc_kernel(const float* __restrict__ g, int npix, int na, int nynx, int no,
                                         __nv_bfloat16* __restrict__ out, int cs) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int used = na * no;
  for (int pix = blockIdx.x * wpb + (threadIdx.x >> 5); pix < npix; pix += gridDim.x * wpb) {
    const int b = pix / nynx, yx = pix - b * nynx;
    const float* gb = g + (size_t)b * na * nynx * no + (size_t)yx * no;
    __nv_bfloat16* o = out + (size_t)pix * cs;
    for (int ch = 2 * lane; ch < cs; ch += 64) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = ch + e;
        v[e] = 0.f;
        if (c < used) {
          const int a = c / no;
          v[e] = gb[(size_t)a * nynx * no + (c - a * no)];
        }
      }
      *reinterpret_cast<__nv_bfloat162*>(o + ch) = __floats2bfloat162_rn(v[0], v[1]);
    }
  }
}
__global__ void head_logits_to_train_kernel(const __nv_bfloat16* __restrict__ logits, int cs, int npix, int na, int nynx,
                                            int no, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int used = na * no;
  for (int pix = blockIdx.x * wpb + (threadIdx.x >> 5); pix < npix; pix += gridDim.x * wpb) {
    const int b = pix / nynx, yx = pix - b * nynx;
    float* ob = out + (size_t)b * na * nynx * no + (size_t)yx * no;
    const __nv_bfloat16* in = logits + (size_t)pix * cs;
    for (int ch = 2 * lane; ch < used; ch += 64) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(in + ch));
      const int a0 = ch / no;
      ob[(size_t)a0 * nynx * no + (ch - a0 * no)] = v.x;
      if (ch + 1 < used) {
        const int a1 = (ch + 1) / no;
        ob[(size_t)a1 * nynx * no + (ch + 1 - a1 * no)] = v.y;
      }
    }
  }
}

// per-channel sum over pixels of a bf16 NHWC tensor (bias gradient of the head convs)
__global__ void channel_sum_kernel(const __nv_bfloat16* __restrict__ g, long long npix, int C, int cs, double* sum) {
  extern __shared__ float sm[];
  const int tpp = C >> 3;
  const int lanes = blockDim.x / tpp;
  const int cg = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  const bool active = pl < lanes;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  if (active)
    for (long long p = (long long)blockIdx.x * lanes + pl; p < npix; p += (long long)gridDim.x * lanes) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(g + p * cs + cg * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] += f[j];
    }
  double* dst[1] = {sum};
  block_channel_reduce<1>(acc, tpp, lanes, cg, pl, active, sm, dst, C);
}

// Fused SGD (nesterov momentum, weight decay) + EMA over one flat fp32 parameter range
// (scripts/train/yolo_trainer.py:332-338 optimizer step + ema.update, scripts/utils/torch_utils.py:405-416).
__global__ void sgd_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, float* ema,
                               long long n, float lr, float momentum, float wd, int nesterov, float ema_decay,
                               const float* __restrict__ inv_scale) {
  const float is = inv_scale ? *inv_scale : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = g[i] * is + wd * p[i];
    const float b = momentum * mom[i] + d;  // torch.optim.SGD: buf = momentum*buf + d_p (dampening 0)
    mom[i] = b;
    d = nesterov ? d + momentum * b : b;
    const float w = p[i] - lr * d;
    p[i] = w;
    if (ema) ema[i] = ema_decay * ema[i] + (1.0f - ema_decay) * w;
  }
}

// The same update with the reference's three parameter groups (yolo_trainer.py:149-168: BatchNorm weights, other weights
// with weight decay, biases) resolved per element: group[i] selects the learning rate / weight decay of element i, so the
// whole model -- whatever the interleaving of the groups in memory -- is ONE launch, also during warm-up when the bias
// group follows its own learning-rate ramp (yolo_trainer.py:194-221).
struct SgdGroups {
  float lr[4], wd[4];
};
__global__ void sgd_ema_groups_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom, float* ema,
                                      const uint8_t* __restrict__ group, long long n, SgdGroups gp, float momentum, int nesterov,
                                      float ema_decay, float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = group[i] & 3;
    const float w0 = p[i];
    float d = g[i] * grad_scale + gp.wd[k] * w0;
    const float b = momentum * mom[i] + d;
    mom[i] = b;
    d = nesterov ? d + momentum * b : b;
    const float w = w0 - gp.lr[k] * d;
    p[i] = w;
    if (ema) ema[i] = ema_decay * ema[i] + (1.0f - ema_decay) * w;
  }
}


// ------------------------------------------------------------------------------------------------
// Weight re-packing for the training engine: every packed bf16 operand (forward [Cout_pad][KH*KW*Cin], the flipped /
// transposed / parity-split data-gradient weights) is a pure index permutation of its fp32 OIHW parameter, padded with
// zeros. `ay2_repack_weights` applies ALL of them in one launch from a device table of segments
// {src parameter, dst operand, int32 gather index per dst element (-1 = zero), first global element}: after an optimizer
// step the ~1,000 small permute / cast / copy kernels of the per-layer refresh become one.
// ------------------------------------------------------------------------------------------------
struct RepackSeg {
  const float* src;
  __nv_bfloat16* dst;
  const int32_t* idx;
  long long begin;  // first element of this segment in the concatenated destination space (multiple of 8)
};
__global__ void repack_weights_kernel(const RepackSeg* __restrict__ segs, int nseg, long long total8) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total8; g += (long long)gridDim.x * blockDim.x) {
    const long long e = g * 8;
    int lo = 0, hi = nseg - 1;  // last segment whose begin <= e
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (segs[mid].begin <= e) lo = mid;
      else hi = mid - 1;
    }
    const RepackSeg sg = segs[lo];
    const long long off = e - sg.begin;
    const int4 i0 = *reinterpret_cast<const int4*>(sg.idx + off), i1 = *reinterpret_cast<const int4*>(sg.idx + off + 4);
    const int ix[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ix[j] >= 0 ? sg.src[ix[j]] : 0.f;
    *reinterpret_cast<uint4*>(sg.dst + off) = pack8(v);
  }
}

static int ew_grid(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace ay2

using namespace ay2;

#define AY2_ST static_cast<cudaStream_t>(stream)
#define AY2_BF(p) static_cast<__nv_bfloat16*>(p)
#define AY2_CBF(p) static_cast<const __nv_bfloat16*>(p)

// Grid of a grid-stride streaming kernel: ONE wave of resident blocks (occupancy x SMs), never more blocks than pixel
// groups. Every block pays a prologue (per-channel constants) and, in the reducing kernels, C double-precision atomics:
// with the former 8-16 blocks per SM the 20x20 / 40x40 layers spent more time there than streaming (26 MB in 35-65 us).
template <typename K>
static int wave_grid(K kernel, int threads, size_t smem, long long groups) {
  static int per_sm = 0, sms = 0;  // one instantiation (and one cache) per kernel type/pointer
  static const void* cached = nullptr;
  if (cached != reinterpret_cast<const void*>(kernel)) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    cached = reinterpret_cast<const void*>(kernel);
  }
  const long long cap = (long long)per_sm * sms;
  return (int)(groups < 1 ? 1 : (groups < cap ? groups : cap));
}

static int red_cfg(int c, int* threads, int* lanes, size_t* smem) {
  const int tpp = c / 8;
  if (c % 8 != 0 || tpp < 1 || tpp > 256) return -1;
  *threads = 256;
  *lanes = 256 / tpp;
  *smem = (size_t)(*lanes) * tpp * 8 * sizeof(float);
  return 0;
}

extern "C" int ay2_bn_stats(const void* z, int64_t npix, int32_t c, int32_t cstride, double* sum, double* sumsq, void* stream) {
  AY2_REQUIRE(z && sum && sumsq, "ay2_bn_stats: null pointer");
  int threads, lanes;
  size_t smem;
  AY2_REQUIRE(red_cfg(c, &threads, &lanes, &smem) == 0, "ay2_bn_stats: channels=%d unsupported", c);
  const int blocks = wave_grid(bn_stats_kernel, threads, smem, (npix + lanes - 1) / lanes);
  bn_stats_kernel<<<blocks, threads, smem, AY2_ST>>>(AY2_CBF(z), npix, c, cstride, sum, sumsq);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_bn_finalize(const double* sum, const double* sumsq, int64_t n, int32_t c, float eps, float momentum,
                               float* running_mean, float* running_var, float* mean, float* invstd, void* stream) {
  AY2_REQUIRE(sum && sumsq && mean && invstd, "ay2_bn_finalize: null pointer");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, AY2_ST>>>(sum, sumsq, n, c, eps, momentum, running_mean, running_var, mean,
                                                          invstd);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_bn_act_fwd(const void* z, int64_t npix, int32_t c, int32_t z_cstride, const float* mean,
                              const float* invstd, const float* gamma, const float* beta, int32_t act, void* y,
                              int32_t y_cstride, const void* residual, int32_t res_cstride, void* stream) {
  AY2_REQUIRE(z && y && mean && invstd && gamma && beta, "ay2_bn_act_fwd: null pointer");
  int threads, lanes;
  size_t smem;
  AY2_REQUIRE(red_cfg(c, &threads, &lanes, &smem) == 0, "ay2_bn_act_fwd: channels=%d unsupported", c);
  bn_act_fwd_kernel<<<wave_grid(bn_act_fwd_kernel, 256, 0, (npix + lanes - 1) / lanes), 256, 0, AY2_ST>>>(AY2_CBF(z), npix, c, z_cstride, mean, invstd, gamma,
                                                                     beta, act, AY2_BF(y), y_cstride, AY2_CBF(residual),
                                                                     res_cstride);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

static int bn_act_bwd_impl(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                           const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act, double* s1,
                           double* s2, void* dz, int32_t dz_cstride, int phases, int64_t npix_norm, void* stream,
                           float* dbeta_acc = nullptr, float* dgamma_acc = nullptr, bool sums_zeroed = false) {
  AY2_REQUIRE(dy && z && mean && invstd && gamma && beta && s1 && s2 && dz, "ay2_bn_act_bwd: null pointer");
  int threads, lanes;
  size_t smem;
  AY2_REQUIRE(red_cfg(c, &threads, &lanes, &smem) == 0, "ay2_bn_act_bwd: channels=%d unsupported", c);
  if (phases & 1) {
    if (sums_zeroed) {
      // the caller cleared every layer's sums with one fill before the backward pass
    } else if (s2 == s1 + c) {
      AY2_CHECK_CUDA(cudaMemsetAsync(s1, 0, sizeof(double) * 2 * c, AY2_ST));
    } else {
      AY2_CHECK_CUDA(cudaMemsetAsync(s1, 0, sizeof(double) * c, AY2_ST));
      AY2_CHECK_CUDA(cudaMemsetAsync(s2, 0, sizeof(double) * c, AY2_ST));
    }
    const int blocks = wave_grid(bn_act_bwd_reduce_kernel, threads, smem, (npix + lanes - 1) / lanes);
    bn_act_bwd_reduce_kernel<<<blocks, threads, smem, AY2_ST>>>(AY2_CBF(dy), dy_cstride, AY2_CBF(z), z_cstride, npix, c,
                                                                    mean, invstd, gamma, beta, act, s1, s2);
    AY2_CHECK_LAUNCH();
    count_launch();
  }
  if (phases & 2) {
    bn_act_bwd_apply_kernel<<<wave_grid(bn_act_bwd_apply_kernel, 256, 0, (npix + lanes - 1) / lanes), 256, 0, AY2_ST>>>(
        AY2_CBF(dy), dy_cstride, AY2_CBF(z), z_cstride, npix, npix_norm, c, mean, invstd, gamma, beta, act, s1, s2, AY2_BF(dz),
        dz_cstride, dbeta_acc, dgamma_acc);
    AY2_CHECK_LAUNCH();
    count_launch();
  }
  return AY2_OK;
}

extern "C" int ay2_bn_act_bwd(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                              const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act,
                              double* s1, double* s2, void* dz, int32_t dz_cstride, void* stream) {
  return bn_act_bwd_impl(dy, dy_cstride, z, z_cstride, npix, c, mean, invstd, gamma, beta, act, s1, s2, dz, dz_cstride, 3, npix, stream);
}

extern "C" int ay2_bn_act_bwd_grads(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                                    const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act,
                                    double* s1, double* s2, void* dz, int32_t dz_cstride, float* dbeta_acc, float* dgamma_acc,
                                    int32_t sums_zeroed, void* stream) {
  AY2_REQUIRE(dbeta_acc && dgamma_acc, "ay2_bn_act_bwd_grads: null gradient pointer");
  return bn_act_bwd_impl(dy, dy_cstride, z, z_cstride, npix, c, mean, invstd, gamma, beta, act, s1, s2, dz, dz_cstride, 3, npix, stream,
                         dbeta_acc, dgamma_acc, sums_zeroed != 0);
}

extern "C" int ay2_bn_act_bwd_phase(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                                    const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act,
                                    double* s1, double* s2, void* dz, int32_t dz_cstride, int32_t phase, int64_t npix_total,
                                    void* stream) {
  AY2_REQUIRE(phase == 1 || phase == 2, "ay2_bn_act_bwd_phase: phase must be 1 (reduce) or 2 (apply)");
  return bn_act_bwd_impl(dy, dy_cstride, z, z_cstride, npix, c, mean, invstd, gamma, beta, act, s1, s2, dz, dz_cstride, phase,
                         npix_total, stream);
}

extern "C" int ay2_add_slices(const void* src, int32_t src_cstride, void* dst, int32_t dst_cstride, int64_t npix, int32_t c,
                              int32_t accumulate, void* stream) {
  AY2_REQUIRE(src && dst && c % 8 == 0, "ay2_add_slices: bad arguments");
  add_slices_kernel<<<ew_grid(npix * (c / 8), 256), 256, 0, AY2_ST>>>(AY2_CBF(src), src_cstride, AY2_BF(dst), dst_cstride,
                                                                     npix, c, accumulate);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_upsample2x_bwd(const void* dy, int32_t dy_cstride, int32_t batch, int32_t h, int32_t w, int32_t c,
                                  void* dx, int32_t dx_cstride, int32_t accumulate, void* stream) {
  AY2_REQUIRE(dy && dx && c % 8 == 0, "ay2_upsample2x_bwd: bad arguments");
  upsample2x_bwd_kernel<<<ew_grid((long long)batch * h * w * (c / 8), 256), 256, 0, AY2_ST>>>(
      AY2_CBF(dy), dy_cstride, batch, h, w, c, AY2_BF(dx), dx_cstride, accumulate);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_maxpool_bwd(const void* x, int32_t x_cstride, const void* dy, int32_t dy_cstride, int32_t batch, int32_t h,
                               int32_t w, int32_t c, int32_t k, void* dx, int32_t dx_cstride, int32_t accumulate,
                               void* stream) {
  AY2_REQUIRE(x && dy && dx && k % 2 == 1, "ay2_maxpool_bwd: bad arguments");
  const size_t tile_bytes = (size_t)h * w * (16 + 16 + 16 + 32 + 16);  // x, dy, row maxima, fp32 sums, row argmax
  if (c % 8 == 0 && x_cstride % 8 == 0 && dy_cstride % 8 == 0 && dx_cstride % 8 == 0 && h * w < 0xffff &&
      tile_bytes <= 200 * 1024 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
    static size_t attr_bytes = 0;
    if (tile_bytes > 48 * 1024 && tile_bytes > attr_bytes) {
      AY2_CHECK_CUDA(cudaFuncSetAttribute(maxpool_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_bytes = 200 * 1024;
    }
    maxpool_bwd_tile_kernel<<<batch * (c / 8), 256, tile_bytes, AY2_ST>>>(AY2_CBF(x), x_cstride, AY2_CBF(dy), dy_cstride, h, w, c,
                                                                         k, AY2_BF(dx), dx_cstride, accumulate);
    AY2_CHECK_LAUNCH();
    count_launch();
    return AY2_OK;
  }
  maxpool_bwd_kernel<<<ew_grid((long long)batch * h * w * c, 256), 256, 0, AY2_ST>>>(
      AY2_CBF(x), x_cstride, AY2_CBF(dy), dy_cstride, batch, h, w, c, k, AY2_BF(dx), dx_cstride, accumulate);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_head_grad_to_nhwc(const float* grad, int32_t batch, int32_t na, int32_t ny, int32_t nx, int32_t no,
                                     void* out, int32_t out_cstride, void* stream) {
  AY2_REQUIRE(grad && out && na * no <= out_cstride, "ay2_head_grad_to_nhwc: bad arguments");
  AY2_REQUIRE(out_cstride % 2 == 0 && (long long)batch * ny * nx < (1ll << 31), "ay2_head_grad_to_nhwc: bad layout");
  head_grad_to_nhwc_kernel<<<ew_grid((long long)batch * ny * nx * 32, 256), 256, 0, AY2_ST>>>(
      grad, batch * ny * nx, na, ny * nx, no, AY2_BF(out), out_cstride);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_head_logits_to_train(const void* logits, int32_t cstride, int32_t batch, int32_t na, int32_t ny,
                                        int32_t nx, int32_t no, float* out, void* stream) {
  AY2_REQUIRE(logits && out && na * no <= cstride, "ay2_head_logits_to_train: bad arguments");
  AY2_REQUIRE(cstride % 2 == 0 && (long long)batch * ny * nx < (1ll << 31), "ay2_head_logits_to_train: bad layout");
  head_logits_to_train_kernel<<<ew_grid((long long)batch * ny * nx * 32, 256), 256, 0, AY2_ST>>>(
      AY2_CBF(logits), cstride, batch * ny * nx, na, ny * nx, no, out);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_channel_sum(const void* g, int64_t npix, int32_t c, int32_t cstride, double* sum, void* stream) {
  AY2_REQUIRE(g && sum, "ay2_channel_sum: null pointer");
  int threads, lanes;
  size_t smem;
  AY2_REQUIRE(red_cfg(c, &threads, &lanes, &smem) == 0, "ay2_channel_sum: channels=%d unsupported", c);
  long long blocks = (npix + lanes - 1) / lanes;
  if (blocks > 148 * 8) blocks = 148 * 8;
  channel_sum_kernel<<<(int)blocks, threads, smem, AY2_ST>>>(AY2_CBF(g), npix, c, cstride, sum);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_repack_weights(const void* segments, int32_t nseg, int64_t total, void* stream) {
  AY2_REQUIRE(segments && nseg > 0 && total > 0 && total % 8 == 0, "ay2_repack_weights: bad arguments");
  repack_weights_kernel<<<ew_grid(total / 8, 256), 256, 0, AY2_ST>>>(static_cast<const RepackSeg*>(segments), nseg, total / 8);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_sgd_ema_step_groups(float* param, const float* grad, float* momentum_buf, float* ema, const uint8_t* group,
                                       int64_t n, const float* lr4, const float* wd4, float momentum, int32_t nesterov,
                                       float ema_decay, float grad_scale, void* stream) {
  AY2_REQUIRE(param && grad && momentum_buf && group && lr4 && wd4 && n >= 0, "ay2_sgd_ema_step_groups: bad arguments");
  if (n == 0) return AY2_OK;
  SgdGroups gp;
  for (int k = 0; k < 4; ++k) gp.lr[k] = lr4[k], gp.wd[k] = wd4[k];  // host arrays: they change every warm-up step
  sgd_ema_groups_kernel<<<ew_grid(n, 256), 256, 0, AY2_ST>>>(param, grad, momentum_buf, ema, group, n, gp, momentum, nesterov,
                                                             ema_decay, grad_scale);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_sgd_ema_step(float* param, const float* grad, float* momentum_buf, float* ema, int64_t n, float lr,
                                float momentum, float weight_decay, int32_t nesterov, float ema_decay,
                                const float* inv_scale, void* stream) {
  AY2_REQUIRE(param && grad && momentum_buf && n >= 0, "ay2_sgd_ema_step: bad arguments");
  if (n == 0) return AY2_OK;
  sgd_ema_kernel<<<ew_grid(n, 256), 256, 0, AY2_ST>>>(param, grad, momentum_buf, ema, n, lr, momentum, weight_decay,
                                                      nesterov, ema_decay, inv_scale);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
