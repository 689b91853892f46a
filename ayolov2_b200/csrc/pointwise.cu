// HBM-bound data-movement kernels of the YOLOv5 forward: input space-to-depth, SPPF/SPP pooling,
// nearest 2x upsample into a concat slice, and the YOLOHead decode. All coalesced 16-byte accesses.
#include "ay2_common.h"
#include "ay2_ptx.cuh"
#include "head_math.cuh"

namespace ay2 {

// ------------------------------------------------------------------------------------------------
// NCHW (u8 | f32) [B,3,H,W] -> NHWC bf16 [B,H/2,W/2,16], channel = (dy*2+dx)*3 + c (12 used, 4 zero).
// One thread per output pixel: reads 2 adjacent input pixels for each of 3 channels x 2 rows
// (warp-coalesced along x), writes one 32-byte pixel.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void space_to_depth_kernel(const T* __restrict__ img, int B, int H, int W, float scale,
                                      uint4* __restrict__ out, int out_row_pixels, int out_x_offset) {
  const int OW = W >> 1, OH = H >> 1;
  const long long total = (long long)B * OH * OW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const int oy = (int)((idx / OW) % OH);
    const int b = (int)(idx / ((long long)OW * OH));
    float v[16];
#pragma unroll
    for (int i = 12; i < 16; ++i) v[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const T* p = img + (((long long)b * 3 + c) * H + (2 * oy + dy)) * W + 2 * ox;
        float a0, a1;
        if constexpr (sizeof(T) == 1) {
          const uchar2 u = *reinterpret_cast<const uchar2*>(p);
          a0 = (float)u.x;
          a1 = (float)u.y;
        } else {
          const float2 f = *reinterpret_cast<const float2*>(p);
          a0 = f.x;
          a1 = f.y;
        }
        v[(dy * 2 + 0) * 3 + c] = a0 * scale;
        v[(dy * 2 + 1) * 3 + c] = a1 * scale;
      }
    }
    uint4 o0, o1;
    o0.x = pack_bf16x2(v[0], v[1]);
    o0.y = pack_bf16x2(v[2], v[3]);
    o0.z = pack_bf16x2(v[4], v[5]);
    o0.w = pack_bf16x2(v[6], v[7]);
    o1.x = pack_bf16x2(v[8], v[9]);
    o1.y = pack_bf16x2(v[10], v[11]);
    o1.z = pack_bf16x2(v[12], v[13]);
    o1.w = pack_bf16x2(v[14], v[15]);
    const long long opix = ((long long)b * OH + oy) * out_row_pixels + ox + out_x_offset;
    out[opix * 2] = o0;
    out[opix * 2 + 1] = o1;
  }
}

// ------------------------------------------------------------------------------------------------
// SPPF/SPP: three stride-1 "same" max pools (windows k1<k2<k3, odd) of one NHWC bf16 tensor.
// One CTA per (image, 8-channel chunk): the plane is staged in shared memory, separable max
// (rows then columns). Out-of-image taps are -inf like nn.MaxPool2d padding.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// CPC: 16-byte channel chunks per CTA. With 2, neighbouring threads read / write the two halves of a 32-byte sector, so
// every DRAM sector the strided NHWC plane touches is used whole (with 1 half of each sector was wasted both ways).
template <int CPC>
__global__ void sppf_pool_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int C, int cstride, int r1,
                                 int r2, int r3, __nv_bfloat16* __restrict__ o1, __nv_bfloat16* __restrict__ o2,
                                 __nv_bfloat16* __restrict__ o3) {
  extern __shared__ uint4 sp[];
  const int groups = C / (8 * CPC);
  const int b = blockIdx.x / groups;
  const int ch = (blockIdx.x - b * groups) * 8 * CPC;
  const int HW = H * W, N = HW * CPC;
  uint4* X = sp;          // [H][W][CPC]
  uint4* R1 = X + N;      // row-max with radius r1
  uint4* R2 = R1 + N;
  uint4* R3 = R2 + N;
  const uint32_t ninf2 = 0xFF80FF80u;  // bf16 -inf pair
  const uint4 NINF = make_uint4(ninf2, ninf2, ninf2, ninf2);
  const long long base = (long long)b * HW * cstride + ch;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    X[i] = *reinterpret_cast<const uint4*>(in + base + (long long)(i / CPC) * cstride + (i % CPC) * 8);
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int x = (i / CPC) % W;
    uint4 m = X[i];
    int d = 1;
    for (; d <= r1; ++d) {
      if (x - d >= 0) m = bf16x8_max(m, X[i - d * CPC]);
      if (x + d < W) m = bf16x8_max(m, X[i + d * CPC]);
    }
    R1[i] = m;
    for (; d <= r2; ++d) {
      if (x - d >= 0) m = bf16x8_max(m, X[i - d * CPC]);
      if (x + d < W) m = bf16x8_max(m, X[i + d * CPC]);
    }
    R2[i] = m;
    for (; d <= r3; ++d) {
      if (x - d >= 0) m = bf16x8_max(m, X[i - d * CPC]);
      if (x + d < W) m = bf16x8_max(m, X[i + d * CPC]);
    }
    R3[i] = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int y = (i / CPC) / W;
    uint4 m1 = NINF, m2 = NINF, m3 = NINF;
    for (int d = -r3; d <= r3; ++d) {
      const int yy = y + d;
      if (yy < 0 || yy >= H) continue;
      const int j = i + d * W * CPC;
      m3 = bf16x8_max(m3, R3[j]);
      if (d >= -r2 && d <= r2) m2 = bf16x8_max(m2, R2[j]);
      if (d >= -r1 && d <= r1) m1 = bf16x8_max(m1, R1[j]);
    }
    const long long o = base + (long long)(i / CPC) * cstride + (i % CPC) * 8;
    *reinterpret_cast<uint4*>(o1 + o) = m1;
    *reinterpret_cast<uint4*>(o2 + o) = m2;
    *reinterpret_cast<uint4*>(o3 + o) = m3;
  }
}

// ------------------------------------------------------------------------------------------------
// nearest 2x upsample, NHWC bf16, into a (possibly wider) destination buffer
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int C, int ics,
                                  __nv_bfloat16* __restrict__ out, int ocs) {
  // one thread per 16-byte chunk of an INPUT pixel: one load, four stores (the 2x2 output pixels), 32-bit index math
  const unsigned chunks = C >> 3;
  const unsigned total = (unsigned)B * H * W * chunks;  // < 2^31 for every supported shape (checked by the launcher)
  const unsigned OW = W * 2;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned c = (idx % chunks) << 3;
    const unsigned pix = idx / chunks;          // (b * H + y) * W + x
    const unsigned x = pix % W;
    const unsigned row = pix / W;               // b * H + y
    const uint4 v = *reinterpret_cast<const uint4*>(in + (size_t)pix * ics + c);
    __nv_bfloat16* o = out + ((size_t)(2 * row) * OW + 2 * x) * ocs + c;  // output row 2 * (b * H + y) == b * 2H + 2y
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + ocs) = v;
    *reinterpret_cast<uint4*>(o + (size_t)OW * ocs) = v;
    *reinterpret_cast<uint4*>(o + (size_t)OW * ocs + ocs) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// YOLOHead decode. One warp per pixel: lanes stride over the na*no logits of the pixel (coalesced
// bf16 reads), each output row of `no` floats is written with consecutive lanes -> consecutive floats.
//   y = sigmoid(t);  xy = (y*2 - 0.5 + grid) * stride;  wh = (y*2)^2 * anchor;  rest = y
// ------------------------------------------------------------------------------------------------
__global__ void head_decode_kernel(const __nv_bfloat16* __restrict__ logits, int B, int ny, int nx, int cstride,
                                   int na, int no, float stride_px, const float* __restrict__ anchor_wh,
                                   float* __restrict__ pred, long long total_rows, long long row_offset,
                                   float* __restrict__ raw) {
  // One warp per pixel. Each lane loads 8 consecutive logits (one 16-byte read), decodes them and parks the
  // results in shared memory; the warp then streams each anchor's `no` floats out with consecutive lanes.
  extern __shared__ float dsm[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int nch = na * no;
  const int nch8 = (nch + 7) & ~7;
  float* vbuf = dsm + (size_t)wib * 2 * nch8;  // decoded values
  float* rbuf = vbuf + nch8;                   // raw logits (fp32)
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long npix = (long long)B * ny * nx;
  for (long long pix = warp_id; pix < npix; pix += nwarps) {
    const int x = (int)(pix % nx);
    const int y = (int)((pix / nx) % ny);
    const int b = (int)(pix / ((long long)nx * ny));
    const __nv_bfloat16* lp = logits + pix * cstride;
    for (int c0 = lane * 8; c0 < nch8; c0 += 256) {
      const uint4 q = *reinterpret_cast<const uint4*>(lp + c0);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
      float t[8], sg[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        t[2 * j] = f.x;
        t[2 * j + 1] = f.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sg[j] = head_sigmoid(t[j]);
      *reinterpret_cast<float4*>(vbuf + c0) = make_float4(sg[0], sg[1], sg[2], sg[3]);
      *reinterpret_cast<float4*>(vbuf + c0 + 4) = make_float4(sg[4], sg[5], sg[6], sg[7]);
      if (raw) {
        *reinterpret_cast<float4*>(rbuf + c0) = make_float4(t[0], t[1], t[2], t[3]);
        *reinterpret_cast<float4*>(rbuf + c0 + 4) = make_float4(t[4], t[5], t[6], t[7]);
      }
    }
    __syncwarp();
    // box channels: 4 per anchor, fixed up by the first 4*na lanes (xy grid/stride, wh anchor)
    if (lane < 4 * na) {
      const int a = lane >> 2, o = lane & 3;
      const float sgm = vbuf[a * no + o];
      float v;
      if (o == 0) v = head_xy(sgm, (float)x, stride_px);
      else if (o == 1) v = head_xy(sgm, (float)y, stride_px);
      else v = head_wh(sgm, anchor_wh[a * 2 + (o - 2)]);
      vbuf[a * no + o] = v;
    }
    __syncwarp();
    for (int a = 0; a < na; ++a) {
      const long long cell = (long long)a * ny * nx + (long long)y * nx + x;
      float* dst = pred + ((long long)b * total_rows + row_offset + cell) * no;
      for (int o = lane; o < no; o += 32) dst[o] = vbuf[a * no + o];
      if (raw) {
        float* rdst = raw + (((long long)b * na) * ny * nx + cell) * no;
        for (int o = lane; o < no; o += 32) rdst[o] = rbuf[a * no + o];
      }
    }
    __syncwarp();
  }
}

}  // namespace ay2

using namespace ay2;

static int grid_for(long long total, int threads, int per_sm) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = 148LL * per_sm;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

extern "C" int ay2_space_to_depth(const void* img, int32_t dtype, int32_t batch, int32_t h, int32_t w, float scale,
                                  void* out, int32_t out_row_pixels, int32_t out_x_offset, void* stream) {
  AY2_REQUIRE(img && out, "ay2_space_to_depth: null pointer");
  AY2_REQUIRE(h % 2 == 0 && w % 2 == 0 && h > 0 && w > 0, "space_to_depth needs even H,W (got %dx%d)", h, w);
  AY2_REQUIRE(dtype == AY2_DT_U8 || dtype == AY2_DT_F32, "space_to_depth dtype %d unsupported", dtype);
  AY2_REQUIRE(out_x_offset >= 0 && out_row_pixels >= w / 2 + out_x_offset, "space_to_depth output row too short");
  const long long total = (long long)batch * (h / 2) * (w / 2);
  const int threads = 256;
  const int blocks = grid_for(total, threads, 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AY2_DT_U8)
    space_to_depth_kernel<uint8_t><<<blocks, threads, 0, st>>>(static_cast<const uint8_t*>(img), batch, h, w, scale,
                                                               static_cast<uint4*>(out), out_row_pixels, out_x_offset);
  else
    space_to_depth_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(img), batch, h, w, scale,
                                                             static_cast<uint4*>(out), out_row_pixels, out_x_offset);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_sppf_pool(const void* in, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t cstride,
                             int32_t k1, int32_t k2, int32_t k3, void* out1, void* out2, void* out3, void* stream) {
  AY2_REQUIRE(in && out1 && out2 && out3, "ay2_sppf_pool: null pointer");
  AY2_REQUIRE(c % 8 == 0 && cstride % 8 == 0, "sppf_pool channels must be multiples of 8");
  AY2_REQUIRE(k1 % 2 == 1 && k2 % 2 == 1 && k3 % 2 == 1 && k1 <= k2 && k2 <= k3, "sppf_pool windows %d,%d,%d invalid",
              k1, k2, k3);
  const int cpc = (c % 16 == 0 && (size_t)8 * h * w * sizeof(uint4) <= 100 * 1024) ? 2 : 1;
  const size_t smem = (size_t)4 * cpc * h * w * sizeof(uint4);
  AY2_REQUIRE(smem <= 200 * 1024, "sppf_pool plane %dx%d too large for shared memory", h, w);
  static DeviceOnce once;
  if (once.first()) {
    AY2_CHECK_CUDA(cudaFuncSetAttribute(sppf_pool_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    AY2_CHECK_CUDA(cudaFuncSetAttribute(sppf_pool_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const int n = h * w * cpc;
  const int threads = n >= 512 ? 512 : ((n + 31) / 32) * 32;
  auto kern = cpc == 2 ? sppf_pool_kernel<2> : sppf_pool_kernel<1>;
  kern<<<batch * (c / (8 * cpc)), threads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), h, w, c, cstride, k1 / 2, k2 / 2, k3 / 2,
      static_cast<__nv_bfloat16*>(out1), static_cast<__nv_bfloat16*>(out2), static_cast<__nv_bfloat16*>(out3));
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_upsample2x(const void* in, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t in_cstride,
                              void* out, int32_t out_cstride, void* stream) {
  AY2_REQUIRE(in && out, "ay2_upsample2x: null pointer");
  AY2_REQUIRE(c % 8 == 0 && in_cstride % 8 == 0 && out_cstride % 8 == 0, "upsample2x channels must be multiples of 8");
  const long long total = (long long)batch * h * w * (c / 8);
  AY2_REQUIRE(total < (1ll << 31), "upsample2x: tensor too large for 32-bit indexing");
  const int threads = 256;
  upsample2x_kernel<<<grid_for(total, threads, 16), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), batch, h, w, c, in_cstride, static_cast<__nv_bfloat16*>(out), out_cstride);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_head_decode(const void* logits, int32_t batch, int32_t ny, int32_t nx, int32_t cstride, int32_t na,
                               int32_t no, float stride_px, const float* anchor_wh_px, float* pred, int64_t total_rows,
                               int64_t row_offset, float* raw, void* stream) {
  AY2_REQUIRE(logits && anchor_wh_px && pred, "ay2_head_decode: null pointer");
  AY2_REQUIRE(na * no <= cstride, "head_decode: na*no=%d exceeds channel stride %d", na * no, cstride);
  AY2_REQUIRE(cstride % 8 == 0 && ((na * no + 7) & ~7) <= cstride, "head_decode: channel stride %d too small", cstride);
  const long long npix = (long long)batch * ny * nx;
  const int threads = 256;
  const size_t smem = (size_t)(threads / 32) * 2 * ((na * no + 7) & ~7) * sizeof(float);
  AY2_REQUIRE(smem <= 48 * 1024, "head_decode: na*no=%d too large", na * no);
  head_decode_kernel<<<grid_for(npix * 32, threads, 8), threads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), batch, ny, nx, cstride, na, no, stride_px, anchor_wh_px, pred,
      total_rows, row_offset, raw);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
