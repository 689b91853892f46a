// Data-movement kernels of the split-precision ("bf16x3") verification mode and the general YOLOHead decode.
//
// In that mode (include/ay2.h, ay2_conv_desc::x3) an activation tensor of C channels is stored as three planes
// [hi | lo | hi] of C channels each, hi = bf16(v), lo = bf16(v - hi): hi + lo carries 16 mantissa bits and is exact in
// fp32. These kernels are the non-convolution layers on that layout; they are verification infrastructure (simple one
// thread per element code), not the benchmarked path.
#include "ay2_common.h"
#include "ay2_ptx.cuh"
#include "head_math.cuh"

namespace ay2 {

__device__ __forceinline__ void split_store(__nv_bfloat16* plane0, long long plane_stride, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  plane0[0] = h;
  plane0[plane_stride] = __float2bfloat16_rn(v - __bfloat162float(h));
  plane0[2 * plane_stride] = h;
}
__device__ __forceinline__ float split_load(const __nv_bfloat16* plane0, long long plane_stride) {
  return __bfloat162float(plane0[0]) + __bfloat162float(plane0[plane_stride]);
}

// NCHW (u8 | f32) [B,3,H,W] -> [B, H/2, row_pixels, 48]: per pixel the planes [hi16 | lo16 | hi16] of the 16-channel
// space-to-depth pixel (channel = (dy*2+dx)*3 + c, 12 used). v = a / divisor in fp32 (255 for uint8 images).
template <typename T>
__global__ void s2d_x3_kernel(const T* __restrict__ img, int B, int H, int W, float divisor, __nv_bfloat16* __restrict__ out,
                              int out_row_pixels, int out_x_offset) {
  const int OW = W >> 1, OH = H >> 1;
  const long long total = (long long)B * OH * OW * 12;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % 12);
    const long long pix = idx / 12;
    const int ox = (int)(pix % OW);
    const int oy = (int)((pix / OW) % OH);
    const int b = (int)(pix / ((long long)OW * OH));
    const int c = ch % 3, q = ch / 3, dy = q >> 1, dx = q & 1;
    const float a = (float)img[(((long long)b * 3 + c) * H + (2 * oy + dy)) * W + 2 * ox + dx];
    const long long opix = ((long long)b * OH + oy) * out_row_pixels + ox + out_x_offset;
    split_store(out + opix * 48 + ch, 16, __fdiv_rn(a, divisor));
  }
}

// three stride-1 "same" max pools (radii r1 <= r2 <= r3) of a split-precision segment; -inf padding like nn.MaxPool2d
__global__ void sppf_x3_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int C, int cstride, int r1, int r2,
                               int r3, __nv_bfloat16* __restrict__ o1, __nv_bfloat16* __restrict__ o2,
                               __nv_bfloat16* __restrict__ o3) {
  const long long total = (long long)B * H * W * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long pix = idx / C;
    const int x = (int)(pix % W), y = (int)((pix / W) % H);
    const int b = (int)(pix / ((long long)W * H));
    float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    for (int dy = -r3; dy <= r3; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -r3; dx <= r3; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        const float v = split_load(in + (((long long)b * H + yy) * W + xx) * cstride + c, C);
        const int r = max(abs(dy), abs(dx));
        m3 = fmaxf(m3, v);
        if (r <= r2) m2 = fmaxf(m2, v);
        if (r <= r1) m1 = fmaxf(m1, v);
      }
    }
    const long long o = pix * cstride + c;
    split_store(o1 + o, C, m1);
    split_store(o2 + o, C, m2);
    split_store(o3 + o, C, m3);
  }
}

// YOLOHead decode, one thread per output element. lo_offset > 0: logits = plane(0) + plane(lo_offset) (split precision).
// flags bit 0: emit x1 y1 x2 y2 instead of x y w h (YOLOHead.out_xyxy, set by export); bit 1: exact sigmoid (fp32 expf and
// IEEE division) instead of the 2-MUFU head_sigmoid shared with the fused NMS path.
__global__ void head_decode2_kernel(const __nv_bfloat16* __restrict__ logits, int lo_offset, int B, int ny, int nx, int cstride,
                                    int na, int no, float stride_px, const float* __restrict__ anchor_wh, int flags,
                                    float* __restrict__ pred, long long total_rows, long long row_offset, float* __restrict__ raw) {
  const long long total = (long long)B * na * ny * nx;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % nx);
    const int y = (int)((idx / nx) % ny);
    const int a = (int)((idx / ((long long)nx * ny)) % na);
    const int b = (int)(idx / ((long long)nx * ny * na));
    const __nv_bfloat16* lp = logits + (((long long)b * ny + y) * nx + x) * cstride + a * no;
    const long long cell = (long long)a * ny * nx + (long long)y * nx + x;
    float* dst = pred + ((long long)b * total_rows + row_offset + cell) * no;
    float* rdst = raw ? raw + (((long long)b * na) * ny * nx + cell) * no : nullptr;
    float box[4];
    for (int o = 0; o < no; ++o) {
      float t = __bfloat162float(lp[o]);
      if (lo_offset) t += __bfloat162float(lp[o + lo_offset]);
      if (rdst) rdst[o] = t;
      const float s = (flags & 2) ? __fdiv_rn(1.0f, 1.0f + expf(-t)) : head_sigmoid(t);
      if (o < 4) {
        box[o] = o == 0 ? head_xy(s, (float)x, stride_px) : (o == 1 ? head_xy(s, (float)y, stride_px) : head_wh(s, anchor_wh[a * 2 + (o - 2)]));
      } else {
        dst[o] = s;
      }
    }
    if (flags & 1) {  // xy - wh/2, xy + wh/2
      const float hw = __fmul_rn(box[2], 0.5f), hh = __fmul_rn(box[3], 0.5f);
      dst[0] = __fsub_rn(box[0], hw), dst[1] = __fsub_rn(box[1], hh), dst[2] = __fadd_rn(box[0], hw), dst[3] = __fadd_rn(box[1], hh);
    } else {
      dst[0] = box[0], dst[1] = box[1], dst[2] = box[2], dst[3] = box[3];
    }
  }
}

static inline int grid_1d(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  return (int)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

}  // namespace ay2

using namespace ay2;

extern "C" int ay2_space_to_depth_x3(const void* img, int32_t dtype, int32_t batch, int32_t h, int32_t w, float divisor, void* out,
                                     int32_t out_row_pixels, int32_t out_x_offset, void* stream) {
  AY2_REQUIRE(img && out, "ay2_space_to_depth_x3: null pointer");
  AY2_REQUIRE(h % 2 == 0 && w % 2 == 0 && divisor > 0.f, "ay2_space_to_depth_x3: even image size and a positive divisor needed");
  AY2_REQUIRE(out_row_pixels >= w / 2 + out_x_offset, "ay2_space_to_depth_x3: output row too short");
  const long long total = (long long)batch * (h / 2) * (w / 2) * 12;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AY2_DT_U8)
    s2d_x3_kernel<uint8_t><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const uint8_t*>(img), batch, h, w, divisor,
                                                                  static_cast<__nv_bfloat16*>(out), out_row_pixels, out_x_offset);
  else if (dtype == AY2_DT_F32)
    s2d_x3_kernel<float><<<grid_1d(total, 256), 256, 0, st>>>(static_cast<const float*>(img), batch, h, w, divisor,
                                                                static_cast<__nv_bfloat16*>(out), out_row_pixels, out_x_offset);
  else
    AY2_REQUIRE(false, "ay2_space_to_depth_x3: dtype %d unsupported", dtype);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_sppf_pool_x3(const void* in, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t cstride, int32_t k1,
                                int32_t k2, int32_t k3, void* out1, void* out2, void* out3, void* stream) {
  AY2_REQUIRE(in && out1 && out2 && out3, "ay2_sppf_pool_x3: null pointer");
  AY2_REQUIRE(k1 % 2 == 1 && k2 % 2 == 1 && k3 % 2 == 1 && k1 <= k2 && k2 <= k3, "sppf_pool_x3 windows %d,%d,%d invalid", k1, k2, k3);
  AY2_REQUIRE(cstride >= 3 * c, "sppf_pool_x3: three planes of %d channels do not fit the channel stride %d", c, cstride);
  const long long total = (long long)batch * h * w * c;
  sppf_x3_kernel<<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), batch, h, w, c, cstride, k1 / 2, k2 / 2, k3 / 2, static_cast<__nv_bfloat16*>(out1),
      static_cast<__nv_bfloat16*>(out2), static_cast<__nv_bfloat16*>(out3));
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_head_decode2(const void* logits, int32_t lo_offset, int32_t batch, int32_t ny, int32_t nx, int32_t cstride,
                                int32_t na, int32_t no, float stride_px, const float* anchor_wh_px, int32_t flags, float* pred,
                                int64_t total_rows, int64_t row_offset, float* raw, void* stream) {
  AY2_REQUIRE(logits && anchor_wh_px && pred, "ay2_head_decode2: null pointer");
  AY2_REQUIRE(na * no + (lo_offset > 0 ? lo_offset : 0) <= cstride, "head_decode2: na*no=%d (+ lo plane at %d) exceeds channel stride %d",
              na * no, lo_offset, cstride);
  const long long total = (long long)batch * na * ny * nx;
  head_decode2_kernel<<<grid_1d(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(logits), lo_offset, batch, ny, nx, cstride, na, no, stride_px, anchor_wh_px, flags, pred,
      total_rows, row_offset, raw);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
