// Input side of the path on the GPU: letterbox + BGR->RGB + HWC->CHW + collate for a whole batch (one launch per kind of image: plain copy / resize).
//
// Replaces, per image on the host in the reference, LoadImages._letterbox (scripts/data_loader/data_loader.py:395-459:
// cv2.resize(INTER_LINEAR) to the unpadded size + cv2.copyMakeBorder(114)), the `transpose((2, 0, 1))[::-1]` of
// :388-389 and the torch.stack of collate_fn (:461-477, :905-909); optionally also prepare_img + the stem's
// space-to-depth (ay2_space_to_depth) so that the uint8 NCHW batch never exists.
//
// Input: the loaded images as they are (ragged HWC BGR uint8, anywhere in one device arena) + a table of per-image
// geometry computed by the host (letterbox_geometry is a dozen scalar operations per image). A thread owns a 4 x 2
// block of output pixels: the two interpolation rows and four interpolation columns are shared inside the block.
// cv2's 8-bit INTER_LINEAR is integer arithmetic (11-bit taps, see oracle/input_oracle.py), reproduced bit for bit:
// the only floating point is the tap position, computed in the same precision and operation order (no contraction).
// HBM-bound: 3 bytes read and 3 (uint8) or 8 (bf16 space-to-depth, 12 of 16 channels used) bytes written per pixel.
#include "ay2_common.h"
#include "ay2_ptx.cuh"

namespace ay2 {

struct LetterboxParams {
  const uint8_t* arena;
  const ay2_letterbox_image* table;
  void* out;
  int H, W;
  int out_row_pixels, out_x_offset;
  float scale;
  uint8_t color[3];  // BGR like the source
};

enum { LB_COPY = 0, LB_AREA = 1, LB_LINEAR = 2 };
__device__ __forceinline__ int image_mode(const ay2_letterbox_image& im) {
  if (im.dst_h == im.src_h && im.dst_w == im.src_w) return LB_COPY;
  if (im.src_h == 2 * im.dst_h && im.src_w == 2 * im.dst_w) return LB_AREA;  // cv2 reroutes the exact 2 x 2 decimation
  return LB_LINEAR;
}

struct Tap {
  int s0, s1, c0, c1;  // source indices and 11-bit weights
};

// resize.cpp: fx = (float)((d + 0.5) * scale - 0.5); s = floor(fx); fx -= s, with scale = 1 / (dst / src) in double
// (the host puts it into the table: two double divisions per tap were a third of the interpolating kernel's instructions)
__device__ __forceinline__ void tap_position(int d, double scale, int& s, float& f) {
  f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
  const float fl = floorf(f);
  s = (int)fl;
  f = __fsub_rn(f, fl);
}

__device__ __forceinline__ Tap make_tap(float f, int s0, int s1) {
  Tap t;
  t.s0 = s0, t.s1 = s1;
  t.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));  // saturate_cast<short>: round half to even
  t.c1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  return t;
}

// columns collapse to one tap outside [0, w - 1); rows are clamped instead
__device__ __forceinline__ Tap column_tap(int d, int ssize, double scale) {
  int s;
  float f;
  tap_position(d, scale, s, f);
  if (s < 0) f = 0.f, s = 0;
  if (s >= ssize - 1) f = 0.f, s = ssize - 1;
  return make_tap(f, s, min(s + 1, ssize - 1));
}

__device__ __forceinline__ Tap row_tap(int d, int ssize, double scale) {
  int s;
  float f;
  tap_position(d, scale, s, f);
  return make_tap(f, min(max(s, 0), ssize - 1), min(max(s + 1, 0), ssize - 1));
}

// the thread's 4 x 2 pixels (BGR) -> the collated uint8 NCHW RGB tensor, or the stem's space-to-depth bf16 image
template <int KIND>
__device__ __forceinline__ void write_block(const LetterboxParams& p, int b, int x0, int y0, const uint8_t (&px)[2][4][3]) {
  if constexpr (KIND == AY2_LB_NCHW_U8) {
    uint8_t* out = static_cast<uint8_t*>(p.out);
#pragma unroll
    for (int c = 0; c < 3; ++c)  // output channel c (RGB) = source channel 2 - c (BGR)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uchar4 v = make_uchar4(px[r][0][2 - c], px[r][1][2 - c], px[r][2][2 - c], px[r][3][2 - c]);
        *reinterpret_cast<uchar4*>(out + (((size_t)b * 3 + c) * p.H + y0 + r) * p.W + x0) = v;
      }
  } else {
    // ay2_space_to_depth's layout: [B, H/2, out_row_pixels, 16] bf16, channel (dy*2+dx)*3 + c, 4 zero channels
    uint4* out = static_cast<uint4*>(p.out);
    const size_t opix = ((size_t)b * (p.H >> 1) + (y0 >> 1)) * p.out_row_pixels + (x0 >> 1) + p.out_x_offset;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[12];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
#pragma unroll
          for (int c = 0; c < 3; ++c) v[(dy * 2 + dx) * 3 + c] = (float)px[dy][2 * h + dx][2 - c] * p.scale;
      uint4 o0, o1;
      o0.x = pack_bf16x2(v[0], v[1]), o0.y = pack_bf16x2(v[2], v[3]), o0.z = pack_bf16x2(v[4], v[5]), o0.w = pack_bf16x2(v[6], v[7]);
      o1.x = pack_bf16x2(v[8], v[9]), o1.y = pack_bf16x2(v[10], v[11]), o1.z = 0u, o1.w = 0u;
      out[(opix + h) * 2] = o0;
      out[(opix + h) * 2 + 1] = o1;
    }
  }
}

// Two kernels over the same grid, because they want different register budgets: the copy kernel (images that enter at
// their final size -- the validation set after `_load_image`) is a pure streaming kernel that needs many warps in flight;
// the interpolating kernel carries six taps per thread. Each exits at once on the other's images.
template <int KIND, bool RESIZE>
__global__ void __launch_bounds__(256) letterbox_collate_kernel(LetterboxParams p) {
  const int b = blockIdx.z;
  const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
  const ay2_letterbox_image im = p.table[b];
  const int mode = image_mode(im);
  if ((mode != LB_COPY) != RESIZE) return;
  if (x0 >= p.W || y0 >= p.H) return;
  const uint8_t* src = p.arena + im.src_offset;

  uint8_t px[2][4][3];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) px[r][j][c] = p.color[c];

  const int ry = y0 - im.top, rx = x0 - im.left;  // position inside the resized image
  if (ry + 1 >= 0 && ry < im.dst_h && rx + 3 >= 0 && rx < im.dst_w) {
    if constexpr (!RESIZE) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (ry + r < 0 || ry + r >= im.dst_h) continue;
        const uint8_t* q = src + (size_t)(ry + r) * im.src_row_bytes + rx * 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rx + j < 0 || rx + j >= im.dst_w) continue;
#pragma unroll
          for (int c = 0; c < 3; ++c) px[r][j][c] = __ldg(q + 3 * j + c);
        }
      }
    } else if (mode == LB_LINEAR) {
      Tap ct[4], rt[2];
#pragma unroll
      for (int j = 0; j < 4; ++j) ct[j] = column_tap(min(max(rx + j, 0), im.dst_w - 1), im.src_w, im.scale_x);
#pragma unroll
      for (int r = 0; r < 2; ++r) rt[r] = row_tap(min(max(ry + r, 0), im.dst_h - 1), im.src_h, im.scale_y);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (ry + r < 0 || ry + r >= im.dst_h) continue;
        const uint8_t* l0 = src + (size_t)rt[r].s0 * im.src_row_bytes;
        const uint8_t* l1 = src + (size_t)rt[r].s1 * im.src_row_bytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rx + j < 0 || rx + j >= im.dst_w) continue;
          const int o0 = ct[j].s0 * 3, o1 = ct[j].s1 * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int h0 = (int)__ldg(l0 + o0 + c) * ct[j].c0 + (int)__ldg(l0 + o1 + c) * ct[j].c1;  // horizontal pass
            const int h1 = (int)__ldg(l1 + o0 + c) * ct[j].c0 + (int)__ldg(l1 + o1 + c) * ct[j].c1;
            const int v = (((rt[r].c0 * (h0 >> 4)) >> 16) + ((rt[r].c1 * (h1 >> 4)) >> 16) + 2) >> 2;  // vertical pass
            px[r][j][c] = (uint8_t)min(max(v, 0), 255);
          }
        }
      }
    } else {  // LB_AREA
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (ry + r < 0 || ry + r >= im.dst_h) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rx + j < 0 || rx + j >= im.dst_w) continue;
          const uint8_t* q0 = src + (size_t)(2 * (ry + r)) * im.src_row_bytes + (2 * (rx + j)) * 3;
          const uint8_t* q1 = q0 + im.src_row_bytes;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            px[r][j][c] = (uint8_t)(((int)__ldg(q0 + c) + (int)__ldg(q0 + 3 + c) + (int)__ldg(q1 + c) + (int)__ldg(q1 + 3 + c) + 2) >> 2);
        }
      }
    }
  }
  write_block<KIND>(p, b, x0, y0, px);
}

// ------------------------------------------------------------------------------------------------
// LoadImages._load_image after the decode (scripts/data_loader/data_loader.py:320-329): long side -> img_size with
// cv2.resize INTER_AREA (shrinking, no augmentation) or INTER_LINEAR, ragged images -> ragged images in the same arena.
// A thread owns one output pixel. INTER_AREA as OpenCV computes it for 8-bit images (oracle/input_oracle.py):
//   integer ratios: integer cell sum, (a + b + c + d + 2) >> 2 for 2 x 2, else round_half_even(sum * fp32(1 / area));
//   other ratios:   per source line the fp32 sum of its cells in source order (buf += S * alpha; partial left cell, full
//                   cells, partial right cell, weights from double arithmetic rounded once to fp32), then the fp32 sum of
//                   the lines (sum += beta * buf) -- separate multiplies and adds (the CPU build has no FMA), half-even.
struct AreaCells {
  int s1, s2;          // full cells [s1, s2)
  bool left, right;    // partial cells at s1 - 1 / s2
  float wl, wm, wr;    // their weights
};
__device__ __forceinline__ AreaCells area_cells(int d, int ssize, double scale) {
  AreaCells a;
  const double fs1 = __dmul_rn((double)d, scale), fs2 = __dadd_rn(fs1, scale);
  const double cell = fmin(scale, __dsub_rn((double)ssize, fs1));
  int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
  s2 = min(s2, ssize - 1);
  s1 = min(s1, s2);
  a.s1 = s1, a.s2 = s2;
  const double dl = __dsub_rn((double)s1, fs1), dr = __dsub_rn(fs2, (double)s2);
  a.left = dl > 1e-3;
  a.right = dr > 1e-3;
  a.wl = __double2float_rn(__ddiv_rn(dl, cell));
  a.wm = __double2float_rn(__ddiv_rn(1.0, cell));
  a.wr = __double2float_rn(__ddiv_rn(fmin(fmin(dr, 1.0), cell), cell));
  return a;
}

__device__ __forceinline__ void area_line(const uint8_t* line, const AreaCells& x, float beta, float (&sum)[3]) {
  float buf[3] = {0.f, 0.f, 0.f};
  if (x.left) {
    const uint8_t* q = line + (x.s1 - 1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)__ldg(q + c), x.wl));
  }
  for (int sx = x.s1; sx < x.s2; ++sx) {
    const uint8_t* q = line + sx * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)__ldg(q + c), x.wm));
  }
  if (x.right) {
    const uint8_t* q = line + x.s2 * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn((float)__ldg(q + c), x.wr));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) sum[c] = __fadd_rn(sum[c], __fmul_rn(beta, buf[c]));
}

__device__ __forceinline__ uint8_t round_u8(float v) { return (uint8_t)min(max(__float2int_rn(v), 0), 255); }

__global__ void __launch_bounds__(256) load_resize_kernel(uint8_t* arena, const ay2_load_resize_image* table) {
  const ay2_load_resize_image im = table[blockIdx.z];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= im.dst_h * im.dst_w) return;
  const int dy = idx / im.dst_w, dx = idx - dy * im.dst_w;
  const uint8_t* src = arena + im.src_offset;
  uint8_t* out = arena + im.dst_offset + (size_t)dy * im.dst_row_bytes + dx * 3;
  const bool half = im.src_h == 2 * im.dst_h && im.src_w == 2 * im.dst_w;
  if (half) {  // both interpolations run the 2 x 2 area filter on an exact decimation
    const uint8_t* q0 = src + (size_t)(2 * dy) * im.src_row_bytes + (2 * dx) * 3;
    const uint8_t* q1 = q0 + im.src_row_bytes;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = (uint8_t)(((int)__ldg(q0 + c) + (int)__ldg(q0 + 3 + c) + (int)__ldg(q1 + c) + (int)__ldg(q1 + 3 + c) + 2) >> 2);
    return;
  }
  if (im.mode == AY2_LR_LINEAR) {
    const Tap ct = column_tap(dx, im.src_w, im.scale_x), rt = row_tap(dy, im.src_h, im.scale_y);
    const uint8_t* l0 = src + (size_t)rt.s0 * im.src_row_bytes;
    const uint8_t* l1 = src + (size_t)rt.s1 * im.src_row_bytes;
    const int o0 = ct.s0 * 3, o1 = ct.s1 * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)__ldg(l0 + o0 + c) * ct.c0 + (int)__ldg(l0 + o1 + c) * ct.c1;
      const int h1 = (int)__ldg(l1 + o0 + c) * ct.c0 + (int)__ldg(l1 + o1 + c) * ct.c1;
      out[c] = (uint8_t)min(max((((rt.c0 * (h0 >> 4)) >> 16) + ((rt.c1 * (h1 >> 4)) >> 16) + 2) >> 2, 0), 255);
    }
    return;
  }
  // INTER_AREA
  const int isx = __double2int_rn(im.scale_x), isy = __double2int_rn(im.scale_y);
  const double eps = 2.220446049250313e-16;  // DBL_EPSILON
  if (fabs(im.scale_x - (double)isx) < eps && fabs(im.scale_y - (double)isy) < eps) {
    int total[3] = {0, 0, 0};
    for (int ky = 0; ky < isy; ++ky) {
      const uint8_t* q = src + (size_t)(dy * isy + ky) * im.src_row_bytes + (dx * isx) * 3;
      for (int kx = 0; kx < isx; ++kx)
#pragma unroll
        for (int c = 0; c < 3; ++c) total[c] += (int)__ldg(q + kx * 3 + c);
    }
    const float inv = __fdiv_rn(1.0f, (float)(isx * isy));
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = round_u8(__fmul_rn((float)total[c], inv));
    return;
  }
  const AreaCells x = area_cells(dx, im.src_w, im.scale_x), y = area_cells(dy, im.src_h, im.scale_y);
  float sum[3] = {0.f, 0.f, 0.f};
  if (y.left) area_line(src + (size_t)(y.s1 - 1) * im.src_row_bytes, x, y.wl, sum);
  for (int sy = y.s1; sy < y.s2; ++sy) area_line(src + (size_t)sy * im.src_row_bytes, x, y.wm, sum);
  if (y.right) area_line(src + (size_t)y.s2 * im.src_row_bytes, x, y.wr, sum);
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c] = round_u8(sum[c]);
}

// YoloTrainer.multi_scale (scripts/train/yolo_trainer.py:223-248) with prepare_img (abstract_trainer.py:252-261) fused in:
// out = F.interpolate(in * pre, size, mode="bilinear", align_corners=False) as fp32 NCHW, in the operation order of
// torch's own kernel (source index = max(ratio * (dst + 0.5) - 0.5, 0), ratio = in / out in fp32; lambda = frac;
// value = l0h * (l0w * v00 + l1w * v01) + l1h * (l0w * v10 + l1w * v11)). A thread owns one output pixel, three channels.
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const T* __restrict__ in, int H, int W, float* __restrict__ out,
                                                              int H2, int W2, float pre, float rh, float rw, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x2 = (int)(idx % W2), y2 = (int)((idx / W2) % H2), b = (int)(idx / ((long long)W2 * H2));
    const float h1r = fmaxf(rh * ((float)y2 + 0.5f) - 0.5f, 0.0f), w1r = fmaxf(rw * ((float)x2 + 0.5f) - 0.5f, 0.0f);
    const int h1 = (int)h1r, w1 = (int)w1r;
    const int h1p = h1 < H - 1 ? 1 : 0, w1p = w1 < W - 1 ? 1 : 0;
    const float h1l = h1r - (float)h1, h0l = 1.0f - h1l, w1l = w1r - (float)w1, w0l = 1.0f - w1l;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T* p = in + (((size_t)b * 3 + c) * H + h1) * W + w1;
      const float v00 = (float)p[0] * pre, v01 = (float)p[w1p] * pre;
      const float v10 = (float)p[(size_t)h1p * W] * pre, v11 = (float)p[(size_t)h1p * W + w1p] * pre;
      out[(((size_t)b * 3 + c) * H2 + y2) * W2 + x2] = h0l * (w0l * v00 + w1l * v01) + h1l * (w0l * v10 + w1l * v11);
    }
  }
}

// LoadImagesAndLabels.collate_fn (data_loader.py:905-907): label column 0 = index of the image the row belongs to
__global__ void collate_labels_kernel(float* labels, const int* offsets, int batch, int total) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int lo = 0, hi = batch;  // largest i with offsets[i] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= t) lo = mid; else hi = mid;
  }
  labels[(size_t)t * 6] = (float)lo;
}

}  // namespace ay2

using namespace ay2;

extern "C" int ay2_letterbox_collate(const uint8_t* arena, const ay2_letterbox_image* table, int32_t batch, int32_t kinds,
                                     int32_t out_h, int32_t out_w, uint32_t color_bgr, int32_t out_kind, void* out,
                                     int32_t out_row_pixels, int32_t out_x_offset, float scale, void* stream) {
  AY2_REQUIRE(arena && table && out && batch >= 0, "ay2_letterbox_collate: bad arguments");
  AY2_REQUIRE(out_h > 0 && out_w > 0 && out_h % 2 == 0 && out_w % 4 == 0,
              "ay2_letterbox_collate: the network input must have an even height and a width that is a multiple of 4 (got %dx%d)", out_h, out_w);
  AY2_REQUIRE(out_kind == AY2_LB_NCHW_U8 || out_kind == AY2_LB_S2D_BF16, "ay2_letterbox_collate: output kind %d unknown", out_kind);
  if (out_kind == AY2_LB_S2D_BF16)
    AY2_REQUIRE(out_x_offset >= 0 && out_row_pixels >= out_w / 2 + out_x_offset, "ay2_letterbox_collate: output row too short");
  AY2_REQUIRE(batch <= 65535, "ay2_letterbox_collate: batch %d too large", batch);
  if (batch == 0) return AY2_OK;
  if ((kinds & (AY2_LB_HAS_COPY | AY2_LB_HAS_RESIZE)) == 0) kinds = AY2_LB_HAS_COPY | AY2_LB_HAS_RESIZE;  // unknown: run both
  LetterboxParams p;
  p.arena = arena, p.table = table, p.out = out, p.H = out_h, p.W = out_w;
  p.out_row_pixels = out_row_pixels, p.out_x_offset = out_x_offset, p.scale = scale;
  p.color[0] = color_bgr & 255u, p.color[1] = (color_bgr >> 8) & 255u, p.color[2] = (color_bgr >> 16) & 255u;
  const dim3 block(32, 8), grid(ceil_div(out_w, 128), ceil_div(out_h, 16), batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool u8 = out_kind == AY2_LB_NCHW_U8;
  if (kinds & AY2_LB_HAS_COPY) {
    if (u8) letterbox_collate_kernel<AY2_LB_NCHW_U8, false><<<grid, block, 0, st>>>(p);
    else letterbox_collate_kernel<AY2_LB_S2D_BF16, false><<<grid, block, 0, st>>>(p);
    AY2_CHECK_LAUNCH();
    count_launch();
  }
  if (kinds & AY2_LB_HAS_RESIZE) {
    if (u8) letterbox_collate_kernel<AY2_LB_NCHW_U8, true><<<grid, block, 0, st>>>(p);
    else letterbox_collate_kernel<AY2_LB_S2D_BF16, true><<<grid, block, 0, st>>>(p);
    AY2_CHECK_LAUNCH();
    count_launch();
  }
  return AY2_OK;
}

extern "C" int ay2_collate_labels(float* labels, const int32_t* offsets, int32_t batch, int32_t total, void* stream) {
  AY2_REQUIRE(batch >= 0 && total >= 0 && (total == 0 || (labels && offsets && batch > 0)), "ay2_collate_labels: bad arguments");
  if (total == 0) return AY2_OK;
  collate_labels_kernel<<<ceil_div(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(labels, offsets, batch, total);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_resize_bilinear(const void* img, int32_t dtype, int32_t batch, int32_t h, int32_t w, float pre_scale, float* out,
                                   int32_t out_h, int32_t out_w, void* stream) {
  AY2_REQUIRE(img && out && batch >= 0 && h > 0 && w > 0 && out_h > 0 && out_w > 0, "ay2_resize_bilinear: bad arguments");
  AY2_REQUIRE(dtype == AY2_DT_U8 || dtype == AY2_DT_F32, "ay2_resize_bilinear: dtype %d unsupported", dtype);
  const long long total = (long long)batch * out_h * out_w;
  if (total == 0) return AY2_OK;
  const float rh = (float)h / (float)out_h, rw = (float)w / (float)out_w;  // area_pixel_compute_scale, align_corners = false
  const int threads = 256;
  const long long want = (total + threads - 1) / threads;
  const int blocks = (int)(want < 148ll * 32 ? want : 148ll * 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == AY2_DT_U8)
    resize_bilinear_kernel<uint8_t><<<blocks, threads, 0, st>>>(static_cast<const uint8_t*>(img), h, w, out, out_h, out_w, pre_scale, rh, rw, total);
  else
    resize_bilinear_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(img), h, w, out, out_h, out_w, pre_scale, rh, rw, total);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_load_resize(uint8_t* arena, const ay2_load_resize_image* table, int32_t count, int32_t max_dst_pixels, void* stream) {
  AY2_REQUIRE(count >= 0 && max_dst_pixels >= 0 && (count == 0 || (arena && table)), "ay2_load_resize: bad arguments");
  AY2_REQUIRE(count <= 65535, "ay2_load_resize: %d images in one call", count);
  if (count == 0 || max_dst_pixels == 0) return AY2_OK;
  load_resize_kernel<<<dim3(ceil_div(max_dst_pixels, 256), 1, count), 256, 0, static_cast<cudaStream_t>(stream)>>>(arena, table);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
