// The non-default suppression rules of the reference, batched over images on the device.
//
//   scripts/utils/metrics.py:388-431  non_max_suppression(nms_type = "batched_nms" | "fast_nms" | "matrix_nms" | "merge_nms")
//   scripts/utils/nms.py:63-110       batched_nms(nms_type = ...) of the val2 path
//
// The reference walks the images on the host and materialises an n x n IoU matrix per image. Here every rule works on one
// CANDIDATE TABLE for the whole batch -- [B][cap] rows {x1, y1, x2, y2, conf, cls, -, -} in the reference's candidate
// order (row order of the prediction tensor, then class for multi_label: metrics.py:337-368 / nms.py:44-55) -- and never
// forms a matrix:
//   fast    a row survives iff no EARLIER row overlaps it:   keep_j = max_{i<j} IoU(i, j) < thr        (column maximum)
//   matrix  gaussian decay of the score:   m_i = max_{k<i} IoU(k, i),  decay_j = min_i exp(-(IoU(i,j)^2 - m_i^2) / 0.5)
//           with IoU(i, j) = 0 for i >= j (the reference's upper-triangular matrix), nothing is removed
//   merge   greedy NMS (the batched kernel of nms.cu), then every kept box becomes the score-weighted mean of the
//           candidates that overlap it, and kept boxes without a second supporter are dropped
//   batched torchvision's coordinate trick: greedy NMS with the class offset (max coordinate of the image + 1); runs on
//           the batched kernel of nms.cu through a per-image offset scale (ay2_nms_batched_scaled)
// IoU values are bit-identical to the reference expression (iou_value); exp / the weighted mean differ from torch by fp32
// rounding only. Launches per batch are fixed (table, column maxima, emit); there is no host synchronisation.
#include "ay2_common.h"
#include "nms_common.cuh"

namespace ay2 {

constexpr int kVThreads = 1024;
constexpr int kVTile = 128;
constexpr int kRow = 8;  // floats per table row

// exclusive prefix of one int per thread over a 1024-thread block; *total = block sum (valid for every thread)
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  __syncthreads();  // s_warp may still be read by the previous call
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  *total = s_warp[31];
  return incl - v + (wid ? s_warp[wid - 1] : 0);
}

// ---------------------------------------------------------------------------------------------------------------
// Candidate table in reference order. One CTA per image walks the rows in chunks of 1024; a block scan gives every
// row its output position, so the table order is the order of the reference's boolean-mask / nonzero() selections.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kVThreads) nmsv_table_kernel(const float* __restrict__ pred, int n, int no, float conf_thres,
                                                               int multi_label, const uint8_t* __restrict__ class_mask, int cap,
                                                               int max_nms, float* __restrict__ table, int* __restrict__ counts,
                                                               float* __restrict__ max_coord, int* __restrict__ flags) {
  __shared__ int s_warp[32];
  __shared__ float s_max[32];
  const int b = blockIdx.x, nc = no - 5;
  const float* ip = pred + (long long)b * n * no;
  float* tp = table + (long long)b * cap * kRow;
  int base = 0;
  float cmax = -INFINITY;
  for (int r0 = 0; r0 < n; r0 += kVThreads) {
    const int row = r0 + threadIdx.x;
    int cnt = 0, bidx = 0;
    float obj = 0.f, best = -INFINITY;
    const float* rp = ip + (long long)row * no;
    if (row < n) {
      obj = rp[4];
      if (obj > conf_thres) {  // metrics.py:313,337 (the val2 rule needs no such test: conf_c <= obj for class scores in [0, 1])
        if (multi_label) {
          for (int c = 0; c < nc; ++c)
            cnt += (__fmul_rn(rp[5 + c], obj) > conf_thres && (!class_mask || class_mask[c])) ? 1 : 0;
        } else {
          for (int c = 0; c < nc; ++c) {  // first arg-max (torch.max keeps the lowest index among equals)
            const float cf = __fmul_rn(rp[5 + c], obj);
            if (cf > best) {
              best = cf;
              bidx = c;
            }
          }
          cnt = (best > conf_thres && (!class_mask || class_mask[bidx])) ? 1 : 0;
        }
      }
    }
    int total;
    int pos = base + block_exclusive_scan(cnt, s_warp, &total);
    if (cnt) {
      const float4 bx = xywh_to_xyxy(make_float4(rp[0], rp[1], rp[2], rp[3]));
      cmax = fmaxf(cmax, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
      if (multi_label) {
        for (int c = 0; c < nc; ++c) {
          const float cf = __fmul_rn(rp[5 + c], obj);
          if (cf > conf_thres && (!class_mask || class_mask[c])) {
            if (pos < cap) {
              float4* o = reinterpret_cast<float4*>(tp + (long long)pos * kRow);
              o[0] = bx;
              o[1] = make_float4(cf, (float)c, 0.f, 0.f);
            }
            ++pos;
          }
        }
      } else if (pos < cap) {
        float4* o = reinterpret_cast<float4*>(tp + (long long)pos * kRow);
        o[0] = bx;
        o[1] = make_float4(best, (float)bidx, 0.f, 0.f);
      }
    }
    base += total;
  }
  // block maximum of the candidate coordinates (torchvision's batched_nms offsets classes by max coordinate + 1)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = cmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_max[0];
    for (int w = 1; w < 32; ++w) m = fmaxf(m, s_max[w]);
    counts[b] = base < cap ? base : cap;
    max_coord[b] = m;
    int f = 0;
    if (base > cap) f |= 1;       // the table is truncated: the caller must retry with a larger capacity
    if (base > max_nms) f |= 2;   // metrics.py:378-379 applies: keep the max_nms best rows (the caller re-ranks the table)
    if (f) atomicOr(flags, f);
  }
}

__device__ __forceinline__ float4 offset_box(const float4 bx, const float cls, const float scale) {
  const float off = __fmul_rn(cls, scale);  // `c * max_wh` (metrics.py:398) / `cls * 4096` (nms.py:75,84)
  return make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
}

// colmax[j] = max_{i<j} IoU(i, j) over class-offset boxes: the column maximum of the reference's triu_(diagonal=1) matrix
__global__ void __launch_bounds__(kVTile) nmsv_colmax_kernel(const float* __restrict__ table, const int* __restrict__ counts, int cap,
                                                             float class_offset, float* __restrict__ colmax) {
  __shared__ float4 s_box[kVTile];
  __shared__ float s_area[kVTile];
  const int b = blockIdx.y;
  const int n = counts[b];
  const int j0 = blockIdx.x * kVTile;
  if (j0 >= n) return;
  const float* tp = table + (long long)b * cap * kRow;
  const int j = j0 + threadIdx.x;
  float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
  float aj = 0.f;
  if (j < n) {
    const float4* r = reinterpret_cast<const float4*>(tp + (long long)j * kRow);
    bj = offset_box(r[0], r[1].y, class_offset);
    aj = box_area(bj);
  }
  float m = 0.0f;
  for (int i0 = 0; i0 < j0 + kVTile && i0 < n; i0 += kVTile) {
    const int i = i0 + threadIdx.x;
    if (i < n) {
      const float4* r = reinterpret_cast<const float4*>(tp + (long long)i * kRow);
      const float4 bi = offset_box(r[0], r[1].y, class_offset);
      s_box[threadIdx.x] = bi;
      s_area[threadIdx.x] = box_area(bi);
    }
    __syncthreads();
    const int lim = min(kVTile, min(n, j) - i0);  // rows i < j only
    for (int ii = 0; ii < lim; ++ii) m = fmaxf(m, iou_value(s_box[ii], s_area[ii], bj, aj));
    __syncthreads();
  }
  if (j < n) colmax[(long long)b * cap + j] = m;
}

// fast NMS: rows with colmax < thr, in table order, first out_cap of them
__global__ void __launch_bounds__(kVThreads) nmsv_fast_emit_kernel(const float* __restrict__ table, const int* __restrict__ counts,
                                                                   int cap, const float* __restrict__ colmax, float thr,
                                                                   float* __restrict__ out, int out_cap, int* __restrict__ out_count) {
  __shared__ int s_warp[32];
  const int b = blockIdx.x;
  const int n = counts[b];
  const float* tp = table + (long long)b * cap * kRow;
  int base = 0;
  for (int j0 = 0; j0 < n && base < out_cap; j0 += kVThreads) {
    const int j = j0 + threadIdx.x;
    const int keep = (j < n && colmax[(long long)b * cap + j] < thr) ? 1 : 0;
    int total;
    const int pos = base + block_exclusive_scan(keep, s_warp, &total);
    if (keep && pos < out_cap) {
      const float4* r = reinterpret_cast<const float4*>(tp + (long long)j * kRow);
      float* o = out + ((long long)b * out_cap + pos) * 6;
      o[0] = r[0].x, o[1] = r[0].y, o[2] = r[0].z, o[3] = r[0].w, o[4] = r[1].x, o[5] = r[1].y;
    }
    base += total;
  }
  if (threadIdx.x == 0) out_count[b] = base < out_cap ? base : out_cap;
}

// matrix NMS: conf_j *= min_i exp(-(IoU(i,j)^2 - m_i^2) / 0.5), IoU(i, j) = 0 for i >= j; the first out_cap rows are emitted
__global__ void __launch_bounds__(kVTile) nmsv_decay_kernel(const float* __restrict__ table, const int* __restrict__ counts, int cap,
                                                            float class_offset, const float* __restrict__ colmax,
                                                            float* __restrict__ out, int out_cap, int* __restrict__ out_count) {
  __shared__ float4 s_box[kVTile];
  __shared__ float s_area[kVTile], s_m2[kVTile];
  const int b = blockIdx.y;
  const int n = counts[b];
  const int j0 = blockIdx.x * kVTile;
  if (blockIdx.x == 0 && threadIdx.x == 0) out_count[b] = n < out_cap ? n : out_cap;
  if (j0 >= n || j0 >= out_cap) return;
  const float* tp = table + (long long)b * cap * kRow;
  const int j = j0 + threadIdx.x;
  float4 raw = make_float4(0.f, 0.f, 0.f, 0.f), bj = raw;
  float conf = 0.f, cls = 0.f, aj = 0.f;
  if (j < n) {
    const float4* r = reinterpret_cast<const float4*>(tp + (long long)j * kRow);
    raw = r[0];
    conf = r[1].x;
    cls = r[1].y;
    bj = offset_box(raw, cls, class_offset);
    aj = box_area(bj);
  }
  float d = INFINITY;
  for (int i0 = 0; i0 < n; i0 += kVTile) {
    const int i = i0 + threadIdx.x;
    if (i < n) {
      const float4* r = reinterpret_cast<const float4*>(tp + (long long)i * kRow);
      const float4 bi = offset_box(r[0], r[1].y, class_offset);
      s_box[threadIdx.x] = bi;
      s_area[threadIdx.x] = box_area(bi);
      const float m = colmax[(long long)b * cap + i];
      s_m2[threadIdx.x] = __fmul_rn(m, m);
    }
    __syncthreads();
    const int lim = min(kVTile, n - i0);
    for (int ii = 0; ii < lim; ++ii) {
      const float iou = (i0 + ii < j) ? iou_value(s_box[ii], s_area[ii], bj, aj) : 0.0f;
      const float t = __fdiv_rn(__fsub_rn(__fmul_rn(iou, iou), s_m2[ii]), 0.5f);
      d = fminf(d, expf(-t));
    }
    __syncthreads();
  }
  if (j < n && j < out_cap) {
    float* o = out + ((long long)b * out_cap + j) * 6;
    o[0] = raw.x, o[1] = raw.y, o[2] = raw.z, o[3] = raw.w, o[4] = __fmul_rn(conf, d), o[5] = cls;
  }
}

// merge NMS: det rows (the greedy survivors, in kept order) <- weighted box means; rows supported by one box only are dropped
__global__ void __launch_bounds__(kVThreads) nmsv_merge_kernel(const float* __restrict__ table, const int* __restrict__ counts,
                                                               int cap, float class_offset, float thr, int n_min_excl,
                                                               int n_max_excl, float* __restrict__ det, int det_cap,
                                                               int* __restrict__ det_count) {
  __shared__ float4 s_new[kVThreads];
  __shared__ unsigned char s_keep[kVThreads];
  __shared__ int s_warp[32];
  const int b = blockIdx.x;
  const int n = counts[b];
  const int kept = min(det_count[b], min(det_cap, kVThreads));
  if (!(n > n_min_excl && n < n_max_excl)) return;  // metrics.py:420 `if 1 < n < 3e3` (nms.py merges unconditionally)
  const float* tp = table + (long long)b * cap * kRow;
  float* dp = det + (long long)b * det_cap * 6;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = wid; k < kept; k += kVThreads / 32) {
    const float4 kb = offset_box(make_float4(dp[k * 6 + 0], dp[k * 6 + 1], dp[k * 6 + 2], dp[k * 6 + 3]), dp[k * 6 + 5], class_offset);
    const float ka = box_area(kb);
    double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, sw = 0;
    int hits = 0;
    for (int j = lane; j < n; j += 32) {
      const float4* r = reinterpret_cast<const float4*>(tp + (long long)j * kRow);
      const float4 raw = r[0];
      const float4 bj = offset_box(raw, r[1].y, class_offset);
      if (iou_value(kb, ka, bj, box_area(bj)) > thr) {
        const double w = r[1].x;
        sw += w, sx1 += w * raw.x, sy1 += w * raw.y, sx2 += w * raw.z, sy2 += w * raw.w;
        ++hits;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      sx1 += __shfl_xor_sync(0xffffffffu, sx1, o);
      sy1 += __shfl_xor_sync(0xffffffffu, sy1, o);
      sx2 += __shfl_xor_sync(0xffffffffu, sx2, o);
      sy2 += __shfl_xor_sync(0xffffffffu, sy2, o);
      hits += __shfl_xor_sync(0xffffffffu, hits, o);
    }
    if (lane == 0) {
      const float fw = static_cast<float>(sw);  // torch: mm(weights, boxes).float() / weights.sum(1) in fp32
      s_new[k] = make_float4(__fdiv_rn(static_cast<float>(sx1), fw), __fdiv_rn(static_cast<float>(sy1), fw),
                             __fdiv_rn(static_cast<float>(sx2), fw), __fdiv_rn(static_cast<float>(sy2), fw));
      s_keep[k] = hits > 1 ? 1 : 0;
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  float conf = 0.f, cls = 0.f;
  int keep = 0;
  if (t < kept) {
    conf = dp[t * 6 + 4];
    cls = dp[t * 6 + 5];
    keep = s_keep[t];
  }
  int total;
  const int pos = block_exclusive_scan(keep, s_warp, &total);
  __syncthreads();  // every row has been read before any is overwritten
  if (keep) {
    float* o = dp + pos * 6;
    const float4 nb = s_new[t];
    o[0] = nb.x, o[1] = nb.y, o[2] = nb.z, o[3] = nb.w, o[4] = conf, o[5] = cls;
  }
  if (t == 0) det_count[b] = total;
}

}  // namespace ay2

using namespace ay2;

extern "C" int ay2_nms_candidate_table(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, float* table,
                                       int32_t cap, int32_t* counts, float* max_coord, int32_t* flags, void* stream) {
  AY2_REQUIRE(pred && p && table && counts && max_coord && flags, "ay2_nms_candidate_table: null pointer");
  AY2_REQUIRE(p->batch > 0 && p->n > 0 && p->no > 5 && cap > 0, "ay2_nms_candidate_table: bad shape (batch=%d n=%d no=%d cap=%d)",
              p->batch, p->n, p->no, cap);
  AY2_REQUIRE((reinterpret_cast<uintptr_t>(table) & 15) == 0, "ay2_nms_candidate_table: table must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AY2_CHECK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
  nmsv_table_kernel<<<p->batch, kVThreads, 0, st>>>(pred, p->n, p->no, p->conf_thres, p->multi_label, class_mask, cap, p->max_nms,
                                                    table, counts, max_coord, flags);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

static int variant_checks(const float* table, const int32_t* counts, int32_t batch, int32_t cap, const void* out, const void* out_count) {
  AY2_REQUIRE(table && counts && out && out_count, "ay2_nms variant: null pointer");
  AY2_REQUIRE(batch > 0 && cap > 0, "ay2_nms variant: bad shape (batch=%d cap=%d)", batch, cap);
  return AY2_OK;
}

extern "C" int ay2_nms_fast(const float* table, const int32_t* counts, int32_t batch, int32_t cap, float class_offset,
                            float iou_thres, float* colmax_ws, float* out_det, int32_t out_cap, int32_t* out_count, void* stream) {
  int rc = variant_checks(table, counts, batch, cap, out_det, out_count);
  if (rc != AY2_OK) return rc;
  AY2_REQUIRE(colmax_ws && out_cap > 0, "ay2_nms_fast: missing scratch / output capacity");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nmsv_colmax_kernel<<<dim3(ceil_div(cap, kVTile), batch), kVTile, 0, st>>>(table, counts, cap, class_offset, colmax_ws);
  AY2_CHECK_LAUNCH();
  nmsv_fast_emit_kernel<<<batch, kVThreads, 0, st>>>(table, counts, cap, colmax_ws, iou_thres, out_det, out_cap, out_count);
  AY2_CHECK_LAUNCH();
  count_launch(2);
  return AY2_OK;
}

extern "C" int ay2_nms_matrix(const float* table, const int32_t* counts, int32_t batch, int32_t cap, float class_offset,
                              float* colmax_ws, float* out_det, int32_t out_cap, int32_t* out_count, void* stream) {
  int rc = variant_checks(table, counts, batch, cap, out_det, out_count);
  if (rc != AY2_OK) return rc;
  AY2_REQUIRE(colmax_ws && out_cap > 0, "ay2_nms_matrix: missing scratch / output capacity");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(ceil_div(cap, kVTile), batch);
  nmsv_colmax_kernel<<<grid, kVTile, 0, st>>>(table, counts, cap, class_offset, colmax_ws);
  AY2_CHECK_LAUNCH();
  nmsv_decay_kernel<<<grid, kVTile, 0, st>>>(table, counts, cap, class_offset, colmax_ws, out_det, out_cap, out_count);
  AY2_CHECK_LAUNCH();
  count_launch(2);
  return AY2_OK;
}

extern "C" int ay2_nms_merge(const float* table, const int32_t* counts, int32_t batch, int32_t cap, float class_offset,
                             float iou_thres, int32_t n_min_excl, int32_t n_max_excl, float* det, int32_t det_cap,
                             int32_t* det_count, void* stream) {
  int rc = variant_checks(table, counts, batch, cap, det, det_count);
  if (rc != AY2_OK) return rc;
  AY2_REQUIRE(det_cap >= 1 && det_cap <= kVThreads, "ay2_nms_merge: det_cap=%d unsupported (1..%d)", det_cap, kVThreads);
  nmsv_merge_kernel<<<batch, kVThreads, 0, static_cast<cudaStream_t>(stream)>>>(table, counts, cap, class_offset, iou_thres,
                                                                                n_min_excl, n_max_excl, det, det_cap, det_count);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
