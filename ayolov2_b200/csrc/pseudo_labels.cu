// Teacher detections -> pseudo-labels of the knowledge-distillation trainer, for a whole batch in one launch.
//
// Replaces SoftTeacherTrainer.prepare_labels_for_augmention / filter_invalid and the non-augmenting branch of
// get_pseudo_labeled_batch (scripts/train/kd_trainer.py:385-397, 436-487): per image, on the host in the reference,
// keep detections with score > thr and box width / height > min_size, divide by the image size, clip to [0, 1],
// xyxy -> xywh with the validity correction of scripts/utils/general.py:250-295, prepend class and image index.
// Input is the NMS output buffer as the NMS kernels leave it ([batch][max_det][6] + counts), output the (N, 6) label
// tensor ComputeLoss consumes, rows in image order then detection order (an ordered compaction: one CTA, chunked scan --
// batch * max_det is ~19,000 slots).
#include "ay2_common.h"

namespace ay2 {

__device__ __forceinline__ float unit_clip(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
// numpy: float32 array /= int64 array computes in double and rounds once to float32
__device__ __forceinline__ float div_like_numpy(float v, float size) { return __double2float_rn(__ddiv_rn((double)v, (double)size)); }

__global__ void __launch_bounds__(1024) pseudo_labels_kernel(const float* __restrict__ det, const int* __restrict__ counts, int batch,
                                                             int max_det, float thr, float min_size, int use_min_size, float width,
                                                             float height, float* __restrict__ labels, int* __restrict__ out_counts) {
  __shared__ int warp_sums[32];
  __shared__ int base, chunk_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  for (int b = tid; b < batch; b += blockDim.x) out_counts[b] = 0;
  __syncthreads();
  const int total = batch * max_det;
  for (int start = 0; start < total; start += blockDim.x) {
    const int slot = start + tid;
    const int b = slot < total ? slot / max_det : 0;
    const int d = slot - b * max_det;
    bool keep = false;
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f, cls = 0.f;
    if (slot < total && d < min(counts[b], max_det)) {
      const float* r = det + (size_t)slot * 6;
      x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3], cls = r[5];
      keep = r[4] > thr;                                                               // kd_trainer.py:473
      if (use_min_size) keep = keep && __fsub_rn(x2, x1) > min_size && __fsub_rn(y2, y1) > min_size;  // :479-482
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sums[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      const int v = warp_sums[lane];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      warp_sums[lane] = incl - v;  // exclusive prefix of the warp totals
      if (lane == 31) chunk_total = incl;
    }
    __syncthreads();
    const int pos = base + warp_sums[warp] + __popc(m & ((1u << lane) - 1u));
    if (keep) {
      // kd_trainer.py:458-461: normalise, clip; general.py:281-293: centre / size, validity correction, clip(1e-12, 1)
      const float a = unit_clip(div_like_numpy(x1, width)), bq = unit_clip(div_like_numpy(y1, height));
      const float c = unit_clip(div_like_numpy(x2, width)), e = unit_clip(div_like_numpy(y2, height));
      const float xc = __fdiv_rn(__fadd_rn(a, c), 2.0f), yc = __fdiv_rn(__fadd_rn(bq, e), 2.0f);
      float w = __fsub_rn(c, a), h = __fsub_rn(e, bq);
      w = __fadd_rn(w, __fmul_rn(fminf(__fsub_rn(xc, __fdiv_rn(w, 2.0f)), 0.0f), 2.0f));
      w = __fsub_rn(w, __fmul_rn(__fsub_rn(fmaxf(__fadd_rn(xc, __fdiv_rn(w, 2.0f)), 1.0f), 1.0f), 2.0f));
      h = __fadd_rn(h, __fmul_rn(fminf(__fsub_rn(yc, __fdiv_rn(h, 2.0f)), 0.0f), 2.0f));
      h = __fsub_rn(h, __fmul_rn(__fsub_rn(fmaxf(__fadd_rn(yc, __fdiv_rn(h, 2.0f)), 1.0f), 1.0f), 2.0f));
      float* o = labels + (size_t)pos * 6;
      o[0] = (float)b, o[1] = cls;
      o[2] = fminf(fmaxf(xc, 1e-12f), 1.0f), o[3] = fminf(fmaxf(yc, 1e-12f), 1.0f);
      o[4] = fminf(fmaxf(w, 1e-12f), 1.0f), o[5] = fminf(fmaxf(h, 1e-12f), 1.0f);
      atomicAdd(&out_counts[b], 1);
    }
    __syncthreads();
    if (tid == 0) base += chunk_total;
    __syncthreads();
  }
  if (tid == 0) out_counts[batch] = base;
}

}  // namespace ay2

using namespace ay2;

extern "C" int ay2_pseudo_labels(const float* det, const int32_t* counts, int32_t batch, int32_t max_det, float score_thr,
                                 float min_size, int32_t use_min_size, float width, float height, float* labels,
                                 int32_t* out_counts, void* stream) {
  AY2_REQUIRE(det && counts && labels && out_counts && batch >= 0 && max_det > 0, "ay2_pseudo_labels: bad arguments");
  AY2_REQUIRE(width > 0.f && height > 0.f, "ay2_pseudo_labels: image size %g x %g invalid", width, height);
  AY2_REQUIRE((long long)batch * max_det < (1ll << 30), "ay2_pseudo_labels: batch * max_det too large");
  pseudo_labels_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(det, counts, batch, max_det, score_thr, min_size, use_min_size,
                                                                          width, height, labels, out_counts);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
