// Validation statistics on the GPU: detections -> "correct at IoU threshold k" matrix, for a whole batch in one launch.
//
// Replaces the per-image host loop of YoloValidator.statistics_per_image / process_batch
// (scripts/utils/train_utils.py:294-401), which moves every image's IoU matches to the CPU (.cpu().numpy(), :319-329)
// right after NMS. Per image:
//   1. detections and labels are mapped from the letterboxed network input back to the native image
//      (scale_coords + clip_coords, scripts/utils/general.py:203-230,324-358; labels go through xywh2xyxy first,
//      general.py:316-319) -- optional, the plain process_batch(detections, labels) contract skips it;
//   2. iou = box_iou(labels, detections) (scripts/utils/metrics.py:138-164), candidates = iou >= iouv[0] & same class;
//   3. the reference sorts the candidate pairs by IoU (descending), keeps the first pair of every detection, then -- in
//      DETECTION order, the second sort is commented out (:324) -- the first pair of every label. Equivalent, without
//      the sort: every detection picks its best label, every label keeps the lowest-index detection that picked it;
//   4. correct[d][k] = iou(d) >= iouv[k] for the matched detections.
// One CTA per image; a thread owns one detection (strided). fp32 arithmetic in the reference's operation order.
#include "ay2_common.h"

namespace ay2 {

struct MatchParams {
  const float* det;      // [B][max_det][6] x1 y1 x2 y2 conf cls  (the NMS output buffer)
  const int* counts;     // [B]
  const float* labels;   // [T][6] image, class, then xyxy (scale == 0) or xywh pixels of the network input (scale == 1)
  const float* meta;     // [B][5] gain, pad_x, pad_y, native width, native height (scale == 1)
  const float* iouv;     // [niou]
  unsigned char* correct;  // [B][max_det][niou]
  int batch, max_det, nt, niou, scale, lcap;
};

__device__ __forceinline__ float4 to_native(float4 b, float gain, float px, float py, float w0, float h0) {
  // general.py:354-357: x -= pad_x, y -= pad_y, / gain, clamp to [0, w0] x [0, h0]
  b.x = __fdiv_rn(__fsub_rn(b.x, px), gain);
  b.z = __fdiv_rn(__fsub_rn(b.z, px), gain);
  b.y = __fdiv_rn(__fsub_rn(b.y, py), gain);
  b.w = __fdiv_rn(__fsub_rn(b.w, py), gain);
  b.x = fminf(fmaxf(b.x, 0.0f), w0);
  b.z = fminf(fmaxf(b.z, 0.0f), w0);
  b.y = fminf(fmaxf(b.y, 0.0f), h0);
  b.w = fminf(fmaxf(b.w, 0.0f), h0);
  return b;
}

__global__ void val_match_kernel(MatchParams p) {
  extern __shared__ unsigned char vm_smem[];
  float4* lbox = reinterpret_cast<float4*>(vm_smem);            // [lcap]
  float* lcls = reinterpret_cast<float*>(lbox + p.lcap);        // [lcap]
  int* lfirst = reinterpret_cast<int*>(lcls + p.lcap);          // [lcap] lowest detection index that picked the label
  int* best_l = lfirst + p.lcap;                                // [max_det]
  float* best_iou = reinterpret_cast<float*>(best_l + p.max_det);  // [max_det]
  __shared__ int s_nl;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int nd = min(p.counts[b], p.max_det);
  float gain = 1.f, px = 0.f, py = 0.f, w0 = 0.f, h0 = 0.f;
  if (p.scale) {
    const float* m = p.meta + b * 5;
    gain = m[0], px = m[1], py = m[2], w0 = m[3], h0 = m[4];
  }
  if (tid == 0) s_nl = 0;
  __syncthreads();
  // labels of this image, in their original order (warp 0: ordered compaction)
  if (tid < 32) {
    int base = 0;
    for (int t0 = 0; t0 < p.nt; t0 += 32) {
      const int t = t0 + lane;
      const bool mine = t < p.nt && static_cast<int>(p.labels[t * 6]) == b;
      const unsigned mk = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        const int pos = base + __popc(mk & ((1u << lane) - 1u));
        if (pos < p.lcap) {
          const float* L = p.labels + t * 6;
          float4 bx;
          if (p.scale) {  // xywh -> xyxy (general.py:316-319), then the native-image mapping
            const float hw = __fdiv_rn(L[4], 2.0f), hh = __fdiv_rn(L[5], 2.0f);
            bx = to_native(make_float4(__fsub_rn(L[2], hw), __fsub_rn(L[3], hh), __fadd_rn(L[2], hw), __fadd_rn(L[3], hh)), gain, px, py,
                           w0, h0);
          } else {
            bx = make_float4(L[2], L[3], L[4], L[5]);
          }
          lbox[pos] = bx;
          lcls[pos] = L[1];
          lfirst[pos] = 0x7fffffff;
        }
      }
      base += __popc(mk);
    }
    if (lane == 0) s_nl = min(base, p.lcap);
  }
  __syncthreads();
  const int nl = s_nl;
  const float thr0 = p.iouv[0];
  for (int d = tid; d < nd; d += blockDim.x) {
    const float* D = p.det + ((size_t)b * p.max_det + d) * 6;
    float4 db = make_float4(D[0], D[1], D[2], D[3]);
    if (p.scale) db = to_native(db, gain, px, py, w0, h0);
    const float dc = D[5];
    const float area_d = __fmul_rn(__fsub_rn(db.z, db.x), __fsub_rn(db.w, db.y));
    int bl = -1;
    float bi = -1.0f;
    for (int l = 0; l < nl; ++l) {
      if (lcls[l] != dc) continue;
      const float4 a = lbox[l];
      const float area_l = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
      const float w = fmaxf(__fsub_rn(fminf(a.z, db.z), fmaxf(a.x, db.x)), 0.0f);
      const float h = fmaxf(__fsub_rn(fminf(a.w, db.w), fmaxf(a.y, db.y)), 0.0f);
      const float inter = __fmul_rn(w, h);
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_l, area_d), inter));
      if (iou >= thr0 && iou >= bi) {  // best IoU; equal IoUs keep the later label (the reference's reversed argsort)
        bi = iou;
        bl = l;
      }
    }
    best_l[d] = bl;
    best_iou[d] = bi;
    if (bl >= 0) atomicMin(&lfirst[bl], d);
  }
  __syncthreads();
  for (int d = tid; d < p.max_det; d += blockDim.x) {
    const bool matched = d < nd && best_l[d] >= 0 && lfirst[best_l[d]] == d;
    const float iou = matched ? best_iou[d] : -1.0f;
    unsigned char* c = p.correct + ((size_t)b * p.max_det + d) * p.niou;
    for (int k = 0; k < p.niou; ++k) c[k] = iou >= p.iouv[k] ? 1 : 0;
  }
}

}  // namespace ay2

using namespace ay2;

extern "C" int ay2_match_detections(const float* det, const int32_t* counts, int32_t batch, int32_t max_det, const float* labels,
                                    int32_t nt, int32_t labels_cap, const float* meta, const float* iouv, int32_t niou,
                                    uint8_t* correct, void* stream) {
  AY2_REQUIRE(det && counts && iouv && correct && batch >= 0 && max_det > 0 && niou > 0, "ay2_match_detections: bad arguments");
  AY2_REQUIRE(nt == 0 || labels, "ay2_match_detections: labels missing");
  if (batch == 0) return AY2_OK;
  MatchParams p;
  p.det = det, p.counts = counts, p.labels = labels, p.meta = meta, p.iouv = iouv, p.correct = correct;
  p.batch = batch, p.max_det = max_det, p.nt = nt, p.niou = niou, p.scale = meta != nullptr;
  p.lcap = labels_cap > 0 ? labels_cap : 1;
  const size_t smem = (size_t)p.lcap * (16 + 4 + 4) + (size_t)max_det * 8;
  AY2_REQUIRE(smem <= 200 * 1024, "ay2_match_detections: %d labels per image / %d detections do not fit in shared memory", p.lcap, max_det);
  if (smem > 48 * 1024) AY2_CHECK_CUDA(cudaFuncSetAttribute(val_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  val_match_kernel<<<batch, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
