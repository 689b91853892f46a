// Device helpers shared by the NMS translation units (nms.cu, nms_variants.cu).
#pragma once
#include <cuda_runtime.h>

namespace ay2 {

// IoU > iou_thres exactly as torchvision's CPU kernel evaluates it: fp32 arithmetic, w/h clamped at 0, RN division,
// comparison against the double threshold. The double comparison is folded into `thr` = the largest float <= iou_thres
// ((double)x > T  <=>  x > thr for every float x). The division is avoided for all but a sliver of pairs:
// with q = inter/uni (real), inter > uni*thr*(1+2^-20) => q > thr + ulp(thr) => RN(q) > thr, and
// inter < uni*thr*(1-2^-20) => q < thr => RN(q) <= thr (RN is monotone, thr is a float); the roundings of hi/lo are
// < 2^-22 relative as long as uni is a normal positive number far from overflow (`fin`). Branch-free up to that sliver.
struct IouThr {
  float thr, hi, lo;
};
__device__ __forceinline__ IouThr make_iou_thr(double iou_thres) {
  IouThr t;
  t.thr = static_cast<float>(iou_thres);
  if (static_cast<double>(t.thr) > iou_thres) t.thr = nextafterf(t.thr, -INFINITY);
  const bool filt = t.thr > 1e-30f && t.thr < 1e30f;
  t.hi = filt ? __fmul_rn(t.thr, 1.000001f) : INFINITY;
  t.lo = filt ? __fmul_rn(t.thr, 0.999999f) : -INFINITY;
  return t;
}
__device__ __forceinline__ bool iou_gt(const float4 a, const float aa, const float4 b, const float ab, const IouThr t) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
  const bool fin = (__float_as_uint(uni) - 0x0D800000u) < 0x64000000u;  // 2^-100 <= uni < 2^100 (positive, finite)
  if (fin && inter > __fmul_rn(uni, t.hi)) return true;
  if (fin && inter < __fmul_rn(uni, t.lo)) return false;
  return __fdiv_rn(inter, uni) > t.thr;
}


// scripts/utils/general.py:316-319 with ratio = wh = 1, pad = 0: 1*1*(x -+ w/2) + 0, every operation rounded in fp32
__device__ __forceinline__ float4 xywh_to_xyxy(const float4 r) {
  float4 bx;
  bx.x = __fadd_rn(__fsub_rn(r.x, __fmul_rn(r.z, 0.5f)), 0.0f);
  bx.y = __fadd_rn(__fsub_rn(r.y, __fmul_rn(r.w, 0.5f)), 0.0f);
  bx.z = __fadd_rn(__fadd_rn(r.x, __fmul_rn(r.z, 0.5f)), 0.0f);
  bx.w = __fadd_rn(__fadd_rn(r.y, __fmul_rn(r.w, 0.5f)), 0.0f);
  return bx;
}

// metrics.py:138-164 box_iou, same fp32 operation order (IEEE division): bit-identical to the torch expression
__device__ __forceinline__ float iou_value(const float4 a, const float aa, const float4 b, const float ab) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
}
__device__ __forceinline__ float box_area(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

}  // namespace ay2
