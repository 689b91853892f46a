// YOLOHead decode arithmetic shared by the dense decode kernel and the fused logits->NMS path, written with
// explicit rounding intrinsics so both paths produce bit-identical scores and boxes.
//   y = sigmoid(t); xy = (y*2 - 0.5 + grid) * stride; wh = (y*2)^2 * anchor_px      
#pragma once
#include <cuda_runtime.h>

namespace ay2 {

// ex2.approx + rcp.approx: 2 MUFU ops, ~2 ulp; the SAME function feeds the dense decode and the fused NMS filter
__device__ __forceinline__ float head_sigmoid(float t) { return __fdividef(1.0f, __fadd_rn(1.0f, __expf(-t))); }
__device__ __forceinline__ float head_xy(float s, float g, float stride) {
  return __fmul_rn(__fadd_rn(__fmaf_rn(s, 2.0f, -0.5f), g), stride);  // s*2 is exact, so the fma == mul, sub
}
__device__ __forceinline__ float head_wh(float s, float anchor_px) {
  const float q = __fmul_rn(s, 2.0f);
  return __fmul_rn(__fmul_rn(q, q), anchor_px);
}

}  // namespace ay2
