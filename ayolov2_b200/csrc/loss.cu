// YOLO detection loss (CIoU box + BCE objectness + BCE class) forward and analytic backward on the GPU.
//
// Reference: scripts/loss/losses.py:223-391 (ComputeLoss.__call__ / build_targets) and
// scripts/utils/metrics.py:60-135 (bbox_iou, CIoU branch); default configuration (fl_gamma 0, gr 1,
// autobalance off). The reference runs ~60 tiny kernels with boolean-mask compactions (host syncs) per level;
// here the whole loss is 4 kernels + one memset, no host synchronisation:
//   1. loss_assign : one thread per candidate (level, offset o, anchor a, target t). A candidate is a match iff
//                    the anchor ratio test passes (losses.py:356-357) and its neighbour-cell condition holds
//                    (:364-367). Counts matches per level and resolves duplicate cells the way the reference's
//                    sequential index_put does (last candidate in (o, a, t) order wins, :273) via atomicMax.
//   2. loss_match  : one warp per match. Gathers the 5+nc logits, CIoU (fp32), class BCE over lanes; accumulates
//                    the partial sums and scatters d(loss)/d(logits) with atomics (duplicates accumulate like
//                    autograd's index backward); the owning match writes tobj = clamp(iou, 0).
//   3. loss_obj    : dense BCE(pred[..., 4], tobj) over every cell of a level (+ its gradient).
//   4. loss_final  : lbox/lobj/lcls/loss exactly as losses.py:294-300.
// Gradients are scaled by a device-resident scalar (autograd's grad_output) so AMP / DDP scaling needs no sync.
#include <string.h>

#include "ay2_common.h"
#include "ay2_ptx.cuh"

namespace ay2 {

struct LossLevel {
  const float* pred;  // (bs, na, ny, nx, no)
  float* grad;        // same shape or nullptr
  float* tobj;        // (bs, na, ny, nx)
  int* owner;         // (bs, na, ny, nx)
  int ny, nx;
  float balance;
};

struct LossKernelParams {
  LossLevel lv[AY2_LOSS_MAX_LEVELS];
  const float* targets;  // (nt, 6)
  const float* anchors;  // (nl, na, 2) grid units
  const float* gscale;   // device scalar or nullptr (= 1)
  int nl, na, nc, no, bs, nt;
  float anchor_t, hbox, hobj, hcls, cls_pw, obj_pw, cp, cn, fl_gamma, fl_alpha;
  double* acc;  // [nl][3] : sum(1 - ciou), sum(cls bce), sum(obj bce)
  int* count;   // [nl]
};

struct Match {
  bool valid;
  int b, cls, gi, gj, a;
  float tx, ty, tw, th;  // tbox
  float aw, ah;
  int rank;
};

// python-style float remainder for positive divisor 1
__device__ __forceinline__ float mod1(float x) {
  float r = fmodf(x, 1.0f);
  if (r != 0.0f && r < 0.0f) r += 1.0f;
  return r;
}

__device__ __forceinline__ Match make_match(const LossKernelParams& p, int level, int cand) {
  Match m;
  m.valid = false;
  const int per_o = p.na * p.nt;
  const int o = cand / per_o;
  const int a = (cand - o * per_o) / p.nt;
  const int t = cand - o * per_o - a * p.nt;
  const float* tg = p.targets + (size_t)t * 6;
  const float nx = (float)p.lv[level].nx, ny = (float)p.lv[level].ny;
  const float gx = tg[2] * nx, gy = tg[3] * ny, gw = tg[4] * nx, gh = tg[5] * ny;
  const float aw = p.anchors[(level * p.na + a) * 2 + 0], ah = p.anchors[(level * p.na + a) * 2 + 1];
  const float rw = __fdiv_rn(gw, aw), rh = __fdiv_rn(gh, ah);
  const float mr = fmaxf(fmaxf(rw, __fdiv_rn(1.0f, rw)), fmaxf(rh, __fdiv_rn(1.0f, rh)));
  if (!(mr < p.anchor_t)) return m;
  float offx = 0.f, offy = 0.f;
  const float g = 0.5f;
  if (o == 1) {
    if (!(mod1(gx) < g && gx > 1.0f)) return m;
    offx = g;
  } else if (o == 2) {
    if (!(mod1(gy) < g && gy > 1.0f)) return m;
    offy = g;
  } else if (o == 3) {
    const float gxi = nx - gx;
    if (!(mod1(gxi) < g && gxi > 1.0f)) return m;
    offx = -g;
  } else if (o == 4) {
    const float gyi = ny - gy;
    if (!(mod1(gyi) < g && gyi > 1.0f)) return m;
    offy = -g;
  }
  const int gi_raw = (int)(gx - offx);  // .long(): truncation toward zero
  const int gj_raw = (int)(gy - offy);
  m.valid = true;
  m.b = (int)tg[0];
  m.cls = (int)tg[1];
  m.a = a;
  m.gi = min(max(gi_raw, 0), p.lv[level].nx - 1);
  m.gj = min(max(gj_raw, 0), p.lv[level].ny - 1);
  m.tx = gx - (float)gi_raw;
  m.ty = gy - (float)gj_raw;
  m.tw = gw;
  m.th = gh;
  m.aw = aw;
  m.ah = ah;
  m.rank = cand;  // (o, a, t) lexicographic == the reference's row order after the repeats/masks
  return m;
}

__global__ void loss_assign_kernel(LossKernelParams p) {
  const int level = blockIdx.y;
  const int ncand = 5 * p.na * p.nt;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncand; c += gridDim.x * blockDim.x) {
    const Match m = make_match(p, level, c);
    if (!m.valid) continue;
    if (m.b < 0 || m.b >= p.bs) continue;
    atomicAdd(&p.count[level], 1);
    const LossLevel& L = p.lv[level];
    const size_t cell = (((size_t)m.b * p.na + m.a) * L.ny + m.gj) * L.nx + m.gi;
    atomicMax(&L.owner[cell], m.rank);
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
// One element of BCEWithLogits(pos_weight) and its derivative; with gamma > 0 the FocalLoss wrapper (losses.py:64-114):
// el * alpha_t * (1 - p_t)^gamma, p_t = t p + (1 - t)(1 - p), alpha_t = t a + (1 - t)(1 - a); d/dx by the product rule,
// d p_t / dx = (2t - 1) p (1 - p).
__device__ __forceinline__ float bce_elem(float x, float t, float pw, float gamma, float alpha, bool want_grad, float* grad) {
  const float bce = pw * t * softplusf_(-x) + (1.0f - t) * softplusf_(x);
  if (gamma <= 0.0f) {
    if (want_grad) *grad = sigmoidf_(x) * (1.0f - t + pw * t) - pw * t;
    return bce;
  }
  const float s = sigmoidf_(x);
  const float q = 1.0f - (t * s + (1.0f - t) * (1.0f - s));
  const float af = t * alpha + (1.0f - t) * (1.0f - alpha);
  const float m = powf(q, gamma);
  if (want_grad) {
    const float dbce = s * (1.0f - t + pw * t) - pw * t;
    const float dm = q > 0.0f ? -gamma * powf(q, gamma - 1.0f) * (2.0f * t - 1.0f) * s * (1.0f - s) : 0.0f;
    *grad = af * (dbce * m + bce * dm);
  }
  return bce * af * m;
}
// d min(a,b)/da and d max(a,b)/da with torch's even split on ties
__device__ __forceinline__ float dmin_a(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float dmax_a(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

__global__ void loss_match_kernel(LossKernelParams p) {
  const int level = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ncand = 5 * p.na * p.nt;
  const LossLevel& L = p.lv[level];
  const int n = p.count[level];
  const float gs = p.gscale ? *p.gscale : 1.0f;
  const float eps = 1e-7f;
  for (int c = warp; c < ncand; c += nwarps) {
    const Match m = make_match(p, level, c);
    if (!m.valid || m.b < 0 || m.b >= p.bs) continue;
    const size_t cell = (((size_t)m.b * p.na + m.a) * L.ny + m.gj) * L.nx + m.gi;
    const float* ps = L.pred + cell * p.no;
    float* gp = L.grad ? L.grad + cell * p.no : nullptr;
    // ---------------- box: CIoU(pbox, tbox) (metrics.py:84-130, x1y1x2y2=False)
    const float s0 = sigmoidf_(ps[0]), s1 = sigmoidf_(ps[1]), s2 = sigmoidf_(ps[2]), s3 = sigmoidf_(ps[3]);
    const float px = s0 * 2.0f - 0.5f, py = s1 * 2.0f - 0.5f;
    const float pw = (s2 * 2.0f) * (s2 * 2.0f) * m.aw, ph = (s3 * 2.0f) * (s3 * 2.0f) * m.ah;
    const float b1x1 = px - pw / 2, b1x2 = px + pw / 2, b1y1 = py - ph / 2, b1y2 = py + ph / 2;
    const float b2x1 = m.tx - m.tw / 2, b2x2 = m.tx + m.tw / 2, b2y1 = m.ty - m.th / 2, b2y2 = m.ty + m.th / 2;
    const float iw_raw = fminf(b1x2, b2x2) - fmaxf(b1x1, b2x1), ih_raw = fminf(b1y2, b2y2) - fmaxf(b1y1, b2y1);
    const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
    const float inter = iw * ih;
    const float w1 = b1x2 - b1x1, h1 = b1y2 - b1y1 + eps, w2 = b2x2 - b2x1, h2 = b2y2 - b2y1 + eps;
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(b1x2, b2x2) - fminf(b1x1, b2x1), ch = fmaxf(b1y2, b2y2) - fminf(b1y1, b2y1);
    const float c2 = cw * cw + ch * ch + eps;
    const float sx = b2x1 + b2x2 - b1x1 - b1x2, sy = b2y1 + b2y2 - b1y1 - b1y2;
    const float rho2 = (sx * sx + sy * sy) / 4;
    const float kv = 4.0f / (3.14159265358979323846f * 3.14159265358979323846f);
    const float q1 = w1 / h1, dA = atanf(w2 / h2) - atanf(q1);
    const float v = kv * dA * dA;
    const float alpha = v / (v - iou + (1 + eps));
    const float ciou = iou - (rho2 / c2 + v * alpha);
    if (lane == 0) {
      atomicAdd(&p.acc[level * 3 + 0], (double)(1.0f - ciou));
      if (L.owner[cell] == m.rank) L.tobj[cell] = fmaxf(ciou, 0.f);
    }
    if (gp && lane == 0) {
      // reverse-mode through the CIoU expression; upstream d(total)/d(ciou) = -hbox * bs * gs / n
      const float up = -p.hbox * (float)p.bs * gs / (float)n;
      const float g_iou = up, g_rho2 = -up / c2, g_c2 = up * rho2 / (c2 * c2), g_v = -up * alpha;
      float g_inter = g_iou / uni;
      const float g_uni = -g_iou * inter / (uni * uni);
      float g_w1 = g_uni * h1, g_h1 = g_uni * w1;
      g_inter -= g_uni;
      const float g_iw = iw_raw >= 0.f ? g_inter * ih : 0.f, g_ih = ih_raw >= 0.f ? g_inter * iw : 0.f;
      float g_x1 = -g_iw * dmax_a(b1x1, b2x1), g_x2 = g_iw * dmin_a(b1x2, b2x2);
      float g_y1 = -g_ih * dmax_a(b1y1, b2y1), g_y2 = g_ih * dmin_a(b1y2, b2y2);
      const float g_cw = g_c2 * 2 * cw, g_ch = g_c2 * 2 * ch;
      g_x2 += g_cw * dmax_a(b1x2, b2x2);
      g_x1 -= g_cw * dmin_a(b1x1, b2x1);
      g_y2 += g_ch * dmax_a(b1y2, b2y2);
      g_y1 -= g_ch * dmin_a(b1y1, b2y1);
      const float g_sx = g_rho2 * sx / 2, g_sy = g_rho2 * sy / 2;
      g_x1 -= g_sx;
      g_x2 -= g_sx;
      g_y1 -= g_sy;
      g_y2 -= g_sy;
      const float g_A1 = -g_v * kv * 2 * dA;
      const float g_q = g_A1 / (1 + q1 * q1);
      g_w1 += g_q / h1;
      g_h1 -= g_q * w1 / (h1 * h1);
      g_x2 += g_w1;
      g_x1 -= g_w1;
      g_y2 += g_h1;
      g_y1 -= g_h1;
      const float g_px = g_x1 + g_x2, g_py = g_y1 + g_y2, g_pw = (g_x2 - g_x1) / 2, g_ph = (g_y2 - g_y1) / 2;
      atomicAdd(gp + 0, g_px * 2 * s0 * (1 - s0));
      atomicAdd(gp + 1, g_py * 2 * s1 * (1 - s1));
      atomicAdd(gp + 2, g_pw * m.aw * 8 * s2 * s2 * (1 - s2));
      atomicAdd(gp + 3, g_ph * m.ah * 8 * s3 * s3 * (1 - s3));
    }
    // ---------------- class BCE over lanes (losses.py:276-279), pos_weight = cls_pw
    if (p.nc > 1) {
      float sum = 0.f;
      const float gsc = p.hcls * (float)p.bs * gs / ((float)n * (float)p.nc);
      for (int k = lane; k < p.nc; k += 32) {
        const float x = ps[5 + k];
        const float t = k == m.cls ? p.cp : p.cn;
        float g = 0.f;
        sum += bce_elem(x, t, p.cls_pw, p.fl_gamma, p.fl_alpha, gp != nullptr, &g);
        if (gp) atomicAdd(gp + 5 + k, gsc * g);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) atomicAdd(&p.acc[level * 3 + 1], (double)sum);
    }
  }
}

__global__ void loss_obj_kernel(LossKernelParams p, int level) {
  const LossLevel& L = p.lv[level];
  const size_t ncell = (size_t)p.bs * p.na * L.ny * L.nx;
  const float gs = p.gscale ? *p.gscale : 1.0f;
  const float gsc = p.hobj * L.balance * (float)p.bs * gs / (float)ncell;
  float sum = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x) {
    const float x = L.pred[i * p.no + 4];
    const float t = L.tobj[i];
    float g = 0.f;
    sum += bce_elem(x, t, p.obj_pw, p.fl_gamma, p.fl_alpha, L.grad != nullptr, &g);
    // the matched cells already hold box/cls gradients in other channels; channel 4 is written only here
    if (L.grad) L.grad[i * p.no + 4] = gsc * g;
  }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(&p.acc[level * 3 + 2], (double)s);
  }
}

__global__ void loss_final_kernel(LossKernelParams p, float* out5) {
  double lbox = 0, lobj = 0, lcls = 0;
  for (int i = 0; i < p.nl; ++i) {
    const int n = p.count[i];
    const double ncell = (double)p.bs * p.na * p.lv[i].ny * p.lv[i].nx;
    if (n > 0) {
      lbox += p.acc[i * 3 + 0] / n;
      if (p.nc > 1) lcls += p.acc[i * 3 + 1] / ((double)n * p.nc);
    }
    lobj += p.acc[i * 3 + 2] / ncell * p.lv[i].balance;
  }
  lbox *= p.hbox;
  lobj *= p.hobj;
  lcls *= p.hcls;
  const double loss = lbox + lobj + lcls;
  out5[0] = (float)(loss * p.bs);
  out5[1] = (float)lbox;
  out5[2] = (float)lobj;
  out5[3] = (float)lcls;
  out5[4] = (float)loss;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace ay2

using namespace ay2;

extern "C" size_t ay2_yolo_loss_workspace_bytes(const ay2_loss_params* p) {
  if (!p) return 0;
  size_t total = align256(sizeof(double) * AY2_LOSS_MAX_LEVELS * 3 + sizeof(int) * AY2_LOSS_MAX_LEVELS);
  for (int i = 0; i < p->nl; ++i) {
    const size_t ncell = (size_t)p->bs * p->na * p->ny[i] * p->nx[i];
    total += align256(ncell * sizeof(float)) + align256(ncell * sizeof(int));
  }
  return total;
}

extern "C" int ay2_yolo_loss(const ay2_loss_params* hp, const float* const* preds, float* const* grads,
                             const float* targets, const float* anchors, const float* gscale, void* workspace,
                             size_t workspace_bytes, float* out5, void* stream) {
  AY2_REQUIRE(hp && preds && anchors && workspace && out5, "ay2_yolo_loss: null pointer");
  AY2_REQUIRE(hp->nl >= 1 && hp->nl <= AY2_LOSS_MAX_LEVELS, "nl=%d unsupported", hp->nl);
  AY2_REQUIRE(hp->nt == 0 || targets, "targets pointer missing");
  AY2_REQUIRE(workspace_bytes >= ay2_yolo_loss_workspace_bytes(hp), "loss workspace too small");
  AY2_REQUIRE((long long)5 * hp->na * hp->nt < (1LL << 31), "too many (offset, anchor, target) candidates");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LossKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  kp.acc = reinterpret_cast<double*>(ws);
  kp.count = reinterpret_cast<int*>(ws + sizeof(double) * AY2_LOSS_MAX_LEVELS * 3);
  size_t off = align256(sizeof(double) * AY2_LOSS_MAX_LEVELS * 3 + sizeof(int) * AY2_LOSS_MAX_LEVELS);
  const size_t head = off;
  size_t tobj_bytes = 0;
  for (int i = 0; i < hp->nl; ++i) {
    const size_t ncell = (size_t)hp->bs * hp->na * hp->ny[i] * hp->nx[i];
    kp.lv[i].pred = preds[i];
    kp.lv[i].grad = grads ? grads[i] : nullptr;
    kp.lv[i].tobj = reinterpret_cast<float*>(ws + off);
    off += align256(ncell * sizeof(float));
    kp.lv[i].ny = hp->ny[i];
    kp.lv[i].nx = hp->nx[i];
    kp.lv[i].balance = hp->balance[i];
    AY2_REQUIRE(preds[i], "pred level %d is null", i);
  }
  tobj_bytes = off - head;
  const size_t owner_off = off;
  for (int i = 0; i < hp->nl; ++i) {
    const size_t ncell = (size_t)hp->bs * hp->na * hp->ny[i] * hp->nx[i];
    kp.lv[i].owner = reinterpret_cast<int*>(ws + off);
    off += align256(ncell * sizeof(int));
  }
  kp.targets = targets;
  kp.anchors = anchors;
  kp.gscale = gscale;
  kp.nl = hp->nl;
  kp.na = hp->na;
  kp.nc = hp->nc;
  kp.no = hp->nc + 5;
  kp.bs = hp->bs;
  kp.nt = hp->nt;
  kp.anchor_t = hp->anchor_t;
  kp.hbox = hp->box;
  kp.hobj = hp->obj;
  kp.hcls = hp->cls;
  kp.cls_pw = hp->cls_pw;
  kp.obj_pw = hp->obj_pw;
  kp.cp = hp->cp;
  kp.cn = hp->cn;
  kp.fl_gamma = hp->fl_gamma;
  kp.fl_alpha = hp->fl_alpha;
  // accumulators, counts, tobj <- 0 ; owner <- -1 (0xFF bytes)
  AY2_CHECK_CUDA(cudaMemsetAsync(ws, 0, head + tobj_bytes, st));
  AY2_CHECK_CUDA(cudaMemsetAsync(ws + owner_off, 0xFF, off - owner_off, st));
  int launches = 0;
  if (hp->nt > 0) {
    const int ncand = 5 * hp->na * hp->nt;
    const int threads = 128;
    int bx = (ncand + threads - 1) / threads;
    if (bx > 1024) bx = 1024;
    loss_assign_kernel<<<dim3(bx, hp->nl), threads, 0, st>>>(kp);
    AY2_CHECK_LAUNCH();
    int bw = (ncand + (threads / 32) - 1) / (threads / 32);
    if (bw > 148 * 8) bw = 148 * 8;
    loss_match_kernel<<<dim3(bw, hp->nl), threads, 0, st>>>(kp);
    AY2_CHECK_LAUNCH();
    launches += 2;
  }
  for (int i = 0; i < hp->nl; ++i) {
    const size_t ncell = (size_t)hp->bs * hp->na * hp->ny[i] * hp->nx[i];
    const int threads = 256;
    long long bx = (long long)((ncell + threads - 1) / threads);
    if (bx > 148 * 8) bx = 148 * 8;
    loss_obj_kernel<<<(int)bx, threads, 0, st>>>(kp, i);
    AY2_CHECK_LAUNCH();
    ++launches;
  }
  loss_final_kernel<<<1, 1, 0, st>>>(kp, out5);
  AY2_CHECK_LAUNCH();
  count_launch(launches + 1);
  return AY2_OK;
}
