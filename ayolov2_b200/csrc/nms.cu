// Batched non-max suppression for the whole batch in three launches, no host synchronisation.
//
// Reference: scripts/utils/metrics.py:285-443 (non_max_suppression, nms_type="nms") which loops over
// images on the host and calls torchvision.ops.nms (:385). Semantics reproduced bit-for-bit on identical
// fp32 inputs:
//   candidates  : rows with obj > conf (:313,337); conf_c = cls_c * obj (:353, fp32 multiply);
//                 best class = first arg-max (:363-364) or every class above conf when multi_label (:360-361)
//   box         : xywh -> xyxy as x -/+ w/2 (scripts/utils/general.py:316-319), class offset cls*max_wh added
//                 in fp32 (:383-384) for the suppression test only
//   order       : descending score, ties by candidate order (torchvision's stable sort); cap max_nms (:378-379)
//   suppression : greedy, IoU = inter / (area_i + area_j - inter) in fp32, suppressed iff IoU > iou_thres
//   output      : first max_det kept rows [x1,y1,x2,y2,conf,cls] in kept order (:386-388)
//
// Kernel 1 (filter): one warp per 32 rows reads only the objectness column, then cooperatively scores the
//   rows that pass and appends 64-bit keys  (~score_bits << 32 | row*nc + cls)  to a per-image list.
// Kernel 2 (sort + scan): one CTA per image. Bitonic sort of the keys (shared memory up to 8192 keys, in
//   global memory above), then the greedy scan in chunks of 256 sorted candidates: each chunk is first
//   tested against the boxes kept so far, then a 256x256 bit matrix of within-chunk overlaps is built and
//   resolved word by word by one warp. The scan stops as soon as max_det boxes are kept, which is what
//   makes greedy NMS cheap: later candidates cannot change the first max_det decisions.
#include <stdlib.h>
#include <string.h>

#include "ay2_common.h"
#include "ay2_ptx.cuh"
#include "head_math.cuh"
#include "nms_common.cuh"

namespace ay2 {

constexpr int kNmsThreads = 1024;
constexpr int kSortSmemKeys = 8192;
constexpr int kChunk = 256;
constexpr int kChunkWords = kChunk / 32;
constexpr int kMaxDetCap = 1024;
constexpr int kFastN = 2048;         // segmented path: at most this many candidates per image
constexpr int kCountNc = 128;        // counting sort by class up to this many classes (bitonic sort above)
constexpr int kGroups = 4;           // warp groups that resolve multi-block segments concurrently ...
constexpr int kGroupWarps = 8;       // ... of this many warps each (kGroups * kGroupWarps * 32 == kNmsThreads)

// Per-warp staging of candidate keys in shared memory: one global atomicAdd per flush instead of one per candidate
// (2,800 same-address atomics per image serialise in L2 at ~70 ns each: 0.2 ms; staged: ~30 per image).
constexpr int kStage = 96;
struct WarpStage {
  unsigned long long* buf;  // [kStage] shared memory, private to the warp
  int n;                    // warp-uniform fill count
};
__device__ __forceinline__ void stage_flush(WarpStage& s, unsigned long long* kb, int* cnt, int cap, int lane) {
  if (s.n == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(cnt, s.n);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int i = lane; i < s.n; i += 32)
    if (base + i < cap) kb[base + i] = s.buf[i];
  __syncwarp();
  s.n = 0;
}
// all lanes call; lanes with ok contribute `key`
__device__ __forceinline__ void stage_push(WarpStage& s, bool ok, unsigned long long key, unsigned long long* kb, int* cnt,
                                           int cap, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (!m) return;
  const int c = __popc(m);
  if (s.n + c > kStage) stage_flush(s, kb, cnt, cap, lane);
  if (ok) s.buf[s.n + __popc(m & ((1u << lane) - 1))] = key;
  s.n += c;
  __syncwarp();
}

__device__ __forceinline__ void group_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <typename Ptr>
__device__ __forceinline__ void bitonic_sort(Ptr A, int N) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool asc = (i & k) == 0;
        const unsigned long long a = A[i], b = A[l];
        if ((a > b) == asc) {
          A[i] = b;
          A[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Bitonic sort of N <= 2 * blockDim.x keys with two keys per thread held in registers: the j = 1 stages are in-thread,
// j = 2..32 are warp shuffles, only j >= 64 exchange through shared memory. (A pure shared-memory bitonic sort moves
// 32 KB per stage for 2048 keys -- 66 stages at the 128 B/clk of one SM are 9 us before any barrier cost; measured 18 us.)
// N is a power of two >= 2; A holds the keys on entry and the sorted keys on return.
__device__ __forceinline__ void bitonic_cmpx(unsigned long long& a, const unsigned long long o, const bool keep_min) {
  const bool lt = a < o;
  a = (lt == keep_min) ? a : o;
}
__device__ __forceinline__ void bitonic_sort_regs(unsigned long long* A, const int N) {
  const int t = threadIdx.x;
  const bool act = 2 * t < N;
  unsigned long long a0 = ~0ull, a1 = ~0ull;
  if (act) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&A[2 * t]);
    a0 = v.x;
    a1 = v.y;
  }
  for (int k = 2; k <= N; k <<= 1) {
    const bool up = ((2 * t) & k) == 0;
    for (int j = k >> 1; j >= 64; j >>= 1) {
      __syncthreads();
      if (act) *reinterpret_cast<ulonglong2*>(&A[2 * t]) = make_ulonglong2(a0, a1);
      __syncthreads();
      if (act) {
        const int m = j >> 1;
        const ulonglong2 o = *reinterpret_cast<const ulonglong2*>(&A[2 * (t ^ m)]);
        const bool keep_min = ((t & m) == 0) == up;
        bitonic_cmpx(a0, o.x, keep_min);
        bitonic_cmpx(a1, o.y, keep_min);
      }
    }
    for (int j = (k >> 1) < 32 ? (k >> 1) : 32; j >= 2; j >>= 1) {
      const int m = j >> 1;
      const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, a0, m);
      const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, a1, m);
      const bool keep_min = ((t & m) == 0) == up;
      bitonic_cmpx(a0, o0, keep_min);
      bitonic_cmpx(a1, o1, keep_min);
    }
    const unsigned long long lo = a0 < a1 ? a0 : a1, hi = a0 < a1 ? a1 : a0;
    a0 = up ? lo : hi;
    a1 = up ? hi : lo;
  }
  __syncthreads();
  if (act) *reinterpret_cast<ulonglong2*>(&A[2 * t]) = make_ulonglong2(a0, a1);
  __syncthreads();
}

// Where a candidate's xywh box comes from: the dense fp32 prediction tensor, or (fused path) the bf16 head
// logits of the pyramid level the row belongs to, decoded on the fly with the shared head arithmetic.
struct BoxSource {
  const float* pred;  // dense mode: (B, n, no); nullptr in logits mode
  int n, no;
  int nl, na;
  const __nv_bfloat16* logits[AY2_NMS_MAX_LEVELS];
  int ny[AY2_NMS_MAX_LEVELS], nx[AY2_NMS_MAX_LEVELS], cstride[AY2_NMS_MAX_LEVELS];
  int row_off[AY2_NMS_MAX_LEVELS + 1];
  float stride_px[AY2_NMS_MAX_LEVELS];
  float anchor_px[AY2_NMS_MAX_LEVELS][AY2_NMS_MAX_ANCHORS][2];
};

__device__ __forceinline__ float4 load_xywh(const BoxSource& s, int b, int row) {
  if (s.pred) {
    const float* rp = s.pred + ((long long)b * s.n + row) * s.no;
    return make_float4(rp[0], rp[1], rp[2], rp[3]);
  }
  int l = 0;
  while (l + 1 < s.nl && row >= s.row_off[l + 1]) ++l;
  const int r = row - s.row_off[l];
  const int plane = s.ny[l] * s.nx[l];
  const int a = r / plane;
  const int pix = r - a * plane;
  const int y = pix / s.nx[l], x = pix - y * s.nx[l];
  const __nv_bfloat16* lp = s.logits[l] + ((long long)b * plane + pix) * s.cstride[l] + a * s.no;
  float4 o;
  o.x = head_xy(head_sigmoid(__bfloat162float(lp[0])), (float)x, s.stride_px[l]);
  o.y = head_xy(head_sigmoid(__bfloat162float(lp[1])), (float)y, s.stride_px[l]);
  o.z = head_wh(head_sigmoid(__bfloat162float(lp[2])), s.anchor_px[l][a][0]);
  o.w = head_wh(head_sigmoid(__bfloat162float(lp[3])), s.anchor_px[l][a][1]);
  return o;
}

// ---------------------------------------------------------------------------------------------------------------
// Candidate generation in two balanced phases.
//   phase A (rows)  : every row's objectness is tested once (coalesced over rows / pixels); passing row indices are
//                     appended to a per-image list. Objectness is spatially clustered, so scoring the rows inside this
//                     loop would leave a few warps with ~100 serial row visits (measured: 0.2 ms of tail).
//   phase B (score) : one warp per listed row, grid-strided over the list -> every warp gets the same amount of work.
//                     conf_c = cls_c * obj, best class (first arg-max) or every class above conf (multi_label).
// Dense mode reads the fp32 prediction tensor; logits mode decodes the bf16 head logits with the shared head arithmetic
// (bit-identical scores, no dense (B, 25200, 85) tensor).
// ---------------------------------------------------------------------------------------------------------------
struct U32Stage {
  unsigned* buf;
  int n;
};
__device__ __forceinline__ void rows_flush(U32Stage& s, unsigned* list, int* cnt, int cap, int lane) {
  if (s.n == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(cnt, s.n);
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int i = lane; i < s.n; i += 32)
    if (base + i < cap) list[base + i] = s.buf[i];
  __syncwarp();
  s.n = 0;
}
__device__ __forceinline__ void rows_push(U32Stage& s, bool ok, unsigned v, unsigned* list, int* cnt, int cap, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if (!m) return;
  const int c = __popc(m);
  if (s.n + c > kStage) rows_flush(s, list, cnt, cap, lane);
  if (ok) s.buf[s.n + __popc(m & ((1u << lane) - 1))] = v;
  s.n += c;
  __syncwarp();
}

__global__ void nms_rows_kernel(BoxSource src, ay2_nms_params p, unsigned* __restrict__ rows, int* __restrict__ row_counts) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp_in_grid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned* list = rows + (long long)b * p.n;
  __shared__ unsigned stage_mem[8][kStage];
  U32Stage st{stage_mem[threadIdx.x >> 5], 0};
  if (src.pred) {
    const float* ip = src.pred + (long long)b * p.n * p.no;
    const int groups = (p.n + 31) / 32;
    for (int g = warp_in_grid; g < groups; g += nwarps) {
      const int row = g * 32 + lane;
      const bool ok = row < p.n && ip[(long long)row * p.no + 4] > p.conf_thres;
      rows_push(st, ok, (unsigned)row, list, &row_counts[b], p.n, lane);
    }
  } else {
    for (int l = 0; l < src.nl; ++l) {
      const int plane = src.ny[l] * src.nx[l];
      const int groups = (plane + 31) / 32;
      const __nv_bfloat16* base = src.logits[l] + (long long)b * plane * src.cstride[l];
      for (int g = warp_in_grid; g < groups; g += nwarps) {
        const int pix = g * 32 + lane;
        for (int a = 0; a < src.na; ++a) {
          bool ok = false;
          if (pix < plane)
            ok = head_sigmoid(__bfloat162float(base[(long long)pix * src.cstride[l] + a * src.no + 4])) > p.conf_thres;
          rows_push(st, ok, (unsigned)(src.row_off[l] + a * plane + pix), list, &row_counts[b], p.n, lane);
        }
      }
    }
  }
  rows_flush(st, list, &row_counts[b], p.n, lane);
}

__global__ void nms_score_rows_kernel(BoxSource src, ay2_nms_params p, const uint8_t* __restrict__ class_mask,
                                      const unsigned* __restrict__ rows, const int* __restrict__ row_counts,
                                      unsigned long long* __restrict__ keys, long long key_stride, int* __restrict__ counts) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warp_in_grid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nc = p.no - 5;
  const unsigned* list = rows + (long long)b * p.n;
  unsigned long long* kb = keys + (long long)b * key_stride;
  const int nrows = min(row_counts[b], p.n);
  __shared__ unsigned long long stage_mem[8][kStage];
  WarpStage st{stage_mem[threadIdx.x >> 5], 0};
  for (int e = warp_in_grid; e < nrows; e += nwarps) {
    const int row = (int)list[e];
    float robj;
    float cv[4];  // class scores of this lane: classes lane, lane+32, lane+64, lane+96 (more classes: extra passes below)
    const float* fp = nullptr;
    const __nv_bfloat16* hp = nullptr;
    if (src.pred) {
      fp = src.pred + ((long long)b * p.n + row) * p.no;
      robj = fp[4];
    } else {
      int l = 0;
      while (l + 1 < src.nl && row >= src.row_off[l + 1]) ++l;
      const int r = row - src.row_off[l];
      const int plane = src.ny[l] * src.nx[l];
      const int a = r / plane;
      const int pix = r - a * plane;
      hp = src.logits[l] + ((long long)b * plane + pix) * src.cstride[l] + a * src.no;
      robj = head_sigmoid(__bfloat162float(hp[4]));
    }
    (void)cv;
    if (p.multi_label) {
      for (int c0 = 0; c0 < nc; c0 += 32) {
        const int c = c0 + lane;
        float conf = 0.f;
        bool ok = false;
        if (c < nc) {
          const float cs = fp ? fp[5 + c] : head_sigmoid(__bfloat162float(hp[5 + c]));
          conf = __fmul_rn(cs, robj);
          ok = conf > p.conf_thres && (!class_mask || class_mask[c]);
        }
        stage_push(st, ok, (static_cast<unsigned long long>(~__float_as_uint(conf)) << 32) | static_cast<unsigned>(row * nc + c),
                   kb, &counts[b], p.max_candidates, lane);
      }
    } else {
      // first arg-max over classes (torch.max(1) keeps the first maximal index)
      float best = -INFINITY;
      int bidx = 0x7fffffff;
      for (int c = lane; c < nc; c += 32) {
        const float cs = fp ? fp[5 + c] : head_sigmoid(__bfloat162float(hp[5 + c]));
        const float conf = __fmul_rn(cs, robj);
        if (conf > best) {
          best = conf;
          bidx = c;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) {
          best = ob;
          bidx = oi;
        }
      }
      stage_push(st, lane == 0 && best > p.conf_thres && (!class_mask || class_mask[bidx]),
                 (static_cast<unsigned long long>(~__float_as_uint(best)) << 32) | static_cast<unsigned>(row * nc + bidx), kb,
                 &counts[b], p.max_candidates, lane);
    }
  }
  stage_flush(st, kb, &counts[b], p.max_candidates, lane);
}

// Where one resolved image (or one class group of it) goes: rows [max_det][6], optionally the rows' 64-bit keys (the merge of
// class groups needs the global order), and the row count.
struct NmsDst {
  float* det;
  unsigned long long* keys;
  int* count;
};

// Resolves image b -- or, with ngroups > 1, only its candidates whose class c has c % ngroups == group (classes never
// suppress each other under per-class offsets, so greedy NMS splits exactly along classes; the caller merges the groups'
// kept lists by key). Sets *wide_flag when a box leaves the disjoint-class window (the split is then invalid).
__device__ __noinline__ void nms_image(const BoxSource& src, ay2_nms_params p, unsigned long long* __restrict__ keys,
                                       long long key_stride, const int* __restrict__ counts, const NmsDst dst,
                                       int* __restrict__ overflow, long long* __restrict__ trace,
                                       const float* __restrict__ wh_scale, const int b, const int ngroups, const int group,
                                       int* __restrict__ wide_flag, const int dense, unsigned long long* __restrict__ keys2,
                                       int* __restrict__ count_back) {
#define AY2_NMS_MARK(k)                                                    \
  do {                                                                     \
    if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[k] = clock64(); \
  } while (0)
  AY2_NMS_MARK(0);
  extern __shared__ __align__(16) uint8_t sm[];
  unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm);             // [kSortSmemKeys]
  float4* cbo = reinterpret_cast<float4*>(skeys + kSortSmemKeys);                    // chunk boxes + class offset
  float4* cbox = cbo + kChunk;                                                       // chunk boxes (output form)
  float4* kbo = cbox + kChunk;                                                       // kept boxes + offset [kMaxDetCap]
  float* carea = reinterpret_cast<float*>(kbo + kMaxDetCap);                         // [kChunk]
  float* cconf = carea + kChunk;                                                     // [kChunk]
  float* karea = cconf + kChunk;                                                     // [kMaxDetCap]
  int* ccls = reinterpret_cast<int*>(karea + kMaxDetCap);                            // [kChunk]
  int* kcls = ccls + kChunk;                                                         // [kMaxDetCap]
  int* alive = kcls + kMaxDetCap;                                                    // [kChunk]
  int* kept_idx = alive + kChunk;                                                    // [kChunk]
  unsigned* mask = reinterpret_cast<unsigned*>(kept_idx + kChunk);                   // [kChunk][kChunkWords]
  __shared__ int s_kept, s_new, s_wide;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wid = tid >> 5;
  const int nwarp = blockDim.x >> 5;
  const int nc = p.no - 5;
  __syncthreads();  // a second call in the same CTA (fallback after a failed class split) re-uses the shared arrays
  // per-image class-offset scale (torchvision's batched_nms coordinate trick: max coordinate of the image + 1). A scale
  // that small fails the disjoint-window test below, so such images take the single-segment (exact, all-pairs) route.
  if (wh_scale) p.max_wh = __fadd_rn(wh_scale[b], 1.0f);
  // dense-slot mode (fused head): the key array holds one slot per ROW, ~0 where the row is no candidate
  int n = dense ? p.n : counts[b];
  if (n > p.max_candidates) {
    n = p.max_candidates;
    if (tid == 0) atomicExch(overflow, 1);
  }
  if (tid == 0) {
    s_kept = 0;
    s_wide = 0;
  }
  if (n == 0) {
    if (tid == 0) *dst.count = 0;
    return;
  }
  // ---------------------------------------------------------------- sort (ascending key == descending score)
  int npow2 = 2;
  while (npow2 < n) npow2 <<= 1;
  unsigned long long* gk = keys + (long long)b * key_stride;
  const unsigned long long* sorted;
  if (ngroups > 1 || dense) {
    // Compacting load: keep the live slots (dense mode) of this CTA's class group; order is irrelevant before the sort.
    // Up to kSortSmemKeys keys land in shared memory, any excess in the global scratch list.
    __shared__ int s_take;
    unsigned long long* g2 = keys2 + (long long)b * key_stride;
    if (tid == 0) s_take = 0;
    __syncthreads();
    // (kLoadAhead independent loads per thread per trip: the dense slot array is 25,200 keys per image, and with one load
    // in flight the 25 dependent L2 round trips of this loop were ~15 % of the kernel)
    constexpr int kLoadAhead = 4;
    for (int i0 = 0; i0 < n; i0 += kLoadAhead * blockDim.x) {
      unsigned long long kv[kLoadAhead];
#pragma unroll
      for (int u = 0; u < kLoadAhead; ++u) {
        const int i = i0 + u * blockDim.x + tid;
        kv[u] = i < n ? gk[i] : ~0ull;
      }
#pragma unroll
      for (int u = 0; u < kLoadAhead; ++u) {
        const int i = i0 + u * blockDim.x + tid;
        const unsigned long long key = kv[u];
        const bool mine = i < n && (!dense || key != ~0ull) &&
                          (ngroups == 1 || static_cast<int>((static_cast<unsigned>(key) % nc) % ngroups) == group);
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_take, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (mine) {
          const int pos = base + __popc(m & ((1u << lane) - 1u));
          if (pos < kSortSmemKeys) skeys[pos] = key;
          else if (ngroups == 1) g2[pos] = key;
        }
      }
    }
    __syncthreads();
    n = s_take;
    if (count_back && tid == 0 && n) atomicAdd(count_back, n);  // the candidate count, for the callers that report it
    if (n == 0) {
      if (tid == 0) *dst.count = 0;
      return;
    }
    npow2 = 2;
    while (npow2 < n) npow2 <<= 1;
    if (n > kSortSmemKeys) {
      if (ngroups > 1) {  // a class group too large for the shared-memory sort: the whole image takes the single-CTA route
        if (tid == 0) {
          *dst.count = 0;
          if (wide_flag) atomicExch(wide_flag, 1);
        }
        return;
      }
      for (int i = tid; i < npow2; i += blockDim.x)
        if (i < kSortSmemKeys) g2[i] = skeys[i];
        else if (i >= n) g2[i] = ~0ull;
      __syncthreads();
      bitonic_sort(g2, npow2);
      sorted = g2;
    } else {
      for (int i = n + tid; i < npow2; i += blockDim.x) skeys[i] = ~0ull;
      __syncthreads();
      if (npow2 <= 2 * kNmsThreads) bitonic_sort_regs(skeys, npow2);
      else bitonic_sort(skeys, npow2);
      sorted = skeys;
    }
  } else if (npow2 <= kSortSmemKeys) {
    for (int i = tid; i < npow2; i += blockDim.x) skeys[i] = i < n ? gk[i] : ~0ull;
    __syncthreads();
    if (npow2 <= 2 * kNmsThreads) bitonic_sort_regs(skeys, npow2);
    else bitonic_sort(skeys, npow2);
    sorted = skeys;
  } else {
    for (int i = n + tid; i < npow2; i += blockDim.x) gk[i] = ~0ull;
    __syncthreads();
    bitonic_sort(gk, npow2);
    sorted = gk;
  }
  AY2_NMS_MARK(1);
  const bool fits_fast = npow2 <= kFastN;  // judged on the unclamped count: the key buffer's upper half must be free
  if (n > p.max_nms) n = p.max_nms;
  const IouThr thr = make_iou_thr(p.iou_thres);

  // ---------------------------------------------------------------- segmented path (n <= kFastN)
  // With per-class box offsets (metrics.py:383-384) boxes of different classes cannot intersect as long as every
  // box lies inside a max_wh-wide window (checked below; exact, not an approximation), so greedy NMS decomposes
  // into independent per-class greedy scans. Candidates are re-ordered by (class, score rank) into segments (one
  // segment for everything when agnostic or when a box leaves the window) and each segment is resolved 32
  // candidates (one block) at a time:
  //   D. all warps: for every block, which earlier candidates of the same block overlap each candidate (<= 31 tests);
  //   S. per block, in order: every candidate is tested against the segment's kept list so far (the only tests
  //      whose outcome can matter), then one warp resolves the block's internal greedy order from D's bits and
  //      appends the survivors to the kept list. One-block segments need one warp; longer ones are shared out to
  //      four 8-warp groups (the kept list is split over the group's warps).
  // Survivors are compacted in global score order -- exactly the rows, in exactly the order, of the sequential
  // reference. A segment stops once it holds max_det survivors: its later rows cannot reach the output.
  unsigned long long* key2 = skeys + kFastN;                                    // [kFastN]  (class, rank) keys
  float4* sbo = reinterpret_cast<float4*>(skeys + 2 * kFastN);                  // [kFastN]  offset boxes, segment order
  uint8_t* fr = reinterpret_cast<uint8_t*>(skeys + kSortSmemKeys);              // overlay of the chunk-path region
  float4* rbox = reinterpret_cast<float4*>(fr);                                 // [kFastN] output boxes, score order
  float* sarea = reinterpret_cast<float*>(rbox + kFastN);                       // [kFastN]
  int* st0 = reinterpret_cast<int*>(sarea + kFastN);                            // [kFastN] segment start of row t
  int* segend = st0 + kFastN;                                                   // [kFastN] indexed by segment start
  int* fseg = segend + kFastN;                                                  // [kFastN] one-block segments from the
                                                                                //   front, longer ones from the back
  int* klist = fseg + kFastN;                                                   // [kFastN] kept rows, per segment at t0
  unsigned* sdin = reinterpret_cast<unsigned*>(klist + kFastN);                 // [kFastN] in-block overlap bits
  int* blist = reinterpret_cast<int*>(sdin + kFastN);                           // [kFastN] first rows of all blocks
  unsigned char* ssupp = reinterpret_cast<unsigned char*>(blist + kFastN);      // [kFastN] by score rank
  unsigned short* ccnt = reinterpret_cast<unsigned short*>(klist);              // [64][kCountNc] counting-sort table
                                                                                //   (klist + sdin, before they are used)
  __shared__ int s_nsmall, s_nbig, s_nblk, s_bignext, s_scan[kNmsThreads / 32], s_cstart[kCountNc], s_ctot[kCountNc];
  __shared__ int s_gseg[kGroups], s_gk[kGroups];
  __shared__ unsigned s_gsup[kGroups];
  bool fast = fits_fast;
  if (fast) {
    if (tid == 0) {
      s_nsmall = 0;
      s_nbig = 0;
      s_nblk = 0;
      s_bignext = 0;
    }
    if (tid < kGroups) s_gsup[tid] = 0;
    const float lo = -0.25f * p.max_wh, hi = 0.75f * p.max_wh - 2.0f;
    {  // both rows of a thread are fetched before either is used: the gather is pure latency
      float4 rr[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tid + h * kNmsThreads;
        if (i < n) rr[h] = load_xywh(src, b, static_cast<unsigned>(sorted[i]) / nc);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tid + h * kNmsThreads;
        if (i < n) {
          const float4 r = rr[h];
          // general.py:316-319 with ratio = wh = 1, pad = 0  (1*1*(x -+ w/2) + 0)
          float4 bx;
          bx.x = __fadd_rn(__fsub_rn(r.x, __fmul_rn(r.z, 0.5f)), 0.0f);
          bx.y = __fadd_rn(__fsub_rn(r.y, __fmul_rn(r.w, 0.5f)), 0.0f);
          bx.z = __fadd_rn(__fadd_rn(r.x, __fmul_rn(r.z, 0.5f)), 0.0f);
          bx.w = __fadd_rn(__fadd_rn(r.y, __fmul_rn(r.w, 0.5f)), 0.0f);
          if (!(bx.x >= lo && bx.y >= lo && bx.z <= hi && bx.w <= hi)) s_wide = 1;
          rbox[i] = bx;
          ssupp[i] = 0;
        }
      }
    }
    __syncthreads();
    AY2_NMS_MARK(2);
    const bool partition = !p.agnostic && !s_wide;
    // ---- (class, rank) order -> key2[t] (low word = score rank), st0[t], segend[], segment lists
    auto add_segment = [&](int t0, int t1) {
      segend[t0] = t1;
      if (t1 - t0 <= 32) fseg[atomicAdd(&s_nsmall, 1)] = t0;
      else fseg[kFastN - 1 - atomicAdd(&s_nbig, 1)] = t0;
    };
    if (!partition) {
      for (int t = tid; t < n; t += blockDim.x) {
        key2[t] = static_cast<unsigned>(t);
        st0[t] = 0;
      }
      if (tid == 0) add_segment(0, n);
      __syncthreads();
    } else if (nc <= kCountNc) {
      // stable counting sort by class: rank inside a 32-candidate chunk by __match_any_sync, chunk totals per class in
      // a [chunk][class] table, running sum over chunks per class, exclusive scan over classes
      const int chunks = (n + 31) >> 5;
      for (int i = tid; i < chunks * kCountNc / 2; i += blockDim.x) reinterpret_cast<unsigned*>(ccnt)[i] = 0u;
      __syncthreads();
      unsigned mycls[2];
      int myrank[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tid + h * kNmsThreads;
        const bool valid = i < n;
        mycls[h] = valid ? static_cast<unsigned>(sorted[i]) % nc : 0xffffffffu;
        if ((i & ~31) < n) {  // warp-uniform: the chunk exists
          const unsigned peers = __match_any_sync(0xffffffffu, mycls[h]);
          myrank[h] = __popc(peers & ((1u << lane) - 1u));
          if (valid && lane == __ffs(peers) - 1) ccnt[(i >> 5) * kCountNc + mycls[h]] = static_cast<unsigned short>(__popc(peers));
        }
      }
      __syncthreads();
      if (tid < nc) {
        int run = 0;
        for (int w = 0; w < chunks; ++w) {
          const int v = ccnt[w * kCountNc + tid];
          ccnt[w * kCountNc + tid] = static_cast<unsigned short>(run);
          run += v;
        }
        s_ctot[tid] = run;
      }
      __syncthreads();
      if (wid == 0) {
        constexpr int per = kCountNc / 32;
        int v[per], sum = 0;
#pragma unroll
        for (int e = 0; e < per; ++e) {
          const int c = lane * per + e;
          v[e] = c < nc ? s_ctot[c] : 0;
          sum += v[e];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        int base = incl - sum;
#pragma unroll
        for (int e = 0; e < per; ++e) {
          const int c = lane * per + e;
          if (c < nc) {
            s_cstart[c] = base;
            if (v[e] > 0) add_segment(base, base + v[e]);
          }
          base += v[e];
        }
      }
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tid + h * kNmsThreads;
        if (i < n) {
          const int t0 = s_cstart[mycls[h]];
          const int pos = t0 + ccnt[(i >> 5) * kCountNc + mycls[h]] + myrank[h];
          key2[pos] = (static_cast<unsigned long long>(mycls[h]) << 32) | static_cast<unsigned>(i);
          st0[pos] = t0;
        }
      }
      __syncthreads();
    } else {
      int n2 = 2;
      while (n2 < n) n2 <<= 1;
      for (int i = tid; i < n2; i += blockDim.x)
        key2[i] = i < n ? ((static_cast<unsigned long long>(static_cast<unsigned>(sorted[i]) % nc) << 32) | static_cast<unsigned>(i)) : ~0ull;
      __syncthreads();
      bitonic_sort_regs(key2, n2);  // (class, score rank) ascending
      for (int t = tid; t < n; t += blockDim.x) {
        const unsigned c = static_cast<unsigned>(key2[t] >> 32);
        int a = 0, z = t;  // first index of class c
        while (a < z) {
          const int mid = (a + z) >> 1;
          if (static_cast<unsigned>(key2[mid] >> 32) < c) a = mid + 1;
          else z = mid;
        }
        st0[t] = a;
        if (t == a) {
          int a2 = t + 1, z2 = n;  // first index whose class differs
          while (a2 < z2) {
            const int mid = (a2 + z2) >> 1;
            if (static_cast<unsigned>(key2[mid] >> 32) == c) a2 = mid + 1;
            else z2 = mid;
          }
          add_segment(t, a2);
        }
      }
      __syncthreads();
    }
    AY2_NMS_MARK(3);
    // ---- segment-ordered offset boxes, block list
    for (int t = tid; t < n; t += blockDim.x) {
      const int i = static_cast<int>(static_cast<unsigned>(key2[t]));
      const float4 bx = rbox[i];
      const float off = p.agnostic ? 0.0f : __fmul_rn(static_cast<float>(static_cast<unsigned>(sorted[i]) % nc), p.max_wh);
      float4 bo;
      bo.x = __fadd_rn(bx.x, off);
      bo.y = __fadd_rn(bx.y, off);
      bo.z = __fadd_rn(bx.z, off);
      bo.w = __fadd_rn(bx.w, off);
      sbo[t] = bo;
      sarea[t] = __fmul_rn(__fsub_rn(bo.z, bo.x), __fsub_rn(bo.w, bo.y));
      if (((t - st0[t]) & 31) == 0) blist[atomicAdd(&s_nblk, 1)] = t;
    }
    __syncthreads();
    AY2_NMS_MARK(4);
    // ---- D: in-block overlaps. sdin[c] bit r <=> row (block start + r) precedes c in its block and IoU > thr
    const int nblk = s_nblk;
    for (int bi = wid; bi < nblk; bi += nwarp) {
      const int tb = blist[bi];
      const int t1 = segend[st0[tb]];
      const int c = tb + lane;
      const bool valid = c < t1;
      const float4 bc = valid ? sbo[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float ac = valid ? sarea[c] : 0.f;
      const int rows = min(31, t1 - tb - 1);
      unsigned din = 0;
      for (int r = 0; r < rows; ++r) {
        const bool hit = valid && lane > r && iou_gt(sbo[tb + r], sarea[tb + r], bc, ac, thr);
        din |= (hit ? 1u : 0u) << r;
      }
      if (valid) sdin[c] = din;
    }
    __syncthreads();
    AY2_NMS_MARK(5);
    // ---- S (one-block segments): one warp each
    const unsigned lt_mask = (1u << lane) - 1u;
    const int nsmall = s_nsmall, nbig = s_nbig;
    for (int si = wid; si < nsmall; si += nwarp) {
      const int t0 = fseg[si];
      const int sl = segend[t0] - t0;
      const bool valid = lane < sl;
      const unsigned din = valid ? sdin[t0 + lane] : 0u;
      unsigned avail = sl >= 32 ? 0xffffffffu : ((1u << sl) - 1u);
      unsigned keep = 0;
      while (avail) {  // warp-uniform
        const int i = __ffs(avail) - 1;
        keep |= 1u << i;
        avail &= ~(1u << i);
        avail &= ~__ballot_sync(0xffffffffu, (din >> i) & 1u);
      }
      if (valid && !((keep >> lane) & 1u)) ssupp[static_cast<unsigned>(key2[t0 + lane])] = 1;
    }
    // ---- S (longer segments): 8-warp groups take segments from the list; per block: test against the kept list
    //      (split over the group's warps), then warp 0 of the group resolves the block and extends the list
    const int grp = wid / kGroupWarps, gw = wid % kGroupWarps;
    for (;;) {
      if (gw == 0 && lane == 0) {
        s_gseg[grp] = atomicAdd(&s_bignext, 1);
        s_gk[grp] = 0;
      }
      group_bar_sync(1 + grp, kGroupWarps * 32);
      const int sgi = s_gseg[grp];
      if (sgi >= nbig) break;  // group-uniform
      const int t0 = fseg[kFastN - 1 - sgi], t1 = segend[t0];
      const int nw = (t1 - t0 + 31) >> 5;
      for (int wb = 0; wb < nw; ++wb) {
        const int c = t0 + (wb << 5) + lane;
        const bool valid = c < t1;
        const int kc = s_gk[grp];
        if (kc >= p.max_det) {  // group-uniform: the rest of the segment cannot reach the output
          if (valid && gw == 0) ssupp[static_cast<unsigned>(key2[c])] = 1;
          continue;
        }
        const float4 bc = valid ? sbo[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float ac = valid ? sarea[c] : 0.f;
        bool sup = false;
        for (int i = gw; i < kc; i += 2 * kGroupWarps) {  // two independent tests per trip
          const int i2 = i + kGroupWarps;
          const bool has2 = i2 < kc;
          const int r0 = klist[t0 + i], r1 = klist[t0 + (has2 ? i2 : i)];
          const bool h0 = iou_gt(sbo[r0], sarea[r0], bc, ac, thr);
          const bool h1 = iou_gt(sbo[r1], sarea[r1], bc, ac, thr);
          sup |= h0 | (has2 & h1);
        }
        const unsigned word = __ballot_sync(0xffffffffu, valid && sup);
        if (lane == 0 && word) atomicOr(&s_gsup[grp], word);
        group_bar_sync(1 + grp, kGroupWarps * 32);
        if (gw == 0) {
          const unsigned supw = s_gsup[grp];
          const int left = t1 - t0 - (wb << 5);
          const unsigned din = valid ? sdin[c] : 0u;
          unsigned avail = ~supw & (left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
          unsigned keep = 0;
          while (avail) {  // warp-uniform
            const int i = __ffs(avail) - 1;
            keep |= 1u << i;
            avail &= ~(1u << i);
            avail &= ~__ballot_sync(0xffffffffu, (din >> i) & 1u);
          }
          if (valid) {
            if ((keep >> lane) & 1u) klist[t0 + kc + __popc(keep & lt_mask)] = c;
            else ssupp[static_cast<unsigned>(key2[c])] = 1;
          }
          if (lane == 0) {
            s_gk[grp] = kc + __popc(keep);
            s_gsup[grp] = 0;
          }
        }
        group_bar_sync(1 + grp, kGroupWarps * 32);
      }
    }
    __syncthreads();
    AY2_NMS_MARK(6);
    // compaction in score order: exclusive prefix count of kept flags
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += ssupp[i] ? 0 : 1;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[wid] = incl;
    AY2_NMS_MARK(14);
    __syncthreads();
    AY2_NMS_MARK(15);
    if (wid == 0) {
      int v = lane < nwarp ? s_scan[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (lane < nwarp) s_scan[lane] = v;
    }
    __syncthreads();
    int k = incl - cnt + (wid > 0 ? s_scan[wid - 1] : 0);
    AY2_NMS_MARK(12);
    for (int i = i0; i < i1 && k < p.max_det; ++i) {
      if (ssupp[i]) continue;
      const unsigned long long key = sorted[i];
      const unsigned idx = static_cast<unsigned>(key);
      const float4 bx = rbox[i];
      float* o = dst.det + (long long)k * 6;
      o[0] = bx.x;
      o[1] = bx.y;
      o[2] = bx.z;
      o[3] = bx.w;
      o[4] = __uint_as_float(~static_cast<unsigned>(key >> 32));
      o[5] = static_cast<float>(idx % nc);
      if (dst.keys) dst.keys[k] = key;
      ++k;
    }
    AY2_NMS_MARK(13);
    __syncthreads();
    if (tid == 0) {
      *dst.count = min(s_scan[nwarp - 1], p.max_det);
      if (s_wide && wide_flag) atomicExch(wide_flag, 1);
    }
    AY2_NMS_MARK(7);
    if (trace && b == 0 && tid == 0) {
      trace[8] = n;
      trace[9] = nsmall;
      trace[10] = nbig;
      trace[11] = nblk;
    }
    return;
  }
  __syncthreads();
  if (tid == 0) s_wide = 0;  // the chunk path recomputes it
  __syncthreads();

  // ---------------------------------------------------------------- greedy scan (general path)
  for (int cs = 0; cs < n; cs += kChunk) {
    const int kc = s_kept;
    if (kc >= p.max_det) break;
    const int cn = min(kChunk, n - cs);
    if (tid < kChunk) {
      int al = 0;
      if (tid < cn) {
        const unsigned long long key = sorted[cs + tid];
        const unsigned idx = static_cast<unsigned>(key);
        const float score = __uint_as_float(~static_cast<unsigned>(key >> 32));
        const int row = idx / nc;
        const int cls = idx - row * nc;
        const float4 r = load_xywh(src, b, row);
        // general.py:316-319 with ratio = wh = 1, pad = 0  (1*1*(x -+ w/2) + 0)
        float4 bx;
        bx.x = __fadd_rn(__fsub_rn(r.x, __fmul_rn(r.z, 0.5f)), 0.0f);
        bx.y = __fadd_rn(__fsub_rn(r.y, __fmul_rn(r.w, 0.5f)), 0.0f);
        bx.z = __fadd_rn(__fadd_rn(r.x, __fmul_rn(r.z, 0.5f)), 0.0f);
        bx.w = __fadd_rn(__fadd_rn(r.y, __fmul_rn(r.w, 0.5f)), 0.0f);
        const float off = p.agnostic ? 0.0f : __fmul_rn(static_cast<float>(cls), p.max_wh);
        float4 bo;
        bo.x = __fadd_rn(bx.x, off);
        bo.y = __fadd_rn(bx.y, off);
        bo.z = __fadd_rn(bx.z, off);
        bo.w = __fadd_rn(bx.w, off);
        cbox[tid] = bx;
        cbo[tid] = bo;
        carea[tid] = __fmul_rn(__fsub_rn(bo.z, bo.x), __fsub_rn(bo.w, bo.y));
        cconf[tid] = score;
        ccls[tid] = cls;
        al = 1;
        // Boxes of different classes are max_wh apart after the offset: they can only intersect if some box
        // reaches outside the window [-max_wh/4, 3*max_wh/4 - 2] (2 px of slack for the rounding of the offset
        // add). If none does, cross-class pairs are skipped without evaluating the IoU -- an exact shortcut.
        const float lo = -0.25f * p.max_wh, hi = 0.75f * p.max_wh - 2.0f;
        if (!(bx.x >= lo && bx.y >= lo && bx.z <= hi && bx.w <= hi)) s_wide = 1;
      }
      alive[tid] = al;
    }
    __syncthreads();
    const bool by_class = !p.agnostic && !s_wide;
    // (1) chunk vs. boxes kept so far: warp per kept box, lanes over candidates
    for (int k = wid; k < kc; k += nwarp) {
      const float4 kb4 = kbo[k];
      const float ka = karea[k];
      const int kcl = kcls[k];
      for (int c = lane; c < cn; c += 32) {
        if (!alive[c]) continue;
        if (by_class && ccls[c] != kcl) continue;
        if (iou_gt(kb4, ka, cbo[c], carea[c], thr)) alive[c] = 0;
      }
    }
    __syncthreads();
    // (2) within-chunk overlap bit matrix: mask[i][w] bit jj <=> j = 32w+jj > i and IoU(i,j) > thr
    for (int t = tid; t < kChunk * kChunkWords; t += blockDim.x) {
      const int i = t / kChunkWords;
      const int w = t - i * kChunkWords;
      unsigned bits = 0;
      if (i < cn && alive[i] && (w * 32 + 31) > i) {
        const float4 bi = cbo[i];
        const float ai = carea[i];
        const int ci = ccls[i];
        for (int jj = 0; jj < 32; ++jj) {
          const int j = w * 32 + jj;
          if (j <= i || j >= cn || !alive[j]) continue;
          if (by_class && ccls[j] != ci) continue;
          if (iou_gt(bi, ai, cbo[j], carea[j], thr)) bits |= 1u << jj;
        }
      }
      mask[t] = bits;
    }
    __syncthreads();
    // (3) resolve the chunk, one 32-candidate word at a time
    if (tid < 32) {
      unsigned removed = 0xffffffffu;
      if (lane < kChunkWords) {
        unsigned a = 0;
        for (int jj = 0; jj < 32; ++jj) a |= (alive[lane * 32 + jj] ? 1u : 0u) << jj;
        removed = ~a;
      }
      int total = kc, nnew = 0;
      for (int w = 0; w < kChunkWords && total < p.max_det; ++w) {
        unsigned avail = ~__shfl_sync(0xffffffffu, removed, w);
        while (avail && total < p.max_det) {
          const int i = __ffs(avail) - 1;
          const int gi = w * 32 + i;
          avail &= ~(1u << i);
          if (lane == 0) kept_idx[nnew] = gi;
          ++nnew;
          ++total;
          const unsigned m = lane < kChunkWords ? mask[gi * kChunkWords + lane] : 0u;
          removed |= m;
          avail &= ~__shfl_sync(0xffffffffu, m, w);
        }
      }
      if (lane == 0) s_new = nnew;
    }
    __syncthreads();
    // (4) append the newly kept boxes and emit their output rows
    const int nnew = s_new;
    if (tid < nnew) {
      const int gi = kept_idx[tid];
      const int k = kc + tid;
      kbo[k] = cbo[gi];
      karea[k] = carea[gi];
      kcls[k] = ccls[gi];
      float* o = dst.det + (long long)k * 6;
      if (dst.keys) dst.keys[k] = sorted[cs + gi];
      const float4 bx = cbox[gi];
      o[0] = bx.x;
      o[1] = bx.y;
      o[2] = bx.z;
      o[3] = bx.w;
      o[4] = cconf[gi];
      o[5] = static_cast<float>(ccls[gi]);
    }
    __syncthreads();
    if (tid == 0) s_kept = kc + nnew;
    __syncthreads();
  }
  if (tid == 0) {
    *dst.count = s_kept;
    if (s_wide && wide_flag) atomicExch(wide_flag, 1);
  }
}

// One CTA per (image, class group). With one group the CTA resolves the whole image straight into the output. With G > 1
// (64 images leave more than half of the 148 SMs idle, and a group's sort / scan is a fraction of the image's) each CTA
// resolves its classes into a partial list; the LAST CTA of an image to finish (ticket counter) merges the partial lists
// by key -- every row's rank is its own index plus the number of smaller keys in the other lists -- or, if the class
// split was invalid for this image (a box outside the disjoint-class window, more candidates than the shared-memory
// sort holds, a max_nms cut), resolves the whole image again by itself. Tickets and flags are left zeroed for the next launch.
__global__ void __launch_bounds__(kNmsThreads, 1)
    nms_sort_scan_kernel(BoxSource src, ay2_nms_params p, unsigned long long* __restrict__ keys, long long key_stride,
                         const int* __restrict__ counts, float* __restrict__ out_det, int* __restrict__ out_count,
                         int* __restrict__ overflow, long long* __restrict__ trace, const float* __restrict__ wh_scale, int G,
                         float* __restrict__ part_det, unsigned long long* __restrict__ part_keys, int* __restrict__ part_count,
                         int* __restrict__ done, int* __restrict__ wide, int dense, unsigned long long* __restrict__ keys2,
                         int* __restrict__ counts_rw) {
  const int b = blockIdx.x / G, g = blockIdx.x - b * G;
  const NmsDst whole{out_det + (long long)b * p.max_det * 6, nullptr, out_count + b};
  int* count_back = dense ? counts_rw + b : nullptr;
  if (G == 1) {
    nms_image(src, p, keys, key_stride, counts, whole, overflow, trace, wh_scale, b, 1, 0, nullptr, dense, keys2, count_back);
    return;
  }
  const int tid = threadIdx.x;
  const int n_all = dense ? 0 : counts[b];  // (dense mode: unknown before the scan; an oversized group reports itself)
  const bool splittable = n_all <= kSortSmemKeys && n_all <= p.max_nms && n_all <= p.max_candidates && (!dense || p.n <= p.max_nms);
  const long long slot = (long long)b * G + g;
  if (splittable) {
    const NmsDst part{part_det + slot * p.max_det * 6, part_keys + slot * p.max_det, part_count + slot};
    nms_image(src, p, keys, key_stride, counts, part, overflow, trace, wh_scale, b, G, g, wide + b, dense, keys2, count_back);
  } else if (tid == 0) {
    atomicExch(wide + b, 1);
  }
  __shared__ int s_ticket;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(done + b, 1);
  __syncthreads();
  if (s_ticket != G - 1) return;
  __threadfence();
  if (*reinterpret_cast<volatile int*>(wide + b)) {
    nms_image(src, p, keys, key_stride, counts, whole, overflow, trace, wh_scale, b, 1, 0, nullptr, dense, keys2,
              splittable ? nullptr : count_back);  // (the groups have already reported the candidate count, if they ran)
  } else {
    int total = 0;
    for (int h = 0; h < G; ++h) total += *reinterpret_cast<volatile int*>(part_count + (long long)b * G + h);
    for (int e = tid; e < G * p.max_det; e += blockDim.x) {
      const int h = e / p.max_det, i = e - h * p.max_det;
      const int cnt = *reinterpret_cast<volatile int*>(part_count + (long long)b * G + h);
      if (i >= cnt) continue;
      const unsigned long long* mine = part_keys + ((long long)b * G + h) * p.max_det;
      const unsigned long long key = __ldcg(mine + i);
      int rank = i;
      for (int o = 0; o < G; ++o) {
        if (o == h) continue;
        const unsigned long long* other = part_keys + ((long long)b * G + o) * p.max_det;
        int lo = 0, hi = *reinterpret_cast<volatile int*>(part_count + (long long)b * G + o);
        while (lo < hi) {  // number of keys of list o below `key` (keys are unique: score bits, then candidate index)
          const int mid = (lo + hi) >> 1;
          if (__ldcg(other + mid) < key) lo = mid + 1;
          else hi = mid;
        }
        rank += lo;
      }
      if (rank < p.max_det) {
        const float* r = part_det + (((long long)b * G + h) * p.max_det + i) * 6;
        float* o6 = whole.det + (long long)rank * 6;
#pragma unroll
        for (int q = 0; q < 6; ++q) o6[q] = __ldcg(r + q);
      }
    }
    if (tid == 0) *whole.count = min(total, p.max_det);
  }
  __syncthreads();
  if (tid == 0) {
    done[b] = 0;
    wide[b] = 0;
  }
}

constexpr size_t kNmsChunkBytes = sizeof(float4) * (2 * kChunk + kMaxDetCap) + sizeof(float) * (2 * kChunk + kMaxDetCap) +
                                  sizeof(int) * (3 * kChunk + kMaxDetCap) + sizeof(unsigned) * kChunk * kChunkWords;
constexpr size_t kNmsFastBytes = sizeof(float4) * kFastN + sizeof(float) * kFastN + 6 * sizeof(int) * kFastN + kFastN;
static_assert(sizeof(unsigned long long) * kSortSmemKeys >= sizeof(unsigned long long) * 2 * kFastN + sizeof(float4) * kFastN,
              "the key buffer holds the sorted keys, the (class, rank) keys and the segment-ordered boxes");
static_assert(kFastN <= 2 * kNmsThreads && kGroups * kGroupWarps * 32 == kNmsThreads, "two keys per thread; warp groups");
static_assert(2 * sizeof(int) * kFastN >= sizeof(unsigned short) * (kFastN / 32) * kCountNc && kCountNc % 32 == 0,
              "the counting-sort table overlays klist + sdin");
constexpr size_t kNmsSmemBytes = sizeof(unsigned long long) * kSortSmemKeys +
                                 (kNmsChunkBytes > kNmsFastBytes ? kNmsChunkBytes : kNmsFastBytes);

static long long key_stride_for(const ay2_nms_params* p) {
  long long s = 2;
  while (s < p->max_candidates) s <<= 1;
  return s;
}

}  // namespace ay2

using namespace ay2;

// workspace layout: [counts int32 x B][overflow int32][row_counts int32 x B][done int32 x B][wide int32 x B][pad to 256]
//                   [keys u64 x B x key_stride][keys2 u64 x B x key_stride][rows u32 x B x n]
//                   [partial keys u64 x B x G x max_det][partial rows f32 x B x G x max_det x 6][partial counts int32 x B x G]
constexpr int kMaxGroups = 4;  // class groups (CTAs) per image
static size_t nms_head_ints(const ay2_nms_params* p) { return 4 * (size_t)p->batch + 1; }
static size_t nms_head_bytes(const ay2_nms_params* p) { return ((sizeof(int) * nms_head_ints(p) + 255) / 256) * 256; }
static size_t nms_rows_bytes(const ay2_nms_params* p) { return ((sizeof(unsigned) * (size_t)p->batch * (size_t)p->n + 255) / 256) * 256; }
extern "C" size_t ay2_nms_workspace_bytes(const ay2_nms_params* p) {
  if (!p) return 0;
  const size_t parts = (size_t)p->batch * kMaxGroups;
  return nms_head_bytes(p) + 2 * sizeof(unsigned long long) * (size_t)p->batch * (size_t)key_stride_for(p) + nms_rows_bytes(p) +
         parts * p->max_det * (sizeof(unsigned long long) + 6 * sizeof(float)) + parts * sizeof(int);
}

namespace ay2 {
NmsWorkspaceView nms_workspace_view(const ay2_nms_params* p, void* workspace) {
  NmsWorkspaceView v;
  v.counts = static_cast<int*>(workspace);
  v.overflow = v.counts + p->batch;
  v.row_counts = v.overflow + 1;
  v.done = v.row_counts + p->batch;
  v.wide = v.done + p->batch;
  v.keys = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + nms_head_bytes(p));
  v.key_stride = key_stride_for(p);
  v.keys2 = v.keys + (size_t)p->batch * v.key_stride;
  v.rows = reinterpret_cast<unsigned*>(v.keys2 + (size_t)p->batch * v.key_stride);
  const size_t parts = (size_t)p->batch * kMaxGroups;
  v.part_keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(v.rows) + nms_rows_bytes(p));
  v.part_det = reinterpret_cast<float*>(v.part_keys + parts * p->max_det);
  v.part_count = reinterpret_cast<int*>(v.part_det + parts * p->max_det * 6);
  return v;
}
}  // namespace ay2

static int nms_generate_candidates(const BoxSource& src, const ay2_nms_params* p, const uint8_t* class_mask,
                                   const NmsWorkspaceView& v, cudaStream_t st) {
  AY2_CHECK_CUDA(cudaMemsetAsync(v.counts, 0, sizeof(int) * nms_head_ints(p), st));
  nms_rows_kernel<<<dim3(16, p->batch), 256, 0, st>>>(src, *p, v.rows, v.row_counts);
  AY2_CHECK_LAUNCH();
  nms_score_rows_kernel<<<dim3(16, p->batch), 256, 0, st>>>(src, *p, class_mask, v.rows, v.row_counts, v.keys, v.key_stride,
                                                            v.counts);
  AY2_CHECK_LAUNCH();
  return AY2_OK;
}

static int nms_common_checks(const ay2_nms_params* p, const void* workspace, size_t workspace_bytes, const void* out_det,
                             const void* out_count) {
  AY2_REQUIRE(p && workspace && out_det && out_count, "ay2_nms: null pointer");
  AY2_REQUIRE(p->no > 5 && p->n > 0 && p->batch > 0, "ay2_nms: bad shape (batch=%d n=%d no=%d)", p->batch, p->n, p->no);
  AY2_REQUIRE(p->max_det >= 1 && p->max_det <= kMaxDetCap, "max_det=%d unsupported (1..%d)", p->max_det, kMaxDetCap);
  AY2_REQUIRE(p->max_candidates >= 1, "max_candidates must be positive");
  AY2_REQUIRE((long long)p->n * (p->no - 5) < (1LL << 32), "n*nc does not fit the 32-bit candidate index");
  AY2_REQUIRE(workspace_bytes >= ay2_nms_workspace_bytes(p), "NMS workspace too small (%zu < %zu)", workspace_bytes,
              ay2_nms_workspace_bytes(p));
  return AY2_OK;
}

static int nms_sort_scan_launch(const BoxSource& src, const ay2_nms_params* p, const NmsWorkspaceView& v, float* out_det,
                                int32_t* out_count, int32_t* overflow_flag, cudaStream_t st, const float* wh_scale = nullptr,
                                bool dense = false) {
  // the opt-in to > 48 KB of dynamic shared memory is per (function, device): one flag per device, not per process
  static DeviceOnce once;
  if (once.first())
    AY2_CHECK_CUDA(
        cudaFuncSetAttribute(nms_sort_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNmsSmemBytes));
  // AY2_NMS_TRACE=1: clock64 stamps of image 0's phases (debugging aid; synchronises, never set in production)
  static const bool want_trace = getenv("AY2_NMS_TRACE") != nullptr;
  static long long* trace = nullptr;
  if (want_trace && !trace) {
    AY2_CHECK_CUDA(cudaMalloc(&trace, 16 * sizeof(long long)));
    AY2_CHECK_CUDA(cudaMemset(trace, 0, 16 * sizeof(long long)));
  }
  // class groups per image: as many CTAs as the SMs can hold at once (one CTA per SM), at most kMaxGroups; greedy NMS only
  // splits along classes when classes are separated by offsets (not agnostic) and there is more than one class
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const int env_groups = getenv("AY2_NMS_GROUPS") ? atoi(getenv("AY2_NMS_GROUPS")) : 0;
  int G = env_groups > 0 ? env_groups : sms / p->batch;
  if (G > kMaxGroups) G = kMaxGroups;
  if (G < 1 || p->agnostic || p->no - 5 < 2 || wh_scale) G = 1;
  nms_sort_scan_kernel<<<p->batch * G, kNmsThreads, kNmsSmemBytes, st>>>(src, *p, v.keys, v.key_stride, v.counts, out_det, out_count,
                                                                         v.overflow, trace, wh_scale, G, v.part_det, v.part_keys,
                                                                         v.part_count, v.done, v.wide, dense ? 1 : 0, v.keys2, v.counts);
  AY2_CHECK_LAUNCH();
  if (trace) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs == cudaStreamCaptureStatusNone) {
      long long h[16];
      AY2_CHECK_CUDA(cudaStreamSynchronize(st));
      AY2_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[ay2 nms trace] img0 n=%lld segments %lld one-block + %lld longer, %lld blocks | clocks: sort %lld, boxes %lld, "
              "class order %lld, offset boxes %lld, in-block tests %lld, resolve %lld, emit %lld (scan %lld = count %lld + barrier %lld + rest, write %lld, tail %lld)\n",
              h[8], h[9], h[10], h[11], h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[7] - h[6],
              h[12] - h[6], h[14] - h[6], h[15] - h[14], h[13] - h[12], h[7] - h[13]);
    }
  }
  if (overflow_flag) AY2_CHECK_CUDA(cudaMemcpyAsync(overflow_flag, v.overflow, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return AY2_OK;
}

static int nms_batched_impl(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, const float* max_coord,
                            void* workspace, size_t workspace_bytes, float* out_det, int32_t* out_count, int32_t* overflow_flag,
                            void* stream) {
  AY2_REQUIRE(pred, "ay2_nms_batched: null prediction pointer");
  int rc = nms_common_checks(p, workspace, workspace_bytes, out_det, out_count);
  if (rc != AY2_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BoxSource src;
  memset(&src, 0, sizeof(src));
  src.pred = pred;
  src.n = p->n;
  src.no = p->no;
  const NmsWorkspaceView v = nms_workspace_view(p, workspace);
  rc = nms_generate_candidates(src, p, class_mask, v, st);
  if (rc != AY2_OK) return rc;
  rc = nms_sort_scan_launch(src, p, v, out_det, out_count, overflow_flag, st, max_coord);
  if (rc != AY2_OK) return rc;
  count_launch(3);
  return AY2_OK;
}

extern "C" int ay2_nms_batched(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, void* workspace,
                               size_t workspace_bytes, float* out_det, int32_t* out_count, int32_t* overflow_flag,
                               void* stream) {
  return nms_batched_impl(pred, p, class_mask, nullptr, workspace, workspace_bytes, out_det, out_count, overflow_flag, stream);
}

extern "C" int ay2_nms_batched_scaled(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, const float* max_coord,
                                      void* workspace, size_t workspace_bytes, float* out_det, int32_t* out_count,
                                      int32_t* overflow_flag, void* stream) {
  AY2_REQUIRE(max_coord, "ay2_nms_batched_scaled: null per-image coordinate maxima");
  return nms_batched_impl(pred, p, class_mask, max_coord, workspace, workspace_bytes, out_det, out_count, overflow_flag, stream);
}

static int box_source_from_levels(const ay2_head_levels* hl, const ay2_nms_params* p, BoxSource* out) {
  AY2_REQUIRE(hl, "ay2_nms: null level table");
  AY2_REQUIRE(hl->nl >= 1 && hl->nl <= AY2_NMS_MAX_LEVELS && hl->na >= 1 && hl->na <= AY2_NMS_MAX_ANCHORS,
              "ay2_nms: nl=%d na=%d unsupported", hl->nl, hl->na);
  BoxSource& src = *out;
  memset(&src, 0, sizeof(src));
  src.no = p->no;
  src.nl = hl->nl;
  src.na = hl->na;
  int rows = 0;
  for (int l = 0; l < hl->nl; ++l) {
    AY2_REQUIRE(hl->logits[l] && hl->cstride[l] >= hl->na * p->no, "level %d: bad logits pointer / channel stride", l);
    src.logits[l] = static_cast<const __nv_bfloat16*>(hl->logits[l]);
    src.ny[l] = hl->ny[l];
    src.nx[l] = hl->nx[l];
    src.cstride[l] = hl->cstride[l];
    src.stride_px[l] = hl->stride_px[l];
    src.row_off[l] = rows;
    rows += hl->na * hl->ny[l] * hl->nx[l];
    for (int a = 0; a < hl->na; ++a) {
      src.anchor_px[l][a][0] = hl->anchor_px[l][a][0];
      src.anchor_px[l][a][1] = hl->anchor_px[l][a][1];
    }
  }
  src.row_off[hl->nl] = rows;
  src.n = rows;
  AY2_REQUIRE(rows == p->n, "level table covers %d rows but params.n = %d", rows, p->n);
  return AY2_OK;
}

extern "C" int ay2_nms_candidates_begin(const ay2_nms_params* p, void* workspace, size_t workspace_bytes, void* stream) {
  AY2_REQUIRE(p && workspace, "ay2_nms_candidates_begin: null pointer");
  AY2_REQUIRE(workspace_bytes >= ay2_nms_workspace_bytes(p), "NMS workspace too small (%zu < %zu)", workspace_bytes,
              ay2_nms_workspace_bytes(p));
  const NmsWorkspaceView v = nms_workspace_view(p, workspace);
  AY2_CHECK_CUDA(cudaMemsetAsync(v.counts, 0, sizeof(int) * nms_head_ints(p), static_cast<cudaStream_t>(stream)));
  if (nms_dense_slots(p))  // row-indexed key slots: every slot starts as "no candidate" (~0); only the n used slots per image
    AY2_CHECK_CUDA(cudaMemset2DAsync(v.keys, sizeof(unsigned long long) * v.key_stride, 0xFF, sizeof(unsigned long long) * p->n,
                                     p->batch, static_cast<cudaStream_t>(stream)));
  return AY2_OK;
}

extern "C" int ay2_nms_from_candidates(const ay2_head_levels* hl, const ay2_nms_params* p, void* workspace,
                                       size_t workspace_bytes, float* out_det, int32_t* out_count, int32_t* overflow_flag,
                                       void* stream) {
  int rc = nms_common_checks(p, workspace, workspace_bytes, out_det, out_count);
  if (rc != AY2_OK) return rc;
  BoxSource src;
  rc = box_source_from_levels(hl, p, &src);
  if (rc != AY2_OK) return rc;
  rc = nms_sort_scan_launch(src, p, nms_workspace_view(p, workspace), out_det, out_count, overflow_flag,
                            static_cast<cudaStream_t>(stream), nullptr, nms_dense_slots(p));
  if (rc != AY2_OK) return rc;
  count_launch(1);
  return AY2_OK;
}

extern "C" int ay2_nms_from_logits(const ay2_head_levels* hl, const ay2_nms_params* p, const uint8_t* class_mask,
                                   void* workspace, size_t workspace_bytes, float* out_det, int32_t* out_count,
                                   int32_t* overflow_flag, void* stream) {
  int rc = nms_common_checks(p, workspace, workspace_bytes, out_det, out_count);
  if (rc != AY2_OK) return rc;
  BoxSource src;
  rc = box_source_from_levels(hl, p, &src);
  if (rc != AY2_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const NmsWorkspaceView v = nms_workspace_view(p, workspace);
  rc = nms_generate_candidates(src, p, class_mask, v, st);
  if (rc != AY2_OK) return rc;
  rc = nms_sort_scan_launch(src, p, v, out_det, out_count, overflow_flag, st);
  if (rc != AY2_OK) return rc;
  count_launch(3);
  return AY2_OK;
}

// ------------------------------------------------------------------------------------------------
// Pairwise IoU matrix (scripts/utils/metrics.py:138-164 box_iou): out[i][j] = inter / (area1[i] + area2[j] - inter),
// inter = clamp(min(rb) - max(lt), 0).prod(). Same fp32 operation order as the reference expression (explicit
// round-to-nearest intrinsics, IEEE division), so the matrix is bit-identical to torch's on identical inputs.
// ------------------------------------------------------------------------------------------------
namespace ay2 {
__global__ void box_iou_kernel(const float4* __restrict__ b1, int n, const float4* __restrict__ b2, int m,
                               float* __restrict__ out) {
  const long long total = (long long)n * m;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / m), j = (int)(idx - (long long)i * m);
    const float4 a = b1[i], b = b2[j];
    const float area1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    out[idx] = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area1, area2), inter));
  }
}
}  // namespace ay2

extern "C" int ay2_box_iou(const float* box1, int32_t n, const float* box2, int32_t m, float* out, void* stream) {
  AY2_REQUIRE(n >= 0 && m >= 0, "ay2_box_iou: negative size");
  if (n == 0 || m == 0) return AY2_OK;
  AY2_REQUIRE(box1 && box2 && out, "ay2_box_iou: null pointer");
  AY2_REQUIRE(((reinterpret_cast<uintptr_t>(box1) | reinterpret_cast<uintptr_t>(box2)) & 15) == 0, "ay2_box_iou: boxes must be 16-byte aligned");
  const long long total = (long long)n * m;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  box_iou_kernel<<<(int)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(box1), n,
                                                                                   reinterpret_cast<const float4*>(box2), m, out);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

// ------------------------------------------------------------------------------------------------
// torchvision.ops.nms on an arbitrary box list (metrics.py:385 and the "batched_nms" / "merge_nms" nms_type branches,
// :391-431): `order` = indices by descending score (stable), greedy suppression IoU > thr with the same exact test as
// the batched kernel (iou_gt). Phase 1 fills the upper-triangular suppression bit matrix over sorted positions,
// phase 2 walks it once (one CTA; the walk is inherently sequential) and emits the kept ORIGINAL indices in score order.
// ------------------------------------------------------------------------------------------------
namespace ay2 {
__global__ void nms_boxes_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ order, int n, double iou_thres,
                                      unsigned long long* __restrict__ mask) {
  const int words = (n + 63) / 64;
  const IouThr t = make_iou_thr(iou_thres);
  const long long total = (long long)n * words;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / words), w = (int)(idx - (long long)i * words);
    unsigned long long bits = 0;
    if (w * 64 + 63 > i) {
      const float4 a = boxes[order[i]];
      const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
      for (int k = 0; k < 64; ++k) {
        const int j = w * 64 + k;
        if (j <= i || j >= n) continue;
        const float4 b = boxes[order[j]];
        const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
        if (iou_gt(a, aa, b, ab, t)) bits |= 1ull << k;
      }
    }
    mask[idx] = bits;
  }
}

__global__ void nms_boxes_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ order, int n,
                                      int* __restrict__ keep, int* __restrict__ count) {
  extern __shared__ unsigned long long remv[];
  const int words = (n + 63) / 64;
  for (int w = threadIdx.x; w < words; w += blockDim.x) remv[w] = 0;
  __shared__ int nkeep;
  if (threadIdx.x == 0) nkeep = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    const bool dead = (remv[i >> 6] >> (i & 63)) & 1ull;  // uniform: every thread reads the same word
    __syncthreads();
    if (!dead) {
      if (threadIdx.x == 0) keep[nkeep++] = order[i];
      for (int w = (i >> 6) + threadIdx.x; w < words; w += blockDim.x) remv[w] |= mask[(long long)i * words + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = nkeep;
}
}  // namespace ay2

extern "C" int ay2_nms_boxes(const float* boxes, const int32_t* order, int32_t n, double iou_thres, unsigned long long* mask_ws,
                             int32_t* keep, int32_t* count, void* stream) {
  AY2_REQUIRE(n >= 0 && count, "ay2_nms_boxes: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    AY2_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    return AY2_OK;
  }
  AY2_REQUIRE(boxes && order && mask_ws && keep, "ay2_nms_boxes: null pointer");
  AY2_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "ay2_nms_boxes: boxes must be 16-byte aligned");
  const int words = (n + 63) / 64;
  AY2_REQUIRE((size_t)words * 8 <= 96 * 1024, "ay2_nms_boxes: n = %d exceeds the scan kernel's shared-memory bit vector", n);
  const long long total = (long long)n * words;
  long long blocks = (total + 127) / 128;
  if (blocks > 148 * 16) blocks = 148 * 16;
  nms_boxes_mask_kernel<<<(int)blocks, 128, 0, st>>>(reinterpret_cast<const float4*>(boxes), order, n, iou_thres, mask_ws);
  AY2_CHECK_LAUNCH();
  count_launch();
  if ((size_t)words * 8 > 48 * 1024)
    AY2_CHECK_CUDA(cudaFuncSetAttribute(nms_boxes_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 8));
  nms_boxes_scan_kernel<<<1, 256, words * 8, st>>>(mask_ws, order, n, keep, count);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
