// Fused Conv2d + folded-BN bias + SiLU (+ residual) as an implicit GEMM on the sm_100a tensor cores.
//
//   out[b, oy, ox, n] = act( bias[n] + sum_{kh,kw,c} in[b, oy*s + kh - p, ox*s + kw - p, c] * w[n, kh, kw, c] ) (+ res)
//
// GEMM view: M = output pixels, N = Cout, K = KH*KW*Cin. One CTA tile is 128 output pixels x BLOCK_N
// channels. The 128 pixels of a tile are NB spatial boxes of BH x BW pixels (BH*BW*NB = 128), so that the
// A operand of one filter tap is NB *tiled* 4-D TMA boxes [CK ch][BW][BH][1] of the NHWC input, shifted by
// the tap offset; TMA's out-of-bounds zero fill implements the convolution padding and every ragged edge
// (partial boxes, M tail), and the same boxes written back with TMA stores implement the clipped output
// write. Stride-2 convolutions read through four parity views of the input (one tensor map per (row,col)
// parity, strides doubled), which makes every tap a stride-1 box again.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles):
//   warp 0 : TMA producer   (A boxes + B tile per K chunk into a NSTAGES ring, mbarrier full/empty)
//   warp 1 : tcgen05.mma issuer (one lane), accumulators double-buffered in TMEM
//   warp 2 : TMEM allocate / free
//   warps 4-7 : epilogue: tcgen05.ld -> +bias -> SiLU -> (+residual) -> bf16 -> swizzled smem -> TMA store
#include <stdlib.h>
#include <string.h>

#include "ay2_common.h"
#include "ay2_ptx.cuh"
#include "head_math.cuh"

namespace ay2 {

struct ConvKernelParams {
  CUtensorMap tmA[4];  // input views: [0] for stride 1; [ph*2+pw] parity views for stride 2
  CUtensorMap tmB;     // weights [Ktot][cout_pad] (K innermost)
  CUtensorMap tmOut;   // output  [C][W][H][B]
  CUtensorMap tmRes;   // residual (same geometry as output) if has_res
  // split-precision mode (x3, see ConvCfg): the output segment is three planes of `cout` channels [hi | lo | hi]; tmOut is
  // plane 0, tmOutX[0..1] planes 1 and 2; the residual is read as hi (tmRes) + lo (tmResLo)
  CUtensorMap tmOutX[2];
  CUtensorMap tmResLo;
  const float* bias;   // [cout_pad]
  int num_m_tiles, num_n_tiles;
  int boxes_x, boxes_per_img;
  int BH, BW, NB;
  int kh, kw, stride, pad, pad_w, stride_w;
  int cin_chunks;  // Cin / CK
  int cin;         // K extent per tap in the packed weights
  int split_chunks;  // > 0: channel chunks >= split_chunks come from the second input view tmA[3] (stride-1 convs only):
                     // the consumer-side form of torch.cat([a, b], 1) when a and b live in different buffers
  int cout_pad;
  int act, has_res;
  int csize;  // CTAs per cluster (1 or 2): with 2, the pair works on neighbouring M tiles of the same N tile and each
              // CTA fetches half of every weight (B) tile, multicast into both CTAs' shared memory
  HeadCandParams hc;  // detect head only: NMS candidates are scored and appended from the staged output tile
  // halo kernel (conv_halo_kernel): work items are 16-row bands x pairs of 8-pixel-wide half tiles
  int hl_pairs_x, hl_bands_y, hl_w_halves;
  int experiment;           // diagnostics only (AY2_CONV_EXPERIMENT): bit 0 = the MMA warp skips the tcgen05.mma instructions
  unsigned long long* dbg;  // optional timeline buffer (tools/conv_timeline.py): every CTA records %globaltimer per phase
};

__device__ __forceinline__ unsigned long long cv_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CV_DBG(slot)                                                \
  do {                                                              \
    if (p.dbg) p.dbg[(size_t)blockIdx.x * 16 + (slot)] = cv_now();  \
  } while (0)

constexpr int kSmemPerSm = 227 * 1024;

// X3 = split-precision verification mode ("bf16x3"): every activation v is stored as hi = bf16(v), lo = bf16(v - hi) in three
// channel planes [hi | lo | hi] and every weight as [w_hi | w_hi | w_lo] along K, so the unchanged main loop accumulates
// x_hi w_hi + x_lo w_hi + x_hi w_lo in fp32: fp32-equivalent products (the dropped x_lo w_lo term is 2^-16 relative) on
// the same TMA / tcgen05 pipeline. Only the epilogue differs: exact SiLU, hi/lo split, three plane stores.
// PAIR = the CTA-pair form (ay2_ptx.cuh "CTA pair"): a cluster of two CTAs computes two neighbouring M tiles of one N tile
// with ONE M = 256 tcgen05.mma.cta_group::2 per K step. Each CTA stages its own A tile and HALF of the weight tile, so
// the weight bytes cross L2 -> SM (and occupy shared memory) once per pair: the stage shrinks from (128 + N) to
// (128 + N / 2) operand rows and the ring gets deeper. Accumulators, epilogue and stores stay per CTA.
template <int BLOCK_N, int CK, bool X3 = false, bool PAIR = false>
struct ConvCfg {
  static constexpr bool kX3 = X3;
  static constexpr bool kPair = PAIR;
  static constexpr int BN = BLOCK_N;
  static constexpr int SWA = CK * 2;                 // operand row bytes == swizzle span
  static constexpr int A_BYTES = 128 * SWA;
  static constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * SWA;  // per CTA
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OC = BLOCK_N < 64 ? BLOCK_N : 64;  // channels per output slab
  static constexpr int SWO = OC * 2;                       // output slab row bytes == swizzle span
  static constexpr int SLAB_BYTES = 128 * SWO;
  static constexpr int NSLAB = BLOCK_N / OC;
  static constexpr int STAGING_BYTES = 128 * BLOCK_N * 2 * (X3 ? 2 : 1);  // x3: hi slabs, then lo slabs
  // bias slice + barriers (2 NSTAGES + 4 + 4 <= 24) + tmem ptr (256 B) + head-candidate lists: (pixel, anchor) entries
  // (768 B), their keys (3072 B) and images (768 B)
  static constexpr int TAIL_BYTES = BLOCK_N * 4 + 256 + 768 + 3072 + 768 + 256;
  // Small tiles are latency-bound per tile (TMA round trip, TMEM drain, store hand-off): co-residency of
  // several CTAs per SM interleaves independent tile streams. TMEM: CTAS_PER_SM * 2 * BLOCK_N <= 512 columns.
  static constexpr int CTAS_PER_SM = (X3 || PAIR) ? 1 : (BLOCK_N <= 64 ? 3 : (BLOCK_N == 128 ? 2 : 1));
  // Epilogue warps: one group of 4 warps covers the 128 TMEM lanes and owns ONE output slab (<= 64 columns): it drains
  // it, stores it with TMA and waits for its own store only -- groups never synchronise with each other, and every
  // scheduler has EPI_GROUPS x CTAS_PER_SM epilogue warps to interleave (the drain is latency-bound per warp).
  static constexpr int EPI_GROUPS = NSLAB;
  static constexpr int EPI_THREADS = 128 * EPI_GROUPS;
  static constexpr int THREADS = 128 + EPI_THREADS;
  static constexpr int COLS_PER_GROUP = BLOCK_N / EPI_GROUPS;
  static constexpr int SMEM_BUDGET = kSmemPerSm / CTAS_PER_SM - 1024;  // 1 KB per CTA is reserved by the system
  static constexpr int NSTAGES_RAW = (SMEM_BUDGET - 1024 - STAGING_BYTES - TAIL_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGES = NSTAGES_RAW > 8 ? 8 : NSTAGES_RAW;
  static constexpr int SMEM_BYTES = 1024 + NSTAGES * STAGE_BYTES + STAGING_BYTES + TAIL_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // 64..512, power of two
  static_assert(NSTAGES >= 2, "not enough shared memory for a pipeline");
  static_assert(!(X3 && PAIR), "the split-precision verification mode runs single CTAs");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two");
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Detect-head epilogue: NMS candidate generation from the staged bf16 output tile (metrics.py:313-364).
// The tile holds all na*no logits of 128 pixels; epilogue thread `et` owns one pixel, the epilogue groups split the
// anchors. A row passes when obj > conf (:313,337); its score is conf_c = cls_c * obj (:353) with the first arg-max
// class (:363-364) or every class above conf (multi_label, :360-361) -- the same arithmetic (head_math.cuh) on the
// same stored bf16 logits as the stand-alone kernels in nms.cu, so the keys are bit-identical to theirs. Keys are
// appended to the per-image list with one atomicAdd per warp and image.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void head_cand_push(const HeadCandParams& h, bool ok, int b, unsigned long long key, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, ok);
  while (m) {  // one round per distinct image among the passing lanes (a warp's 32 pixels rarely span two images)
    const int leader = __ffs(m) - 1;
    const int bl = __shfl_sync(0xffffffffu, b, leader);
    const unsigned same = __ballot_sync(0xffffffffu, ok && b == bl);
    int base = 0;
    if (lane == leader) base = atomicAdd(&h.counts[bl], __popc(same));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok && b == bl) {
      const int pos = base + __popc(same & ((1u << lane) - 1u));
      if (pos < h.max_candidates) h.keys[(long long)bl * h.key_stride + pos] = key;
    }
    m &= ~same;
  }
}

// Three phases per tile. (1) Every epilogue thread owns one pixel and tests the objectness of its group's anchors; passing
// (pixel, anchor) pairs are appended to a shared-memory list (warp-aggregated) together with their image and row number.
// (2) After a barrier the list is walked ONE ENTRY PER THREAD: the thread streams the entry's class logits out of the
// staged tile with 16-byte shared-memory reads (32 lanes reading 32 different 128-byte rows of a swizzled slab hit every
// bank group equally: no conflicts beyond the 4 wavefronts 512 bytes need) -- a warp's instruction stream serves 32 rows
// at once (r02 measurements: a warp per row issued ~3000 clk per row; 80 sigmoids per row before that).
//   Best class (metrics.py:362-364) WITHOUT a sigmoid per class: conf_c = sigmoid(z_c) * obj is non-decreasing in the logit
//   z_c, so only classes whose logit lies within a small window of the row's largest logit can attain the maximal conf;
//   the window (1/16 below the maximum, or everything above 11 where fp32 sigmoids saturate into ties) is wide enough that
//   classes outside it differ from the maximum by > 16 ulps of the sigmoid, far beyond the 2-ulp error of the fast
//   sigmoid. Pass 1 finds the largest logit, pass 2 evaluates exact conf values for the window only (usually one class),
//   first arg-max among them.
//   multi_label (metrics.py:359-361) needs every conf_c > conf: classes are pre-filtered in logit space by the same
//   monotonicity (sigmoid(z) * obj > T  =>  z > logit(T / obj) - margin) and evaluated exactly when they may pass.
// (3) The tile's keys are appended with one global atomicAdd per warp and image (single-label; multi_label appends per
// class step): one atomic per ROW would serialise in L2 on the image's counter.
template <class Cfg>
__device__ __forceinline__ void head_candidates(const ConvKernelParams& p, const uint8_t* staging, int3 epix, const int* box_s,
                                                int et, int egrp, int lane, int eall, unsigned short* list, int* cnt,
                                                unsigned long long* keys_s, short* img_s) {
  const HeadCandParams& h = p.hc;
  const int plane = h.out_h * h.out_w;
  const int nc = h.no - 5;
  auto logit = [&](int px, int ch) -> float {  // channel ch of pixel px in the swizzled staging slabs
    const uint8_t* slab = staging + (ch / Cfg::OC) * Cfg::SLAB_BYTES;
    const int cc = ch % Cfg::OC;
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(slab + swizzled_offset<Cfg::SWO>(px, cc >> 3) + (cc & 7) * 2));
  };
  // ---- phase 1: objectness test, list of passing (pixel, anchor) pairs with their (image, row)
  // (image, origin) of this thread's box come from the group-0 leader's per-tile index arithmetic (box_s); the thread's
  // place inside its box (epix = box, row, column) does not change from tile to tile: no divisions here
  const int b0 = box_s[epix.x * 3], oy0 = box_s[epix.x * 3 + 1] + epix.y, ox0 = box_s[epix.x * 3 + 2] + epix.z;
  const bool inside = b0 < h.batch && oy0 < h.out_h && ox0 < h.out_w;
  for (int a = egrp; a < h.na; a += Cfg::EPI_GROUPS) {
    const bool pass = inside && head_sigmoid(logit(et, a * h.no + 4)) > h.conf_thres;
    const unsigned mk = __ballot_sync(0xffffffffu, pass);
    if (mk) {
      int base = 0;
      if (lane == 0) base = atomicAdd(cnt, __popc(mk));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) {
        const int e = base + __popc(mk & ((1u << lane) - 1u));
        list[e] = static_cast<unsigned short>(et | (a << 8));
        img_s[e] = static_cast<short>(b0);
        keys_s[e] = static_cast<unsigned>(h.row_off + a * plane + oy0 * h.out_w + ox0);  // the row; phase 2 turns it into the key
      }
    }
  }
  named_bar_sync(1, Cfg::EPI_THREADS);
  const int n = *reinterpret_cast<volatile int*>(cnt);
  // ---- phase 2: EIGHT lanes per list entry (four entries per warp and pass). A lane holds at most two 16-byte chunks
  // (16 logits) of the entry's class range in registers: one shared-memory round trip, a 3-step shuffle for the largest
  // logit, the exact conf of the window classes from the registers, a 3-step shuffle for the first arg-max. (One thread
  // per entry made a ~4000-clk dependent chain that three warps walked while thirteen waited at the barrier below;
  // one warp per entry issued ~3000 clk per row.)
  constexpr int kWarps = Cfg::EPI_THREADS / 32;
  const int sub = lane & 7, slot = lane >> 3, wall = eall >> 5;
  const int passes = (n + 4 * kWarps - 1) / (4 * kWarps);
  for (int ps = 0; ps < passes; ++ps) {
    const int e = (ps * kWarps + wall) * 4 + slot;
    const bool have = e < n;
    if (!__ballot_sync(0xffffffffu, have)) continue;  // warp-uniform
    const int ent = have ? list[e] : 0;
    const int px = ent & 0xff, a = ent >> 8;
    const int c0 = a * h.no;
    const int b = have ? img_s[e] : 0;
    const unsigned row = have ? static_cast<unsigned>(keys_s[e]) : 0u;
    const float obj = have ? head_sigmoid(logit(px, c0 + 4)) : 0.0f;
    const int ch_lo = c0 + 5, ch_hi = c0 + h.no;  // class channels [ch_lo, ch_hi)
    const int k_lo = ch_lo >> 3, k_hi = (ch_hi + 7) >> 3;
    float z[16];
    int chb[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int k = k_lo + sub + 8 * q;
      chb[q] = 8 * k;
      uint4 v = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // bf16 -inf
      if (have && k < k_hi) {
        const uint8_t* slab = staging + (k / (Cfg::OC / 8)) * Cfg::SLAB_BYTES;
        v = *reinterpret_cast<const uint4*>(slab + swizzled_offset<Cfg::SWO>(px, k % (Cfg::OC / 8)));
      }
      const __nv_bfloat16* hv = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ch = chb[q] + i;
        z[8 * q + i] = (ch >= ch_lo && ch < ch_hi) ? __bfloat162float(hv[i]) : -INFINITY;
      }
    }
    // (a head with more than 16 x 8 = 128 class channels per anchor would need more chunks per lane)
    if (h.multi_label) {
      // sigmoid(z) * obj > T needs sigmoid(z) > T / obj; in logit space, with a margin far above the fast sigmoid's error
      const float r = have ? __fdividef(h.conf_thres, obj) : 2.0f;
      const float zmin = r >= 1.0f ? INFINITY : (r <= 0.0f ? -INFINITY : __logf(__fdividef(r, 1.0f - r)) - 0.0625f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = chb[i >> 3] + (i & 7) - ch_lo;
        float conf = 0.0f;
        bool ok = false;
        if (z[i] >= zmin) {  // (-inf for channels outside the class range / lanes without an entry)
          conf = __fmul_rn(head_sigmoid(z[i]), obj);
          ok = conf > h.conf_thres && (!h.class_mask || h.class_mask[c]);
        }
        if (__ballot_sync(0xffffffffu, ok))
          head_cand_push(h, ok, b, (static_cast<unsigned long long>(~__float_as_uint(conf)) << 32) | (row * nc + c), lane);
      }
    } else {
      // Arg-max in LOGIT space: the lane's first largest logit (ascending class order inside a lane), then the entry's
      // over its 8 lanes (larger logit, then lower class). Below 11 two different bf16 logits give different fp32 scores
      // (their sigmoids differ by >= 1e-6 relative, far above the 2^-22 error of ex2 / rcp and the product's rounding), and
      // equal logits give equal scores, so this IS the first arg-max of the scores (metrics.py:363-364) and only the
      // winner's score is ever computed. Above 11 the sigmoid saturates (distinct logits may round to one score, and the
      // lower class must win): those rows take the window scan below.
      float zb = z[0];
      int ib = 0;
#pragma unroll
      for (int i = 1; i < 16; ++i)
        if (z[i] > zb) {
          zb = z[i];
          ib = i;
        }
      int cb = chb[ib >> 3] + (ib & 7) - ch_lo;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float oz = __shfl_xor_sync(0xffffffffu, zb, o);
        const int oc = __shfl_xor_sync(0xffffffffu, cb, o);
        if (oz > zb || (oz == zb && oc < cb)) {
          zb = oz;
          cb = oc;
        }
      }
      float best;
      int bidx;
      if (!__ballot_sync(0xffffffffu, zb > 10.9375f)) {  // warp-uniform: no entry of this pass is near saturation
        best = __fmul_rn(head_sigmoid(zb), obj);
        bidx = cb;
      } else {
        const float zwin = fminf(zb - 0.0625f, 11.0f);  // zb is the entry's largest logit in all of its lanes
        best = -INFINITY;
        bidx = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 16; ++i) {  // ascending class order inside a chunk; the reduction below orders across lanes by index
          if (z[i] >= zwin) {
            const float conf = __fmul_rn(head_sigmoid(z[i]), obj);
            const int c = chb[i >> 3] + (i & 7) - ch_lo;
            if (conf > best || (conf == best && c < bidx)) {
              best = conf;
              bidx = c;
            }
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {  // first arg-max over the entry's 8 lanes: larger score, then lower class
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          if (ob > best || (ob == best && oi < bidx)) {
            best = ob;
            bidx = oi;
          }
        }
      }
      if (have && sub == 0) {
        const bool ok = best > h.conf_thres && (!h.class_mask || h.class_mask[bidx]);
        const unsigned long long key = (static_cast<unsigned long long>(~__float_as_uint(best)) << 32) | (row * nc + bidx);
        if (h.dense_slots) {
          // row-indexed slot: no counter, no atomic, nothing to wait for
          if (ok) h.keys[(long long)b * h.key_stride + row] = key;
        } else {
          keys_s[e] = ok ? key : ~0ull;
        }
      }
    }
  }
  if (!h.multi_label && !h.dense_slots) {
    // ---- phase 3: append the tile's keys, one atomicAdd per warp and image (a tile rarely spans two images)
    named_bar_sync(1, Cfg::EPI_THREADS);
    for (int e0 = (eall >> 5) * 32; e0 < n; e0 += Cfg::EPI_THREADS) {
      const int e = e0 + lane;
      const unsigned long long key = e < n ? keys_s[e] : ~0ull;
      head_cand_push(h, key != ~0ull, e < n ? img_s[e] : 0, key, lane);
    }
  }
}

// Drain this warp group's slab of one accumulator: TMEM -> +bias -> act (-> + residual already staged in the output
// slab) -> bf16 -> swizzled staging slab. Bias comes from shared memory with explicit ld.shared (a pointer derived from
// the aligned dynamic-smem base loses its address space and compiles to generic loads). Activation / residual are
// compile-time: no predicated-off residual code, no branch per column group.
template <class Cfg, bool SILU, bool RES>
__device__ __forceinline__ void drain_accumulator(uint32_t taddr, uint32_t slab, uint32_t bias_u32, int et) {
  constexpr int BLK = Cfg::OC < 32 ? Cfg::OC : 32;  // columns per tcgen05.ld
#pragma unroll 1
  for (int cc = 0; cc < Cfg::OC; cc += BLK) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + cc, v);
    tmem_ld_wait();
    const uint32_t ba = bias_u32 + cc * 4;
#pragma unroll
    for (int g = 0; g < BLK / 8; ++g) {
      float bb[8], f[8];
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[0]), "=f"(bb[1]), "=f"(bb[2]), "=f"(bb[3]) : "r"(ba + g * 32));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[4]), "=f"(bb[5]), "=f"(bb[6]), "=f"(bb[7]) : "r"(ba + g * 32 + 16));
      // two columns per instruction (FADD2 / FMUL2 / FFMA2: each lane rounds exactly like the scalar form, so the result
      // is bit-identical): the drain is bound by instruction issue on the narrow-N layers (~6 instructions per element)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 x = __fadd2_rn(make_float2(__uint_as_float(v[g * 8 + 2 * i]), __uint_as_float(v[g * 8 + 2 * i + 1])),
                              make_float2(bb[2 * i], bb[2 * i + 1]));
        if (SILU) {  // x * sigmoid(x) = h + h * tanh(h), h = x / 2 (silu_f)
          const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
          float tx, ty;
          asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(h.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(h.y));
          x = __ffma2_rn(h, make_float2(tx, ty), h);
        }
        f[2 * i] = x.x;
        f[2 * i + 1] = x.y;
      }
      const uint32_t dst = slab + swizzled_offset<Cfg::SWO>(et, cc / 8 + g);
      if (RES) {
        uint4 rv;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(dst));
        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 rf = __bfloat1622float2(r2[i]);
          f[2 * i] += rf.x;
          f[2 * i + 1] += rf.y;
        }
      }
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16x2(f[0], f[1])),
                   "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7]))
                   : "memory");
    }
  }
}

// Split-precision drain: TMEM -> +bias -> exact SiLU (-> + residual hi + lo already staged) -> hi = bf16(v), lo = bf16(v - hi)
// -> the hi slab and the lo slab (same swizzled layout, `lo_delta` bytes apart).
template <class Cfg, bool SILU, bool RES>
__device__ __forceinline__ void drain_accumulator_x3(uint32_t taddr, uint32_t slab, uint32_t lo_delta, uint32_t bias_u32, int et) {
  constexpr int BLK = Cfg::OC < 32 ? Cfg::OC : 32;
#pragma unroll 1
  for (int cc = 0; cc < Cfg::OC; cc += BLK) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(taddr + cc, v);
    tmem_ld_wait();
    const uint32_t ba = bias_u32 + cc * 4;
#pragma unroll
    for (int g = 0; g < BLK / 8; ++g) {
      float bb[8], f[8];
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[0]), "=f"(bb[1]), "=f"(bb[2]), "=f"(bb[3]) : "r"(ba + g * 32));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[4]), "=f"(bb[5]), "=f"(bb[6]), "=f"(bb[7]) : "r"(ba + g * 32 + 16));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = __uint_as_float(v[g * 8 + i]) + bb[i];
        f[i] = SILU ? __fdividef(x, 1.0f + __expf(-x)) : x;
      }
      const uint32_t dst = slab + swizzled_offset<Cfg::SWO>(et, cc / 8 + g);
      if (RES) {
        uint4 rh, rl;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rh.x), "=r"(rh.y), "=r"(rh.z), "=r"(rh.w) : "r"(dst));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rl.x), "=r"(rl.y), "=r"(rl.z), "=r"(rl.w) : "r"(dst + lo_delta));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rh);
        const __nv_bfloat162* l2 = reinterpret_cast<const __nv_bfloat162*>(&rl);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __bfloat1622float2(h2[i]), b = __bfloat1622float2(l2[i]);
          f[2 * i] += a.x + b.x;  // hi + lo is exact in fp32 (at most 17 significant bits)
          f[2 * i + 1] += a.y + b.y;
        }
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        const float2 hf = __bfloat1622float2(h);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = pack_bf16x2(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
      }
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + lo_delta), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
  }
}

// taddr: TMEM address (lane quadrant + first column) of the slab; slab: its staging bytes; bias_u32: its bias slice
template <class Cfg>
__device__ __forceinline__ void drain_dispatch(const ConvKernelParams& p, uint32_t taddr, uint32_t slab, uint32_t bias_u32, int et) {
  if constexpr (Cfg::kX3) {
    constexpr uint32_t lo_delta = 128 * Cfg::BN * 2;  // the lo slabs follow all hi slabs
    if (p.act == AY2_ACT_SILU) {
      if (p.has_res) drain_accumulator_x3<Cfg, true, true>(taddr, slab, lo_delta, bias_u32, et);
      else drain_accumulator_x3<Cfg, true, false>(taddr, slab, lo_delta, bias_u32, et);
    } else {
      if (p.has_res) drain_accumulator_x3<Cfg, false, true>(taddr, slab, lo_delta, bias_u32, et);
      else drain_accumulator_x3<Cfg, false, false>(taddr, slab, lo_delta, bias_u32, et);
    }
    return;
  }
  if (p.act == AY2_ACT_SILU) {
    if (p.has_res) drain_accumulator<Cfg, true, true>(taddr, slab, bias_u32, et);
    else drain_accumulator<Cfg, true, false>(taddr, slab, bias_u32, et);
  } else {
    if (p.has_res) drain_accumulator<Cfg, false, true>(taddr, slab, bias_u32, et);
    else drain_accumulator<Cfg, false, false>(taddr, slab, bias_u32, et);
  }
}

template <int BLOCK_N, int CK, bool X3, bool PAIR>
__global__ void __launch_bounds__(ConvCfg<BLOCK_N, CK, X3, PAIR>::THREADS, ConvCfg<BLOCK_N, CK, X3, PAIR>::CTAS_PER_SM) conv_tc_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = ConvCfg<BLOCK_N, CK, X3, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: swizzle patterns repeat every 1024 B and UMMA descriptors assume base_offset 0
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint8_t* staging = smem + Cfg::NSTAGES * Cfg::STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(staging + Cfg::STAGING_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + BLOCK_N);
  uint64_t* full_bar = bars;                      // [NSTAGES]
  uint64_t* empty_bar = bars + Cfg::NSTAGES;      // [NSTAGES]
  uint64_t* tmem_full = bars + 2 * Cfg::NSTAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint64_t* res_full = tmem_empty + 2;            // [EPI_GROUPS <= 4]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(res_full + 4);
  int* cand_cnt = reinterpret_cast<int*>(tmem_ptr_s + 2);                               // [2], one per tile parity
  unsigned short* cand_list = reinterpret_cast<unsigned short*>(bars) + 128;           // 256 B past the barriers: 128 x 3 entries
  unsigned long long* cand_keys = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [384]
  short* cand_img = reinterpret_cast<short*>(reinterpret_cast<uint8_t*>(bars) + 1024 + 3072);                      // [384]
  int* cand_box = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 1024 + 3072 + 768);                     // [8][3] image, y, x of the tile's boxes

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int num_k_chunks = p.kh * p.kw * p.cin_chunks;

  if (threadIdx.x == 0) {
    CV_DBG(0);  // CTA entry
    if (p.dbg) p.dbg[(size_t)blockIdx.x * 16 + 9] = clock64();
    cand_cnt[0] = cand_cnt[1] = 0;
    for (int i = 0; i < Cfg::NSTAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      // multicast form: every CTA sharing the B tile must release the stage; pair form: the leader's one commit does
      mbar_init(&empty_bar[i], PAIR ? 1 : p.csize);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      // pair form: one arrival per epilogue group of BOTH CTAs, on the leader's barrier
      mbar_init(&tmem_empty[i], PAIR ? 2 * Cfg::EPI_GROUPS : Cfg::EPI_THREADS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_ptr_s, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_ptr_s, Cfg::TMEM_COLS);
  }
  tcgen05_fence_before();
  uint32_t tmem_base = 0;
  if (p.csize > 1) {
    cluster_sync_all();  // peer barriers must be initialised before any multicast can land
    tcgen05_fence_after();
    tmem_base = *tmem_ptr_s;
  } else if (warp == 0) {
    // The TMA producer never touches TMEM: it only ARRIVES on the prologue barrier (its own thread 0 has initialised the
    // mbarriers above) and starts filling the ring while the other warps still wait for the TMEM allocation.
    asm volatile("bar.arrive 0, %0;" ::"r"(Cfg::THREADS) : "memory");
  } else {
    asm volatile("bar.sync 0, %0;" ::"r"(Cfg::THREADS) : "memory");
    tcgen05_fence_after();
    tmem_base = *tmem_ptr_s;
  }
  if (threadIdx.x == 32) CV_DBG(1);  // prologue done (barriers, TMEM)
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the
  // tail of the previous kernel in the stream; nothing below touches global memory before that kernel has completed.
  // The next kernel may start its own prologue as soon as every CTA of this grid is resident (all are: persistent grid).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // work items: (group of csize neighbouring M tiles, N tile); a cluster walks the item list, CTA `crank` takes its M tile
  const int crank = p.csize > 1 ? (int)cluster_ctarank() : 0;
  const int item0 = blockIdx.x / p.csize;
  const int item_step = gridDim.x / p.csize;
  const int total_items = ((p.num_m_tiles + p.csize - 1) / p.csize) * p.num_n_tiles;
  const uint16_t cmask = (uint16_t)((1u << p.csize) - 1);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // All 32 lanes run the loop: warp-uniform control flow lets the compiler keep coordinates and addresses in uniform
    // registers; one elected lane issues the copies (no per-instruction uniformisation loops in the SASS).
    {
      int stage = 0, phase = 0;
      const int box_rows = p.BH * p.BW;
      for (int item = item0; item < total_items; item += item_step) {
        const int mg = item / p.num_n_tiles;
        const int m = mg * p.csize + crank;
        const int n0 = (item - mg * p.num_n_tiles) * BLOCK_N;
        // Box origins of the tile, warp-uniform (uniform registers feed the TMA instructions directly). A tile of ONE box
        // (8 x 16 pixels: every map of 80 x 80 and above) takes the short path: two divisions per tile and one copy per
        // stage instead of sixteen divisions and an eight-way predicated sequence -- for the K = 32 / one-chunk stages of
        // the early layers that sequence was the longest chain in the kernel.
        const bool one_box = p.NB == 1;
        int bx[8], by[8], bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j > 0 && one_box) break;
          const int q = m * p.NB + j;
          const int b = q / p.boxes_per_img;
          const int r = q - b * p.boxes_per_img;
          const int py = r / p.boxes_x;
          bb[j] = b;
          by[j] = py * p.BH;
          bx[j] = (r - py * p.boxes_x) * p.BW;
        }
        for (int kh = 0; kh < p.kh; ++kh) {
          for (int kw = 0; kw < p.kw; ++kw) {
            // input-box origin relative to the output-box origin for this tap
            int dy, dx, view = 0;
            if (p.stride == 1) {
              dy = kh - p.pad;
              dx = kw - p.pad_w;
            } else {
              const int uy = kh - p.pad, ux = kw - p.pad_w;
              const int ph = uy & 1, pw = p.stride_w == 1 ? 0 : (ux & 1);  // stride_w 1: row-parity views only
              dy = (uy - ph) >> 1;
              dx = p.stride_w == 1 ? ux : ((ux - pw) >> 1);
              view = ph * 2 + pw;
            }
            const int kbase = (kh * p.kw + kw) * p.cin;
            for (int cc = 0; cc < p.cin_chunks; ++cc) {
              const bool second = p.split_chunks > 0 && cc >= p.split_chunks;
              const CUtensorMap* tmA = second ? &p.tmA[3] : &p.tmA[view];
              const int ccoord = (second ? cc - p.split_chunks : cc) * CK;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one()) {
                uint8_t* sa = stages + stage * Cfg::STAGE_BYTES;
                uint8_t* sb = sa + Cfg::A_BYTES;
                if constexpr (PAIR) {
                  // both CTAs' copies complete on the LEADER's barrier, which expects the bytes of the two stages
                  if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                  if (one_box) {
                    tma_load_4d_pair(tmA, &full_bar[stage], sa, ccoord, bx[0] + dx, by[0] + dy, bb[0]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      if (j < p.NB)
                        tma_load_4d_pair(tmA, &full_bar[stage], sa + j * box_rows * Cfg::SWA, ccoord, bx[j] + dx, by[j] + dy, bb[j]);
                    }
                  }
                  tma_load_2d_pair(&p.tmB, &full_bar[stage], sb, kbase + cc * CK, n0 + crank * (BLOCK_N / 2));
                } else {
                  mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                  if (one_box) {
                    tma_load_4d(tmA, &full_bar[stage], sa, ccoord, bx[0] + dx, by[0] + dy, bb[0]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      if (j < p.NB)
                        tma_load_4d(tmA, &full_bar[stage], sa + j * box_rows * Cfg::SWA, ccoord, bx[j] + dx, by[j] + dy, bb[j]);
                    }
                  }
                  if (p.csize > 1) {
                    constexpr int HALF_ROWS = BLOCK_N / 2;
                    tma_load_2d_mc(&p.tmB, &full_bar[stage], sb + crank * HALF_ROWS * Cfg::SWA, kbase + cc * CK,
                                   n0 + crank * HALF_ROWS, cmask);
                  } else {
                    tma_load_2d(&p.tmB, &full_bar[stage], sb, kbase + cc * CK, n0);
                  }
                }
              }
              __syncwarp();
              if (++stage == Cfg::NSTAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 && (!PAIR || crank == 0)) {
    // =============================== MMA issuer (pair form: the leader CTA only) ===============================
    {  // warp-uniform loop, one elected lane issues (see the producer)
      constexpr uint32_t idesc = make_idesc_bf16_f32(PAIR ? 256 : 128, BLOCK_N);
      int stage = 0, phase = 0, it = 0;
      for (int item = item0; item < total_items; item += item_step, ++it) {
        const int acc = it & 1;
        const int acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kc = 0; kc < num_k_chunks; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          if (it == 0 && kc == 0 && lane == 0) CV_DBG(2);  // first operands landed
          const uint32_t a_addr = smem_u32(stages + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < CK / 16 && !(p.experiment & 1); ++k) {
              const uint64_t adesc = make_smem_desc_kmajor(a_addr + k * 32, Cfg::SWA);
              const uint64_t bdesc = make_smem_desc_kmajor(b_addr + k * 32, Cfg::SWA);
              if constexpr (PAIR) umma_f16_ss_pair(tmem_d, adesc, bdesc, idesc, (kc | k) != 0 ? 1u : 0u);
              else umma_f16_ss(tmem_d, adesc, bdesc, idesc, (kc | k) != 0 ? 1u : 0u);
            }
            if constexpr (PAIR) {
              umma_commit_pair(&empty_bar[stage], 3);                              // the stage is free in both CTAs
              if (kc == num_k_chunks - 1) umma_commit_pair(&tmem_full[acc], 3);    // both CTAs' accumulators are ready
            } else {
            if (p.csize > 1) umma_commit_mc(&empty_bar[stage], cmask);  // release the stage in every CTA of the cluster
            else umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
            if (kc == num_k_chunks - 1) umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
            }
          }
          __syncwarp();
          if (++stage == Cfg::NSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int eall = threadIdx.x - 128;  // 0 .. EPI_THREADS-1
    const int et = eall & 127;           // accumulator row == TMEM lane
    const int egrp = eall >> 7;          // which column range of the tile this warp group drains
    const int ewarp = warp & 3;          // TMEM lane quadrant a warp may read == warp id % 4
    const int box_rows = p.BH * p.BW;
    int it = 0;
    uint32_t res_phase = 0;
    const bool leader = ewarp == 0;  // first warp of the group: lanes < NB own one box each (residual load, store, bulk group)
    int3 epix;                       // this thread's accumulator row inside the tile: (box, row in box, column in box)
    epix.x = et / box_rows;
    epix.y = (et - epix.x * box_rows) / p.BW;
    epix.z = et - epix.x * box_rows - epix.y * p.BW;
    uint8_t* slab = staging + egrp * Cfg::SLAB_BYTES;
    const int gbar = 2 + egrp;       // named barrier of this group
    // one N tile: the bias slice never changes -> staged once
    const bool bias_once = p.num_n_tiles == 1;
    if (bias_once) {
      for (int i = eall; i < BLOCK_N; i += Cfg::EPI_THREADS) bias_s[i] = p.bias[i];
      named_bar_sync(1, Cfg::EPI_THREADS);
    }
    for (int item = item0; item < total_items; item += item_step, ++it) {
      const int acc = it & 1;
      const int acc_phase = (it >> 1) & 1;
      const int mg = item / p.num_n_tiles;
      const int m = mg * p.csize + crank;
      const int n0 = (item - mg * p.num_n_tiles) * BLOCK_N;
      const int nc0 = n0 + egrp * Cfg::OC;  // first output channel of this group's slab

      int cb = 0, cy = 0, cx = 0;
      if (leader) {
        {  // box j of this tile: lane j does the index arithmetic
          const int q = m * p.NB + (lane & 7);
          cb = q / p.boxes_per_img;
          const int r = q - cb * p.boxes_per_img;
          const int py = r / p.boxes_x;
          cy = py * p.BH;
          cx = (r - py * p.boxes_x) * p.BW;
        }
        if (p.hc.keys && egrp == 0 && lane < 8) {  // for the candidate epilogue below (read after an all-group barrier)
          cand_box[lane * 3] = cb;
          cand_box[lane * 3 + 1] = cy;
          cand_box[lane * 3 + 2] = cx;
        }
        tma_store_wait_read<0>();  // the previous tile's store of this slab has finished reading it
        if (p.has_res) {
          if (lane == 0) mbar_expect_tx(&res_full[egrp], Cfg::SLAB_BYTES * (X3 ? 2 : 1));
          __syncwarp();
          if (lane < p.NB) {
            tma_load_4d(&p.tmRes, &res_full[egrp], slab + lane * box_rows * Cfg::SWO, nc0, cx, cy, cb);
            if constexpr (X3)
              tma_load_4d(&p.tmResLo, &res_full[egrp], slab + 128 * BLOCK_N * 2 + lane * box_rows * Cfg::SWO, nc0, cx, cy, cb);
          }
        }
        if (!bias_once)
          for (int i = lane; i < Cfg::OC; i += 32) bias_s[egrp * Cfg::OC + i] = p.bias[nc0 + i];
      }
      named_bar_sync(gbar, 128);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      if (eall == 0) CV_DBG(it == 0 ? 3 : 5);  // first / last accumulator complete
      if (p.has_res) mbar_wait(&res_full[egrp], res_phase);
      drain_dispatch<Cfg>(p, tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + acc * BLOCK_N + egrp * Cfg::OC,
                          smem_u32(slab), smem_u32(bias_s) + egrp * Cfg::OC * 4, et);
      if (eall == 0 && it == 0) CV_DBG(12);  // first tile drained by this thread
      tcgen05_fence_before();
      if constexpr (!PAIR) mbar_arrive(&tmem_empty[acc]);  // accumulator drained -> back to the MMA warp
      fence_proxy_async_smem();       // staging complete -> visible to the TMA store
      named_bar_sync(gbar, 128);
      if constexpr (PAIR) {  // the group has drained its columns: ONE arrival on the pair leader's barrier
        if (et == 0) mbar_arrive_pair_leader(&tmem_empty[acc]);
      }
      if (leader) {
        if (lane < p.NB) {
          tma_store_4d(&p.tmOut, slab + lane * box_rows * Cfg::SWO, nc0, cx, cy, cb);
          if constexpr (X3) {  // planes [hi | lo | hi] of the output segment
            tma_store_4d(&p.tmOutX[0], slab + 128 * BLOCK_N * 2 + lane * box_rows * Cfg::SWO, nc0, cx, cy, cb);
            tma_store_4d(&p.tmOutX[1], slab + lane * box_rows * Cfg::SWO, nc0, cx, cy, cb);
          }
          tma_store_commit();
        }
        if (eall == 0) CV_DBG(it == 0 ? 4 : 6);  // first / last tile's store issued
      }
      res_phase ^= 1;
      if (!X3 && p.hc.keys) {
        // detect head: score NMS candidates from the whole staged tile (all slabs) while the TMA stores drain it (both
        // only read); no group may start rewriting its slab before every group has finished reading
        named_bar_sync(1, Cfg::EPI_THREADS);
        if (eall == 0) cand_cnt[(it + 1) & 1] = 0;  // the other counter was last read before the barrier above
        head_candidates<Cfg>(p, staging, epix, cand_box, et, egrp, lane, eall, cand_list, &cand_cnt[it & 1], cand_keys, cand_img);
        named_bar_sync(1, Cfg::EPI_THREADS);
      }
    }
    if (leader) {
      tma_store_wait_all<0>();
      if (eall == 0) CV_DBG(7);  // stores drained
    }
  }

  tcgen05_fence_before();
  if (p.csize > 1) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it / arrive on it
  else __syncthreads();
  if (threadIdx.x == 0) {
    CV_DBG(8);  // all roles finished
    if (p.dbg) p.dbg[(size_t)blockIdx.x * 16 + 15] = clock64();
  }
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 convolution from a HALO tile: the nine taps are nine shifted views of ONE input tile.
//
// conv_tc_kernel fetches a 128-pixel A tile per tap (9x the input through L2 -> shared memory) and a weight tile per
// 128 pixels; measured, the 3x3 layers sit at the SM's ingest limit (~45 B/clk) at half the tensor rate. Here one work
// item is a 16 x 16 output tile (two 16 x 8 halves, each one M = 128 accumulator): per 64-channel chunk the
// 18 x 18 x 64 input halo is loaded ONCE (TMA, SWIZZLE_128B: one pixel = one 128-byte row) and every weight tile is
// shared by both halves: 3.1x fewer bytes per FLOP.
//   A operand of tap (ky, kx), half h: the UMMA descriptor's start address is the halo row (ky * HW + kx + 8 h) --
//   an 8-row group is 8 horizontally adjacent pixels = 8 consecutive 128-byte rows, the group stride (SBO) one halo
//   line (HW rows). The 128-byte swizzle is a function of the shared-memory ADDRESS bits [7,10) for TMA writes and
//   UMMA reads alike, so a start address shifted by whole rows addresses the same bytes TMA wrote.
// Tiles at the right edge have one half (HW = 10 instead of 18). Epilogue groups own one (half, 64-column slab).
// ------------------------------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct HaloCfg {
  static constexpr bool kX3 = false;
  static constexpr int BN = BLOCK_N;
  static constexpr int CK = 64, SWA = 128;
  static constexpr int TH = 16, TW = 8;                   // one half tile
  static constexpr int HH = TH + 2;                       // halo lines
  static constexpr int X_SLOT = (HH * (2 * TW + 2) * SWA + 1023) / 1024 * 1024;  // 18 x 18 pixels x 128 B
  // input ring: a narrow-N item (the packed stem: one chunk, three taps) finishes its tile long before the next halo
  // tile's TMA round trip completes, so two slots leave the MMA warp waiting; three fit next to the small weight stages
  static constexpr int NX = BLOCK_N <= 64 ? 3 : 2;
  static constexpr int W_STAGE = BLOCK_N * SWA;
  // 32-channel slabs up to N = 64: (half, slab) = 4 epilogue groups = 16 warps. With one 64-column slab per half the 8
  // epilogue warps (2 per scheduler, ~600 dependent instructions per tile each) were the critical path of the packed stem
  // (ncu: tensor 26 %, DRAM 44 %, XU 40 % -- nothing saturated, 19 % of the warp slots active).
  static constexpr int OC = BLOCK_N <= 64 ? 32 : 64;
  static constexpr int SWO = OC * 2;
  static constexpr int SLAB_BYTES = 128 * SWO;
  static constexpr int NSLAB = BLOCK_N / OC;
  static constexpr int EPI_GROUPS = 2 * NSLAB;            // (half, slab)
  static constexpr int EPI_THREADS = 128 * EPI_GROUPS;
  static constexpr int THREADS = 128 + EPI_THREADS;
  static constexpr int STAGING_BYTES = EPI_GROUPS * SLAB_BYTES;
  static constexpr int TAIL_BYTES = EPI_GROUPS * OC * 4 + 320;  // per-group bias slice + barriers + tmem ptr
  static constexpr int NW_RAW = (kSmemPerSm - 2048 - NX * X_SLOT - STAGING_BYTES - TAIL_BYTES) / W_STAGE;
  static constexpr int NW = NW_RAW > 8 ? 8 : NW_RAW;
  static constexpr int SMEM_BYTES = 1024 + NX * X_SLOT + NW * W_STAGE + STAGING_BYTES + TAIL_BYTES;
  static constexpr int TMEM_COLS = 4 * BLOCK_N;           // two halves, double-buffered: 128 / 256 / 512
  static_assert(NW >= 3, "not enough shared memory for the weight ring");
};

template <int BLOCK_N, int KW = 3>  // KW = 3: 3x3 / pad 1; KW = 1: the 3x1 window form of the packed stem (no halo columns)
__global__ void __launch_bounds__(HaloCfg<BLOCK_N>::THREADS, 1) conv_halo_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = HaloCfg<BLOCK_N>;
  constexpr int NTAPS = 3 * KW, PADW = KW == 3 ? 1 : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xring = smem;
  uint8_t* wring = xring + Cfg::NX * Cfg::X_SLOT;
  uint8_t* staging = wring + Cfg::NW * Cfg::W_STAGE;
  float* bias_s = reinterpret_cast<float*>(staging + Cfg::STAGING_BYTES);  // [EPI_GROUPS][OC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + Cfg::EPI_GROUPS * Cfg::OC);
  uint64_t* xfull = bars;                // [NX <= 4]
  uint64_t* xempty = bars + 4;           // [NX <= 4]
  uint64_t* wfull = bars + 8;            // [8]
  uint64_t* wempty = bars + 16;          // [8]
  uint64_t* tmem_full = bars + 24;       // [2]
  uint64_t* tmem_empty = bars + 26;      // [2]
  uint64_t* res_full = bars + 28;        // [EPI_GROUPS <= 4]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    CV_DBG(0);
    if (p.dbg) p.dbg[(size_t)blockIdx.x * 16 + 9] = clock64();
    for (int i = 0; i < Cfg::NX; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], Cfg::EPI_THREADS);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmA[1]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 2) tmem_alloc(tmem_ptr_s, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  uint32_t tmem_base = 0;
  if (warp == 0) {  // the producer only arrives (see conv_tc_kernel)
    asm volatile("bar.arrive 0, %0;" ::"r"(Cfg::THREADS) : "memory");
  } else {
    asm volatile("bar.sync 0, %0;" ::"r"(Cfg::THREADS) : "memory");
    tcgen05_fence_after();
    tmem_base = *tmem_ptr_s;
  }
  if (threadIdx.x == 32) CV_DBG(1);
  // programmatic dependent launch (opt-in, see ay2_conv_plan_run): no global memory access before the previous kernel is done
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int tiles_per_img = p.hl_pairs_x * p.hl_bands_y;
  const int total_items = tiles_per_img * p.num_m_tiles * p.num_n_tiles;  // num_m_tiles == batch here
  const int cin_chunks = p.cin_chunks;
  // item -> (N tile fastest: the second N tile finds the halo in L2, tile, image)
  // Longest work first: all two-half tiles, then the one-half tiles of the right edge (odd number of halves per row), so
  // that the short items fill the tail of the static round-robin schedule.
  const int full_pairs = p.hl_w_halves >> 1;
  const int tiles2 = p.num_m_tiles * p.hl_bands_y * full_pairs;
  auto decode = [&](int item, int& n0, int& b, int& y0, int& x0, int& nh) {
    int t = item / p.num_n_tiles;
    n0 = (item - t * p.num_n_tiles) * BLOCK_N;
    int band, pair;
    if (t < tiles2) {
      const int per_img = p.hl_bands_y * full_pairs;
      b = t / per_img;
      const int r = t - b * per_img;
      band = r / full_pairs;
      pair = r - band * full_pairs;
      nh = 2;
    } else {
      t -= tiles2;
      b = t / p.hl_bands_y;
      band = t - b * p.hl_bands_y;
      pair = full_pairs;
      nh = 1;
    }
    y0 = band * Cfg::TH;
    x0 = pair * 2 * Cfg::TW;
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    int xs = 0, xph = 0, ws = 0, wph = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      int n0, b, y0, x0, nh;
      decode(item, n0, b, y0, x0, nh);
      const CUtensorMap* tmX = nh == 2 ? &p.tmA[0] : &p.tmA[1];
      const uint32_t xbytes = Cfg::HH * (nh * Cfg::TW + KW - 1) * Cfg::SWA;  // KW = 3: one halo column each side; KW = 1: none
      for (int c = 0; c < cin_chunks; ++c) {
        mbar_wait(&xempty[xs], xph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&xfull[xs], xbytes);
          tma_load_4d(tmX, &xfull[xs], xring + xs * Cfg::X_SLOT, c * Cfg::CK, x0 - PADW, y0 - 1, b);
        }
        __syncwarp();
        if (++xs == Cfg::NX) { xs = 0; xph ^= 1; }
        for (int tap = 0; tap < NTAPS; ++tap) {
          mbar_wait(&wempty[ws], wph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&wfull[ws], Cfg::W_STAGE);
            tma_load_2d(&p.tmB, &wfull[ws], wring + ws * Cfg::W_STAGE, tap * p.cin + c * Cfg::CK, n0);
          }
          __syncwarp();
          if (++ws == Cfg::NW) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = make_idesc_bf16_f32(128, BLOCK_N);
    // descriptor halves: lo = start address >> 4 | LBO (unused for swizzled K-major: 1) << 16;
    //                    hi = SBO >> 4 | version 1 << 14 | base offset << 17 | SWIZZLE_128B (2) << 29
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t x_lo0 = ((smem_u32(xring) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t w_lo0 = ((smem_u32(wring) & 0x3FFFF) >> 4) | (1u << 16);
    int xs = 0, xph = 0, ws = 0, wph = 0, it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int n0, b, y0, x0, nh;
      decode(item, n0, b, y0, x0, nh);
      const int acc = it & 1;
      mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t d0 = tmem_base + acc * 2 * BLOCK_N;
      const int hw = nh * Cfg::TW + KW - 1;                  // halo line in pixels == rows
      const uint32_t a_hi0 = (static_cast<uint32_t>(hw * Cfg::SWA) >> 4) | (1u << 14) | (2u << 29);
      for (int c = 0; c < cin_chunks; ++c) {
        mbar_wait(&xfull[xs], xph);
        const uint32_t a_slot = x_lo0 + xs * (Cfg::X_SLOT >> 4);
        for (int tap = 0; tap < NTAPS; ++tap) {
          const int ky = tap / KW, kx = tap - ky * KW;
          mbar_wait(&wfull[ws], wph);
          tcgen05_fence_after();
          if (c == 0 && tap == 0 && it == 0 && lane == 0) CV_DBG(2);
          const uint32_t a_tap = a_slot + (ky * hw + kx) * (Cfg::SWA >> 4);  // whole rows: 8 address units each
          const uint32_t b_lo = w_lo0 + ws * (Cfg::W_STAGE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < Cfg::CK / 16; ++k) {
              const uint32_t accum = (c | tap | k) != 0 ? 1u : 0u;
              for (int h = 0; h < nh; ++h) {
                const uint32_t a_lo = a_tap + h * (Cfg::TW * Cfg::SWA >> 4) + 2 * k;
                // experiment bit 1: descriptor base offset = row phase of the start address inside the 1024-byte pattern
                const uint32_t a_hi = (p.experiment & 2) ? (a_hi0 | (((a_lo >> 3) & 7u) << 17)) : a_hi0;
                umma_f16_lohi(d0 + h * BLOCK_N, a_lo, a_hi, b_lo + 2 * k, b_hi, idesc, accum);
              }
            }
            umma_commit(&wempty[ws]);
            if (tap == NTAPS - 1) umma_commit(&xempty[xs]);
            if (tap == NTAPS - 1 && c == cin_chunks - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++ws == Cfg::NW) { ws = 0; wph ^= 1; }
        }
        if (++xs == Cfg::NX) { xs = 0; xph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int eall = threadIdx.x - 128;
    const int et = eall & 127;          // accumulator row == TMEM lane == pixel (y = et / 8, x = et % 8) of the half
    const int egrp = eall >> 7;
    const int half = egrp / Cfg::NSLAB, slabi = egrp - half * Cfg::NSLAB;
    const int ewarp = warp & 3;
    const bool leader = ewarp == 0;
    uint8_t* slab = staging + egrp * Cfg::SLAB_BYTES;
    float* bias_g = bias_s + egrp * Cfg::OC;
    const int gbar = 2 + egrp;
    const bool bias_once = p.num_n_tiles == 1;
    if (bias_once && et < Cfg::OC) bias_g[et] = p.bias[slabi * Cfg::OC + et];
    uint32_t res_phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int n0, b, y0, x0, nh;
      decode(item, n0, b, y0, x0, nh);
      const int acc = it & 1;
      const bool active = half < nh;
      const int nc0 = n0 + slabi * Cfg::OC;
      const int xh = x0 + half * Cfg::TW;
      if (leader && active) {
        tma_store_wait_read<0>();  // the previous store of this slab has been read out
        if (p.has_res && lane == 0) {
          mbar_expect_tx(&res_full[egrp], Cfg::SLAB_BYTES);
          tma_load_4d(&p.tmRes, &res_full[egrp], slab, nc0, xh, y0, b);
        }
      }
      if (!bias_once && et < Cfg::OC) bias_g[et] = p.bias[nc0 + et];
      named_bar_sync(gbar, 128);
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      if (eall == 0) CV_DBG(it == 0 ? 3 : 5);
      if (active) {
        if (p.has_res) {
          mbar_wait(&res_full[egrp], res_phase);
          res_phase ^= 1;
        }
        drain_dispatch<Cfg>(p, tmem_base + (static_cast<uint32_t>(ewarp * 32) << 16) + acc * 2 * BLOCK_N + half * BLOCK_N + slabi * Cfg::OC,
                            smem_u32(slab), smem_u32(bias_g), et);
      }
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      fence_proxy_async_smem();
      named_bar_sync(gbar, 128);
      if (leader && active && lane == 0) {
        tma_store_4d(&p.tmOut, slab, nc0, xh, y0, b);
        tma_store_commit();
      }
      if (eall == 0) CV_DBG(it == 0 ? 4 : 6);
    }
    if (leader) {
      tma_store_wait_all<0>();
      if (eall == 0) CV_DBG(7);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    CV_DBG(8);
    if (p.dbg) p.dbg[(size_t)blockIdx.x * 16 + 15] = clock64();
  }
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: plan = encoded tensor maps + launch configuration
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

static CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// NHWC bf16 activation view [C][W][H][B] with explicit element strides; box [boxc][bw][bh][1].
int encode_act_map(CUtensorMap* tm, const void* base, int C, int W, int H, int B, int64_t sW, int64_t sH,
                          int64_t sB, int boxc, int bw, int bh) {
  // L2 promotion: fetch 256 B only when a pixel's channels are one dense run of >= 256 B that this conv consumes
  // entirely; a channel slice of a wider concat buffer would otherwise drag its neighbour's bytes through HBM.
  // (a pixel pitch that is not a multiple of 128 B -- the 3-slice C3 buffer at c_ = 32: 192 B -- puts every other row
  // across two 128-byte lines: promote 64 B there, or the neighbouring slice's bytes come along)
  const bool dense = (int64_t)C == sW && C * 2 >= 256;
  const CUtensorMapL2promotion promo = dense ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                             : (boxc * 2 >= 128 && (sW * 2) % 128 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                                       : CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return AY2_ERR_CUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(boxc * 2), promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d box=%d,%d,%d strides=%lld,%lld,%lld) -> %d", C, W,
              H, B, boxc, bw, bh, (long long)sW, (long long)sH, (long long)sB, (int)r);
    return AY2_ERR_CUDA;
  }
  return AY2_OK;
}

int encode_weight_map(CUtensorMap* tm, const void* base, int Ktot, int rows, int boxk, int boxn) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return AY2_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxk, (cuuint32_t)boxn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for_bytes(boxk * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(weights K=%d rows=%d box=%d,%d) -> %d", Ktot, rows, boxk, boxn, (int)r);
    return AY2_ERR_CUDA;
  }
  return AY2_OK;
}

}  // namespace ay2

using namespace ay2;

struct ay2_conv_plan {
  ConvKernelParams kp;
  ay2_conv_desc desc;
  int block_n, ck, pair;
  int ctas_per_sm, threads;
  int grid, halo;
  size_t smem;
  void (*kernel)(const ConvKernelParams);
};

extern "C" int ay2_conv_block_n(int32_t cout) {
  if (cout <= 32) return 32;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  return 256;
}

template <int BN, int CK, bool X3, bool PAIR = false>
static void bind_kernel(ay2_conv_plan* pl) {
  pl->kernel = conv_tc_kernel<BN, CK, X3, PAIR>;
  pl->smem = ConvCfg<BN, CK, X3, PAIR>::SMEM_BYTES;
  pl->ctas_per_sm = ConvCfg<BN, CK, X3, PAIR>::CTAS_PER_SM;
  pl->threads = ConvCfg<BN, CK, X3, PAIR>::THREADS;
}

static int pick_box(int H, int W, int* bh, int* bw) {
  // candidate boxes (rows multiple of 8 keeps every box on a swizzle-pattern boundary; <= 8 boxes per tile)
  static const int cand[][2] = {{8, 16}, {16, 8}, {4, 32}, {8, 8},  {4, 16}, {16, 4},
                                {2, 32}, {4, 8},  {8, 4},  {2, 16}, {4, 4},  {2, 8}, {1, 16}};
  double best = 1e30;
  int bi = -1;
  for (int i = 0; i < (int)(sizeof(cand) / sizeof(cand[0])); ++i) {
    const int h = cand[i][0], w = cand[i][1];
    const double cover = (double)ceil_div(H, h) * h * (double)ceil_div(W, w) * w;
    // prefer less padding; break ties towards bigger boxes (fewer TMA ops)
    const double cost = cover * (1.0 + 0.002 * (128 / (h * w)));
    if (cost < best) {
      best = cost;
      bi = i;
    }
  }
  *bh = cand[bi][0];
  *bw = cand[bi][1];
  return 0;
}

extern "C" int ay2_conv_plan_create(const ay2_conv_desc* d, const void* in, const void* weight, const float* bias,
                                    const void* residual, void* out, ay2_conv_plan** plan_out) {
  AY2_REQUIRE(d && in && weight && bias && out && plan_out, "ay2_conv_plan_create: null argument");
  AY2_REQUIRE(d->stride == 1 || d->stride == 2, "conv stride %d unsupported (1 or 2)", d->stride);
  AY2_REQUIRE(d->cin % 16 == 0 && d->cin >= 16, "conv cin=%d must be a multiple of 16", d->cin);
  // TMA clips the innermost dimension in 16-byte units, so a slice must own whole groups of 8 channels
  AY2_REQUIRE(d->cout % 8 == 0 && d->cout >= 8, "conv cout=%d must be a multiple of 8 (pad with zero filters)", d->cout);
  AY2_REQUIRE(d->in_cstride % 8 == 0 && d->out_cstride % 8 == 0, "channel strides must be multiples of 8");
  AY2_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh <= 7 && d->kw <= 7, "kernel %dx%d unsupported", d->kh, d->kw);
  AY2_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(weight) & 15) == 0,
              "conv buffers must be 16-byte aligned");
  const int pad_w = d->pad_w < 0 ? d->pad : d->pad_w;
  const int64_t pix_stride = d->in_pix_stride > 0 ? d->in_pix_stride : d->in_cstride;
  const int64_t row_pixels = d->in_row_pixels > 0 ? d->in_row_pixels : d->in_w;
  AY2_REQUIRE(pix_stride % 8 == 0, "in_pix_stride must be a multiple of 8 elements");
  AY2_REQUIRE(d->stride == 1 || (d->in_pix_stride <= 0 && d->in_row_pixels <= 0), "custom input strides need stride 1");
  const int stride_w = d->stride_w > 0 ? d->stride_w : d->stride;
  AY2_REQUIRE(stride_w == d->stride || (d->stride == 2 && stride_w == 1 && d->kw == 2 && pad_w == 1 && !d->x3 && d->cin_split == 0),
              "stride_w != stride is built for the pixel-pair form only (stride 2, stride_w 1, kw 2, pad_w 1)");
  const int exp_oh = (d->in_h + 2 * d->pad - d->kh) / d->stride + 1;
  const int exp_ow = stride_w != d->stride ? d->in_w : (d->in_w + 2 * pad_w - d->kw) / d->stride + 1;  // pair form: left pad only
  // (sub-grid outputs are the dgrad of a strided conv: asymmetric implicit padding, the caller fixes the size)
  AY2_REQUIRE(d->out_pix_stride > 0 || (exp_oh == d->out_h && exp_ow == d->out_w),
              "conv output size %dx%d does not match %dx%d", d->out_h, d->out_w, exp_oh, exp_ow);
  if (d->stride == 2) AY2_REQUIRE(d->in_h % 2 == 0 && (stride_w == 1 || d->in_w % 2 == 0), "stride-2 conv needs even input size");
  const int bn = ay2_conv_block_n(d->cout);
  AY2_REQUIRE(d->cout_pad >= d->cout && d->cout_pad % bn == 0, "cout_pad=%d must be a multiple of %d", d->cout_pad, bn);
  AY2_REQUIRE(d->res_cstride == 0 || residual, "residual stride given without a residual pointer");
  AY2_REQUIRE(d->res_cstride % 8 == 0, "residual channel stride must be a multiple of 8");
  AY2_REQUIRE(!d->x3 || (d->out_pix_stride <= 0 && d->out_cstride >= 3 * d->cout && (d->res_cstride == 0 || d->res_cstride >= 3 * d->cout)),
              "split-precision output needs three planes of cout channels inside the channel stride");

  ay2_conv_plan* pl = new ay2_conv_plan();
  memset(pl, 0, sizeof(*pl));
  pl->desc = *d;
  const int split = d->cin_split;
  AY2_REQUIRE(split == 0 || (d->stride == 1 && d->in2 && split > 0 && split < d->cin && split % 16 == 0 && d->in2_cstride % 8 == 0 &&
                             d->in_pix_stride <= 0 && d->in_row_pixels <= 0),
              "two-source input needs stride 1, a second pointer, and cin_split a multiple of 16 inside (0, cin)");
  int ck = (d->cin % 64 == 0 && split % 64 == 0) ? 64 : ((d->cin % 32 == 0 && split % 32 == 0) ? 32 : 16);
  if (d->x3 && bn == 256 && ck == 64) ck = 32;  // the doubled (hi + lo) staging of a 256-column tile leaves room for 32-channel stages only
  pl->block_n = bn;
  pl->ck = ck;
  ConvKernelParams& kp = pl->kp;
  kp.experiment = getenv("AY2_CONV_EXPERIMENT") ? atoi(getenv("AY2_CONV_EXPERIMENT")) : 0;
  {
    // 3x3 / s1 / p1 over whole 64-channel chunks: the halo kernel (unless the 16 x 8 half tiles would waste > 30 % of
    // the MMA rows on this feature-map size, e.g. 20 x 20)
    static const int env_halo = getenv("AY2_CONV_HALO") ? atoi(getenv("AY2_CONV_HALO")) : 1;
    const int bands = ceil_div(d->out_h, 16), halves = ceil_div(d->out_w, 8);
    // (3x3 with pad 1, or the 3x1 window form of the packed stem: kw = 1, pad_w = 0, custom input pixel / row strides)
    const bool shape_ok = !d->x3 && d->kh == 3 && ((d->kw == 3 && pad_w == 1 && d->in_pix_stride <= 0 && d->in_row_pixels <= 0) ||
                                                   (d->kw == 1 && pad_w == 0)) &&
                          d->stride == 1 && d->pad == 1 && d->cin % 64 == 0 && split == 0 && d->out_pix_stride <= 0 &&
                          d->out_row_pixels <= 0;
    const bool fill_ok = (double)bands * 16 * halves * 8 <= 1.3 * d->out_h * d->out_w;
    if (env_halo && shape_ok && (fill_ok || env_halo == 2)) {
      const int nt = bn < 128 ? bn : 128;
      pl->block_n = nt;
      pl->ck = 64;
      kp.num_m_tiles = d->batch;
      kp.num_n_tiles = d->cout_pad / nt;
      kp.hl_bands_y = bands;
      kp.hl_w_halves = halves;
      kp.hl_pairs_x = ceil_div(halves, 2);
      kp.cin = d->cin;
      kp.cin_chunks = d->cin / 64;
      kp.cout_pad = d->cout_pad;
      kp.act = d->act;
      kp.has_res = d->res_cstride != 0;
      kp.bias = bias;
      kp.csize = 1;
      kp.kw = d->kw;
      kp.pad_w = pad_w;
      const int64_t os = d->out_cstride, rs = d->res_cstride;
      const int W = d->in_w, H = d->in_h;
      const int oc = nt <= 64 ? 32 : 64;  // HaloCfg::OC
      const int xw = d->kw - 1;  // halo columns
      int rc = encode_act_map(&kp.tmA[0], in, d->cin, W, H, d->batch, pix_stride, pix_stride * row_pixels,
                              pix_stride * row_pixels * H, 64, 16 + xw, 18);
      if (rc == AY2_OK)
        rc = encode_act_map(&kp.tmA[1], in, d->cin, W, H, d->batch, pix_stride, pix_stride * row_pixels,
                            pix_stride * row_pixels * H, 64, 8 + xw, 18);
      if (rc == AY2_OK) rc = encode_weight_map(&kp.tmB, weight, 3 * d->kw * d->cin, d->cout_pad, 64, nt);
      if (rc == AY2_OK) rc = encode_act_map(&kp.tmOut, out, d->cout, W, H, d->batch, os, os * W, os * W * H, oc, 8, 16);
      if (rc == AY2_OK && kp.has_res) rc = encode_act_map(&kp.tmRes, residual, d->cout, W, H, d->batch, rs, rs * W, rs * W * H, oc, 8, 16);
      if (rc != AY2_OK) {
        delete pl;
        return rc;
      }
      if (nt == 32) pl->smem = HaloCfg<32>::SMEM_BYTES, pl->threads = HaloCfg<32>::THREADS;
      else if (nt == 64) pl->smem = HaloCfg<64>::SMEM_BYTES, pl->threads = HaloCfg<64>::THREADS;
      else pl->smem = HaloCfg<128>::SMEM_BYTES, pl->threads = HaloCfg<128>::THREADS;
      if (d->kw == 3) pl->kernel = nt == 32 ? conv_halo_kernel<32, 3> : (nt == 64 ? conv_halo_kernel<64, 3> : conv_halo_kernel<128, 3>);
      else pl->kernel = nt == 32 ? conv_halo_kernel<32, 1> : (nt == 64 ? conv_halo_kernel<64, 1> : conv_halo_kernel<128, 1>);
      pl->ctas_per_sm = 1;
      pl->halo = 1;
      cudaError_t e = cudaFuncSetAttribute(pl->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem);
      if (e != cudaSuccess) {
        delete pl;
        set_error("cudaFuncSetAttribute(halo smem=%zu) failed: %s", pl->smem, cudaGetErrorString(e));
        return AY2_ERR_CUDA;
      }
      cudaFuncSetAttribute(pl->kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      int dev = 0, sms = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int items = kp.hl_pairs_x * kp.hl_bands_y * d->batch * kp.num_n_tiles;
      pl->grid = items < sms ? items : sms;
      *plan_out = pl;
      return AY2_OK;
    }
  }
  int bh, bw;
  pick_box(d->out_h, d->out_w, &bh, &bw);
  kp.BH = bh;
  kp.BW = bw;
  kp.NB = 128 / (bh * bw);
  kp.boxes_x = ceil_div(d->out_w, bw);
  const int boxes_y = ceil_div(d->out_h, bh);
  kp.boxes_per_img = kp.boxes_x * boxes_y;
  const long long total_boxes = (long long)kp.boxes_per_img * d->batch;
  kp.num_m_tiles = (int)((total_boxes + kp.NB - 1) / kp.NB);
  kp.num_n_tiles = d->cout_pad / bn;
  kp.kh = d->kh;
  kp.kw = d->kw;
  kp.stride = d->stride;
  kp.pad = d->pad;
  kp.pad_w = pad_w;
  kp.stride_w = stride_w;
  kp.cin = d->cin;
  kp.cin_chunks = d->cin / ck;
  kp.split_chunks = split / ck;
  kp.cout_pad = d->cout_pad;
  kp.act = d->act;
  kp.has_res = d->res_cstride != 0;
  kp.bias = bias;

  int rc = AY2_OK;
  const int64_t cs = d->in_cstride;
  if (d->stride == 1) {
    // pix_stride < cin gives overlapping windows of neighbouring pixels (the packed 16-channel stem)
    rc = encode_act_map(&kp.tmA[0], in, split ? split : d->cin, d->in_w, d->in_h, d->batch, pix_stride, pix_stride * row_pixels,
                        pix_stride * row_pixels * d->in_h, ck, bw, bh);
    if (rc == AY2_OK && split)
      rc = encode_act_map(&kp.tmA[3], d->in2, d->cin - split, d->in_w, d->in_h, d->batch, d->in2_cstride,
                          (int64_t)d->in2_cstride * d->in_w, (int64_t)d->in2_cstride * d->in_w * d->in_h, ck, bw, bh);
  } else if (stride_w == 1) {
    for (int ph = 0; ph < 2 && rc == AY2_OK; ++ph) {  // row-parity views: every column, every other row
      const uint8_t* base = static_cast<const uint8_t*>(in) + (int64_t)ph * d->in_w * cs * 2;
      rc = encode_act_map(&kp.tmA[ph * 2], base, d->cin, d->in_w, d->in_h / 2, d->batch, cs, 2 * cs * d->in_w,
                          cs * d->in_w * d->in_h, ck, bw, bh);
    }
  } else {
    for (int ph = 0; ph < 2 && rc == AY2_OK; ++ph)
      for (int pw = 0; pw < 2 && rc == AY2_OK; ++pw) {
        const uint8_t* base = static_cast<const uint8_t*>(in) + ((int64_t)ph * d->in_w + pw) * cs * 2;
        rc = encode_act_map(&kp.tmA[ph * 2 + pw], base, d->cin, d->in_w / 2, d->in_h / 2, d->batch, 2 * cs,
                            2 * cs * d->in_w, cs * d->in_w * d->in_h, ck, bw, bh);
      }
  }
  // Weight-tile multicast across a 2-CTA cluster is implemented and parity-tested (AY2_CONV_CLUSTER=2), but measured
  // 5 % SLOWER on every yolov5s layer (r01: 2.81 vs 2.65 ms per step): at cluster sizes <= 4 the L2 already dedups
  // neighbouring unicast requests, so multicast saves no LTS bandwidth and only adds lock-step. Off by default.
  static const int env_cluster = getenv("AY2_CONV_CLUSTER") ? atoi(getenv("AY2_CONV_CLUSTER")) : 1;
  kp.csize = (env_cluster == 2 && kp.num_m_tiles >= 2) ? 2 : 1;
  // CTA-pair form (tcgen05 cta_group::2): for the wide tiles whose operand ingest is dominated by the weight tile -- N tile
  // of 128 / 256 over whole 64-channel chunks -- when there are enough M tiles to keep every SM pair busy.
  // AY2_CONV_PAIR: 0 = never, 1 (default) = N >= 128 and K >= 256, 2 = every eligible layer.
  static const int env_pair = getenv("AY2_CONV_PAIR") ? atoi(getenv("AY2_CONV_PAIR")) : 1;
  int dev0 = 0, sms0 = 148;
  cudaGetDevice(&dev0);
  cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
  const int ktot = d->kh * d->kw * d->cin;
  // Measured (r02, bs 64): 3x3 / s2 layers -6 .. -8 us each, 1x1 512 -> 256 -4 us; the HBM-bound 1x1 256 -> 128 layers got
  // SLOWER (61.6 -> 80.8 us @80x80: the pair form runs one CTA per SM where two co-resident CTAs hid the latency), so
  // N = 128 tiles pair up only under a spatial kernel.
  const bool pair = env_pair > 0 && !d->x3 && (bn == 128 || bn == 256) && ck == 64 && kp.num_m_tiles >= sms0 &&
                    (env_pair == 2 || (ktot >= 256 && (bn == 256 || d->kh * d->kw > 1)));
  if (pair) kp.csize = 2;
  pl->pair = pair ? 1 : 0;
  if (rc == AY2_OK) rc = encode_weight_map(&kp.tmB, weight, d->kh * d->kw * d->cin, d->cout_pad, ck, bn / kp.csize);
  const int oc = bn < 64 ? bn : 64;
  // output view: pixel stride / row pitch / image pitch (a parity sub-grid doubles the first and keeps the others)
  const bool sub = d->out_pix_stride > 0;
  const int64_t ops_ = sub ? d->out_pix_stride : d->out_cstride;
  const int64_t orow = d->out_row_pixels > 0 ? (int64_t)d->out_row_pixels * ops_ : ops_ * d->out_w;
  const int64_t oimg = sub ? orow * d->out_h : orow * d->out_h;
  if (rc == AY2_OK) rc = encode_act_map(&kp.tmOut, out, d->cout, d->out_w, d->out_h, d->batch, ops_, orow, oimg, oc, bw, bh);
  for (int pl_i = 0; pl_i < 2 && rc == AY2_OK && d->x3; ++pl_i)  // planes 1 (lo) and 2 (hi again) of the output segment
    rc = encode_act_map(&kp.tmOutX[pl_i], static_cast<uint8_t*>(out) + (size_t)(pl_i + 1) * d->cout * 2, d->cout, d->out_w,
                        d->out_h, d->batch, ops_, orow, oimg, oc, bw, bh);
  if (rc == AY2_OK && kp.has_res) {
    const int64_t rscale = sub ? d->out_pix_stride / d->out_cstride : 1;  // same sub-grid geometry for the residual
    const int64_t rps = (int64_t)d->res_cstride * rscale;
    const int64_t rrow = d->out_row_pixels > 0 ? (int64_t)d->out_row_pixels * rps : rps * d->out_w;
    rc = encode_act_map(&kp.tmRes, residual, d->cout, d->out_w, d->out_h, d->batch, rps, rrow, rrow * d->out_h, oc, bw, bh);
    if (rc == AY2_OK && d->x3)
      rc = encode_act_map(&kp.tmResLo, static_cast<const uint8_t*>(residual) + (size_t)d->cout * 2, d->cout, d->out_w, d->out_h,
                          d->batch, rps, rrow, rrow * d->out_h, oc, bw, bh);
  }
  if (rc != AY2_OK) {
    delete pl;
    return rc;
  }

#define AY2_BIND(BN, CKV)                                         \
  if (bn == BN && ck == CKV) {                                    \
    if constexpr (BN * CKV < 256 * 64) {                          \
      if (d->x3) bind_kernel<BN, CKV, true>(pl);                  \
    }                                                             \
    if constexpr (BN >= 128 && CKV == 64) {                       \
      if (pair) bind_kernel<BN, CKV, false, true>(pl);            \
    }                                                             \
    if (!d->x3 && !pl->kernel) bind_kernel<BN, CKV, false>(pl);   \
  }
  AY2_BIND(32, 16) AY2_BIND(32, 32) AY2_BIND(32, 64)
  AY2_BIND(64, 16) AY2_BIND(64, 32) AY2_BIND(64, 64)
  AY2_BIND(128, 16) AY2_BIND(128, 32) AY2_BIND(128, 64)
  AY2_BIND(256, 16) AY2_BIND(256, 32) AY2_BIND(256, 64)
#undef AY2_BIND
  if (!pl->kernel) {
    delete pl;
    set_error("no conv kernel for block_n=%d ck=%d", bn, ck);
    return AY2_ERR_INVALID;
  }
  cudaError_t e = cudaFuncSetAttribute(pl->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem);
  if (e != cudaSuccess) {
    delete pl;
    set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", pl->smem, cudaGetErrorString(e));
    return AY2_ERR_CUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(pl->kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int items = ((kp.num_m_tiles + kp.csize - 1) / kp.csize) * kp.num_n_tiles;
  const int resident = sms * pl->ctas_per_sm / kp.csize;  // clusters that fit at once
  pl->grid = (items < resident ? items : resident) * kp.csize;
  *plan_out = pl;
  return AY2_OK;
}

extern "C" int ay2_conv_plan_set_head_candidates(ay2_conv_plan* pl, const ay2_nms_params* p, int32_t na, int32_t row_off,
                                                 const uint8_t* class_mask, void* nms_workspace, size_t workspace_bytes) {
  AY2_REQUIRE(pl, "ay2_conv_plan_set_head_candidates: null plan");
  if (!p) {  // switch the fused candidate generation off again
    memset(&pl->kp.hc, 0, sizeof(pl->kp.hc));
    return AY2_OK;
  }
  const ay2_conv_desc& d = pl->desc;
  AY2_REQUIRE(nms_workspace && workspace_bytes >= ay2_nms_workspace_bytes(p), "NMS workspace missing or too small");
  AY2_REQUIRE(!pl->halo, "head candidates are scored by the 1x1 detect convolutions, not by a 3x3 halo plan");
  AY2_REQUIRE(pl->kp.num_n_tiles == 1, "head candidates need all %d output channels in one N tile (block_n=%d)", d.cout,
              pl->block_n);
  AY2_REQUIRE(na >= 1 && p->no > 5 && na * p->no <= d.cout, "head layout na=%d no=%d does not fit cout=%d", na, p->no, d.cout);
  AY2_REQUIRE(p->no - 5 <= 120, "the fused candidate epilogue holds at most 120 classes per anchor in registers (nc=%d)", p->no - 5);
  AY2_REQUIRE(p->batch == d.batch, "NMS batch %d != conv batch %d", p->batch, d.batch);
  AY2_REQUIRE(d.out_pix_stride <= 0, "head candidates are not defined for sub-grid outputs");
  AY2_REQUIRE(row_off >= 0 && row_off + na * d.out_h * d.out_w <= p->n, "level rows [%d, %d) exceed params.n = %d", row_off,
              row_off + na * d.out_h * d.out_w, p->n);
  const NmsWorkspaceView v = nms_workspace_view(p, nms_workspace);
  HeadCandParams& h = pl->kp.hc;
  h.keys = v.keys;
  h.counts = v.counts;
  h.class_mask = class_mask;
  h.key_stride = v.key_stride;
  h.conf_thres = p->conf_thres;
  h.max_candidates = p->max_candidates;
  h.na = na;
  h.no = p->no;
  h.row_off = row_off;
  h.multi_label = p->multi_label;
  h.dense_slots = nms_dense_slots(p) ? 1 : 0;
  h.batch = d.batch;
  h.out_h = d.out_h;
  h.out_w = d.out_w;
  return AY2_OK;
}

extern "C" int ay2_conv_plan_run(const ay2_conv_plan* pl, void* stream) {
  AY2_REQUIRE(pl, "ay2_conv_plan_run: null plan");
  static const bool use_pdl = getenv("AY2_CONV_PDL") && atoi(getenv("AY2_CONV_PDL")) == 1;  // measured: no gain inside a CUDA graph (2.72 vs 2.70 ms/step), off by default
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl->grid);
  cfg.blockDim = dim3(pl->threads);
  cfg.dynamicSmemBytes = pl->smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (pl->kp.csize > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = pl->kp.csize;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (use_pdl) {  // the kernel's prologue may overlap the previous kernel's tail (griddepcontrol.wait guards the data)
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  AY2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pl->kernel, pl->kp));
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_conv_plan_set_debug(ay2_conv_plan* pl, unsigned long long* dbg, int32_t* info4) {
  AY2_REQUIRE(pl, "ay2_conv_plan_set_debug: null plan");
  pl->kp.dbg = dbg;  // device buffer of grid x 16 uint64, or NULL
  if (info4)
    info4[0] = pl->grid, info4[1] = pl->ctas_per_sm, info4[2] = pl->halo ? -pl->block_n : pl->block_n,
    info4[3] = (pl->halo ? pl->kp.hl_pairs_x * pl->kp.hl_bands_y : 1) * pl->kp.num_m_tiles * pl->kp.num_n_tiles;
  return AY2_OK;
}

extern "C" int ay2_conv_plan_info(const ay2_conv_plan* pl, int32_t* out8) {
  AY2_REQUIRE(pl && out8, "ay2_conv_plan_info: null argument");
  out8[0] = pl->grid, out8[1] = pl->ctas_per_sm, out8[2] = pl->block_n, out8[3] = pl->ck, out8[4] = pl->halo, out8[5] = pl->pair;
  out8[6] = pl->kp.csize, out8[7] = (int32_t)pl->smem;
  return AY2_OK;
}

extern "C" int ay2_conv_plan_destroy(ay2_conv_plan* pl) {
  delete pl;
  return AY2_OK;
}

extern "C" double ay2_conv_plan_flops(const ay2_conv_plan* pl) {
  if (!pl) return 0.0;
  const ay2_conv_desc& d = pl->desc;
  return 2.0 * d.batch * d.out_h * d.out_w * (double)d.cout * d.kh * d.kw * d.cin;
}

// ------------------------------------------------------------------------------------------------
// Reference SIMT direct convolution (test infrastructure only)
// ------------------------------------------------------------------------------------------------
namespace ay2 {
__global__ void conv_ref_simt_kernel(ay2_conv_desc d, const __nv_bfloat16* __restrict__ in,
                                     const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                                     const __nv_bfloat16* __restrict__ res, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)d.batch * d.out_h * d.out_w * d.cout;
  const int pad_w = d.pad_w < 0 ? d.pad : d.pad_w;
  const long long pixs = d.in_pix_stride > 0 ? d.in_pix_stride : d.in_cstride;
  const long long rowp = d.in_row_pixels > 0 ? d.in_row_pixels : d.in_w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % d.cout);
    long long pix = idx / d.cout;
    const int ox = (int)(pix % d.out_w);
    const int oy = (int)((pix / d.out_w) % d.out_h);
    const int b = (int)(pix / ((long long)d.out_w * d.out_h));
    float acc = 0.f;
    for (int kh = 0; kh < d.kh; ++kh) {
      const int iy = oy * d.stride + kh - d.pad;
      if (iy < 0 || iy >= d.in_h) continue;
      for (int kw = 0; kw < d.kw; ++kw) {
        const int ix = ox * d.stride + kw - pad_w;
        if (ix < 0 || ix >= d.in_w) continue;
        const __nv_bfloat16* ip = in + (((long long)b * d.in_h + iy) * rowp + ix) * pixs;
        const __nv_bfloat16* wp = w + ((long long)n * d.kh * d.kw + kh * d.kw + kw) * d.cin;
        const int c_first = d.cin_split > 0 ? d.cin_split : d.cin;
        for (int c = 0; c < c_first; ++c) acc += __bfloat162float(ip[c]) * __bfloat162float(wp[c]);
        if (d.cin_split > 0) {
          const __nv_bfloat16* ip2 =
              static_cast<const __nv_bfloat16*>(d.in2) + (((long long)b * d.in_h + iy) * d.in_w + ix) * d.in2_cstride;
          for (int c = c_first; c < d.cin; ++c) acc += __bfloat162float(ip2[c - c_first]) * __bfloat162float(wp[c]);
        }
      }
    }
    acc += bias[n];
    if (d.act == AY2_ACT_SILU) acc = acc / (1.0f + expf(-acc));
    if (d.res_cstride) acc += __bfloat162float(res[pix * d.res_cstride + n]);
    out[pix * d.out_cstride + n] = __float2bfloat16_rn(acc);
  }
}
}  // namespace ay2

extern "C" int ay2_conv_reference_simt(const ay2_conv_desc* d, const void* in, const void* weight, const float* bias,
                                       const void* residual, void* out, void* stream) {
  AY2_REQUIRE(d && in && weight && bias && out, "ay2_conv_reference_simt: null argument");
  const long long total = (long long)d->batch * d->out_h * d->out_w * d->cout;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads < 148 * 32 ? (total + threads - 1) / threads : 148 * 32);
  conv_ref_simt_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      *d, static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(weight), bias,
      static_cast<const __nv_bfloat16*>(residual), static_cast<__nv_bfloat16*>(out));
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}
