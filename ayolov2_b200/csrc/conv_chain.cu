// Fused convolution CHAIN on the sm_100a tensor cores: 1x1 -> 3x3 (-> 1x1), intermediates never leave the SM.
//
//   t   = act1(bias1 + W1 * x)                 1x1, Cin -> C1      (computed on the 18x10 halo of a 16x8 output tile)
//   u   = act2(bias2 + W2 (*) t)               3x3 / stride 1 / pad 1, C1 -> C2
//   out = act3(bias3 + W3 * u)  (+ residual)   1x1, C2 -> C3        (optional; without it `u` (+ residual) is the output)
//
// Two users:
//   * the Tucker-2 chain emitted by scripts/tensor_decomposition/decomposition.py:363-424 (first 1x1 without bias,
//     core kxk without bias, last 1x1 carrying the bias; the enclosing kindle Conv's BN + SiLU fold into stage 3);
//   * kindle Bottleneck  x + conv2_3x3(conv1_1x1(x))  (stage 3 absent, residual = x).
//
// Data flow per tile (one CTA, persistent over tiles):
//   TMA   : x halo tile [18][10][CK1 ch] per 64-channel chunk (OOB zero fill = image border) + the matching W1 chunk
//   MMA 1 : D1[256 halo rows (2 x M128)][C1]  in TMEM                      (tcgen05.mma kind::f16, fp32 accumulate)
//   EPI 1 : TMEM -> +bias1 -> act1 -> zero outside the image (the 3x3's padding applies to t, not to x) -> bf16 ->
//           shared memory T in the UMMA *no-swizzle* K-major core-matrix layout: 8 channels x 16 B per pixel, pixels
//           contiguous inside a plane of 8 channels. A 3x3 tap is then just a shifted START ADDRESS of the same tile:
//           8-row groups (= 8 horizontally adjacent pixels) are 128 contiguous bytes for every shift, the group stride
//           (SBO) is one halo row (160 B), the K stride (LBO) one 8-channel plane.
//   MMA 2 : D2[128 pixels][C2] += T(shifted by tap) x W2[tap]   (9 x C1/16 instructions; W2 streamed by TMA)
//   EPI 2 : either the final epilogue (bias2, act2, + residual, bf16, swizzled staging, TMA store) or
//           TMEM -> bias2 -> act2 -> bf16 -> shared memory U (same no-swizzle layout), then
//   MMA 3 : D3[128][N3 tile] = U x W3, EPI 3 = final epilogue, per N3 tile (D3 re-uses the TMEM columns of D1).
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM alloc, warps 4-7 epilogues.
#include <stdlib.h>
#include <string.h>

#include "ay2_common.h"
#include "ay2_ptx.cuh"

namespace ay2 {

constexpr int CH_TH = 16, CH_TW = 16;  // output tile: two 16x8 halves side by side (one UMMA 8-row group == 8 horizontally
                                       // adjacent pixels); the halves are independent accumulators sharing every weight tile
constexpr int CH_HH = CH_TH + 2, CH_HW = CH_TW + 2;
constexpr int CH_NH = CH_HH * CH_HW;   // 324 halo pixels == rows of the stage-1 GEMM (three M=128 instructions)
constexpr int CH_THREADS = 384;        // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: two epilogue groups
constexpr int CH_EPI_THREADS = 256;
constexpr int CH_MAX_STAGES = 8;
constexpr int CH_U_PLANE = 256 * 16;   // bytes of one 8-channel plane of U (256 pixels x 16 B)

struct ChainParams {
  CUtensorMap tmX;    // input  [C][W][H][B], box [ck1][18][18][1]
  CUtensorMap tmW1;   // [cin][c1],    box [ck1][c1]
  CUtensorMap tmW2;   // [9*c1][c2],   box [ck2][c2]
  CUtensorMap tmW3;   // [c2][c3],     box [ck3][n3]
  CUtensorMap tmOut;  // output [C][W][H][B], box [oc][16][16][1]
  CUtensorMap tmRes;  // residual, same geometry as the output
  const float* bias;  // [c1 + c2 + c3] fp32 (zeros where the reference has no bias)
  unsigned long long* dbg;  // optional timeline buffer (tools/chain_timeline.py): block 0 records %globaltimer per phase
  int tiles_x, tiles_y, num_tiles;
  int in_h, in_w;
  int cin_chunks, ck1, c1, act1;
  int c1_chunks, ck2, c2, act2;
  int c3, n3, n3_tiles, c2_chunks, ck3, act3;  // c3 == 0: no stage 3
  int has_res;
  int res_from_x;    // the residual IS the input (Bottleneck shortcut): read it from the halo tile still sitting in the X ring
  int oc;            // channels per output slab (64 / 32 / 16); swizzle span = 2*oc bytes
  int nx, x_slot;    // X ring: halo chunks of x (slots of round_up(324 rows * ck1 * 2, 1024) bytes)
  int nw, w_slot;    // W ring: W1 / W2 / W3 chunks in consumption order
  int w_resident;    // all weight chunks stay in shared memory (loaded once per CTA): the "ring" has one slot per chunk
  int alias_staging; // the output staging buffer re-uses T's bytes (T is dead once stage 2 has finished)
  int t_plane;       // bytes of one 8-channel plane of T (CH_NH * 16)
  int d2_split;            // stage 2 accumulates in regions R3 / R4 of its own (no stage 3, 5 regions fit): stage 1 of the
                           // NEXT tile then overlaps the final epilogue of this one instead of waiting for it to drain
  int tmem_cols, d2_col;   // TMEM: three regions of d2_col columns: R0 = D1 rows 0..127 / D3 left half, R1 = D1 rows
                           // 128..255 / D2 left / D3 right, R2 = D1 rows 256..383 / D2 right
  int off_W, off_T, off_U, off_staging, off_bias, off_bars;  // byte offsets from the 1024-aligned shared-memory base
};

__device__ __forceinline__ void ch_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptors as (lo, hi) 32-bit halves: everything except the start address is loop-invariant, so
// the issue loop only ever adds a byte offset >> 4 to `lo` (address field = bits [0,14) of lo; no carry can leave it).
//   lo: [0,14) start address >> 4 | [16,30) leading byte offset >> 4
//   hi: [0,14) stride byte offset >> 4 | [14] version = 1 (sm_100) | [29,32) layout (0 none, 2 SW128, 4 SW64, 6 SW32)
// K-major swizzled operand with dense rows of `swz` bytes (what a TMA box with that swizzle writes): LBO unused (1),
// SBO = 8 rows. K-major operand WITHOUT swizzle (canonical ((8,m),(8,2)) : ((16 B, SBO), (2 B, LBO))): a core matrix is
// 8 rows x 16 bytes stored contiguously; LBO = distance between the two core matrices of one K = 16 step, SBO = distance
// between consecutive 8-row groups; the start address only needs 16-byte alignment, which is what lets a 3x3 tap be a
// shifted view of one halo tile. (Measured, tools/micro/mma_rate.cu: both layouts issue at the same rate.)
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFF) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint32_t desc_hi_swz(int swz) {
  return static_cast<uint32_t>((8 * swz) >> 4) | (1u << 14) | ((swz == 128 ? 2u : (swz == 64 ? 4u : 6u)) << 29);
}
__device__ __forceinline__ uint32_t desc_hi_nosw(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }

__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ uint32_t swizzled_offset_rt(uint32_t row, uint32_t chunk16, uint32_t row_bytes) {
  uint32_t off = row * row_bytes + chunk16 * 16;
  const uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
  return off ^ (((off >> 7) & mask) << 4);
}

__device__ __forceinline__ unsigned long long ch_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CH_DBG(it, slot)                                                            \
  do {                                                                              \
    if (p.dbg && blockIdx.x == 0 && (it) < 16) p.dbg[(it) * 16 + (slot)] = ch_now(); \
  } while (0)

// act(v + bias) for 8 consecutive columns, two columns per instruction (FADD2 / FMUL2 / FFMA2 round per lane exactly like
// the scalar form of silu_f: bit-identical results)
__device__ __forceinline__ void ch_bias_act8(const uint32_t* v, const float* bb, int act, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 x = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), make_float2(bb[2 * i], bb[2 * i + 1]));
    if (act == AY2_ACT_SILU) {
      const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
      float tx, ty;
      asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(h.x));
      asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(h.y));
      x = __ffma2_rn(h, make_float2(tx, ty), h);
    }
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}

// 16 accumulator columns of one row: +bias -> act (-> zero) -> bf16 -> two 16-byte stores `plane` bytes apart
// (no-swizzle core-matrix layout of T / U: 8 channels x 16 B per pixel, one plane per 8 channels).
__device__ __forceinline__ void ch_store_planes(const uint32_t* v, uint32_t bias_addr, int act, bool keep, uint32_t dst,
                                                uint32_t plane) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float bb[8];
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[0]), "=f"(bb[1]), "=f"(bb[2]), "=f"(bb[3]) : "r"(bias_addr + g * 32));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[4]), "=f"(bb[5]), "=f"(bb[6]), "=f"(bb[7]) : "r"(bias_addr + g * 32 + 16));
    float f[8];
    ch_bias_act8(v + g * 8, bb, act, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = keep ? f[i] : 0.0f;
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + g * plane), "r"(pack_bf16x2(f[0], f[1])),
                 "r"(pack_bf16x2(f[2], f[3])), "r"(pack_bf16x2(f[4], f[5])), "r"(pack_bf16x2(f[6], f[7]))
                 : "memory");
  }
}

// Final epilogue of one accumulator pair (left / right half of the tile): TMEM -> +bias -> act (-> + residual) -> bf16 ->
// swizzled staging [256 pixels][oc] per slab (pixel = y*16 + x, the order a [oc][16][16] TMA box uses) -> TMA store.
// All 256 epilogue threads call it (group `eg` drains the 16-column blocks [blk_lo, blk_hi)); `acc_full` is the
// MMA->epilogue barrier of this pair, `acc_empty` the way back. taddr0 / taddr1: TMEM addresses of the two halves.
__device__ __forceinline__ void chain_store_tile(const ChainParams& p, uint8_t* staging, uint32_t bias_addr, uint32_t taddr0,
                                                 uint32_t taddr1, int ncols, int act, int n0, int x0, int y0, int b, int et,
                                                 int eall, int blk_lo, int blk_hi, uint64_t* acc_full, uint32_t acc_phase,
                                                 uint64_t* acc_empty, uint64_t* res_full, uint32_t& res_phase,
                                                 const uint8_t* xring, int xslot0) {
  const int swo = p.oc * 2;
  const int slab_bytes = 256 * swo;
  const int oc_shift = p.oc == 64 ? 6 : (p.oc == 32 ? 5 : 4);
  const int nslab = ncols >> oc_shift;
  const bool tma_res = p.has_res && !p.res_from_x;
  if (eall == 0) {
    tma_store_wait_read<0>();  // the previous tile's stores have finished reading the staging buffer
    if (tma_res) {
      // staging aliased onto T: the tensor core may still be reading T until this accumulator is complete
      if (p.alias_staging) mbar_wait(acc_full, acc_phase);
      mbar_expect_tx(res_full, 256 * ncols * 2);
      for (int s = 0; s < nslab; ++s) tma_load_4d(&p.tmRes, res_full, staging + s * slab_bytes, n0 + s * p.oc, x0, y0, b);
    }
  }
  ch_bar_sync(1, CH_EPI_THREADS);
  mbar_wait(acc_full, acc_phase);
  tcgen05_fence_after();
  if (tma_res) {
    mbar_wait(res_full, res_phase);
    res_phase ^= 1;
  }
  const int ty = et >> 3, tx = et & 7;
  const int swa1 = p.ck1 * 2;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const uint32_t taddr = half ? taddr1 : taddr0;
    const int row = ty * CH_TW + half * 8 + tx;             // staging row of this thread's pixel
    const int hrow = (ty + 1) * CH_HW + half * 8 + tx + 1;  // the same pixel inside the halo tile
#pragma unroll 1
    for (int blk = blk_lo; blk < blk_hi; ++blk) {
      const int c0 = blk * 16;
      uint32_t v[16];
      tmem_ld_32x32b_x16(taddr + c0, v);
      tmem_ld_wait();
      uint8_t* slab = staging + (c0 >> oc_shift) * slab_bytes;
      const int chunk0 = (c0 & (p.oc - 1)) >> 3;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float bb[8];
        const uint32_t ba = bias_addr + (n0 + c0 + g * 8) * 4;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[0]), "=f"(bb[1]), "=f"(bb[2]), "=f"(bb[3]) : "r"(ba));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bb[4]), "=f"(bb[5]), "=f"(bb[6]), "=f"(bb[7]) : "r"(ba + 16));
        float f[8];
        ch_bias_act8(v + g * 8, bb, act, f);
        const uint32_t dst = smem_u32(slab) + swizzled_offset_rt(row, chunk0 + g, swo);
        if (p.has_res) {
          uint32_t src = dst;
          if (p.res_from_x) {  // channel c0 + 8g of the input: chunk (c / ck1) of this tile's halo, row hrow
            const int c = c0 + g * 8;
            int slot = xslot0 + c / p.ck1;
            if (slot >= p.nx) slot -= p.nx;
            src = smem_u32(xring + slot * p.x_slot) + swizzled_offset_rt(hrow, (c % p.ck1) / 8, swa1);
          }
          uint4 rv;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(src));
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
  tcgen05_fence_before();
  mbar_arrive(acc_empty);  // accumulators drained
  fence_proxy_async_smem();
  ch_bar_sync(1, CH_EPI_THREADS);
  if (eall == 0) {
    for (int s = 0; s < nslab; ++s) tma_store_4d(&p.tmOut, staging + s * slab_bytes, n0 + s * p.oc, x0, y0, b);
    tma_store_commit();
  }
}

__global__ void __launch_bounds__(CH_THREADS, 2) conv_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xring = smem;
  uint8_t* wring = smem + p.off_W;
  uint8_t* T = smem + p.off_T;
  uint8_t* U = smem + p.off_U;
  uint8_t* staging = smem + p.off_staging;
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bars);
  uint64_t* wfull = bars;                           // [CH_MAX_STAGES] W ring (wfull[0] doubles as "resident weights landed")
  uint64_t* wempty = bars + CH_MAX_STAGES;          // [CH_MAX_STAGES]
  uint64_t* xfull = bars + 2 * CH_MAX_STAGES;       // [4] X ring
  uint64_t* xempty = xfull + 4;                     // [4]
  uint64_t* d1_full = xfull + 8;                    // MMA -> epilogue 1
  uint64_t* t_ready = d1_full + 1;                  // epilogue 1 -> MMA (T written, D1 drained)
  uint64_t* d2_full = d1_full + 2;
  uint64_t* u_ready = d1_full + 3;
  uint64_t* d3_full = d1_full + 4;
  uint64_t* acc_empty = d1_full + 5;                // final epilogue -> MMA: last accumulators drained
  uint64_t* res_full = d1_full + 6;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(d1_full + 7);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int swa1 = p.ck1 * 2, swa2 = p.ck2 * 2, swa3 = p.ck3 * 2;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const bool resident = p.w_resident != 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < CH_MAX_STAGES; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], p.res_from_x ? 1 + CH_EPI_THREADS : 1);  // + the epilogue reading the shortcut from the slot
    }
    mbar_init(d1_full, 1);
    mbar_init(t_ready, CH_EPI_THREADS);
    mbar_init(d2_full, 1);
    mbar_init(u_ready, CH_EPI_THREADS);
    mbar_init(d3_full, 1);
    mbar_init(acc_empty, CH_EPI_THREADS);
    mbar_init(res_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmOut);
  }
  for (int i = threadIdx.x; i < p.c1 + p.c2 + p.c3; i += CH_THREADS) bias_s[i] = p.bias[i];
  if (warp == 2) tmem_alloc(tmem_ptr_s, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  // weight chunks in consumption order: [W1: cin_chunks] [W2: 9 * c1_chunks] [W3: n3_tiles * c2_chunks]; with resident
  // weights chunk i lives in slot i for the whole kernel, otherwise the chunks stream through the nw-slot ring per tile
  if (warp == 0) {
    // =============================== TMA producer ===============================
    // All 32 lanes run the loop (warp-uniform control flow keeps addresses / coordinates in uniform registers);
    // one elected lane issues the copies.
    int ws = 0, wph = 0, xs = 0, xph = 0;
    const int nx = p.nx, nw = p.nw, x_slot = p.x_slot, w_slot = p.w_slot;
    const int cin_chunks = p.cin_chunks, c1_chunks = p.c1_chunks, c2_chunks = p.c2_chunks, n3_tiles = p.n3_tiles;
    const int ck1 = p.ck1, ck2 = p.ck2, ck3 = p.ck3, c1 = p.c1, n3 = p.n3;
    const uint32_t xbytes = CH_NH * swa1, w1bytes = p.c1 * swa1, w2bytes = p.c2 * swa2, w3bytes = p.n3 * swa3;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int ty = r / p.tiles_x;
      const int y0 = ty * CH_TH, x0 = (r - ty * p.tiles_x) * CH_TW;
      const bool load_w = !resident || first;
      if (resident && first && elect_one())
        mbar_expect_tx(&wfull[0], cin_chunks * w1bytes + 9 * c1_chunks * w2bytes + n3_tiles * c2_chunks * w3bytes);
      __syncwarp();
      for (int kc = 0; kc < cin_chunks; ++kc) {  // stage 1: halo chunk of x, W1 chunk
        mbar_wait(&xempty[xs], xph ^ 1);
        if (!resident) mbar_wait(&wempty[ws], wph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&xfull[xs], xbytes);
          tma_load_4d(&p.tmX, &xfull[xs], xring + xs * x_slot, kc * ck1, x0 - 1, y0 - 1, b);
          if (load_w) {
            if (!resident) mbar_expect_tx(&wfull[ws], w1bytes);
            tma_load_2d(&p.tmW1, &wfull[resident ? 0 : ws], wring + ws * w_slot, kc * ck1, 0);
          }
        }
        __syncwarp();
        if (++xs == nx) { xs = 0; xph ^= 1; }
        if (load_w && ++ws == nw) { ws = 0; wph ^= 1; }
      }
      if (load_w) {
        for (int tap = 0; tap < 9; ++tap) {          // stage 2: W2 per tap and channel chunk
          for (int cc = 0; cc < c1_chunks; ++cc) {
            if (!resident) mbar_wait(&wempty[ws], wph ^ 1);
            if (elect_one()) {
              if (!resident) mbar_expect_tx(&wfull[ws], w2bytes);
              tma_load_2d(&p.tmW2, &wfull[resident ? 0 : ws], wring + ws * w_slot, tap * c1 + cc * ck2, 0);
            }
            __syncwarp();
            if (++ws == nw) { ws = 0; wph ^= 1; }
          }
        }
        for (int n = 0; n < n3_tiles; ++n) {         // stage 3: W3 per N tile and channel chunk
          for (int cc = 0; cc < c2_chunks; ++cc) {
            if (!resident) mbar_wait(&wempty[ws], wph ^ 1);
            if (elect_one()) {
              if (!resident) mbar_expect_tx(&wfull[ws], w3bytes);
              tma_load_2d(&p.tmW3, &wfull[resident ? 0 : ws], wring + ws * w_slot, cc * ck3, n * n3);
            }
            __syncwarp();
            if (++ws == nw) { ws = 0; wph ^= 1; }
          }
        }
      }
      first = false;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // Warp-uniform loop, one elected lane issues tcgen05.mma / commit. Descriptors are (lo, hi) pairs whose hi halves
    // and lo bases are loop-invariant: one 32-bit add per operand per instruction.
    const uint32_t idesc1 = make_idesc_bf16_f32(128, p.c1);
    const uint32_t idesc2 = make_idesc_bf16_f32(128, p.c2);
    const uint32_t idesc3 = make_idesc_bf16_f32(128, p.n3 > 0 ? p.n3 : 16);
    const int nx = p.nx, nw = p.nw;
    const int cin_chunks = p.cin_chunks, c1_chunks = p.c1_chunks, c2_chunks = p.c2_chunks, n3_tiles = p.n3_tiles;
    const int ks1 = p.ck1 / 16, ks2 = p.ck2 / 16, ks3 = p.ck3 / 16;
    const uint32_t hi1 = desc_hi_swz(swa1), hi2 = desc_hi_swz(swa2), hi3 = desc_hi_swz(swa3);
    const uint32_t hiT = desc_hi_nosw(CH_HW * 16), hiU = desc_hi_nosw(128);
    const uint32_t x_lo0 = desc_lo(smem_u32(xring), 16), x_step = p.x_slot >> 4, x_blk = (128 * swa1) >> 4;
    const uint32_t w_lo0 = desc_lo(smem_u32(wring), 16), w_step = p.w_slot >> 4;
    const uint32_t t_lo0 = desc_lo(smem_u32(T), p.t_plane), t_kstep = (2 * p.t_plane) >> 4;
    const uint32_t u_lo0 = desc_lo(smem_u32(U), CH_U_PLANE), u_kstep = (2 * CH_U_PLANE) >> 4;
    const uint32_t r1_t = tmem_base + p.d2_col, r2_t = tmem_base + 2 * p.d2_col;
    const bool split = p.d2_split != 0;
    const uint32_t d2l_t = split ? tmem_base + 3 * p.d2_col : r1_t, d2r_t = split ? tmem_base + 4 * p.d2_col : r2_t;
    const bool has3 = p.c3 != 0;
    int ws = 0, wph = 0, xs = 0, xph = 0;
    uint32_t t_phase = 0, u_phase = 0, ae_phase = 0;
    bool acc_pending = false;  // the last accumulators handed to the final epilogue have not been drained yet
    int it = -1;
    if (resident) {
      mbar_wait(&wfull[0], 0);
      tcgen05_fence_after();
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      ++it;
      if (lane == 0) CH_DBG(it, 0);  // tile start (MMA warp)
      if (resident) ws = 0;
      if (acc_pending && !split) {  // D1 shares its TMEM columns with D2 / D3 of the previous tile
        mbar_wait(acc_empty, ae_phase);
        ae_phase ^= 1;
        acc_pending = false;
      }
      // ---- stage 1: D1[j] = Xhalo[j*128 .. +128) x W1^T, j = 0..2 (three independent accumulators)
      for (int kc = 0; kc < cin_chunks; ++kc) {
        mbar_wait(&xfull[xs], xph);
        if (!resident) mbar_wait(&wfull[ws], wph);
        tcgen05_fence_after();
        if (kc == 0 && lane == 0) CH_DBG(it, 1);  // previous accumulators drained, first x chunk landed
        const uint32_t a_lo = x_lo0 + xs * x_step, b_lo = w_lo0 + ws * w_step;
        if (elect_one()) {
          for (int k = 0; k < ks1; ++k) {
            const uint32_t acc = (kc | k) != 0 ? 1u : 0u;
            umma_ss(tmem_base, a_lo + 2 * k, hi1, b_lo + 2 * k, hi1, idesc1, acc);
            umma_ss(r1_t, a_lo + x_blk + 2 * k, hi1, b_lo + 2 * k, hi1, idesc1, acc);
            umma_ss(r2_t, a_lo + 2 * x_blk + 2 * k, hi1, b_lo + 2 * k, hi1, idesc1, acc);
          }
          umma_commit(&xempty[xs]);
          if (!resident) umma_commit(&wempty[ws]);
          if (kc == cin_chunks - 1) umma_commit(d1_full);
        }
        __syncwarp();
        if (++xs == nx) { xs = 0; xph ^= 1; }
        if (++ws == nw) { ws = 0; wph ^= 1; }
      }
      if (lane == 0) CH_DBG(it, 2);  // stage-1 MMAs issued
      // ---- stage 2: D2[left|right] += T(shifted by tap) x W2[tap]^T
      mbar_wait(t_ready, t_phase);
      t_phase ^= 1;
      if (acc_pending) {  // split regions: only D2 waits for the previous tile's final epilogue (already true in practice:
        mbar_wait(acc_empty, ae_phase);  // the epilogue warps wrote this tile's T after draining the previous D2)
        ae_phase ^= 1;
        acc_pending = false;
      }
      tcgen05_fence_after();
      if (lane == 0) CH_DBG(it, 3);  // T ready
      for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t a_tap = t_lo0 + ky * CH_HW + kx;  // one pixel == 16 bytes == 1 address unit
        for (int cc = 0; cc < c1_chunks; ++cc) {
          if (!resident) {
            mbar_wait(&wfull[ws], wph);
            tcgen05_fence_after();
          }
          const uint32_t b_lo = w_lo0 + ws * w_step;
          const uint32_t a_lo = a_tap + cc * ks2 * t_kstep;
          if (elect_one()) {
            for (int k = 0; k < ks2; ++k) {  // left / right half: independent accumulators sharing the weight tile
              const uint32_t acc = (tap | cc | k) != 0 ? 1u : 0u;
              umma_ss(d2l_t, a_lo + k * t_kstep, hiT, b_lo + 2 * k, hi2, idesc2, acc);
              umma_ss(d2r_t, a_lo + 8 + k * t_kstep, hiT, b_lo + 2 * k, hi2, idesc2, acc);
            }
            if (!resident) umma_commit(&wempty[ws]);
            if (tap == 8 && cc == c1_chunks - 1) umma_commit(d2_full);
          }
          __syncwarp();
          if (++ws == nw) { ws = 0; wph ^= 1; }
        }
      }
      if (lane == 0) CH_DBG(it, 4);  // stage-2 MMAs issued (all W2 chunks had landed)
      if (!has3) acc_pending = true;
      // ---- stage 3: D3[n][left|right] = U x W3[n]^T
      if (has3) {
        mbar_wait(u_ready, u_phase);
        u_phase ^= 1;
        tcgen05_fence_after();
        for (int n = 0; n < n3_tiles; ++n) {
          if (acc_pending) {
            mbar_wait(acc_empty, ae_phase);
            ae_phase ^= 1;
            acc_pending = false;
            tcgen05_fence_after();
          }
          for (int cc = 0; cc < c2_chunks; ++cc) {
            if (!resident) {
              mbar_wait(&wfull[ws], wph);
              tcgen05_fence_after();
            }
            const uint32_t b_lo = w_lo0 + ws * w_step;
            const uint32_t a_lo = u_lo0 + cc * ks3 * u_kstep;
            if (elect_one()) {
              for (int k = 0; k < ks3; ++k) {
                const uint32_t acc = (cc | k) != 0 ? 1u : 0u;
                umma_ss(tmem_base, a_lo + k * u_kstep, hiU, b_lo + 2 * k, hi3, idesc3, acc);
                umma_ss(r1_t, a_lo + 128 + k * u_kstep, hiU, b_lo + 2 * k, hi3, idesc3, acc);  // + 128 pixels x 16 B
              }
              if (!resident) umma_commit(&wempty[ws]);
              if (cc == c2_chunks - 1) umma_commit(d3_full);
            }
            __syncwarp();
            if (++ws == nw) { ws = 0; wph ^= 1; }
          }
          acc_pending = true;
        }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogues ===============================
    // Two groups of 4 warps; both cover all 128 TMEM lanes (a warp may read the lane quadrant warp % 4) and split the
    // 16-column blocks of every accumulator between them.
    const int eall = threadIdx.x - 128;  // 0..255
    const int et = eall & 127;           // accumulator row == TMEM lane
    const int eg = eall >> 7;            // epilogue group
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t d1_phase = 0, d2_phase = 0, d3_phase = 0, res_phase = 0;
    const uint32_t bias1 = smem_u32(bias_s);
    const uint32_t bias2 = bias1 + p.c1 * 4;
    const uint32_t bias3 = bias2 + p.c2 * 4;
    auto split = [&](int ncols, int& lo, int& hi) {  // this group's share of the 16-column blocks
      const int nblk = ncols / 16, mid = (nblk + 1) / 2;
      lo = eg ? mid : 0;
      hi = eg ? nblk : mid;
    };
    int b1_lo, b1_hi, b2_lo, b2_hi, b3_lo, b3_hi;
    split(p.c1, b1_lo, b1_hi);
    split(p.c2, b2_lo, b2_hi);
    split(p.n3, b3_lo, b3_hi);
    const uint32_t t_plane = p.t_plane;
    int it = -1;
    int xslot0 = -p.cin_chunks;  // X-ring slot of this tile's first halo chunk
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      ++it;
      xslot0 += p.cin_chunks;
      if (xslot0 >= p.nx) xslot0 -= p.nx;
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int ty = r / p.tiles_x;
      const int y0 = ty * CH_TH, x0 = (r - ty * p.tiles_x) * CH_TW;
      // ---- epilogue 1: D1 -> T
      if (p.alias_staging) {  // T is about to be rewritten: the previous tile's TMA store must have read it out
        if (eall == 0) tma_store_wait_read<0>();
        ch_bar_sync(1, CH_EPI_THREADS);
      }
      mbar_wait(d1_full, d1_phase);
      d1_phase ^= 1;
      tcgen05_fence_after();
      if (eall == 0) CH_DBG(it, 8);  // D1 complete
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        if (pass * 128 + (et & ~31) >= CH_NH) continue;  // warp-uniform: this warp's rows are all beyond the halo
        const int h = pass * 128 + et;
        bool inb = false;
        if (h < CH_NH) {
          const int hy = h / CH_HW, hx = h - hy * CH_HW;
          const int iy = y0 - 1 + hy, ix = x0 - 1 + hx;
          inb = iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
        }
        const uint32_t tbase = tmem_base + lane_off + pass * p.d2_col;
        const uint32_t trow = smem_u32(T) + h * 16;
#pragma unroll 1
        for (int blk = b1_lo; blk < b1_hi; ++blk) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tbase + blk * 16, v);
          tmem_ld_wait();
          if (h < CH_NH) ch_store_planes(v, bias1 + blk * 64, p.act1, inb, trow + blk * 2 * t_plane, t_plane);
        }
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();  // generic-proxy writes of T -> visible to the tensor core's async-proxy reads
      mbar_arrive(t_ready);
      if (eall == 0) CH_DBG(it, 9);  // epilogue 1 done
      if (p.c3 == 0) {
        // ---- epilogue 2 = final
        const int d2reg = p.d2_split ? 3 : 1;
        chain_store_tile(p, staging, bias2, tmem_base + lane_off + d2reg * p.d2_col, tmem_base + lane_off + (d2reg + 1) * p.d2_col, p.c2, p.act2, 0,
                         x0, y0, b, et, eall, b2_lo, b2_hi, d2_full, d2_phase, acc_empty, res_full, res_phase, xring, xslot0);
        d2_phase ^= 1;
        if (p.res_from_x) {  // the shortcut has been read: hand this tile's halo slots back to the producer
          for (int kc = 0; kc < p.cin_chunks; ++kc) {
            int slot = xslot0 + kc;
            if (slot >= p.nx) slot -= p.nx;
            mbar_arrive(&xempty[slot]);
          }
        }
        if (eall == 0) CH_DBG(it, 11);  // final epilogue done (store issued)
      } else {
        // ---- epilogue 2: D2 -> U
        mbar_wait(d2_full, d2_phase);
        d2_phase ^= 1;
        tcgen05_fence_after();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const uint32_t urow = smem_u32(U) + (half * 128 + et) * 16;
          const uint32_t tsrc = tmem_base + lane_off + (1 + half) * p.d2_col;
#pragma unroll 1
          for (int blk = b2_lo; blk < b2_hi; ++blk) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tsrc + blk * 16, v);
            tmem_ld_wait();
            ch_store_planes(v, bias2 + blk * 64, p.act2, true, urow + blk * 2 * CH_U_PLANE, CH_U_PLANE);
          }
        }
        tcgen05_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(u_ready);
        // ---- epilogue 3 per N tile
        for (int n = 0; n < p.n3_tiles; ++n) {
          chain_store_tile(p, staging, bias3, tmem_base + lane_off, tmem_base + lane_off + p.d2_col, p.n3, p.act3, n * p.n3, x0,
                           y0, b, et, eall, b3_lo, b3_hi, d3_full, d3_phase, acc_empty, res_full, res_phase, xring, xslot0);
          d3_phase ^= 1;
        }
      }
    }
    if (eall == 0) tma_store_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace ay2

using namespace ay2;

struct ay2_chain_plan {
  ChainParams kp;
  ay2_chain_desc desc;
  int grid, ctas_per_sm;
  size_t smem;
};

static int chunk_for(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }
static int round_up_i(int v, int m) { return (v + m - 1) / m * m; }

extern "C" int ay2_chain_supported(const ay2_chain_desc* d) {
  if (!d) return 0;
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1) return 0;
  if (d->cin % 16 || d->c1 % 16 || d->c2 % 16 || d->c3 % 16) return 0;
  if (d->cin < 16 || d->c1 < 16 || d->c1 > 128 || d->c2 < 16 || d->c2 > 160) return 0;  // 3 x max(c1, c2, n3) TMEM columns
  if (d->c3 == 0 && d->c2 % 16) return 0;
  if (d->in_cstride % 8 || d->out_cstride % 8 || d->res_cstride % 8) return 0;
  return 1;
}

extern "C" int ay2_chain_plan_create(const ay2_chain_desc* d, const void* in, const void* w1, const void* w2, const void* w3,
                                     const float* bias, const void* residual, void* out, ay2_chain_plan** plan_out) {
  AY2_REQUIRE(d && in && w1 && w2 && bias && out && plan_out, "ay2_chain_plan_create: null argument");
  AY2_REQUIRE(ay2_chain_supported(d),
              "chain cin=%d c1=%d c2=%d c3=%d k=%dx%d s=%d p=%d is outside the fused kernel's envelope (3x3/s1/p1, channels "
              "multiples of 16, c1 <= 128, c2 <= 160)",
              d->cin, d->c1, d->c2, d->c3, d->kh, d->kw, d->stride, d->pad);
  AY2_REQUIRE(d->c3 == 0 || w3, "stage 3 needs its weights");
  AY2_REQUIRE(d->res_cstride == 0 || residual, "residual stride given without a residual pointer");
  AY2_REQUIRE(out != in, "the fused chain cannot run in place (tiles read their neighbours' halo)");
  ay2_chain_plan* pl = new ay2_chain_plan();
  memset(pl, 0, sizeof(*pl));
  pl->desc = *d;
  ChainParams& kp = pl->kp;
  const int H = d->in_h, W = d->in_w;
  kp.in_h = H;
  kp.in_w = W;
  kp.tiles_x = ceil_div(W, CH_TW);
  kp.tiles_y = ceil_div(H, CH_TH);
  kp.num_tiles = kp.tiles_x * kp.tiles_y * d->batch;
  kp.c1 = d->c1;
  kp.act1 = d->act1;
  kp.c2 = d->c2;
  kp.act2 = d->act2;
  kp.c3 = d->c3;
  kp.act3 = d->act3;
  if (d->c3) {
    // N tile of stage 3: the largest divisor of c3 that is a multiple of 16 and <= 128 (staging and W3 slots stay small)
    int n3 = 0;
    for (int cand = 128; cand >= 16; cand -= 16)
      if (d->c3 % cand == 0) {
        n3 = cand;
        break;
      }
    kp.n3 = n3;
    kp.n3_tiles = d->c3 / n3;
  }
  kp.has_res = d->res_cstride != 0;
  const int cout = d->c3 ? d->c3 : d->c2;
  const int ntile = d->c3 ? kp.n3 : d->c2;  // channels per final accumulator tile
  kp.oc = ntile % 64 == 0 ? 64 : (ntile % 32 == 0 ? 32 : 16);
  kp.t_plane = CH_NH * 16;
  kp.bias = bias;
  // TMEM: three regions of max(c1, c2, n3) columns (see ChainParams::d2_col)
  int regW = d->c1 > d->c2 ? d->c1 : d->c2;
  if (kp.n3 > regW) regW = kp.n3;
  kp.d2_col = regW;
  // Without a stage 3, five regions (D1 x 3, D2 x 2) decouple consecutive tiles if they still leave room for two CTAs.
  // Implemented and parity-tested (AY2_CHAIN_SPLIT=1) but measured NEUTRAL (r01: 0.147 ms either way for the 32-channel
  // Bottleneck @160x160): a tile's E1 -> stage 2 -> E2 chain (1.4 + 2.3 + 1.45 us) is the period, and the epilogue warps
  // that would consume the early D1 are still draining the previous tile. Off by default.
  kp.d2_split = (d->c3 == 0 && 5 * regW <= 256 && getenv("AY2_CHAIN_SPLIT") && atoi(getenv("AY2_CHAIN_SPLIT")) == 1) ? 1 : 0;
  int cols = (kp.d2_split ? 5 : 3) * regW, pow2 = 32;
  while (pow2 < cols) pow2 *= 2;
  AY2_REQUIRE(pow2 <= 512, "chain needs %d TMEM columns", cols);
  kp.tmem_cols = pow2;
  // Shared memory: [X ring | W ring | T | U | staging | bias | barriers]. Pick the layout that lets two CTAs share an SM
  // (the stages of one tile are serialised inside a CTA; the co-resident CTA is what overlaps tensor, epilogue and TMA
  // work): first fewer X slots, then staging aliased onto T, then 32-channel instead of 64-channel operand chunks.
  // Small weight sets (<= 48 KB in slots) stay resident in shared memory for the whole kernel instead of streaming.
  const int t_bytes = round_up_i(d->c1 / 8 * kp.t_plane, 1024);
  const int u_bytes = d->c3 ? d->c2 / 8 * CH_U_PLANE : 0;
  const int staging_bytes = 256 * ntile * 2;
  const int bias_bytes = round_up_i((d->c1 + d->c2 + d->c3) * 4, 128);
  const int bars_bytes = (2 * CH_MAX_STAGES + 8 + 8) * 8;
  const bool res_is_input = kp.has_res && residual == in && d->res_cstride == d->in_cstride && d->c3 == 0 && d->c2 == d->cin;
  int ctas = 0;
  static const int env_ctas = getenv("AY2_CHAIN_MAX_CTAS") ? atoi(getenv("AY2_CHAIN_MAX_CTAS")) : 2;
  static const int env_res = getenv("AY2_CHAIN_RESIDENT") ? atoi(getenv("AY2_CHAIN_RESIDENT")) : 1;
  for (int c = env_ctas < 2 ? 1 : 2; c >= 1 && !ctas; --c) {  // 384 threads x <= 80 registers: at most 2 CTAs per SM
    if (kp.tmem_cols * c > 512) continue;
    const int budget = 227 * 1024 / c - 1024;
    for (int small = 0; small < 2 && !ctas; ++small) {
      const int cap = small ? 32 : 64;
      const int ck1 = chunk_for(d->cin) > cap ? cap : chunk_for(d->cin);
      const int ck2 = chunk_for(d->c1) > cap ? cap : chunk_for(d->c1);
      const int ck3 = d->c3 ? (chunk_for(d->c2) > cap ? cap : chunk_for(d->c2)) : 16;
      const int swa1 = ck1 * 2, swa2 = ck2 * 2, swa3 = ck3 * 2;
      const int x_slot = round_up_i(CH_NH * swa1, 1024);
      int w_slot = d->c1 * swa1;
      if (d->c2 * swa2 > w_slot) w_slot = d->c2 * swa2;
      if (d->c3 && kp.n3 * swa3 > w_slot) w_slot = kp.n3 * swa3;
      w_slot = round_up_i(w_slot, 1024);
      const int cin_chunks = d->cin / ck1;
      const int w_chunks = cin_chunks + 9 * (d->c1 / ck2) + (d->c3 ? kp.n3_tiles * (d->c2 / ck3) : 0);
      static const int opts[4][2] = {{2, 0}, {1, 0}, {2, 1}, {1, 1}};  // (tiles of halo chunks in the X ring, staging aliased onto T)
      for (int o = 0; o < 4 && !ctas; ++o) {
        int nx = opts[o][0] * cin_chunks;
        if (nx > 4) nx = cin_chunks <= 4 ? cin_chunks : 4;
        const int alias = opts[o][1];
        // measured (r01): reading the shortcut from the halo slot is slower than a residual TMA load (bank conflicts on the
        // 64/128-byte rows and the slot is held until the last epilogue) -> opt-in only
        static const int env_rfx = getenv("AY2_CHAIN_RES_FROM_X") ? atoi(getenv("AY2_CHAIN_RES_FROM_X")) : 0;
        const int rfx = env_rfx && res_is_input && nx % cin_chunks == 0 ? 1 : 0;  // whole tiles in the X ring
        const int t_region = alias && staging_bytes > t_bytes ? staging_bytes : t_bytes;
        const int fixed = nx * x_slot + t_region + u_bytes + (alias ? 0 : staging_bytes) + bias_bytes + bars_bytes + 2048;
        int nw = (budget - fixed) / w_slot;
        int res = 0;
        if (env_res && !small && w_chunks * w_slot <= 48 * 1024 && nw >= w_chunks) {
          res = 1;
          nw = w_chunks;
        } else {
          if (nw > CH_MAX_STAGES) nw = CH_MAX_STAGES;
          if (nw > w_chunks) nw = w_chunks;  // a deeper ring than one tile's worth buys nothing
        }
        const int need = w_chunks < 4 ? w_chunks : 4;
        if (res || nw >= need || (c == 1 && small && o == 3 && nw >= 2)) {
          ctas = c;
          kp.ck1 = ck1, kp.ck2 = ck2, kp.ck3 = ck3;
          kp.cin_chunks = cin_chunks;
          kp.c1_chunks = d->c1 / ck2;
          kp.c2_chunks = d->c3 ? d->c2 / ck3 : 0;
          kp.nx = nx, kp.nw = nw, kp.alias_staging = alias, kp.w_resident = res, kp.res_from_x = rfx;
          kp.x_slot = x_slot, kp.w_slot = w_slot;
          kp.off_W = nx * x_slot;
          kp.off_T = kp.off_W + nw * w_slot;
          kp.off_U = kp.off_T + t_region;
          kp.off_staging = alias ? kp.off_T : round_up_i(kp.off_U + u_bytes, 1024);
          kp.off_bias = alias ? round_up_i(kp.off_U + u_bytes, 128) : kp.off_staging + staging_bytes;
          kp.off_bars = kp.off_bias + bias_bytes;
        }
      }
    }
  }
  if (!ctas) {
    delete pl;
    set_error("chain cin=%d c1=%d c2=%d c3=%d does not fit in shared memory", d->cin, d->c1, d->c2, d->c3);
    return AY2_ERR_INVALID;
  }
  pl->smem = (size_t)kp.off_bars + bars_bytes + 1024;
  pl->ctas_per_sm = ctas;

  int rc = encode_act_map(&kp.tmX, in, d->cin, W, H, d->batch, d->in_cstride, (int64_t)d->in_cstride * W,
                          (int64_t)d->in_cstride * W * H, kp.ck1, CH_HW, CH_HH);
  if (rc == AY2_OK) rc = encode_weight_map(&kp.tmW1, w1, d->cin, d->c1, kp.ck1, d->c1);
  if (rc == AY2_OK) rc = encode_weight_map(&kp.tmW2, w2, 9 * d->c1, d->c2, kp.ck2, d->c2);
  if (rc == AY2_OK && d->c3) rc = encode_weight_map(&kp.tmW3, w3, d->c2, d->c3, kp.ck3, kp.n3);
  if (rc == AY2_OK)
    rc = encode_act_map(&kp.tmOut, out, cout, W, H, d->batch, d->out_cstride, (int64_t)d->out_cstride * W,
                        (int64_t)d->out_cstride * W * H, kp.oc, CH_TW, CH_TH);
  if (rc == AY2_OK && kp.has_res)
    rc = encode_act_map(&kp.tmRes, residual, cout, W, H, d->batch, d->res_cstride, (int64_t)d->res_cstride * W,
                        (int64_t)d->res_cstride * W * H, kp.oc, CH_TW, CH_TH);
  if (rc != AY2_OK) {
    delete pl;
    return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    delete pl;
    set_error("cudaFuncSetAttribute(chain smem) failed: %s", cudaGetErrorString(e));
    return AY2_ERR_CUDA;
  }
  cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int resident = sms * ctas;
  pl->grid = kp.num_tiles < resident ? kp.num_tiles : resident;
  *plan_out = pl;
  return AY2_OK;
}

extern "C" int ay2_chain_plan_set_debug(ay2_chain_plan* pl, unsigned long long* dbg) {
  AY2_REQUIRE(pl, "ay2_chain_plan_set_debug: null plan");
  pl->kp.dbg = dbg;  // device buffer of (16 x 16 + 64) uint64 (block 0's per-phase %globaltimer for its first 16 tiles), or NULL
  return AY2_OK;
}

extern "C" int ay2_chain_plan_info(const ay2_chain_plan* pl, int32_t* out8) {
  AY2_REQUIRE(pl && out8, "ay2_chain_plan_info: null argument");
  out8[0] = pl->ctas_per_sm, out8[1] = pl->grid, out8[2] = (int)pl->smem, out8[3] = pl->kp.nx, out8[4] = pl->kp.nw;
  out8[5] = pl->kp.alias_staging + 2 * pl->kp.w_resident + 4 * pl->kp.res_from_x, out8[6] = pl->kp.tmem_cols, out8[7] = pl->kp.ck2;
  return AY2_OK;
}

extern "C" int ay2_chain_plan_run(const ay2_chain_plan* pl, void* stream) {
  AY2_REQUIRE(pl, "ay2_chain_plan_run: null plan");
  conv_chain_kernel<<<pl->grid, CH_THREADS, pl->smem, static_cast<cudaStream_t>(stream)>>>(pl->kp);
  AY2_CHECK_LAUNCH();
  count_launch();
  return AY2_OK;
}

extern "C" int ay2_chain_plan_destroy(ay2_chain_plan* pl) {
  delete pl;
  return AY2_OK;
}

extern "C" double ay2_chain_plan_flops(const ay2_chain_plan* pl) {
  if (!pl) return 0.0;
  const ay2_chain_desc& d = pl->desc;
  const double px = (double)d.batch * d.in_h * d.in_w;
  return 2.0 * px * ((double)d.cin * d.c1 + 9.0 * d.c1 * d.c2 + (double)d.c2 * d.c3);
}
