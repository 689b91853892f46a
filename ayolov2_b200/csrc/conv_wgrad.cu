// Convolution weight gradient on the sm_100a tensor cores.
//
//   dW[n, kh, kw, c] = sum_{b, oy, ox} dz[b, oy, ox, n] * x[b, oy*s + kh - p, ox*s + kw - p, c]
//
// GEMM view: M = Cout (128 per CTA), N = Cin tile, K = output pixels (the reduction runs over the whole batch).
// Both operands are stored pixel-major in HBM (NHWC), i.e. the *reduction* index is the slow one: they are
// MN-major UMMA operands. A TMA box [64 ch][BW][BH][1] lands in shared memory as 128 rows (pixels) of 128 bytes
// (channels) with the 128-byte swizzle, which is exactly the canonical MN-major SWIZZLE_128B layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 64 channels contiguous, 8 pixels per 1024-byte atom
// (SBO = 1024), the next 64-channel block LBO = 128 rows * 128 B further on.
// For every filter tap the x tile is the same spatial box shifted by the tap offset (TMA zero fill = padding;
// stride 2 reads through four parity views), and each tap owns N_T accumulator columns in TMEM, so one pass over
// dz produces all KH*KW taps. The pixel range is split across CTAs (split-K); partial sums are added to the fp32
// gradient with red.global.add.f32.
//
// Reference: the conv/BN backward that `scaler.scale(loss).backward()` (scripts/train/yolo_trainer.py:329) runs
// through cuDNN for every kindle Conv.
#include <stdlib.h>
#include <string.h>

#include "ay2_common.h"
#include "ay2_ptx.cuh"

namespace ay2 {

struct WgradParams {
  CUtensorMap tmDz;    // [Cout][OW][OH][B], box [64][BW][BH][1]
  CUtensorMap tmX[4];  // input views, box [XB][BW][BH][1]
  float* dw;           // [cout][taps][cin] fp32, accumulated with atomics
  int cout, cin;
  int num_m_tiles, num_n_tiles, ksplit;
  int boxes_x, boxes_per_img, total_chunks;
  int BH, BW, NB;
  int kh, kw, stride, pad;
  int halo_tile_bytes;  // HALO kernels: bytes of one image's (BH+2) x (BW+2) x XB-channel tile
};

template <int N_T, int XB>  // N_T: input channels per CTA (multiple of XB), XB: channels per x box (32 or 64)
struct WgradCfg {
  static constexpr int P = 128;                       // pixels per K chunk
  static constexpr int A_BYTES = 2 * P * 128;         // two 64-channel blocks of dz
  static constexpr int XROW = XB * 2;                 // bytes per x-tile row == swizzle span
  static constexpr int XBLK_BYTES = P * XROW;         // one XB-channel block of one tap
  static constexpr int NXB = N_T / XB;
  static constexpr int TAP_BYTES = NXB * XBLK_BYTES;
};

// HALO (3x3 / stride 1 / pad 1 only): instead of nine shifted copies of the x tile (one per tap) a chunk loads each box's
// (BH+2) x (BW+2) halo tile ONCE; tap (kh, kw) is the same tile read from a start address shifted by kh halo lines + kw
// pixel rows (SWIZZLE_64B/128B are functions of the absolute shared-memory address for the TMA write and the UMMA read
// alike, as in conv_halo_kernel). A K slice of 16 pixels is two 8-row groups of 8 horizontally adjacent pixels: SBO is
// 8 rows when the box is >= 16 pixels wide and one halo line when it is 8 wide. The three taps of one filter row are one
// pixel row apart, i.e. three N blocks at LBO = one row: one N = 3 * N_T instruction per filter row.
// Per 128-pixel chunk the x side shrinks from 9 * 128 TMA rows to (BH+2) * (BW+2) (180 for an 8 x 16 box): the kernel was
// bound by exactly that fill (training-step kernels).
constexpr int kWgradHaloMax = 20 * 1024;  // bytes reserved per stage for the halo tiles of a chunk (one 8 x 16 box: 11.25 KB; eight 2 x 8: 20 KB)

template <int N_T, int XB, int TAPS, bool HALO = false>
__global__ void __launch_bounds__(256, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgradCfg<N_T, XB>;
  static_assert(!HALO || (TAPS == 9 && N_T == XB), "the halo form is the 3x3 kernel with one channel block per CTA");
  constexpr int STAGE_BYTES = Cfg::A_BYTES + (HALO ? kWgradHaloMax : TAPS * Cfg::TAP_BYTES);
  constexpr int NSTAGES = (220 * 1024) / STAGE_BYTES >= 4 ? 4 : (220 * 1024) / STAGE_BYTES >= 3 ? 3 : 2;
  static_assert((220 * 1024) / STAGE_BYTES >= 2, "stage too large");
  constexpr int TMEM_COLS_RAW = TAPS * N_T;
  constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128
                            : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static_assert(TMEM_COLS_RAW <= 512, "too many accumulator columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + NSTAGES;
  uint64_t* done_bar = bars + 2 * NSTAGES;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.ksplit;
  const int ks = blockIdx.x - tile * p.ksplit;
  const int mt = tile / p.num_n_tiles;
  const int nt = tile - mt * p.num_n_tiles;
  const int m0 = mt * 128, c0 = nt * N_T;
  const int per = (p.total_chunks + p.ksplit - 1) / p.ksplit;
  const int chunk_begin = ks * per;
  const int chunk_end = min(p.total_chunks, chunk_begin + per);
  const int nchunks = max(chunk_end - chunk_begin, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_s, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int box_rows = p.BH * p.BW;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- TMA producer
    int stage = 0, phase = 0;
    for (int ch = chunk_begin; ch < chunk_end; ++ch) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = stages + stage * STAGE_BYTES;
      // One chunk = one spatial box of NB consecutive images: every operand is ONE TMA box with a batch extent of NB
      // (rows land image-major, then box row, then pixel -- the same K order for dz and x), so the single producer thread
      // issues 3 (halo) / 2 + TAPS copies per chunk whatever the map size. Images past the batch are zero-filled.
      const int ig = ch / p.boxes_per_img;
      const int r = ch - ig * p.boxes_per_img;
      const int py = r / p.boxes_x;
      const int oy = py * p.BH, ox = (r - py * p.boxes_x) * p.BW, b = ig * p.NB;
      const bool two = m0 + 64 < p.cout;  // the second 64-channel block of dz exists (else its accumulator rows are never read)
      const uint32_t x_bytes = HALO ? p.NB * p.halo_tile_bytes : STAGE_BYTES - Cfg::A_BYTES;
      mbar_expect_tx(&full_bar[stage], (two ? Cfg::A_BYTES : Cfg::A_BYTES / 2) + x_bytes);
      tma_load_4d(&p.tmDz, &full_bar[stage], sa, m0, ox, oy, b);
      if (two) tma_load_4d(&p.tmDz, &full_bar[stage], sa + Cfg::P * 128, m0 + 64, ox, oy, b);
      if constexpr (HALO) {
        tma_load_4d(&p.tmX[0], &full_bar[stage], sa + Cfg::A_BYTES, c0, ox - 1, oy - 1, b);
      } else {
        constexpr int KW = TAPS == 9 ? 3 : 1;
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          const int kh = t / KW, kw = t - kh * KW;
          int dy, dx, view = 0;
          if (p.stride == 1) {
            dy = kh - p.pad;
            dx = kw - p.pad;
          } else {
            const int uy = kh - p.pad, ux = kw - p.pad;
            const int ph = uy & 1, pw = ux & 1;
            dy = (uy - ph) >> 1;
            dx = (ux - pw) >> 1;
            view = ph * 2 + pw;
          }
          uint8_t* sx = sa + Cfg::A_BYTES + t * Cfg::TAP_BYTES;
#pragma unroll
          for (int xb = 0; xb < Cfg::NXB; ++xb)
            tma_load_4d(&p.tmX[view], &full_bar[stage], sx + xb * Cfg::XBLK_BYTES, c0 + xb * XB, ox + dx, oy + dy, b);
        }
      }
      if (++stage == NSTAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    // instruction descriptor: bf16 x bf16 -> f32, A and B both MN-major (bits 15, 16), M = 128.
    // The taps' x tiles sit TAP_BYTES apart in shared memory and their accumulators N_T columns apart in TMEM, i.e. they
    // are consecutive N blocks of ONE MN-major operand (block stride LBO = XBLK_BYTES = TAP_BYTES / NXB): up to 256 / N_T
    // taps go into one instruction. For the 3x3 kernel (N_T = 32) that is N = 256 + N = 32 per K step instead of nine
    // N = 32 instructions, whose (128 + 32) x 32 B of operand reads per 16 tensor clocks made them shared-memory bound.
    constexpr int TAPS_PER_MMA = TAPS * N_T <= 256 ? TAPS : (256 / N_T >= 1 ? 256 / N_T : 1);
    constexpr int N_BIG = TAPS_PER_MMA * N_T;
    constexpr uint32_t idesc_big = make_idesc_bf16_f32(128, N_BIG) | (1u << 15) | (1u << 16);
    constexpr uint32_t idesc_one = make_idesc_bf16_f32(128, N_T) | (1u << 15) | (1u << 16);
    // halo form: the K slices' start offsets inside the chunk's halo tiles do not depend on the chunk -- compute them
    // once (this thread issues every MMA alone; integer divisions per slice made IT the bottleneck of the kernel)
    [[maybe_unused]] constexpr uint32_t idesc_row = make_idesc_bf16_f32(128, 3 * N_T) | (1u << 15) | (1u << 16);
    [[maybe_unused]] uint32_t halo_off[Cfg::P / 16];
    [[maybe_unused]] uint32_t halo_line = 0;
    [[maybe_unused]] uint64_t halo_desc_hi = 0;
    if constexpr (HALO) {
      halo_line = (p.BW + 2) * Cfg::XROW;                            // one halo line
      const uint32_t sbo = p.BW == 8 ? halo_line : 8 * Cfg::XROW;    // distance between the two 8-pixel groups of a K slice
      halo_desc_hi = (static_cast<uint64_t>(Cfg::XROW >> 4) << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46) |
                     ((Cfg::XROW == 128 ? 2ull : 4ull) << 61);
#pragma unroll
      for (int k = 0; k < Cfg::P / 16; ++k) {
        const int pix = 16 * k;                                       // first pixel (A row) of the slice, image-major order
        const int j = pix / box_rows, rem = pix - j * box_rows;
        const int y = rem / p.BW, x0 = rem - y * p.BW;
        halo_off[k] = j * p.halo_tile_bytes + (y * (p.BW + 2) + x0) * Cfg::XROW;
      }
    }
    int stage = 0, phase = 0;
    for (int i = 0; i < nchunks; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      const uint32_t a_addr = smem_u32(stages + stage * STAGE_BYTES);
      const uint32_t x_addr = a_addr + Cfg::A_BYTES;
      if constexpr (HALO) {
#pragma unroll
        for (int k = 0; k < Cfg::P / 16; ++k) {
          const uint32_t a = a_addr + k * 16 * 128;
          const uint64_t adesc = static_cast<uint64_t>((a & 0x3FFFF) >> 4) | (static_cast<uint64_t>((Cfg::P * 128) >> 4) << 16) |
                                 (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t bb = x_addr + halo_off[k] + kh * halo_line;
            const uint64_t bdesc = halo_desc_hi | static_cast<uint64_t>((bb & 0x3FFFF) >> 4);
            umma_f16_ss(tmem_base + kh * 3 * N_T, adesc, bdesc, idesc_row, (i | k) != 0 ? 1u : 0u);
          }
        }
      } else
#pragma unroll 1
      for (int t = 0; t < TAPS; t += (TAPS - t >= TAPS_PER_MMA ? TAPS_PER_MMA : 1)) {
        const bool big = TAPS - t >= TAPS_PER_MMA;
#pragma unroll
        for (int k = 0; k < Cfg::P / 16; ++k) {
          // MN-major descriptors: LBO = distance between channel blocks, SBO = 8 pixel rows
          uint64_t adesc = 0, bdesc = 0;
          {
            const uint32_t a = a_addr + k * 16 * 128;
            adesc = static_cast<uint64_t>((a & 0x3FFFF) >> 4) | (static_cast<uint64_t>((Cfg::P * 128) >> 4) << 16) |
                    (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            const uint32_t bb = x_addr + t * Cfg::TAP_BYTES + k * 16 * Cfg::XROW;
            bdesc = static_cast<uint64_t>((bb & 0x3FFFF) >> 4) |
                    (static_cast<uint64_t>(Cfg::XBLK_BYTES >> 4) << 16) |
                    (static_cast<uint64_t>((8 * Cfg::XROW) >> 4) << 32) | (1ull << 46) |
                    ((Cfg::XROW == 128 ? 2ull : 4ull) << 61);
          }
          umma_f16_ss(tmem_base + t * N_T, adesc, bdesc, big ? idesc_big : idesc_one, (i | k) != 0 ? 1u : 0u);
        }
      }
      umma_commit(&empty_bar[stage]);
      if (++stage == NSTAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
    umma_commit(done_bar);
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: TMEM -> red.add to dW
    if (nchunks > 0) {
      mbar_wait(done_bar, 0);
      tcgen05_fence_after();
      const int row = threadIdx.x - 128;  // accumulator lane == output channel m0 + row
      const int n = m0 + row;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      const int taps = TAPS;
#pragma unroll 1
      for (int t = 0; t < taps; ++t) {
#pragma unroll 1
        for (int cc = 0; cc < N_T; cc += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + t * N_T + cc, v);
          tmem_ld_wait();
          if (n < p.cout) {
            // 16-byte vector reductions (red.global.add.v4.f32): a thread owns 32 consecutive floats of ITS filter row, so
            // the lanes of a warp never share a sector -- four floats per L2 operation instead of one (cin % 8 == 0)
            float* dst = p.dw + ((size_t)n * taps + t) * p.cin + c0 + cc;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (c0 + cc + i < p.cin)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(__uint_as_float(v[i])),
                             "f"(__uint_as_float(v[i + 1])), "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                             : "memory");
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map4(CUtensorMap* tm, const void* base, int C, int W, int H, int B, int64_t sW, int64_t sH, int64_t sB,
                       int boxc, int bw, int bh, int nb) {
  static PFN_encodeTiledW enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled entry point not available");
      return AY2_ERR_CUDA;
    }
    enc = reinterpret_cast<PFN_encodeTiledW>(ptr);
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sB * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = boxc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(wgrad C=%d W=%d H=%d B=%d box=%d,%d,%d) -> %d", C, W, H, B, boxc, bw, bh, (int)r);
    return AY2_ERR_CUDA;
  }
  return AY2_OK;
}

template <int N_T, int XB, int TAPS, bool HALO = false>
static int launch_wgrad(const WgradParams& kp, int grid, cudaStream_t st) {
  using Cfg = WgradCfg<N_T, XB>;
  constexpr int STAGE_BYTES = Cfg::A_BYTES + (HALO ? kWgradHaloMax : TAPS * Cfg::TAP_BYTES);
  constexpr int NSTAGES = (220 * 1024) / STAGE_BYTES >= 4 ? 4 : (220 * 1024) / STAGE_BYTES >= 3 ? 3 : 2;
  constexpr size_t smem = 1024 + (size_t)NSTAGES * STAGE_BYTES + 256;
  static DeviceOnce once;
  if (once.first())
    AY2_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<N_T, XB, TAPS, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv_wgrad_kernel<N_T, XB, TAPS, HALO><<<grid, 256, smem, st>>>(kp);
  AY2_CHECK_LAUNCH();
  return AY2_OK;
}

static void pick_box_w(int H, int W, int* bh, int* bw, int min_w = 1) {
  static const int cand[][2] = {{8, 16}, {16, 8}, {4, 32}, {8, 8}, {4, 16}, {16, 4}, {2, 32}, {4, 8}, {8, 4}, {2, 16}, {4, 4}, {2, 8}, {1, 16}};
  double best = 1e30;
  int bi = 0;
  for (int i = 0; i < (int)(sizeof(cand) / sizeof(cand[0])); ++i) {
    const int h = cand[i][0], w = cand[i][1];
    if (w < min_w) continue;
    const double cover = (double)ceil_div(H, h) * h * (double)ceil_div(W, w) * w;
    const double cost = cover * (1.0 + 0.002 * (128 / (h * w)));
    if (cost < best) {
      best = cost;
      bi = i;
    }
  }
  *bh = cand[bi][0];
  *bw = cand[bi][1];
}

}  // namespace ay2

using namespace ay2;

// x: bf16 NHWC input of the forward conv (channel slice, stride in_cstride); dz: bf16 NHWC gradient of the conv output
// (before BN/activation); dw: fp32 [cout][kh*kw][cin], ACCUMULATED into (zero it for a fresh gradient).
extern "C" int ay2_conv_wgrad(const ay2_conv_desc* d, const void* x, const void* dz, float* dw, void* stream) {
  AY2_REQUIRE(d && x && dz && dw, "ay2_conv_wgrad: null pointer");
  AY2_REQUIRE(d->stride == 1 || d->stride == 2, "wgrad stride %d unsupported", d->stride);
  AY2_REQUIRE(d->in_cstride % 8 == 0 && d->out_cstride % 8 == 0 && d->cin % 8 == 0 && d->cout % 8 == 0,
              "wgrad: channel counts / strides must be multiples of 8");
  AY2_REQUIRE(d->in_pix_stride <= 0 && d->pad_w < 0, "wgrad: windowed inputs are not supported");
  AY2_REQUIRE(d->in_row_pixels <= 0 || d->stride == 1, "wgrad: padded rows need stride 1");
  if (d->stride == 2) AY2_REQUIRE(d->in_h % 2 == 0 && d->in_w % 2 == 0, "wgrad stride 2 needs even input size");
  const int taps = d->kh * d->kw;
  AY2_REQUIRE(taps == 1 || (d->kh == 3 && d->kw == 3), "wgrad supports 1x1 and 3x3 kernels (got %dx%d)", d->kh, d->kw);
  WgradParams kp;
  memset(&kp, 0, sizeof(kp));
  int bh, bw;
  pick_box_w(d->out_h, d->out_w, &bh, &bw);
  // 3x3 / stride 1 / pad 1: one halo tile per box instead of nine shifted x tiles (boxes at least 8 pixels wide, so that
  // the 8-row groups of the MN-major operand are horizontally adjacent pixels)
  static const int env_halo = [] { const char* e = getenv("AY2_WGRAD_HALO"); return e ? atoi(e) : 1; }();
  bool halo = false;
  if (env_halo && taps == 9 && d->stride == 1 && d->pad == 1) {
    int hh, hw;
    pick_box_w(d->out_h, d->out_w, &hh, &hw, 8);
    const int tile = (hh + 2) * (hw + 2) * 64;  // XB = 32 channels = 64-byte rows; the NB tiles of a box are contiguous
    if ((128 / (hh * hw)) * tile <= kWgradHaloMax) {
      halo = true;
      bh = hh;
      bw = hw;
      kp.halo_tile_bytes = tile;
    }
  }
  kp.BH = bh;
  kp.BW = bw;
  kp.NB = 128 / (bh * bw);
  kp.boxes_x = ceil_div(d->out_w, bw);
  kp.boxes_per_img = kp.boxes_x * ceil_div(d->out_h, bh);
  kp.total_chunks = kp.boxes_per_img * ceil_div(d->batch, kp.NB);  // chunk = (group of NB images, spatial box)
  kp.kh = d->kh;
  kp.kw = d->kw;
  kp.stride = d->stride;
  kp.pad = d->pad;
  kp.cout = d->cout;
  kp.cin = d->cin;
  kp.dw = dw;
  // tiling: 3x3 -> 32 input channels per CTA (9 x 32 accumulator columns); 1x1 -> up to 256
  int n_t, xb;
  if (taps == 9) {
    n_t = 32;
    xb = 32;
  } else {
    n_t = d->cin >= 256 ? 256 : (d->cin >= 128 ? 128 : (d->cin >= 64 ? 64 : 32));
    xb = n_t >= 64 ? 64 : 32;
  }
  kp.num_m_tiles = ceil_div(d->cout, 128);
  kp.num_n_tiles = ceil_div(d->cin, n_t);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ksplit = sms / (kp.num_m_tiles * kp.num_n_tiles);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > kp.total_chunks) ksplit = kp.total_chunks;
  kp.ksplit = ksplit;
  int rc = encode_map4(&kp.tmDz, dz, d->cout, d->out_w, d->out_h, d->batch, d->out_cstride, (int64_t)d->out_cstride * d->out_w,
                       (int64_t)d->out_cstride * d->out_w * d->out_h, 64, bw, bh, kp.NB);
  const int64_t cs = d->in_cstride;
  if (rc == AY2_OK) {
    if (d->stride == 1) {
      const int64_t rowp = d->in_row_pixels > 0 ? d->in_row_pixels : d->in_w;
      rc = encode_map4(&kp.tmX[0], x, d->cin, d->in_w, d->in_h, d->batch, cs, cs * rowp, cs * rowp * d->in_h, xb,
                       halo ? bw + 2 : bw, halo ? bh + 2 : bh, kp.NB);
    } else {
      for (int ph = 0; ph < 2 && rc == AY2_OK; ++ph)
        for (int pw = 0; pw < 2 && rc == AY2_OK; ++pw) {
          const uint8_t* base = static_cast<const uint8_t*>(x) + ((int64_t)ph * d->in_w + pw) * cs * 2;
          rc = encode_map4(&kp.tmX[ph * 2 + pw], base, d->cin, d->in_w / 2, d->in_h / 2, d->batch, 2 * cs, 2 * cs * d->in_w,
                           cs * d->in_w * d->in_h, xb, bw, bh, kp.NB);
        }
    }
  }
  if (rc != AY2_OK) return rc;
  const int grid = kp.num_m_tiles * kp.num_n_tiles * ksplit;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (taps == 9 && halo) rc = launch_wgrad<32, 32, 9, true>(kp, grid, st);
  else if (taps == 9) rc = launch_wgrad<32, 32, 9>(kp, grid, st);
  else if (n_t == 256) rc = launch_wgrad<256, 64, 1>(kp, grid, st);
  else if (n_t == 128) rc = launch_wgrad<128, 64, 1>(kp, grid, st);
  else if (n_t == 64) rc = launch_wgrad<64, 64, 1>(kp, grid, st);
  else rc = launch_wgrad<32, 32, 1>(kp, grid, st);
  if (rc != AY2_OK) return rc;
  count_launch();
  return AY2_OK;
}
