// Host-side helpers shared by the C-ABI translation units: error string, CUDA checks, launch counter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ay2.h"

namespace ay2 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define AY2_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::ay2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AY2_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define AY2_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::ay2::set_error(__VA_ARGS__);  \
      return AY2_ERR_INVALID;         \
    }                                 \
  } while (0)

// Launch-error check that does not synchronise.
#define AY2_CHECK_LAUNCH()                                                                    \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      ::ay2::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return AY2_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute (e.g. the opt-in to > 48 KB of dynamic shared memory) is per (function, DEVICE): call sites keep one
// flag per device, `static DeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`, not one flag per process.
struct DeviceOnce {
  bool done[64] = {};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const bool f = !done[dev];
    done[dev] = true;
    return f;
  }
};

// TMA tensor maps (conv_tc.cu). Activation: NHWC bf16 view [C][W][H][B] with element strides, box [boxc][bw][bh][1],
// swizzle span = boxc*2 bytes. Weights: bf16 [rows][Ktot] (K innermost), box [boxk][boxn].
int encode_act_map(CUtensorMap* tm, const void* base, int C, int W, int H, int B, int64_t sW, int64_t sH, int64_t sB,
                   int boxc, int bw, int bh);
int encode_weight_map(CUtensorMap* tm, const void* base, int Ktot, int rows, int boxk, int boxn);

// NMS workspace layout (nms.cu), shared with the head convolution whose epilogue appends candidates to it (conv_tc.cu)
struct NmsWorkspaceView {
  int* counts;               // [batch] candidates per image
  int* overflow;             // [1]
  int* row_counts;           // [batch] rows that passed the objectness test (two-phase generation only)
  unsigned long long* keys;  // [batch][key_stride]
  unsigned long long* keys2; // [batch][key_stride] compaction scratch of the dense-slot mode (more candidates than the shared-memory sort holds)
  long long key_stride;
  unsigned* rows;            // [batch][n]
  // class-group split of the suppression kernel (nms.cu): per (image, group) kept lists + the per-image ticket / fallback flag
  int* done;                 // [batch]
  int* wide;                 // [batch]
  unsigned long long* part_keys;  // [batch][groups][max_det]
  float* part_det;                // [batch][groups][max_det][6]
  int* part_count;                // [batch][groups]
};
NmsWorkspaceView nms_workspace_view(const ay2_nms_params* p, void* workspace);

// Candidate generation fused into the detect-head convolution epilogue (conv_tc.cu); keys == nullptr: off.
struct HeadCandParams {
  unsigned long long* keys;
  int* counts;
  const uint8_t* class_mask;
  long long key_stride;
  float conf_thres;
  int max_candidates, na, no, row_off, multi_label, batch, out_h, out_w;
  // 1: a candidate's key is stored at slot [image][row] of the key array (pre-filled with ~0 by ay2_nms_candidates_begin) --
  // no counters, no atomics; the suppression kernel compacts the slots while loading. Single-label lists with room for
  // every row (nms_dense_slots()).
  int dense_slots;
};
// The fused head uses row-indexed key slots (above) when this holds.
inline bool nms_dense_slots(const ay2_nms_params* p) { return !p->multi_label && p->max_candidates >= p->n; }

}  // namespace ay2
