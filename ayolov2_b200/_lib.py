"""ctypes binding of libay2.so (include/ay2.h). Plain pointers and sizes only; PyTorch owns all memory.

There is no CPU fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

_LIB_PATH = Path(__file__).resolve().parent / "libay2.so"
_lib: Optional[C.CDLL] = None

ACT_NONE, ACT_SILU = 0, 1
DT_U8, DT_F32 = 0, 1


class ConvDesc(C.Structure):
    """ay2_conv_desc (include/ay2.h)."""

    _fields_ = [
        ("batch", C.c_int32),
        ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("cin", C.c_int32), ("in_cstride", C.c_int32),
        ("out_h", C.c_int32), ("out_w", C.c_int32),
        ("cout", C.c_int32), ("out_cstride", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32),
        ("stride", C.c_int32), ("pad", C.c_int32),
        ("act", C.c_int32), ("res_cstride", C.c_int32),
        ("cout_pad", C.c_int32), ("pad_w", C.c_int32),
        ("in_pix_stride", C.c_int32), ("in_row_pixels", C.c_int32),
        ("out_pix_stride", C.c_int32), ("out_row_pixels", C.c_int32),
        ("cin_split", C.c_int32), ("in2_cstride", C.c_int32), ("x3", C.c_int32), ("stride_w", C.c_int32),
        ("in2", C.c_void_p),
    ]


class ChainDesc(C.Structure):
    """ay2_chain_desc (include/ay2.h)."""

    _fields_ = [
        ("batch", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32),
        ("cin", C.c_int32), ("in_cstride", C.c_int32),
        ("c1", C.c_int32), ("act1", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("c2", C.c_int32), ("act2", C.c_int32),
        ("c3", C.c_int32), ("act3", C.c_int32),
        ("out_cstride", C.c_int32), ("res_cstride", C.c_int32),
    ]


class NmsParams(C.Structure):
    """ay2_nms_params (include/ay2.h)."""

    _fields_ = [
        ("iou_thres", C.c_double),
        ("conf_thres", C.c_float), ("max_wh", C.c_float),
        ("batch", C.c_int32), ("n", C.c_int32), ("no", C.c_int32),
        ("multi_label", C.c_int32), ("agnostic", C.c_int32),
        ("max_det", C.c_int32), ("max_nms", C.c_int32), ("max_candidates", C.c_int32),
    ]


class HeadLevels(C.Structure):
    """ay2_head_levels (include/ay2.h)."""

    _fields_ = [
        ("nl", C.c_int32), ("na", C.c_int32),
        ("logits", C.c_void_p * 5),
        ("ny", C.c_int32 * 5), ("nx", C.c_int32 * 5), ("cstride", C.c_int32 * 5),
        ("stride_px", C.c_float * 5),
        ("anchor_px", (C.c_float * 2) * 8 * 5),
    ]


class LossParams(C.Structure):
    """ay2_loss_params (include/ay2.h)."""

    _fields_ = [
        ("nl", C.c_int32), ("na", C.c_int32), ("nc", C.c_int32), ("bs", C.c_int32), ("nt", C.c_int32),
        ("ny", C.c_int32 * 5), ("nx", C.c_int32 * 5),
        ("balance", C.c_float * 5),
        ("anchor_t", C.c_float), ("box", C.c_float), ("obj", C.c_float), ("cls", C.c_float),
        ("cls_pw", C.c_float), ("obj_pw", C.c_float), ("cp", C.c_float), ("cn", C.c_float),
        ("fl_gamma", C.c_float), ("fl_alpha", C.c_float),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/ay2.h
_PROTOS = {
    "ay2_version": (C.c_int, []),
    "ay2_last_error_string": (C.c_char_p, []),
    "ay2_source_hash": (C.c_char_p, []),
    "ay2_launch_count": (C.c_int64, []),
    "ay2_conv_block_n": (C.c_int, [C.c_int32]),
    "ay2_conv_plan_create": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_void_p)]),
    "ay2_conv_plan_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ay2_conv_plan_destroy": (C.c_int, [C.c_void_p]),
    "ay2_conv_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "ay2_conv_plan_flops": (C.c_double, [C.c_void_p]),
    "ay2_conv_reference_simt": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "ay2_chain_supported": (C.c_int, [C.POINTER(ChainDesc)]),
    "ay2_chain_plan_create": (C.c_int, [C.POINTER(ChainDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ay2_chain_plan_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ay2_chain_plan_destroy": (C.c_int, [C.c_void_p]),
    "ay2_chain_plan_flops": (C.c_double, [C.c_void_p]),
    "ay2_chain_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "ay2_chain_plan_set_debug": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ay2_conv_plan_set_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "ay2_space_to_depth": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "ay2_sppf_pool": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_upsample2x": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_int32, C.c_void_p]),
    "ay2_head_decode": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ay2_space_to_depth_x3": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_void_p]),
    "ay2_sppf_pool_x3": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_head_decode2": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ay2_yolo_loss_workspace_bytes": (C.c_size_t, [C.POINTER(LossParams)]),
    "ay2_yolo_loss": (C.c_int, [C.POINTER(LossParams), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ay2_bn_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_bn_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_bn_act_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "ay2_bn_act_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ay2_bn_act_bwd_grads": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p]),
    "ay2_bn_act_bwd_phase": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int64, C.c_void_p]),
    "ay2_conv_wgrad": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_add_slices": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "ay2_upsample2x_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_void_p]),
    "ay2_maxpool_bwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "ay2_head_grad_to_nhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_void_p]),
    "ay2_head_logits_to_train": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p]),
    "ay2_channel_sum": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ay2_sgd_ema_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                                   C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "ay2_repack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "ay2_sgd_ema_step_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_int32, C.c_float,
                                          C.c_float, C.c_void_p]),
    "ay2_letterbox_collate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "ay2_collate_labels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "ay2_pseudo_labels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_resize_bilinear": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_void_p]),
    "ay2_load_resize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "ay2_nms_workspace_bytes": (C.c_size_t, [C.POINTER(NmsParams)]),
    "ay2_nms_from_logits": (C.c_int, [C.POINTER(HeadLevels), C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_nms_batched": (C.c_int, [C.c_void_p, C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_conv_plan_set_head_candidates": (C.c_int, [C.c_void_p, C.POINTER(NmsParams), C.c_int32, C.c_int32, C.c_void_p,
                                                    C.c_void_p, C.c_size_t]),
    "ay2_box_iou": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ay2_match_detections": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ay2_nms_boxes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_nms_candidate_table": (C.c_int, [C.c_void_p, C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_nms_fast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                               C.c_int32, C.c_void_p, C.c_void_p]),
    "ay2_nms_matrix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_void_p, C.c_void_p]),
    "ay2_nms_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ay2_nms_batched_scaled": (C.c_int, [C.c_void_p, C.POINTER(NmsParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ay2_nms_candidates_begin": (C.c_int, [C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ay2_nms_from_candidates": (C.c_int, [C.POINTER(HeadLevels), C.POINTER(NmsParams), C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libay2.so (once). Raises if it has not been built (`python -m ayolov2_b200._build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "ayolov2_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    # A binary built from other sources than the ones next to it (stale after a checkout) must not be bound silently:
    # struct layouts and kernel semantics are only guaranteed for the matching include/ay2.h + csrc/.
    from . import _build

    have, want = lib.ay2_source_hash().decode(), _build._source_hash()
    if have != want:
        raise RuntimeError(f"{_LIB_PATH} was built from different sources (embedded {have[:12]}, tree {want[:12]}): rebuild it "
                           "with `python -c 'import __graft_entry__ as g; g.build()'`")
    _lib = lib
    return lib


def exported_symbols() -> list:
    return sorted(_PROTOS)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ay2_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"libay2 {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().ay2_launch_count())


def ptr(t) -> int:
    """Device address of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream
