"""A/B timing of single conv plans across builds of libay2.so: `python tools/conv_ab.py lib1.so lib2.so ...`.
Each library is loaded with ctypes (no source-hash check: these are deliberately other builds) and timed on the same shapes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ayolov2_b200._lib import ConvDesc  # noqa: E402

SHAPES = [  # name, B, H, W, cin, cout, k, s, p, split
    ("L1 32->64 3x3/s2 @320", 64, 320, 320, 32, 64, 3, 2, 1, 0),
    ("L4 (32|32)->64 1x1 @160 two-source", 64, 160, 160, 64, 64, 1, 1, 0, 32),
    ("L2 64->64 1x1 @160", 64, 160, 160, 64, 64, 1, 1, 0, 0),
    ("L34 256->128 1x1 @80", 64, 80, 80, 256, 128, 1, 1, 0, 0),
    ("L13 256->256 1x1 @40", 64, 40, 40, 256, 256, 1, 1, 0, 0),
    ("L12 128->256 3x3/s2 @80", 64, 80, 80, 128, 256, 3, 2, 1, 0),
]


def bench(libpath):
    lib = C.CDLL(libpath)
    lib.ay2_conv_block_n.restype = C.c_int
    lib.ay2_conv_plan_create.restype = C.c_int
    lib.ay2_conv_plan_create.argtypes = [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.ay2_conv_plan_run.argtypes = [C.c_void_p, C.c_void_p]
    lib.ay2_last_error_string.restype = C.c_char_p
    out = []
    for name, B, H, W, cin, cout, k, s, p, split in SHAPES:
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        x = torch.randn((B, H, W, cin if not split else split), device="cuda").to(torch.bfloat16)
        x2 = torch.randn((B, H, W, cin - split), device="cuda").to(torch.bfloat16) if split else None
        y = torch.zeros((B, OH, OW, cout), device="cuda", dtype=torch.bfloat16)
        bn = lib.ay2_conv_block_n(cout)
        cout_pad = (cout + bn - 1) // bn * bn
        w = (torch.randn((cout_pad, k * k * cin), device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.zeros(cout_pad, device="cuda")
        d = ConvDesc()
        d.batch, d.in_h, d.in_w, d.cin, d.in_cstride = B, H, W, cin, x.shape[3]
        d.out_h, d.out_w, d.cout, d.out_cstride = OH, OW, cout, cout
        d.kh = d.kw = k
        d.stride, d.pad, d.act, d.cout_pad, d.pad_w = s, p, 1, cout_pad, -1
        if split:
            d.cin_split, d.in2_cstride, d.in2 = split, x2.shape[3], x2.data_ptr()
        h = C.c_void_p()
        rc = lib.ay2_conv_plan_create(C.byref(d), x.data_ptr(), w.data_ptr(), bias.data_ptr(), None, y.data_ptr(), C.byref(h))
        assert rc == 0, lib.ay2_last_error_string()
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            lib.ay2_conv_plan_run(h, st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            lib.ay2_conv_plan_run(h, st)
        b.record()
        torch.cuda.synchronize()
        out.append((name, a.elapsed_time(b) / 20 * 1e3))
    return out


if __name__ == "__main__":
    res = {p: bench(p) for p in sys.argv[1:]}
    for i, (name, *_r) in enumerate(SHAPES):
        print(f"{name:40s} " + "  ".join(f"{os.path.basename(p)}: {res[p][i][1]:7.1f} us" for p in sys.argv[1:]))
