"""The training step's heavy kernels one by one at the real yolov5s layer shapes (batch 128, 640x640): CUDA-event time,
achieved GB/s (streaming BatchNorm kernels) or TFLOP/s (weight gradients) per launch. GPU box only.
  python tools/prof_trainops.py            all ops, timing table
  python tools/prof_trainops.py wgrad      only the weight-gradient launches (e.g. under `ncu -k regex:conv_wgrad`)
  python tools/prof_trainops.py bn         only the BatchNorm kernels"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ayolov2_b200 import ops  # noqa: E402
from ayolov2_b200.ops import ActView  # noqa: E402

B = int(os.environ.get("AY2_PROF_BATCH", "128"))
what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
torch.cuda.profiler.start()  # for `ncu --profile-from-start off`


def act(H, W, C, cs=None, c0=0):
    cs = cs or C
    buf = (torch.randn((B, H, W, cs), device=dev) * 0.5).to(torch.bfloat16)
    return ActView(buf, c0, C)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2] * 1e3  # us


if what in ("all", "bn"):
    print(f"BatchNorm streaming kernels, batch {B} (bytes = algorithmic: every tensor once)")
    for (H, C, cs) in [(320, 32, 32), (160, 64, 64), (160, 32, 64), (80, 128, 128), (80, 64, 128), (40, 256, 256), (40, 128, 256), (20, 512, 512)]:
        z, y, gy, gz = act(H, H, C, cs), act(H, H, C, cs), act(H, H, C, cs), act(H, H, C, cs)
        mean, invstd = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        scratch = torch.zeros(2 * C, dtype=torch.float64, device=dev)
        nbytes = B * H * H * C * 2
        lib, st = ops._lib.load(), ops._lib.current_stream_ptr()
        rows = []
        t = timeit(lambda: lib.ay2_bn_stats(z.ptr(), B * H * H, C, cs, scratch.data_ptr(), scratch.data_ptr() + 8 * C, st))
        rows.append(("stats", t, nbytes))
        t = timeit(lambda: ops.bn_act_fwd(z, mean, invstd, gamma, beta, 1, y, None))
        rows.append(("fwd", t, 2 * nbytes))
        args = (gy.ptr(), cs, z.ptr(), cs, B * H * H, C, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1,
                scratch.data_ptr(), scratch.data_ptr() + 8 * C, gz.ptr(), cs)
        t = timeit(lambda: lib.ay2_bn_act_bwd_phase(*args, 1, B * H * H, st))
        rows.append(("bwd_reduce", t, 2 * nbytes))
        t = timeit(lambda: lib.ay2_bn_act_bwd_phase(*args, 2, B * H * H, st))
        rows.append(("bwd_apply", t, 3 * nbytes))
        print(f"  {H}x{H} C={C} (stride {cs}): " + "; ".join(f"{n} {t:.1f} us {b / t / 1e3:.0f} GB/s" for n, t, b in rows))
        del z, y, gy, gz

if what in ("all", "wgrad"):
    print(f"weight gradients, batch {B}")
    for (H, cin, cout, k, s) in [(320, 16, 32, 3, 1), (160, 32, 32, 3, 1), (80, 64, 64, 3, 1), (40, 128, 128, 3, 1), (20, 256, 256, 3, 1),
                                 (320, 32, 64, 3, 2), (160, 64, 128, 3, 2), (80, 128, 256, 3, 2), (40, 256, 512, 3, 2),
                                 (160, 64, 64, 1, 1), (80, 128, 128, 1, 1), (40, 256, 256, 1, 1), (20, 512, 512, 1, 1)]:
        OH = (H + 2 * (k // 2) - k) // s + 1
        x, dz = act(H, H, cin), act(OH, OH, cout)
        dw = torch.zeros((cout, k * k * cin), device=dev)
        t = timeit(lambda: ops.conv_wgrad(x, dz, dw, k, k, s, k // 2))
        fl = 2.0 * B * OH * OH * cout * cin * k * k
        print(f"  {H}x{H} {cin}->{cout} k{k} s{s}: {t:.1f} us, {fl / t / 1e6:.0f} TFLOP/s, reads {(x.buf.numel() + dz.buf.numel()) * 2 / t / 1e3:.0f} GB/s")
        del x, dz
