"""GPU bring-up driver: runs each conv parity case in its own subprocess under a timeout (a deadlocked
mbarrier pipeline must not hang the box), then the other GPU test files. Writes gpurun_out/bringup.log."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def one_case(i: int) -> None:
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_conv_gpu as T

    case = T.CASES[i]
    try:
        T._run_case(case)
        print(f"CASE {i} {case}: PASS", flush=True)
    except AssertionError as e:
        print(f"CASE {i} {case}: FAIL {str(e)[:400]}", flush=True)
        diagnose(case)


def diagnose(case) -> None:
    """Print where the tensor-core result deviates from the SIMT kernel."""
    import torch
    from ayolov2_b200 import ops

    B, H, W, Cin, Cout, k, s, p, act, use_res, in_slice, out_slice = case
    g = torch.Generator(device="cuda").manual_seed(0)
    OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = ops.ActView(torch.randn((B, H, W, Cin), device="cuda", generator=g).to(torch.bfloat16), 0, Cin)
    cout8 = (Cout + 7) // 8 * 8
    y = ops.ActView(torch.zeros((B, OH, OW, cout8), device="cuda", dtype=torch.bfloat16), 0, Cout)
    w = torch.randn((Cout, Cin, k, k), device="cuda", generator=g) * (1.0 / (Cin * k * k) ** 0.5)
    wp, bp = ops.pack_conv_weight(w, None)
    plan = ops.ConvPlan(x, y, wp, bp, k, k, s, p, 0)
    plan.run()
    torch.cuda.synchronize()
    got = y.tensor().float().clone()
    plan.run_reference_simt()
    torch.cuda.synchronize()
    ref = y.tensor().float()
    bad = (got - ref).abs() > (2.0 ** -6 * ref.abs() + 2e-2)
    print(f"  diag(no act/res/slices): bad {int(bad.sum())}/{bad.numel()}  max|got| {float(got.abs().max()):.3f} "
          f"max|ref| {float(ref.abs().max()):.3f}")
    if bad.any():
        print("  bad per image:", bad.sum((1, 2, 3)).tolist())
        print("  bad per out-row (img 0):", bad[0].sum((1, 2)).tolist()[:40])
        print("  bad per out-col (img 0):", bad[0].sum((0, 2)).tolist()[:40])
        print("  bad per channel (first 64):", bad.sum((0, 1, 2)).tolist()[:64])
        idx = bad.nonzero()[:5].tolist()
        for b_, y_, x_, c_ in idx:
            print(f"   at b={b_} y={y_} x={x_} c={c_}: got {float(got[b_, y_, x_, c_]):.4f} ref {float(ref[b_, y_, x_, c_]):.4f}")


def main() -> None:
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        one_case(int(sys.argv[2]))
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    log = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "w")

    def emit(s):
        print(s, flush=True)
        log.write(s + "\n")
        log.flush()

    # count cases robustly by importing without torch usage
    src = open(os.path.join(ROOT, "tests", "test_conv_gpu.py")).read()
    ncases = src.split("CASES = [")[1].split("]\n")[0].count("),")
    emit(f"{ncases} conv cases")
    for i in range(ncases):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", str(i)], stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True, timeout=180)
            out = r.stdout.strip().splitlines()
            tail = [l for l in out if l.startswith(("CASE", "  ", "   "))] or out[-8:]
            emit("\n".join(tail) + f"   [rc={r.returncode}, {time.time() - t0:.0f}s]")
            if r.returncode != 0:
                emit("\n".join(out[-15:]))
        except subprocess.TimeoutExpired as e:
            emit(f"CASE {i}: TIMEOUT (hang) after 180s; partial output: {(e.stdout or '')[-500:]}")
    for f in ("tests/test_pointwise_gpu.py", "tests/test_nms_gpu.py"):
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", f, "-q", "-m", "gpu", "--no-header", "-rf"],
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
            emit(f"== {f} rc={r.returncode}\n" + "\n".join(r.stdout.strip().splitlines()[-40:]))
        except subprocess.TimeoutExpired as e:
            emit(f"== {f}: TIMEOUT; partial: {(e.stdout or '')[-1500:]}")


if __name__ == "__main__":
    main()
