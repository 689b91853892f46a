"""Per-phase timeline of conv_tc_kernel over all CTAs: python tools/conv_timeline.py CIN COUT K S HW_OUT [B]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ayolov2_b200 import _lib, ops  # noqa: E402

NAMES = ["entry", "prologue done", "first operands landed", "first acc complete", "first store issued",
         "last acc complete", "last store issued", "stores drained", "roles finished", "(clock)",
         "  t0: cols 0-31 done", "  t0: cols 32-63 done", "  t0: thread drained", "  t0: fences done", "  t0: epilogue barrier"]


def timeline(cin, cout, k, s, hw, B=64, act=1):
    x = ops.new_act(B, hw * s, hw * s, cin); x.buf.normal_()
    y = ops.new_act(B, hw, hw, cout)
    w, b = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") / (k * k * cin) ** 0.5, torch.zeros(cout, device="cuda"))
    plan = ops.ConvPlan(x, y, w, b, k, k, s, k // 2, act)
    info = (C.c_int32 * 4)()
    lib = _lib.load()
    lib.ay2_conv_plan_set_debug(plan._h, None, info)
    grid = info[0]
    # realistic state: clocks ramped by a long burst, the input freshly written by a preceding kernel (L2-resident
    # when it fits), this kernel's own output not in L2
    xs = ops.new_act(B, hw * s, hw * s, cin); xs.buf.normal_()
    wi, bi = ops.pack_conv_weight(torch.eye(cin, device="cuda").view(cin, cin, 1, 1), torch.zeros(cin, device="cuda"))
    prev = ops.ConvPlan(xs, x, wi, bi, 1, 1, 1, 0, _lib.ACT_NONE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(300):
        prev.run(); plan.run()
    e0.record()
    for _ in range(20):
        prev.run(); plan.run()
    e1.record(); torch.cuda.synchronize()
    ms_pair = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20):
        prev.run()
    e1.record(); torch.cuda.synchronize()
    ms_plain = ms_pair - e0.elapsed_time(e1) / 20
    dbg = torch.zeros(grid * 16, dtype=torch.int64, device="cuda")
    lib.ay2_conv_plan_set_debug(plan._h, dbg.data_ptr(), info)
    for _ in range(50):
        prev.run(); plan.run()
    e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
    d = dbg.cpu().view(grid, 16)
    t0 = int(d[:, 0].min())
    mhz = ((d[:, 15] - d[:, 9]).double() / (d[:, 8] - d[:, 0]).double()).median() * 1000.0
    print(f"conv {cin}->{cout} k{k} s{s} out {hw}x{hw} B{B}: grid {grid} ({info[1]}/SM) N tile {info[2]} tiles {info[3]} "
          f"({info[3] / grid:.2f}/CTA)  kernel {ms_plain * 1000:.1f} us back to back (pair - producer), SM clock {mhz:.0f} MHz")
    for i, n in enumerate(NAMES):
        if i == 9 or i == 15:
            continue
        col = d[:, i]
        col = col[col > 0] - t0
        if col.numel() == 0:
            continue
        q = torch.quantile(col.double(), torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64)) / 1000.0
        print(f"   {n:24s} min {q[0]:7.2f}  med {q[1]:7.2f}  max {q[2]:7.2f} us")
    lib.ay2_conv_plan_set_debug(plan._h, None, info)


if __name__ == "__main__" and os.environ.get("AY2_TIMELINE_GRAPH") != "1":
    a = [int(v) for v in sys.argv[1:]]
    timeline(*a)


def graph_gaps(cin, cout, k, s, hw, B=64):
    """Three launches (identity 1x1 producer -> conv -> identity 1x1 consumer) inside a CUDA graph: idle gaps between
    the last CTA of one kernel finishing and the first CTA of the next one starting."""
    lib = _lib.load()
    xs = ops.new_act(B, hw * s, hw * s, cin); xs.buf.normal_()
    x = ops.new_act(B, hw * s, hw * s, cin)
    y = ops.new_act(B, hw, hw, cout)
    z = ops.new_act(B, hw, hw, cout)
    wi, bi = ops.pack_conv_weight(torch.eye(cin, device="cuda").view(cin, cin, 1, 1), torch.zeros(cin, device="cuda"))
    wo, bo = ops.pack_conv_weight(torch.eye(cout, device="cuda").view(cout, cout, 1, 1), torch.zeros(cout, device="cuda"))
    w, b = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") / (k * k * cin) ** 0.5, torch.zeros(cout, device="cuda"))
    plans = [ops.ConvPlan(xs, x, wi, bi, 1, 1, 1, 0, _lib.ACT_NONE), ops.ConvPlan(x, y, w, b, k, k, s, k // 2, _lib.ACT_SILU),
             ops.ConvPlan(y, z, wo, bo, 1, 1, 1, 0, _lib.ACT_NONE)]
    info = (C.c_int32 * 4)()
    dbgs = []
    for p in plans:
        lib.ay2_conv_plan_set_debug(p._h, None, info)
        dbgs.append(torch.zeros(info[0] * 16, dtype=torch.int64, device="cuda"))
        lib.ay2_conv_plan_set_debug(p._h, dbgs[-1].data_ptr(), info)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            for p in plans:
                p.run()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(8):
                for p in plans:
                    p.run()
        for _ in range(30):
            g.replay()
        st.synchronize()
    d = [t.cpu().view(-1, 16) for t in dbgs]
    t0 = int(d[0][:, 0].min())
    ev = []
    for i, n in enumerate(["producer", "conv", "consumer"]):
        ev.append((int(d[i][:, 0].min()) - t0, int(d[i][:, 0].max()) - t0, int(d[i][:, 8].median()) - t0, int(d[i][:, 8].max()) - t0))
        print(f"   {n:9s} first CTA in {ev[-1][0] / 1000:7.2f}  last CTA in {ev[-1][1] / 1000:7.2f}  median CTA out {ev[-1][2] / 1000:7.2f}  last CTA out {ev[-1][3] / 1000:7.2f} us")
    print(f"   idle gap producer->conv {(ev[1][0] - ev[0][3]) / 1000:.2f} us, conv->consumer {(ev[2][0] - ev[1][3]) / 1000:.2f} us "
          f"(PDL {'on' if os.environ.get('AY2_CONV_PDL') == '1' else 'off'})")


if __name__ == "__main__" and os.environ.get("AY2_TIMELINE_GRAPH") == "1":
    graph_gaps(*[int(v) for v in sys.argv[1:]])
