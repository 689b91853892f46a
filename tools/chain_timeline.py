"""Per-phase timeline of the fused chain kernel (CTA 0, first tiles): python tools/chain_timeline.py C HW [C3]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ayolov2_b200 import _lib, ops  # noqa: E402

c = int(sys.argv[1]); hw = int(sys.argv[2]); c3 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
B = 64
x = ops.new_act(B, hw, hw, c); x.buf.normal_()
y = ops.new_act(B, hw, hw, c3 or c)
links = [ops.pack_chain_weight(torch.randn(c, c, 1, 1, device="cuda") / c ** 0.5, torch.zeros(c, device="cuda")),
         ops.pack_chain_weight(torch.randn(c, c, 3, 3, device="cuda") / (9 * c) ** 0.5, torch.zeros(c, device="cuda"))]
acts = [1, 1]
if c3:
    links.append(ops.pack_chain_weight(torch.randn(c3, c, 1, 1, device="cuda") / c ** 0.5, torch.zeros(c3, device="cuda")))
    acts = [0, 0, 1]
plan = ops.ChainPlan(x, y, links, acts, residual=None if c3 else x)
info = (C.c_int32 * 8)()
_lib.load().ay2_chain_plan_info(plan._h, info)
print("ctas/SM %d grid %d smem %d nx %d nw %d alias %d tmem %d ck2 %d" % tuple(info))
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
_lib.load().ay2_chain_plan_set_debug(plan._h, dbg.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1), "tiles", B * ((hw + 15) // 16) ** 2, "tiles/CTA", B * ((hw + 15) // 16) ** 2 / info[1])
dall = dbg.cpu()
d = dall[:256].view(16, 16)
ck = dall[128:144].view(4, 4)
base = int(ck[0, 0])
print('S2 iteration clocks (tile 1, taps 0-3): [top, after wait, after MMAs, after commit]')
for r in ck.tolist():
    print('   ', [int(v) - base for v in r])
t0 = int(d[0, 0])
names = {0: "tile start", 1: "x landed", 2: "S1 issued", 3: "T ready", 4: "S2 issued", 8: "D1 done", 9: "E1 done", 11: "E2 done"}
for it in range(6):
    row = sorted((int(d[it, k]) - t0, names[k]) for k in names if int(d[it, k]))
    print(f"tile {it}: " + "  ".join(f"{n}@{t/1000:.2f}us" for t, n in row))
