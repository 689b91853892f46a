"""Per-parameter gradient comparison CUDA train step vs CPU oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from copy import deepcopy
import torch
from ayolov2_b200 import synth
from ayolov2_b200.loss import ComputeLoss
from oracle import loss_oracle, yolo_oracle
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_trainstep_gpu import HYP, _targets

from test_trainstep_gpu import _build
name = sys.argv[1] if len(sys.argv) > 1 else "mini_v6"
hw, bs = ((256, 256), 4) if name.startswith("mini") else ((128, 128), 4)
base = _build(name); base.hyp = dict(HYP)
x = torch.rand((bs, 3, *hw), generator=torch.Generator().manual_seed(3))
targets = _targets(bs, 12, 4)
ref = deepcopy(base).train()
yolo_oracle.SIMULATE_BF16 = True
preds_ref = yolo_oracle.forward_with_grad(ref, x)
yolo_oracle.SIMULATE_BF16 = False
for p_ in preds_ref: p_.retain_grad()
head = ref.model[-1]
g = torch.Generator().manual_seed(7)
G = [torch.randn(p_.shape, generator=g) / p_.numel() ** 0.5 for p_ in preds_ref]
sum((p_ * gg).sum() for p_, gg in zip(preds_ref, G)).backward()
m = deepcopy(base).cuda().train()
preds = m(x.cuda())
for p_ in preds: p_.retain_grad()
sum((p_ * gg.cuda()).sum() for p_, gg in zip(preds, G)).backward()
torch.cuda.synchronize()
for i, (a, b) in enumerate(zip(preds, preds_ref)):
    print("dL/dpred level", i, "rel", float((a.grad.cpu() - b.grad).norm() / b.grad.norm()))
rows = []
for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
    g1, g2 = p.grad.detach().cpu().double(), q.grad.detach().double()
    rel = float((g1 - g2).norm() / (g2.norm() + 1e-20))
    cos = float((g1 * g2).sum() / (g1.norm() * g2.norm() + 1e-30))
    rows.append((n, rel, cos, float(g2.norm()), float(g1.norm())))
for r in reversed(rows):
    print(f"{r[0]:55s} rel {r[1]:8.4f} cos {r[2]:7.4f} |ref| {r[3]:.4e} |got| {r[4]:.4e}")
