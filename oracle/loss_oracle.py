"""CPU oracle for the detection loss. TEST INFRASTRUCTURE ONLY (see oracle/nms_oracle.py header).

Restates scripts/loss/losses.py:168-391 (ComputeLoss.__call__ / build_targets, default configuration:
fl_gamma >= 0 i.e. plain BCE or the FocalLoss wrapper, autobalance on or off, gr = 1, sort_obj_iou off) and scripts/utils/metrics.py:60-135 (bbox_iou, CIoU
branch) in plain PyTorch so that autograd supplies the reference gradients. Pinned by tests/test_oracle_loss.py
against tests/golden/loss_golden.npz (generated from the unmodified reference) and, in the build container,
against the reference itself.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def smooth_bce(eps: float = 0.1) -> Tuple[float, float]:
    """losses.py:16-27."""
    return 1.0 - 0.5 * eps, 0.5 * eps


def bbox_ciou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """metrics.py:84-130 with x1y1x2y2=False, c_iou=True. box1 (4, n) xywh, box2 (n, 4) xywh."""
    box2 = box2.T
    b1_x1, b1_x2 = box1[0] - box1[2] / 2, box1[0] + box1[2] / 2
    b1_y1, b1_y2 = box1[1] - box1[3] / 2, box1[1] + box1[3] / 2
    b2_x1, b2_x2 = box2[0] - box2[2] / 2, box2[0] + box2[2] / 2
    b2_y1, b2_y2 = box2[1] - box2[3] / 2, box2[1] + box2[3] / 2
    inter = (torch.min(b1_x2, b2_x2) - torch.max(b1_x1, b2_x1)).clamp(0) * (
        torch.min(b1_y2, b2_y2) - torch.max(b1_y1, b2_y1)).clamp(0)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(b1_x2, b2_x2) - torch.min(b1_x1, b2_x1)
    ch = torch.max(b1_y2, b2_y2) - torch.min(b1_y1, b2_y1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2) ** 2 + (b2_y1 + b2_y2 - b1_y1 - b1_y2) ** 2) / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


def build_targets(shapes: Sequence[Tuple[int, int]], targets: torch.Tensor, anchors: torch.Tensor, anchor_t: float):
    """losses.py:302-391. shapes: [(ny, nx)] per level; targets (nt, 6) [img, cls, x, y, w, h] normalised;
    anchors (nl, na, 2) in grid units. Returns per level (b, a, gj, gi, tbox, anch, tcls) in the reference's order."""
    na, nt = anchors.shape[1], targets.shape[0]
    out = []
    gain = torch.ones(7)
    ai = torch.arange(na).float().view(na, 1).repeat(1, nt)
    t_all = torch.cat((targets.repeat(na, 1, 1), ai[:, :, None]), 2)
    g = 0.5
    off = torch.tensor([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]]).float() * g
    for i, (ny, nx) in enumerate(shapes):
        anch_i = anchors[i]
        gain[2:6] = torch.tensor([nx, ny, nx, ny]).float()
        t = t_all * gain
        if nt:
            r = t[:, :, 4:6] / anch_i[:, None]
            j = torch.max(r, 1.0 / r).max(2)[0] < anchor_t
            t = t[j]
            gxy = t[:, 2:4]
            gxi = gain[[2, 3]] - gxy
            j, k = ((gxy % 1.0 < g) & (gxy > 1.0)).T
            l, m = ((gxi % 1.0 < g) & (gxi > 1.0)).T
            j = torch.stack((torch.ones_like(j), j, k, l, m))
            t = t.repeat((5, 1, 1))[j]
            offsets = (torch.zeros_like(gxy)[None] + off[:, None])[j]
        else:
            t = t_all[0]
            offsets = 0
        b, c = t[:, :2].long().T
        gxy = t[:, 2:4]
        gwh = t[:, 4:6]
        gij = (gxy - offsets).long()
        gi, gj = gij.T
        a = t[:, 6].long()
        gj = gj.clamp(0, ny - 1)
        gi = gi.clamp(0, nx - 1)
        out.append((b, a, gj, gi, torch.cat((gxy - gij, gwh), 1), anch_i[a], c))
    return out


def bce_logits(x: torch.Tensor, t: torch.Tensor, pos_weight: torch.Tensor, gamma: float, alpha: float = 0.25) -> torch.Tensor:
    """Mean BCEWithLogits(pos_weight); with gamma > 0 the FocalLoss wrapper of losses.py:64-114 (the criterion ComputeLoss
    installs when hyp["fl_gamma"] > 0, :193-196): every element is scaled by alpha_t * (1 - p_t) ** gamma before the mean."""
    if gamma <= 0:
        return F.binary_cross_entropy_with_logits(x, t, pos_weight=pos_weight)
    el = F.binary_cross_entropy_with_logits(x, t, pos_weight=pos_weight, reduction="none")
    prob = torch.sigmoid(x)
    p_t = t * prob + (1 - t) * (1 - prob)
    alpha_t = t * alpha + (1 - t) * (1 - alpha)
    return (el * alpha_t * (1.0 - p_t) ** gamma).mean()


def compute_loss(preds: List[torch.Tensor], targets: torch.Tensor, anchors: torch.Tensor, hyp: Dict[str, float],
                 nc: int, balance: "List[float] | None" = None, ssi: "int | None" = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """losses.py:223-300. preds: list of (bs, na, ny, nx, 5+nc) logits (requires_grad for the backward oracle).
    `balance` (a list, updated IN PLACE) + `ssi` (index of the stride-16 level) switch autobalance on (losses.py:286-292):
    the loss of this call uses the incoming weights, then balance[i] <- 0.9999 balance[i] + 1e-4 / obj_loss_i and the
    list is renormalised by balance[ssi]."""
    cp, cn = smooth_bce(hyp.get("label_smoothing", 0.0))
    gamma = float(hyp.get("fl_gamma", 0.0))
    auto = balance is not None
    if balance is None:
        balance = {3: [4.0, 1.0, 0.4]}.get(len(preds), [4.0, 1.0, 0.25, 0.06, 0.02])
    cls_pw = torch.tensor([hyp["cls_pw"]])
    obj_pw = torch.tensor([hyp["obj_pw"]])
    lcls, lbox, lobj = torch.zeros(1), torch.zeros(1), torch.zeros(1)
    tg = build_targets([(p.shape[2], p.shape[3]) for p in preds], targets, anchors, hyp["anchor_t"])
    for i, pi in enumerate(preds):
        b, a, gj, gi, tbox, anch, tcls = tg[i]
        tobj = torch.zeros_like(pi[..., 0])
        n = b.shape[0]
        if n:
            ps = pi[b, a, gj, gi]
            pxy = ps[:, :2].sigmoid() * 2.0 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anch
            iou = bbox_ciou(torch.cat((pxy, pwh), 1).T, tbox)
            lbox = lbox + (1.0 - iou).mean()
            tobj[b, a, gj, gi] = iou.detach().clamp(0).type(tobj.dtype)  # gr = 1; last write wins on duplicates
            if nc > 1:
                t = torch.full_like(ps[:, 5:], cn)
                t[range(n), tcls] = cp
                lcls = lcls + bce_logits(ps[:, 5:], t, cls_pw, gamma)
        obji = bce_logits(pi[..., 4], tobj, obj_pw, gamma)
        lobj = lobj + obji * balance[i]
        if auto:
            balance[i] = balance[i] * 0.9999 + 0.0001 / obji.detach().item()
    if auto:
        norm = balance[ssi]
        balance[:] = [x / norm for x in balance]
    lbox = lbox * hyp["box"]
    lobj = lobj * hyp["obj"]
    lcls = lcls * hyp["cls"]
    bs = preds[0].shape[0]
    loss = lbox + lobj + lcls
    return loss * bs, torch.cat((lbox, lobj, lcls, loss)).detach()
