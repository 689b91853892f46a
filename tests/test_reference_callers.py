"""Boundary integration: the reference's OWN callers -- `YoloValidator.validation_step` /
`statistics_per_image` / `compute_statistics` (scripts/utils/train_utils.py:334-520) and `YoloTrainer.training_step`
(scripts/train/yolo_trainer.py:289-358) -- executed unmodified from /root/reference against this repo's drop-in API:

  * the model is a `kindle.YOLOModel` of THIS repo (the class the reference's `train.py:137` / `val.py` would import);
  * `non_max_suppression` and `ComputeLoss` in the reference modules are replaced by stand-ins that first BIND the
    reference's call (positional + keyword arguments exactly as its call sites pass them) to the signature of the product
    entry point (`ayolov2_b200.nms.non_max_suppression`, `ayolov2_b200.loss.ComputeLoss.__init__/__call__`) -- a call shape
    the product API does not accept fails the test -- and then compute on the CPU with the pinned oracle, because the
    product kernels need a GPU (there is no CPU fallback). The same applies to the model forward.
So this checks, in the build container, that the reference's call stack drives the drop-in boundary end to end and that
what comes back (tuple shapes, list-of-(n, 6) detections, `(loss, loss_items)`) is what the callers consume. On the GPU
box the same API is exercised with the real kernels (tests/test_model_gpu.py, test_nms_gpu.py, test_loss_gpu.py,
test_boundary_gpu.py). Build container only: the reference tree is not on the GPU box."""
import inspect
import types

import numpy as np
import pytest
import torch

from oracle import loss_oracle, nms_oracle, ref_import, val_oracle, yolo_oracle

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")

HYP = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, anchor_t=4.0, fl_gamma=0.0, label_smoothing=0.0, conf_t=0.1,
           iou_t=0.6, warmup_epochs=3.0, warmup_momentum=0.8, warmup_bias_lr=0.1, momentum=0.937, lrf=0.1, weight_decay=5e-4,
           optimizer="SGD", optimizer_params=dict(lr=0.01, momentum=0.937, nesterov=True))


class _Dataset:
    names = [str(i) for i in range(80)]


class _Loader(list):
    dataset = _Dataset()


def _model():
    """This repo's kindle.YOLOModel, with the forward routed to the CPU oracle after a signature check of the product forward."""
    import kindle
    from ayolov2_b200 import synth

    m = synth.build_model("yolov5n", seed=0)
    assert isinstance(m, kindle.YOLOModel)
    product_forward = inspect.signature(kindle.YOLOModel.forward)

    def forward(self, *a, **k):
        product_forward.bind(self, *a, **k)  # the reference's call shape must fit YOLOModel.forward
        x = a[0]
        if self.training:
            return yolo_oracle.forward_with_grad(self, x.float())
        pred, raw = yolo_oracle.forward(self, x.float())
        return pred, raw
    m.forward = types.MethodType(forward, m)
    m.hyp = dict(HYP)
    m.nc = 80
    return m


def _nms_stand_in(calls):
    from ayolov2_b200 import nms as product

    sig = inspect.signature(product.non_max_suppression)

    def non_max_suppression(*a, **k):
        b = sig.bind(*a, **k)
        b.apply_defaults()
        calls.append(dict(b.arguments))
        kw = {n: b.arguments[n] for n in ("conf_thres", "iou_thres", "classes", "agnostic", "multi_label", "max_det", "nms_type")}
        assert not b.arguments["labels"] or all(len(l) == 0 for l in b.arguments["labels"])
        return nms_oracle.non_max_suppression(b.arguments["prediction"], **kw)
    return non_max_suppression


def _loss_stand_in(calls):
    from ayolov2_b200 import loss as product

    init_sig = inspect.signature(product.ComputeLoss.__init__)
    call_sig = inspect.signature(product.ComputeLoss.__call__)

    class ComputeLoss:
        def __init__(self, *a, **k):
            b = init_sig.bind(self, *a, **k)
            model = b.arguments["model"]
            head = model.model[-1]
            for attr in ("nl", "na", "nc", "anchors", "stride"):  # what the product ComputeLoss reads from the head
                assert hasattr(head, attr), attr
            self.model, self.head = model, head

        def __call__(self, *a, **k):
            b = call_sig.bind(self, *a, **k)
            preds, targets = b.arguments["preds"], b.arguments["targets"]
            calls.append((len(preds), tuple(targets.shape)))
            return loss_oracle.compute_loss(list(preds), targets, self.head.anchors, self.model.hyp, self.head.nc)
    return ComputeLoss


def _batch(bs=2, size=96, seed=0):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randint(0, 256, (bs, 3, size, size), generator=g, dtype=torch.uint8)
    t = torch.zeros(5, 6)
    t[:, 0] = torch.tensor([0, 0, 1, 1, 1])
    t[:, 1] = torch.randint(0, 80, (5,), generator=g).float()
    t[:, 2:4] = 0.2 + 0.6 * torch.rand(5, 2, generator=g)
    t[:, 4:6] = 0.1 + 0.3 * torch.rand(5, 2, generator=g)
    shapes = tuple(((size, size), ((1.0, 1.0), (0.0, 0.0))) for _ in range(bs))
    return imgs, t, tuple(f"img{i}.jpg" for i in range(bs)), shapes


def test_reference_validator_drives_the_drop_in_boundary(monkeypatch):
    import kindle  # noqa: F401  (this repo's shim must be the `kindle` the reference modules see, not ref_import's stub)

    ref_import.load()
    from scripts.utils import train_utils as tu  # the UNMODIFIED reference module

    nms_calls, loss_calls = [], []
    monkeypatch.setattr(tu, "non_max_suppression", _nms_stand_in(nms_calls))
    monkeypatch.setattr(tu, "ComputeLoss", _loss_stand_in(loss_calls))
    model = _model().eval()
    cfg = {"train": {"single_cls": False, "plot": False, "batch_size": 2, "image_size": 96}, "hyper_params": dict(HYP)}
    v = tu.YoloValidator(model, _Loader(), torch.device("cpu"), cfg, compute_loss=True)
    v.init_statistics()
    v.seen = 0
    # make the head fire: the random-init head with the detection-prior biases gives nothing above conf 0.1
    head = model.model[-1]
    with torch.no_grad():
        for conv in head.conv:
            conv.bias.view(head.na, -1)[:, 4] += 6.0
            conv.bias.view(head.na, -1)[:, 5 + 3] += 6.0  # ... and one class score
    batch = _batch()
    v.validation_step(batch, 0)
    # the reference called the product-shaped NMS exactly as train_utils.py:461-469 does
    assert len(nms_calls) == 1 and nms_calls[0]["multi_label"] is True and nms_calls[0]["nms_type"] == "nms"
    assert nms_calls[0]["conf_thres"] == HYP["conf_t"] and nms_calls[0]["iou_thres"] == HYP["iou_t"]
    assert loss_calls == [(3, (5, 6))] and v.loss.shape == (3,) and torch.isfinite(v.loss).all()
    assert v.seen == 2 and len(v.statistics["stats"]) == 2
    # its statistics on those detections equal the oracle's matching of the same detections
    imgs, targets, _, shapes = _batch()
    pred, _ = yolo_oracle.forward(model, imgs.float() / 255.0)
    dets = nms_oracle.non_max_suppression(pred, HYP["conf_t"], HYP["iou_t"], multi_label=True)
    assert sum(d.shape[0] for d in dets) > 0
    for si, (correct, conf, pcls, tcls) in enumerate(v.statistics["stats"]):
        lab = targets[targets[:, 0] == si, 1:].clone()
        lab[:, 1:] *= 96.0
        lab_xyxy = np.concatenate((lab[:, :1].numpy(), nms_oracle.xywh2xyxy(lab[:, 1:].numpy())), 1)
        want = val_oracle.process_batch(dets[si].numpy(), lab_xyxy)
        assert np.array_equal(correct.numpy(), want) and np.array_equal(conf.numpy(), dets[si][:, 4].numpy())
    v.compute_statistics()
    assert "map50" in v.statistics and np.isfinite(v.statistics["map50"])


def test_reference_training_step_drives_the_drop_in_boundary(monkeypatch):
    import kindle  # noqa: F401

    ref_import.load()
    from scripts.train import yolo_trainer as yt  # the UNMODIFIED reference module

    loss_calls = []
    LossCls = _loss_stand_in(loss_calls)
    model = _model().train()
    t = object.__new__(yt.YoloTrainer)  # the constructor needs dataloaders / W&B; training_step itself needs only these
    t.model, t.device, t.cuda = model, torch.device("cpu"), False
    t.cfg_train = {"multi_scale": False, "batch_size": 2, "world_size": 1, "epochs": 3}
    t.cfg_hyp = dict(HYP)
    t.train_dataloader = _Loader([None] * 10)
    t.loss = LossCls(model)
    t.nbs, t.accumulate = 64, 1
    t.num_warmups = 1000
    t.epochs = 3
    t.ema, t.pbar, t.log_dir = None, None, "/tmp"
    t.mloss = torch.zeros(4)
    t.scaler = torch.amp.GradScaler("cpu", enabled=False)
    monkeypatch.setattr(yt, "plot_images", lambda **k: None)
    monkeypatch.setattr(yt, "RANK", -1)
    t.log_dict = lambda d: None
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, nesterov=True)
    for g in opt.param_groups:
        g["initial_lr"] = g["lr"]
    opt.add_param_group({"params": [torch.zeros(1, requires_grad=True)], "initial_lr": 0.01})
    opt.add_param_group({"params": [torch.zeros(1, requires_grad=True)], "initial_lr": 0.01})
    t.optimizer = [opt]
    before = [p.detach().clone() for p in model.parameters()]
    loss = t.training_step(_batch(), 5, 0)
    assert loss.dim() == 0 and torch.isfinite(loss)
    assert loss_calls == [(3, (5, 6))]
    assert t.mloss.shape == (4,)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, model.parameters())), "optimizer step ran"
    # the schedule the reference applied at ni = 5 equals what TrainStep.warmup computes
    from ayolov2_b200.trainer import lr_function, warmup_state

    acc, lrs, mom = warmup_state(5, t.num_warmups, lr_function(0, 3, HYP["lrf"]), 0.01, HYP, 2)
    assert acc == t.accumulate and abs(opt.param_groups[0]["lr"] - lrs[0]) < 1e-12
    assert abs(opt.param_groups[2]["lr"] - lrs[2]) < 1e-12 and abs(opt.param_groups[0]["momentum"] - mom) < 1e-12


def test_every_reference_call_site_binds_to_the_product_signatures():
    """Static half of the boundary check: EVERY call of the drop-in entry points in the reference tree -- validator, trainer,
    KD trainer (scripts/train/kd_trainer.py:378), val.py / val2.py, TTA -- is parsed with `ast` and bound (number of
    positional arguments + keyword names) to the signature of the product function of the same name."""
    import ast
    import glob
    import os

    from ayolov2_b200 import loss as ploss
    from ayolov2_b200 import nms as pnms

    targets = {
        "non_max_suppression": inspect.signature(pnms.non_max_suppression),
        "batched_nms": inspect.signature(pnms.batched_nms),
        "box_iou": inspect.signature(pnms.box_iou),
        "ComputeLoss": inspect.signature(ploss.ComputeLoss),
    }
    root = "/root/reference"
    files = [f for f in glob.glob(os.path.join(root, "**", "*.py"), recursive=True) if "/tests/" not in f]
    seen = {k: [] for k in targets}
    for f in files:
        try:
            tree = ast.parse(open(f, encoding="utf-8").read())
        except SyntaxError:
            continue
        defined_here = {n.name for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
        for node in ast.walk(tree):
            if not isinstance(node, ast.Call):
                continue
            name = node.func.id if isinstance(node.func, ast.Name) else (node.func.attr if isinstance(node.func, ast.Attribute) else None)
            if name not in targets or any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
                continue
            if isinstance(node.func, ast.Attribute) and not (isinstance(node.func.value, ast.Name) and node.func.value.id in ("nms", "metrics", "losses")):
                continue  # torchvision.ops.boxes.batched_nms and friends: not the reference's own function
            if name in defined_here and name == "batched_nms" and f.endswith("scripts/utils/nms.py"):
                pass  # the definition file also calls torchvision's; those were filtered above
            args = [object()] * len(node.args)
            kwargs = {k.arg: object() for k in node.keywords}
            try:
                targets[name].bind(*args, **kwargs)
            except TypeError as e:
                raise AssertionError(f"{os.path.relpath(f, root)}:{node.lineno}: {name}(...) does not bind: {e}")
            seen[name].append(f"{os.path.relpath(f, root)}:{node.lineno}")
    assert any("kd_trainer.py" in s for s in seen["non_max_suppression"]), seen["non_max_suppression"]
    assert any("train_utils.py" in s for s in seen["non_max_suppression"])
    assert len(seen["ComputeLoss"]) >= 2 and len(seen["batched_nms"]) >= 1, seen
