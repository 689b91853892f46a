"""BASELINE.json configs[1] at FULL size (yolov5s, 64 x 3 x 640 x 640 uint8, conf 0.25 / iou 0.45) through the Detector:
size-independent properties, plus the oracle on a slice the CPU finishes in seconds.

  * permutation equivariance: images are independent (eval-mode BN, per-image NMS), so permuting the batch permutes the
    detection lists -- bit for bit, whichever tile / CTA an image lands in;
  * batch-size independence: image i of the bs-64 run == the same image in a bs-2 run (different plans, same per-pixel math);
  * NMS idempotence: the kept boxes of an image survive a second class-wise NMS at the same threshold unchanged;
  * two of the 64 images, full 640 x 640, against the fp32 CPU oracle forward: logits / decoded predictions within the bf16
    tolerance, identical candidate rows outside the tolerance band of the confidence threshold.
"""
import pytest
import torch

from _parity import errs, record

pytestmark = pytest.mark.gpu
CONF, IOU = 0.25, 0.45


def _setup(batch, bias_only=False):
    from ayolov2_b200 import synth
    from ayolov2_b200.detector import Detector

    model = synth.build_model("yolov5s", seed=0).cuda()
    g = torch.Generator().manual_seed(77)
    imgs = torch.randint(0, 256, (64, 3, 640, 640), generator=g, dtype=torch.uint8)
    sample = imgs[:4].cuda().float() / 255.0
    if bias_only:  # well-conditioned head for the logit comparison against the fp32 oracle
        with torch.no_grad():
            synth.calibrate_head_bias_only(model, model(sample)[1])
    else:          # the benchmark's head calibration: ~2,000 candidates per image over all levels and many classes
        synth.calibrate_head(model, lambda: model(sample)[1])
    return model, imgs, Detector(model, batch, 640, 640, conf_thres=CONF, iou_thres=IOU, in_dtype=torch.uint8)


def test_full_size_batch_properties():
    from ayolov2_b200.nms import nms_boxes

    model, imgs, det = _setup(64)
    out = det.detect(imgs.pin_memory())
    assert len(out) == 64 and sum(o.shape[0] for o in out) > 64, "the calibrated head must produce detections"
    # permutation equivariance
    perm = torch.randperm(64, generator=torch.Generator().manual_seed(3))
    out_p = det.detect(imgs[perm].contiguous().pin_memory())
    for j, i in enumerate(perm.tolist()):
        assert torch.equal(out_p[j], out[i]), f"image {i} changed when moved to slot {j}"
    # NMS idempotence (class-wise: boxes offset by class * 4096 like metrics.py:383)
    for o in out[:8]:
        o = o.cuda()
        keep = nms_boxes(o[:, :4] + o[:, 5:6] * 4096.0, o[:, 4], IOU)
        assert keep.numel() == o.shape[0] and torch.equal(keep.sort().values, torch.arange(o.shape[0], device="cuda"))
    # batch-size independence
    from ayolov2_b200.detector import Detector

    det2 = Detector(model, 2, 640, 640, conf_thres=CONF, iou_thres=IOU, in_dtype=torch.uint8)
    out2 = det2.detect(imgs[10:12].contiguous().pin_memory())
    assert torch.equal(out2[0], out[10]) and torch.equal(out2[1], out[11])


def test_full_size_images_against_oracle():
    """EIGHT of the benchmark's 64 images, full 640 x 640 (BASELINE.json configs[1]): head logits and decoded predictions of
    the CUDA forward vs the fp32 CPU oracle at the north star's bf16 bound (1e-2, max-norm AND relative L2; measured errors
    recorded), and candidate-set equality: the rows whose objectness passes conf 0.25 are the same set, except rows whose
    oracle objectness lies within the MEASURED probability error of the threshold (a band far narrower than the tolerance).
    Detections themselves are compared at the candidate level only: on this random-weight workload greedy suppression is
    chaotic under bf16-level perturbations, which is why NMS parity is pinned on IDENTICAL inputs (tests/test_nms_gpu.py)."""
    from oracle import yolo_oracle

    model, imgs, _ = _setup(8, bias_only=True)
    pick = [0, 5, 13, 22, 31, 40, 52, 63]
    x = imgs[pick].float() / 255.0
    got_pred, got_raw = model(x.cuda())
    torch.cuda.synchronize()
    got_pred, got_raw = got_pred.float().cpu(), [r.float().cpu() for r in got_raw]
    want_pred, want_raw = yolo_oracle.forward(model.cpu().float(), x)
    assert got_pred.shape == want_pred.shape == (8, 25200, 85)
    for i, (g, w) in enumerate(zip(got_raw, want_raw)):
        e = errs(g, w)
        record(f"fullsize_b8_of_64/logits_P{i + 3}", **e)
        assert e["max_norm"] < 1e-2 and e["rel_l2"] < 1e-2, e
    eb = errs(got_pred[..., :4], want_pred[..., :4])
    pabs = float((got_pred[..., 4:] - want_pred[..., 4:]).abs().max())
    assert eb["max_norm"] < 1e-2 and eb["rel_l2"] < 1e-2 and pabs < 1e-2, (eb, pabs)
    # candidate sets (objectness > conf)
    go, wo = got_pred[..., 4] > CONF, want_pred[..., 4] > CONF
    band = (want_pred[..., 4] - CONF).abs() <= pabs
    mism = int((go != wo)[~band].sum())
    record("fullsize_b8_of_64/decoded", box_max_norm=eb["max_norm"], box_rel_l2=eb["rel_l2"], prob_max_abs=pabs,
           candidates_oracle=int(wo.sum()), candidates_cuda=int(go.sum()), rows_in_band=int(band.sum()),
           candidate_mismatches_outside_band=mism)
    assert mism == 0 and int(wo.sum()) > 4000


def test_full_size_all_images_against_gpu_run_oracle():
    """ALL 64 images of the benchmark batch at 640 x 640: the same comparison as above with the fp32 oracle executed on the
    GPU (plain torch fp32, TF32 off) as the checker -- the CPU needs minutes for 64 images, the restatement is the same code."""
    from copy import deepcopy

    from oracle import yolo_oracle

    model, imgs, _ = _setup(16, bias_only=True)
    ref = deepcopy(model).float().cuda().eval()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    worst = {}
    try:
        for s in range(0, 64, 16):
            x = (imgs[s:s + 16].float() / 255.0).cuda()
            got_pred, got_raw = model(x)
            with torch.no_grad():
                want_pred, want_raw = yolo_oracle.forward(ref, x)
            for i, (g, w) in enumerate(zip(got_raw, want_raw)):
                e = errs(g.float(), w)
                for k, v in e.items():
                    worst[f"P{i + 3}_{k}"] = max(worst.get(f"P{i + 3}_{k}", 0.0), v)
                assert e["max_norm"] < 1e-2 and e["rel_l2"] < 1e-2, (s, i, e)
            eb = errs(got_pred[..., :4].float(), want_pred[..., :4])
            pabs = float((got_pred[..., 4:].float() - want_pred[..., 4:]).abs().max())
            worst["box_max_norm"] = max(worst.get("box_max_norm", 0.0), eb["max_norm"])
            worst["prob_max_abs"] = max(worst.get("prob_max_abs", 0.0), pabs)
            assert eb["max_norm"] < 1e-2 and pabs < 1e-2, (s, eb, pabs)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    record("fullsize_b64_gpu_run_oracle/worst_of_64", **worst)
