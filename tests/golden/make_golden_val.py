"""Generates tests/golden/val_golden.npz from the UNMODIFIED reference (build container only):
YoloValidator.process_batch (train_utils.py:294-333), scale_coords (general.py:324-358) and ap_per_class
(metrics.py:476-548) on seeded synthetic detections / labels (oracle.val_oracle.synth_case)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import, val_oracle  # noqa: E402


def ref_functions():
    ref_import.load()
    import scripts.utils.train_utils as tu  # type: ignore
    from scripts.utils.general import scale_coords  # type: ignore
    from scripts.utils.metrics import ap_per_class  # type: ignore

    fake = types.SimpleNamespace(iouv=torch.linspace(0.5, 0.95, 10))
    return (lambda d, l: tu.YoloValidator.process_batch(fake, d, l)), scale_coords, ap_per_class


def main():
    process_batch, scale_coords, ap_per_class = ref_functions()
    out = {}
    tps, confs, pcls, tcls = [], [], [], []
    for case in range(4):
        det, lab = val_oracle.synth_case(case, n_det=150 if case else 3, n_lab=30 if case else 2)
        ratio_pad = ((0.8, 0.8), (16.0, 24.0))
        shape0 = (740, 760)
        dn = torch.from_numpy(det.copy())
        scale_coords((640, 640), dn[:, :4], shape0, ratio_pad)
        ln = torch.from_numpy(lab.copy())
        scale_coords((640, 640), ln[:, 1:], shape0, ratio_pad)
        correct = process_batch(dn, ln)
        out[f"c{case}_det"], out[f"c{case}_lab"] = det, lab
        out[f"c{case}_detn"], out[f"c{case}_labn"] = dn.numpy(), ln.numpy()
        out[f"c{case}_correct"] = correct.numpy()
        tps.append(correct.numpy()); confs.append(det[:, 4]); pcls.append(det[:, 5]); tcls.append(lab[:, 0])
    p, r, ap, f1, cls = ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), np.concatenate(tcls))
    out.update(ap_p=p, ap_r=r, ap_ap=ap, ap_f1=f1, ap_cls=cls)
    np.savez_compressed(os.path.join(HERE, "val_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()}, "mAP50", ap[:, 0].mean(), "mAP", ap.mean())


if __name__ == "__main__":
    main()
