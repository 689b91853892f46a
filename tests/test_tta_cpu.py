"""ayolov2_b200.tta (host orchestration, no kernels of its own) == scripts/utils/tta_utils.py::inference_with_tta of the
reference, bit for bit, when both drive the same forward (the CPU oracle). Build container only (needs /root/reference)."""
import pytest
import torch

from ayolov2_b200 import synth, tta  # the kindle shim must be imported before the reference stubs `kindle`
from oracle import ref_import, yolo_oracle


class _OracleModel(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m, self.model, self.stride = m, m.model, m.stride

    def forward(self, xi):
        return yolo_oracle.forward(self.m, xi)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_tta_orchestration_matches_reference():
    model = _OracleModel(synth.build_model("yolov5n", seed=3))
    ref_import.load()
    from scripts.utils.tta_utils import inference_with_tta as ref_tta  # type: ignore

    x = torch.rand((2, 3, 128, 160), generator=torch.Generator().manual_seed(1))
    for s, f in (([1, 0.83, 0.67], [None, 3, None]), ([1, 0.83], [None, 2])):
        a, _ = ref_tta(model, x.clone(), s, f)
        b, _ = tta.inference_with_tta(model, x.clone(), s, f)
        assert a.shape == b.shape and torch.equal(a, b)
