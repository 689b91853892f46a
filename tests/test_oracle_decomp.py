"""Pins for the offline Tucker-2 decomposition (SURVEY.md §8a D1-D3, §8c "Decomposition").

* the reference's own golden numbers (tests/test_tensor_decomposition.py:47-49: 7,266,973 -> 6,329,941 parameters, forward
  loss < 0.015 on its fixture checkpoint) reproduced by (a) the UNMODIFIED reference module running on the tensorly
  restatement in oracle/decomp_oracle.py and (b) this repo's ayolov2_b200.decomposition.decompose_model -- build
  container only (the fixture lives under /root/reference);
* committed golden vectors (tests/golden/decomp_golden.json, generated from the reference by make_golden_decomp.py):
  EVBMF ranks, chain parameter counts and chain outputs on seeded planted-rank layers -- run anywhere.
"""
import importlib.util
import json
import logging
import os
from copy import deepcopy

import pytest
import torch

from ayolov2_b200 import decomposition as dec
from oracle import decomp_oracle, ref_import, yolo_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "decomp_golden.json")))


def _make_layer(case, seed):
    spec = importlib.util.spec_from_file_location("mkgold", os.path.join(HERE, "golden", "make_golden_decomp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cout, cin, k, s, ranks, noise = case
    return mod.make_layer((cout, cin, k, s, tuple(ranks), noise), seed)


@pytest.mark.parametrize("g", GOLD, ids=[f"g{i}" for i in range(len(GOLD))])
def test_product_and_oracle_match_reference_golden(g):
    conv, x = _make_layer(g["case"], g["seed"])
    assert dec.estimate_ranks(conv) == g["ranks"]
    assert decomp_oracle.estimate_ranks(conv.weight.data) == g["ranks"]
    for chain in (dec.tucker_decomposition_conv_layer(conv), decomp_oracle.tucker_chain(conv)):
        assert [type(m).__name__ for m in chain] == ["Conv2d"] * 3
        assert chain[0].kernel_size == (1, 1) and chain[2].kernel_size == (1, 1) and chain[1].stride == conv.stride
        assert chain[0].bias is None and chain[1].bias is None and chain[2].bias is not None  # decomposition.py:384-412
        assert sum(p.numel() for p in chain.parameters()) == g["chain_params"]
        with torch.no_grad():
            y = chain(x)
        n = max(1, y.numel() // 16)
        got = y.flatten()[::n][:16]
        assert torch.allclose(got, torch.tensor(g["out_samples"]), rtol=1e-3, atol=1e-3)
        assert abs(float((y - conv(x)).abs().mean()) - g["mean_abs_diff"]) < 1e-3 * max(1.0, g["mean_abs_diff"])


def test_evbmf_returns_reference_shapes():
    """EVBMF(Y) -> (U[:, :pos], diag(d), V[:, :pos], post) (decomposition.py:81-206)."""
    g = torch.Generator().manual_seed(3)
    Y = torch.randn(24, 5, generator=g) @ torch.randn(5, 200, generator=g) + 0.05 * torch.randn(24, 200, generator=g)
    U, S, V, post = dec.EVBMF(Y)
    assert S.shape == (5, 5) and U.shape == (24, 5) and V.shape == (200, 5)
    assert set(post) >= {"ma", "mb", "sa2", "sb2", "cacb", "sigma2", "F"} and post["sigma2"] > 0


def test_zero_rank_is_rejected_like_tensorly():
    conv = torch.nn.Conv2d(8, 8, 3)
    with pytest.raises(ValueError):
        dec.tucker_decomposition_conv_layer(conv, ranks=[0, 3])


@pytest.mark.skipif(not ref_import.available(), reason="reference tree (fixture checkpoint) not present on this machine")
@pytest.mark.parametrize("impl", ["reference+oracle", "product"])
def test_fixture_decomposition_reproduces_reference_golden_count(impl):
    """tests/test_tensor_decomposition.py:22-49 of the reference, verbatim flow."""
    logging.disable(logging.INFO)
    try:
        import kindle  # noqa: F401  (the shim: the fixture pickle references kindle.* classes)

        fn = decomp_oracle.load_reference().decompose_model if impl == "reference+oracle" else dec.decompose_model
        torch.manual_seed(0)
        test_input = torch.rand((1, 3, 320, 320))
        ckpt = torch.load(decomp_oracle.fixture_checkpoint_path(), weights_only=False)
        model = ckpt["model"].float()
        dm = deepcopy(model)
        fn(dm, loss_thr=0.1, prune_step=0.1)
        model.export().eval()
        dm.export().eval()
        count = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
        assert count(model) == 7266973
        assert count(dm) == 6329941
        with torch.no_grad():
            o = yolo_oracle.forward(model, test_input)[0]
            d = yolo_oracle.forward(dm, test_input)[0]
        assert float((o - d).abs().sum() / o.numel()) < 0.015
    finally:
        logging.disable(logging.NOTSET)
