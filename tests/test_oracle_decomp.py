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
    """EVBMF(Y) -> (U[:, :r], diag(d), V[:, :r], info) (decomposition.py:81-206): rank, shapes, a positive noise variance
    and shrunk singular values below the raw ones; the oracle's EVBMF (a restatement of the reference's) agrees."""
    g = torch.Generator().manual_seed(3)
    Y = torch.randn(24, 5, generator=g) @ torch.randn(5, 200, generator=g) + 0.05 * torch.randn(24, 200, generator=g)
    U, S, V, info = dec.EVBMF(Y)
    assert S.shape == (5, 5) and U.shape == (24, 5) and V.shape == (200, 5)
    assert info["rank"] == 5 and info["sigma2"] > 0
    sv = torch.linalg.svdvals(Y.double())[:5].numpy()
    d = S.diagonal()
    assert (d > 0).all() and (d < sv).all() and (d > 0.9 * sv).all()
    ro, so = decomp_oracle.evbmf_rank_sigma2(Y.numpy())
    assert ro == 5 and abs(so - info["sigma2"]) < 1e-4 * so


def test_batched_rank_search_matches_per_matrix():
    """evb_rank_batch over matrices of different shapes at once == one matrix at a time == the oracle's scipy-bounded search."""
    g = torch.Generator().manual_seed(8)
    mats = []
    for (L, M, r, noise) in [(16, 144, 3, 0.05), (64, 576, 20, 0.02), (32, 288, 1, 0.2), (48, 48, 10, 0.05), (96, 30, 7, 0.03)]:
        mats.append(torch.randn(L, r, generator=g) @ torch.randn(r, M, generator=g) / r ** 0.5 + noise * torch.randn(L, M, generator=g))
    svals = [torch.linalg.svdvals(m.double()) for m in mats]
    shapes = [tuple(m.shape) for m in mats]
    ranks, sig = dec.evb_rank_batch(svals, shapes)
    for i, m in enumerate(mats):
        (r1,), (s1,) = dec.evb_rank_batch([svals[i]], [shapes[i]])
        assert r1 == ranks[i] and abs(s1 - sig[i]) <= 1e-9 * sig[i]
        ro, so = decomp_oracle.evbmf_rank_sigma2(m.numpy())  # scipy's bounded Brent search on the same objective
        assert ro == ranks[i], (i, ro, ranks[i])
        assert abs(so - sig[i]) < 1e-4 * so, (i, so, sig[i])


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
