"""Host-side logic of ayolov2_b200.trainer (no GPU): the warm-up / epoch schedules and parameter groups of
scripts/train/yolo_trainer.py:124-221, and -- in two gloo ranks -- that all-reducing a flat gradient buffer in contiguous
buckets (what TrainStep does on a side stream while the backward still runs) equals one all-reduce of the whole buffer."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from ayolov2_b200.trainer import NOMINAL_BATCH, lr_function, parameter_groups, warmup_state

HYP = dict(warmup_bias_lr=0.1, warmup_momentum=0.8, momentum=0.937, lrf=0.1)


def test_lr_function_matches_reference_formulas():
    for e in (0, 1, 37, 150, 299):
        assert lr_function(e, 300, 0.1) == ((1 + math.cos(e * math.pi / 300)) / 2) * 0.9 + 0.1          # yolo_trainer.py:136-138
        assert lr_function(e, 300, 0.1, linear=True) == (1 - e / 299) * 0.9 + 0.1                         # :130-133
    assert lr_function(0, 300, 0.1) == 1.0 and abs(lr_function(300, 300, 0.1) - 0.1) < 1e-12


@pytest.mark.parametrize("bs", [16, 64, 128])
def test_warmup_state(bs):
    nw = 1000.0
    for ni in (0, 1, 250, 999, 1000):
        acc, lrs, mom = warmup_state(ni, nw, 0.97, 0.01, HYP, bs)
        assert acc == max(1, np.interp(ni, [0, nw], [1, NOMINAL_BATCH / bs]).round())                    # :200-204
        assert lrs[0] == lrs[1] == np.interp(ni, [0, nw], [0.0, 0.01 * 0.97])                            # :206-216
        assert lrs[2] == np.interp(ni, [0, nw], [0.1, 0.01 * 0.97])
        assert mom == np.interp(ni, [0, nw], [0.8, 0.937])                                               # :217-221
    assert warmup_state(1000, nw, 1.0, 0.01, HYP, 16)[0] == 4 and warmup_state(0, nw, 1.0, 0.01, HYP, 16)[0] == 1


def test_parameter_groups_follow_the_reference_rule():
    """yolo_trainer.py:149-160 on this repo's kindle model: every parameter lands in exactly one group, BN weights in 0."""
    from ayolov2_b200 import synth

    m = synth.build_model("yolov5n", seed=0)
    g = parameter_groups(m)
    params = dict(m.named_parameters())
    assert set(g) == {id(p) for p in params.values()}
    pg0, pg1, pg2 = [], [], []
    for _, v in m.named_modules():
        if hasattr(v, "bias") and isinstance(v.bias, torch.Tensor):
            pg2.append(v.bias)
        if isinstance(v, nn.BatchNorm2d):
            pg0.append(v.weight)
        elif hasattr(v, "weight") and isinstance(v.weight, torch.Tensor):
            pg1.append(v.weight)
    for grp, ps in enumerate((pg0, pg1, pg2)):
        assert all(g[id(p)] == grp for p in ps)
    assert len(pg0) + len(pg1) + len(pg2) == len(params)


def _bucket_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(rank)
    flat = torch.randn(10007, generator=g)
    whole = flat.clone()
    dist.all_reduce(whole)
    cuts = [(7000, 10007), (2048, 7000), (0, 2048)]  # execution order of a backward: last layers first
    for lo, hi in cuts:
        dist.all_reduce(flat[lo:hi])
    ret[rank] = bool(torch.equal(flat, whole))
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_whole_buffer_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29650 + os.getpid() % 200
    mp.spawn(_bucket_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_weight_repack_gather_tables_reproduce_the_layout_functions():
    """train_engine refreshes every packed bf16 conv operand with ONE gather launch (ay2_repack_weights) whose index tables
    come from running the layout functions on index-valued tensors (ops.gather_index_of). On the CPU: gathering a random
    weight through each table equals applying the layout function itself -- forward layout, stride-1 and stride-2 dgrad."""
    import torch

    from ayolov2_b200 import ops

    g = torch.Generator().manual_seed(0)
    for cout, cin, k, s, p in [(32, 16, 3, 1, 1), (64, 32, 3, 2, 1), (24, 40, 1, 1, 0), (256, 128, 3, 2, 1)]:
        w = torch.randn((cout, cin, k, k), generator=g)
        fns = [ops.conv_weight_layout]
        n_d = len(ops.dgrad_weight_layouts(w, s, p, _round8(cout)))
        fns += [lambda t, i=i: ops.dgrad_weight_layouts(t, s, p, _round8(cout))[i] for i in range(n_d)]
        assert n_d == (1 if s == 1 else 4)
        for fn in fns:
            want = fn(w)
            idx = ops.gather_index_of(fn, tuple(w.shape), "cpu")
            assert idx.dtype == torch.int32 and idx.shape == want.shape
            flat = torch.cat((w.reshape(-1), torch.zeros(1)))
            got = flat[idx.reshape(-1).long().clamp_min(-1)].reshape(want.shape)  # index -1 -> the appended zero
            assert torch.equal(got, want)
            assert int((idx < 0).sum()) == int(want.numel() - (want != 0).sum())  # only the padding is -1 (w has no zeros)


def _round8(n):
    return (n + 7) // 8 * 8


def test_first_writer_overwrites_rule():
    """train_engine._GradSite: the first writer of a gradient slice in the (static) backward order overwrites, later ones
    accumulate; a writer that is partly new, partly old accumulates onto a per-step zero fill of that buffer."""
    import types

    from ayolov2_b200.train_engine import TrainEngine

    eng = types.SimpleNamespace(_cover={}, _needs_zero=set())
    claim = lambda lo, hi, key=1: TrainEngine._claim(eng, key, lo, hi)  # noqa: E731
    assert claim(0, 64) is False          # first writer of [0, 64): overwrite
    assert claim(0, 32) is True           # inside what was written: accumulate
    assert claim(64, 128) is False        # fresh slice of the same buffer (concat neighbour): overwrite
    assert claim(32, 96) is True          # covered by the union of the two
    assert not eng._needs_zero
    assert claim(96, 160) is True         # [128, 160) was never written: accumulate, and the buffer needs its zero fill
    assert eng._needs_zero == {1}
    assert claim(0, 8, key=2) is False and eng._needs_zero == {1}  # other buffers are independent


def test_pixel_pair_weight_is_the_same_convolution():
    """ops.pixel_pair_weight: a 3x3 / stride-2 / pad-1 conv over 32 channels == a 3x2-tap conv over pairs of pixels (64-channel
    "pixels"), stride 2 over rows and 1 over pairs, one pair of zero padding on the left only. Checked with torch on the CPU."""
    import torch
    import torch.nn.functional as F

    from ayolov2_b200 import ops

    g = torch.Generator().manual_seed(1)
    x = torch.randn((2, 32, 12, 20), generator=g)
    w = torch.randn((48, 32, 3, 3), generator=g)
    want = F.conv2d(x, w, stride=2, padding=1)
    # NCHW view of the pairs: channel index = (pixel of the pair) * 32 + c, exactly the NHWC buffer viewed as [B, H, W/2, 64]
    xp = x.view(2, 32, 12, 10, 2).permute(0, 4, 1, 2, 3).reshape(2, 64, 12, 10)
    xp = F.pad(xp, (1, 0, 1, 1))  # left pair, top / bottom rows
    got = F.conv2d(xp, ops.pixel_pair_weight(w), stride=(2, 1))
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-4, rtol=1e-5)
