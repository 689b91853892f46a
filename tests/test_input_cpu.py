"""Host half of the GPU input side (ayolov2_b200/data_loader.py): geometry vs the oracle (and through it the reference's
_letterbox, data_loader.py:428-455), arena / table packing, the no-CPU-fallback rule. CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ayolov2_b200 import data_loader as dl  # noqa: E402
from oracle import input_oracle  # noqa: E402


def test_geometry_matches_oracle_on_many_shapes():
    rng = np.random.default_rng(0)
    for t in range(2000):
        h, w = (int(v) for v in rng.integers(1, 1500, 2))
        new_shape = [(640, 640), (480, 640), (384, 672), (128, 96)][t % 4]
        kw = [dict(auto=False), dict(auto=True), dict(auto=False, scale_up=False), dict(auto=False, scale_fill=True),
              dict(auto=True, scale_up=False, stride=64)][t % 5]
        assert dl.letterbox_geometry((h, w), new_shape, **kw) == input_oracle.letterbox_geometry((h, w), new_shape, **kw), ((h, w), new_shape, kw)


def test_pack_batch_table_and_arena():
    shapes = [(128, 96), (60, 45), (50, 128), (256, 256)]
    imgs = input_oracle.synth_images(5, shapes)
    pb = dl.pack_batch(imgs, (128, 128), paths=["a", "b", "c", "d"])
    assert pb.batch == 4 and pb.out_shape == (128, 128) and pb.paths == ("a", "b", "c", "d")
    raw = pb.arena.numpy()
    rec = raw[:pb.table_bytes].view(dl._REC)
    for i, im in enumerate(imgs):
        (uw, uh), ratio, pad, (top, bottom, left, right) = input_oracle.letterbox_geometry(im.shape[:2], (128, 128), auto=False)
        r = rec[i]
        assert (r["src_h"], r["src_w"], r["src_row_bytes"]) == (im.shape[0], im.shape[1], 3 * im.shape[1])
        assert (r["dst_h"], r["dst_w"], r["top"], r["left"]) == (uh, uw, top, left)
        assert r["src_offset"] % 16 == (16 - pb.table_bytes % 16) % 16  # images start on 16-byte boundaries of the arena
        start = pb.table_bytes + int(r["src_offset"])
        assert np.array_equal(raw[start:start + im.size].reshape(im.shape), im)
        assert pb.shapes[i] == (im.shape[:2], ((1.0, 1.0), pad)) and pb.ratios[i] == ratio
    # the reference's shapes tuple with a different native size (data_loader.py:391)
    pb2 = dl.pack_batch(imgs[:1], (128, 128), orig_shapes=[(256, 192)])
    assert pb2.shapes[0][0] == (256, 192) and pb2.shapes[0][1][0] == (0.5, 0.5)


def test_pack_batch_rejects_what_the_batch_cannot_hold():
    im = input_oracle.synth_images(1, [(100, 60)])[0]
    with pytest.raises(ValueError, match="auto=True"):
        dl.pack_batch([im], (128, 128), auto=True)
    with pytest.raises(TypeError):
        dl.pack_batch([im.astype(np.float32)], (128, 128))


def test_collate_fn_mirrors_the_reference_return_values():
    imgs = input_oracle.synth_images(2, [(96, 128), (33, 47)])
    pb, paths, shapes = dl.collate_fn([(imgs[0], "x.jpg", (192, 256)), (imgs[1], "y.jpg", (33, 47))], new_shape=(96, 128))
    assert paths == ("x.jpg", "y.jpg") and shapes == pb.shapes and shapes[0][0] == (192, 256)
    _, ref_shapes = input_oracle.load_and_collate(imgs, (96, 128))
    assert shapes[1] == ref_shapes[1]


def test_no_cpu_fallback():
    imgs = input_oracle.synth_images(2, [(96, 128)])
    pb = dl.pack_batch(imgs, (96, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pb.to_device(torch.device("cpu"))


class _LoadedImages(torch.utils.data.Dataset):
    """What the reference's LoadImages would return if __getitem__ stopped right after _load_image (data_loader.py:373)."""

    def __init__(self):
        self.shapes = [(96, 128), (72, 128), (96, 100), (33, 47), (95, 127)]
        self.imgs = input_oracle.synth_images(8, self.shapes)

    def __len__(self):
        return len(self.imgs)

    def __getitem__(self, i):
        return self.imgs[i], f"img{i}.jpg", self.imgs[i].shape[:2]


def test_collate_fn_runs_inside_dataloader_workers():
    """The host half is pure CPU and its result pickles: it can be the DataLoader's collate_fn in worker processes, like the
    reference's (data_loader.py:461-477); the device half runs in the consumer."""
    import functools

    ds = _LoadedImages()
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, num_workers=1, drop_last=False,
                                         collate_fn=functools.partial(dl.collate_fn, new_shape=(96, 128)))
    seen = 0
    for pb, paths, shapes in loader:
        assert isinstance(pb, dl.PackedBatch) and pb.batch == len(paths) == len(shapes) and not pb.arena.is_cuda
        want = dl.pack_batch(ds.imgs[seen:seen + pb.batch], (96, 128))
        assert torch.equal(pb.arena, want.arena) and pb.shapes == want.shapes and pb.kinds == want.kinds
        assert paths == tuple(f"img{i}.jpg" for i in range(seen, seen + pb.batch))
        seen += pb.batch
    assert seen == len(ds)


def test_pack_batch_with_decode_side_resize():
    """img_size: `_load_image`'s resize is planned by the host (data_loader.py:320-329) -- table of the resize kernel, scratch
    space behind the uploaded bytes, letterbox records that read the resized images, the reference's shapes tuples."""
    shapes = [(150, 97), (64, 50), (40, 31), (128, 128), (97, 150)]
    imgs = input_oracle.synth_images(6, shapes)
    pb = dl.pack_batch(imgs, (64, 64), img_size=64)
    raw = pb.arena.numpy()
    assert pb.host_bytes == raw.size and pb.n_load == 4 and pb.scratch_bytes > 0
    rec = raw[:pb.table_bytes].view(dl._REC)
    lstart = pb.table_bytes + pb.load_table_offset
    lrec = raw[lstart:lstart + pb.n_load * dl._LOAD_REC.itemsize].view(dl._LOAD_REC)
    _, ref_shapes = input_oracle.load_and_collate(imgs, (64, 64), img_size=64)
    assert pb.shapes == ref_shapes
    k = 0
    for i, im in enumerate(imgs):
        h0, w0 = im.shape[:2]
        want = input_oracle.load_image_resize(im, 64)
        assert (rec[i]["src_h"], rec[i]["src_w"]) == want.shape[:2]
        if want is im:
            assert rec[i]["src_offset"] + pb.table_bytes < pb.host_bytes
            continue
        r = lrec[k]
        k += 1
        assert (r["src_h"], r["src_w"], r["dst_h"], r["dst_w"]) == (h0, w0, want.shape[0], want.shape[1])
        assert r["mode"] == (dl.LR_AREA if want.shape[0] < h0 else dl.LR_LINEAR)
        assert r["dst_offset"] == rec[i]["src_offset"] and r["dst_offset"] + pb.table_bytes >= pb.host_bytes  # in the scratch space
        assert r["dst_offset"] + pb.table_bytes + 3 * want.shape[0] * want.shape[1] <= pb.host_bytes + pb.scratch_bytes
        start = pb.table_bytes + int(r["src_offset"])
        assert np.array_equal(raw[start:start + im.size].reshape(im.shape), im)
        assert r["scale_x"] == 1.0 / (want.shape[1] / w0) and r["scale_y"] == 1.0 / (want.shape[0] / h0)
    assert k == pb.n_load and pb.max_dst_pixels == max(64 * 41, 64 * 49, 64 * 64, 41 * 64)
    aug = dl.pack_batch(imgs[:1], (64, 64), img_size=64, augmentation=True)
    l2 = aug.arena.numpy()[aug.table_bytes + aug.load_table_offset:][:64].view(dl._LOAD_REC)
    assert l2[0]["mode"] == dl.LR_LINEAR
