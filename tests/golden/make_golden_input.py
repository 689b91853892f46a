"""Generates tests/golden/input_golden.npz from the UNMODIFIED reference (build container only; needs cv2):
LoadImages._letterbox (data_loader.py:395-459), the HWC BGR -> CHW RGB step (:388-389), LoadImages.collate_fn (:461-477)
and LoadImagesAndLabels.collate_fn (:888-909) on seeded synthetic images (oracle.input_oracle.synth_images).
The fixture is kept small (network input 96 x 128 / 128 x 128) -- the algorithm has no size-dependent branch."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import input_oracle, ref_import  # noqa: E402

# (network input (H, W), image shapes, letterbox keyword arguments)
CASES = [
    ((128, 128), [(128, 96), (96, 128), (128, 128), (127, 90), (60, 45), (50, 128), (256, 256), (151, 87)], dict(auto=False)),
    ((96, 128), [(96, 128), (72, 128), (96, 100), (33, 47), (192, 256), (95, 127)], dict(auto=False)),
    ((128, 128), [(128, 96), (60, 45), (150, 100)], dict(auto=False, scale_up=False)),
    ((128, 128), [(100, 77), (128, 64)], dict(auto=False, scale_fill=True)),
]


# decoded image shapes and img_size for the `_load_image` fixture (data_loader.py:294-350): INTER_AREA for the general, the
# integer (3 x 3, 2 x 1 -> general in y) and the 2 x 2 ratios, INTER_LINEAR for the up-scale and under augmentation
LOAD_CASES = [((150, 97), 64, False), ((97, 150), 64, False), ((192, 96), 64, False), ((128, 128), 64, False),
              ((40, 31), 64, False), ((150, 97), 64, True), ((64, 50), 64, False), ((131, 200), 100, False)]


def reference_load_image(dl, im, img_size, augmentation, tmpdir):
    """The unmodified LoadImages._load_image on a PNG (lossless) of `im`."""
    import cv2

    path = os.path.join(tmpdir, "im.png")
    cv2.imwrite(path, im)
    fake = types.SimpleNamespace(imgs=[None], img_npy=[None], img_files=[path], img_size=img_size,
                                 augmentation=(lambda x: x) if augmentation else None, cache_images=None)
    out, hw0, hw = dl.LoadImages._load_image(fake, 0)
    return out, hw0, hw


def main():
    import tempfile

    dl = ref_import.load_data_loader()
    fake = types.SimpleNamespace(img_size=128, stride=32)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for li, ((h, w), size, aug) in enumerate(LOAD_CASES):
            im = input_oracle.synth_images(300 + li, [(h, w)])[0]
            res, hw0, hw = reference_load_image(dl, im, size, aug, tmp)
            assert tuple(hw0) == (h, w) and tuple(hw) == res.shape[:2]
            out[f"load{li}_in"], out[f"load{li}_out"] = im, res
    for ci, (new_shape, shapes, kw) in enumerate(CASES):
        imgs = input_oracle.synth_images(100 + ci, shapes)
        items, geo = [], []
        for im in imgs:
            lb, ratio, pad = dl.LoadImages._letterbox(fake, im, new_shape=new_shape, **kw)
            chw = np.ascontiguousarray(lb.transpose((2, 0, 1))[::-1])  # data_loader.py:388-389
            h, w = im.shape[:2]
            items.append((torch.from_numpy(chw), f"img{len(items)}.jpg", ((h, w), ((1.0, 1.0), pad))))
            geo.append([ratio[0], ratio[1], pad[0], pad[1]])
        stacked, paths, shp = dl.LoadImages.collate_fn(items)
        out[f"c{ci}_batch"] = stacked.numpy()
        out[f"c{ci}_geo"] = np.asarray(geo, np.float64)
        for k, im in enumerate(imgs):
            out[f"c{ci}_img{k}"] = im
    rng = np.random.default_rng(7)
    labels = [rng.random((n, 6)).astype(np.float32) for n in (3, 0, 5, 1)]
    batch = [(torch.zeros(3, 8, 8, dtype=torch.uint8), torch.from_numpy(l.copy()), f"p{i}", ((8, 8), ((1.0, 1.0), (0.0, 0.0))))
             for i, l in enumerate(labels)]
    _, lab, _, _ = dl.LoadImagesAndLabels.collate_fn(batch)
    for i, l in enumerate(labels):
        out[f"lab_in{i}"] = l
    out["lab_out"] = lab.numpy()
    np.savez_compressed(os.path.join(HERE, "input_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if "batch" in k or k == "lab_out"})


if __name__ == "__main__":
    main()
