"""Generate tests/golden/aux/pyramid.npz by executing the UNMODIFIED reference dataset code
(build container only):  python tests/golden/make_golden_pyramid.py

``MonoDataset.preprocess`` (datasets/mono_dataset.py:92-111) is called unbound-style on a
``KITTIRAWDataset`` constructed without data (its __init__ only stores sizes and builds the
``transforms.Resize`` chain), i.e. the real Pillow Lanczos resampling + torchvision ToTensor.
The scale -1 input is given at the network resolution, so scale 0 equals it (PIL returns a copy
for an unchanged size) and scales 1-3 are the 2x chain the device kernels restate.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import reference_runner  # noqa: E402


def main():
    assert reference_runner.available()
    reference_runner._install_stubs()
    sys.path.insert(0, reference_runner.REFERENCE_ROOT)
    import PIL.Image as Image
    from datasets import kitti_dataset
    rng = np.random.default_rng(7)
    blob = {}
    for name, (N, H, W) in {"small": (3, 32, 64), "odd_tiles": (2, 96, 160), "tiny": (1, 8, 16)}.items():
        ds = kitti_dataset.KITTIRAWDataset("/nonexistent", ["x 0 l"], H, W, [0], 4, is_train=False)
        # a smooth image with noise and saturated patches (exercises clip8 at both ends)
        frames = np.zeros((N, H, W, 3), dtype=np.uint8)
        for n in range(N):
            base = rng.integers(0, 256, (H // 4 + 1, W // 4 + 1, 3)).astype(np.float32)
            img = np.kron(base, np.ones((4, 4, 1), dtype=np.float32))[:H, :W]
            img += rng.normal(0, 25, img.shape)
            img[: H // 4, : W // 4] = 255
            img[-H // 4:, -W // 4:] = 0
            img[H // 2, :] = rng.integers(0, 2, (W, 3)) * 255          # 0/255 alternation: ringing beyond [0,255]
            frames[n] = np.clip(img, 0, 255).astype(np.uint8)
        blob[name + "|frames"] = frames
        for n in range(N):
            inputs = {("color", 0, -1): Image.fromarray(frames[n])}
            ds.preprocess(inputs, (lambda x: x))
            for s in range(4):
                blob["%s|ref|%d|%d" % (name, n, s)] = inputs[("color", 0, s)].numpy()
    path = os.path.join(HERE, "aux", "pyramid.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KB", "PIL", Image.__version__, "torch", torch.__version__)


if __name__ == "__main__":
    main()
