"""Generate tests/golden/aux/depth_metrics.npz by executing the UNMODIFIED reference's
Trainer.compute_depth_losses (trainer.py:624-652) in the build container (needs /root/reference).

    python tests/golden/make_golden_depth.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_runner  # noqa: E402


def make_inputs(seed=7, B=2, H=48, W=160):
    g = torch.Generator().manual_seed(seed)
    # smooth-ish predicted depth in [0.5, 60] (some values beyond the 80 m clamp after scaling)
    low = torch.rand(B, 1, H // 8, W // 8, generator=g)
    pred = torch.nn.functional.interpolate(low, [H, W], mode="bilinear", align_corners=False) * 59.5 + 0.5
    # sparse LiDAR-like ground truth at 375x1242: ~6 % of the pixels carry a depth in [1, 90]
    gt = torch.rand(B, 1, 375, 1242, generator=g) * 89.0 + 1.0
    keep = torch.rand(B, 1, 375, 1242, generator=g) < 0.06
    return pred, (gt * keep).float()


def main():
    assert reference_runner.available(), "needs the reference tree"
    pred, gt = make_inputs()
    r32 = reference_runner.run_depth_losses(pred, gt, torch.float32)
    r64 = reference_runner.run_depth_losses(pred, gt, torch.float64)
    path = os.path.join(HERE, "aux", "depth_metrics.npz")
    np.savez_compressed(path, pred=pred.numpy(), gt=gt.numpy(), ref_f32=r32.numpy(), ref_f64=r64.numpy())
    print("%s %.1f KB" % (path, os.path.getsize(path) / 1024), r32.tolist(), r64.tolist())


if __name__ == "__main__":
    main()
