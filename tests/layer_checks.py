"""Layer-level drop-ins (layers.py signatures) against the float64 CPU oracle, on any device."""
import torch

import common
from oracle import photometric_oracle as po
from ssde_b200 import layers as L
from ssde_b200 import synthetic


def _leaf(t, dev):
    return t.to(dev).clone().requires_grad_(True), t.double().clone().requires_grad_(True)


def run(device, B=2, H=16, W=24):
    dev = torch.device(device)
    g = torch.Generator().manual_seed(0)
    err = common.rel_err

    # disp_to_depth (layers.py:16-25)
    disp, d64 = _leaf(torch.rand(B, 1, H, W, generator=g), dev)
    sc, dep = L.disp_to_depth(disp, 0.1, 100.0)
    sc64, dep64 = po.disp_to_depth(d64, 0.1, 100.0)
    (sc.sum() + (dep * dep).sum()).backward()
    (sc64.sum() + (dep64 * dep64).sum()).backward()
    assert err(dep.detach().cpu(), dep64.detach()) < 1e-6
    assert err(sc.detach().cpu(), sc64.detach()) < 1e-6
    assert err(disp.grad.cpu(), d64.grad) < 1e-5

    # BackprojectDepth + Project3D (layers.py:139-193)
    intr = synthetic.intrinsics(B, H, W, 1)
    inv_K, K = intr[("inv_K", 0)], intr[("K", 0)]
    T0 = synthetic.pose_matrix(torch.randn(B, 3, generator=g) * 0.01, torch.randn(B, 3, generator=g) * 0.05, False)
    depth, dp64 = _leaf(1 + 4 * torch.rand(B, 1, H, W, generator=g), dev)
    T, T64 = _leaf(T0, dev)
    cam = L.BackprojectDepth(B, H, W).to(dev)(depth, inv_K.to(dev))
    grid = L.Project3D(B, H, W).to(dev)(cam, K.to(dev), T)
    w = torch.randn(grid.shape, generator=g)
    (grid * w.to(dev)).sum().backward()
    cam64 = po.backproject(dp64, inv_K.double())
    grid64 = po.project(cam64, K.double(), T64, H, W)
    (grid64 * w.double()).sum().backward()
    assert cam.shape == (B, 4, H * W) and grid.shape == (B, H, W, 2)
    assert err(cam.detach().cpu(), cam64.detach()) < 1e-6
    assert err(grid.detach().cpu(), grid64.detach()) < 1e-5
    assert err(depth.grad.cpu(), dp64.grad) < 1e-4
    assert err(T.grad.cpu(), T64.grad) < 1e-4

    # SSIM (layers.py:218-248)
    x, x64 = _leaf(torch.rand(B, 3, H, W, generator=g), dev)
    y, y64 = _leaf(torch.rand(B, 3, H, W, generator=g), dev)
    w = torch.randn(B, 3, H, W, generator=g)
    out = L.SSIM().to(dev)(x, y)
    (out * w.to(dev)).sum().backward()
    out64 = po.ssim(x64, y64)
    (out64 * w.double()).sum().backward()
    assert err(out.detach().cpu(), out64.detach()) < 1e-4
    assert err(x.grad.cpu(), x64.grad) < 2e-4 and err(y.grad.cpu(), y64.grad) < 2e-4

    # get_smooth_loss (layers.py:202-215)
    dsp, dsp64 = _leaf(torch.rand(B, 1, H, W, generator=g), dev)
    img, img64 = _leaf(torch.rand(B, 3, H, W, generator=g), dev)
    sl = L.get_smooth_loss(dsp, img)
    (sl * 3).backward()
    sl64 = po.smooth_loss(dsp64, img64)
    (sl64 * 3).backward()
    assert sl.dim() == 0
    assert err(sl.detach().cpu(), sl64.detach()) < 1e-5
    assert err(dsp.grad.cpu(), dsp64.grad) < 1e-4 and err(img.grad.cpu(), img64.grad) < 1e-4

    # transformation_from_parameters (layers.py:28-103)
    for invert in (False, True):
        aa, aa64 = _leaf(torch.randn(B, 1, 3, generator=g) * 0.1, dev)
        tr, tr64 = _leaf(torch.randn(B, 1, 3, generator=g), dev)
        w = torch.randn(B, 4, 4, generator=g)
        M = L.transformation_from_parameters(aa, tr, invert)
        M64 = po.transformation_from_parameters(aa64, tr64, invert)
        (M * w.to(dev)).sum().backward()
        (M64 * w.double()).sum().backward()
        assert err(M.detach().cpu(), M64.detach()) < 1e-6
        assert err(aa.grad.cpu(), aa64.grad) < 1e-4 and err(tr.grad.cpu(), tr64.grad) < 1e-5

    # F.interpolate bilinear, align_corners=False (trainer.py:474-475, :574-576) incl. non-integer ratios
    for (h, w, Hh, Ww) in ((H // 2, W // 2, H, W), (H // 8, W // 8, H, W), (5, 7, 16, 24), (H, W, H, W)):
        xs, xs64 = _leaf(torch.rand(B, 2, h, w, generator=g), dev)
        wgt = torch.randn(B, 2, Hh, Ww, generator=g)
        up = L.interpolate_bilinear(xs, [Hh, Ww])
        up64 = torch.nn.functional.interpolate(xs64, [Hh, Ww], mode="bilinear", align_corners=False)
        (up * wgt.to(dev)).sum().backward()
        (up64 * wgt.double()).sum().backward()
        assert up.shape == (B, 2, Hh, Ww)
        # fp32 source coordinates (ATen computes them in fp32 too): ~ulp(scale) * size * slope
        assert err(up.detach().cpu(), up64.detach()) < 3e-6
        assert err(xs.grad.cpu(), xs64.grad) < 1e-5

    # nn.BCELoss()(mask, ones) (trainer.py:582)
    from ssde_b200 import functional as Fn
    m, m64 = _leaf(torch.sigmoid(3 * torch.randn(B, 2, H, W, generator=g)), dev)
    bce = Fn.bce_against_ones(m)
    bce64 = torch.nn.functional.binary_cross_entropy(m64, torch.ones_like(m64))
    (bce * 0.2).backward()
    (bce64 * 0.2).backward()
    assert bce.dim() == 0 and err(bce.detach().cpu(), bce64.detach()) < 1e-6
    assert err(m.grad.cpu(), m64.grad) < 1e-5


def depth_metrics(device):
    """functional.depth_metrics / trainer_hooks.compute_depth_losses against the float64 oracle on the
    reference-generated fixture (trainer.py:624-652)."""
    import os
    import numpy as np
    from types import SimpleNamespace
    from oracle import photometric_oracle as po
    from ssde_b200 import functional as Fn, trainer_hooks
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux", "depth_metrics.npz"))
    pred, gt = torch.from_numpy(z["pred"]), torch.from_numpy(z["gt"])
    want = po.compute_depth_losses(pred, gt, torch.float64)
    got = Fn.depth_metrics(pred.to(device), gt.to(device)).cpu().double()
    rel = ((got - want).abs() / want.abs()).max().item()
    assert rel < 2e-5, (got.tolist(), want.tolist())
    assert torch.allclose(got, torch.from_numpy(z["ref_f64"]), rtol=2e-5, atol=0)      # the reference itself
    # drop-in method: same dictionary entries as Trainer.compute_depth_losses
    losses = {}
    ns = SimpleNamespace(depth_metric_names=list(trainer_hooks.DEPTH_METRIC_NAMES))
    trainer_hooks.compute_depth_losses(ns, {"depth_gt": gt.to(device)}, {("depth", 0, 0): pred.to(device)}, losses)
    assert sorted(losses) == sorted(trainer_hooks.DEPTH_METRIC_NAMES)
    assert abs(float(losses["de/abs_rel"]) - want[0].item()) / want[0].item() < 2e-5
    # a ground truth at another resolution / with nothing inside the crop must not crash
    empty = Fn.depth_metrics(pred.to(device), torch.zeros_like(gt).to(device)).cpu()
    assert empty.shape == (7,)
