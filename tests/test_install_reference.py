"""``trainer_hooks.install()`` against the REAL reference modules (build container only; skipped where
/root/reference is absent, e.g. on the GPU box).  The reference's own ``Trainer.generate_images_pred`` /
``compute_losses`` / ``log`` are called unbound on a SimpleNamespace before and after patching
(trainer.py:24,465-622,666-698); the kernels run on the host-thread emulation build."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

import common
from oracle import reference_runner
from ssde_b200 import synthetic, trainer_hooks

pytestmark = pytest.mark.skipif(not reference_runner.available(), reason="reference tree not present")

PATCHED = ("generate_images_pred", "compute_reprojection_loss", "compute_losses", "compute_depth_losses", "log")
SYMBOLS = ("BackprojectDepth", "Project3D", "SSIM", "disp_to_depth", "get_smooth_loss", "transformation_from_parameters")


class _Writer:
    """Stand-in for tensorboardX.SummaryWriter: records what Trainer.log writes."""
    def __init__(self):
        self.scalars, self.images = {}, {}

    def add_scalar(self, name, value, step):
        self.scalars[name] = float(value)

    def add_image(self, name, img, step):
        assert isinstance(img, torch.Tensor) and img.dim() == 3, name
        self.images[name] = img.detach().float().cpu()


def _namespace(mod, layers, opt, n_flat, dtype=torch.float32):
    ns = SimpleNamespace(opt=opt, device=torch.device("cpu"), num_scales=len(opt.scales), step=0)
    ns.ssim = layers.SSIM()
    ns.backproject_depth, ns.project_3d = {}, {}
    for s in opt.scales:
        h, w = opt.height // 2 ** s, opt.width // 2 ** s
        ns.backproject_depth[s] = layers.BackprojectDepth(n_flat, h, w)
        ns.project_3d[s] = layers.Project3D(n_flat, h, w)
    ns.compute_reprojection_loss = lambda pred, target: mod.Trainer.compute_reprojection_loss(ns, pred, target)
    ns.writers = {"train": _Writer()}
    return ns


@pytest.mark.parametrize("variant", ["trainer", "fusion", "fusion_v3", "gru"])
def test_install_patches_the_reference_trainer(emu_lib, variant):
    mod, layers = reference_runner.load(variant)
    cls = mod.Trainer
    saved = {k: cls.__dict__[k] for k in PATCHED if k in cls.__dict__}
    saved_syms = {k: getattr(mod, k) for k in SYMBOLS if hasattr(mod, k)}
    B, H, W = (3, 32, 64) if variant == "gru" else (2, 32, 64)
    kw = dict(batch_size=1, len_sequence=3) if variant == "gru" else dict(batch_size=B)
    opt = synthetic.make_options(H, W, **kw)
    inputs, outputs = synthetic.make_batch(B, H, W, seed=41, full_res_disp=(variant == "fusion"))
    inp_ref = synthetic.to_sequence_layout(inputs, opt.len_sequence) if variant == "gru" else inputs
    noise_seed = 9

    def run(patched):
        o = SimpleNamespace(**vars(opt))
        if patched:
            o.pml_noise = "host"          # tie-break noise from the CPU generator, exactly like trainer.py:594
        ns = _namespace(mod, layers, o, B)
        inp = {k: v.clone() for k, v in inp_ref.items()}
        out = {k: v.clone() for k, v in outputs.items()}
        for s in opt.scales:
            out[("disp", s)].requires_grad_(True)
        for f in (-1, 1):
            out[("cam_T_cam", 0, f)].requires_grad_(True)
        torch.manual_seed(noise_seed)
        cls.generate_images_pred(ns, inp, out)
        # the reference moves every entry of both dictionaries between the two calls (trainer.py:369-373)
        for k, v in list(out.items()):
            out[k] = v.to(ns.device)
        losses = cls.compute_losses(ns, inp, out)
        losses["loss"].backward()
        if variant != "gru":   # trainer_gru.py logs through 4-tuple keys of its own dataset; same outputs entries
            cls.log(ns, "train", inp, out, losses)
        return ns, out, losses

    try:
        ns_ref, out_ref, loss_ref = run(False)
        patched_cls = trainer_hooks.install(mod)
        assert patched_cls is cls
        for sym in saved_syms:
            assert getattr(mod, sym).__module__.startswith(("ssde_b200", "self-supervised")), sym
        ns_new, out_new, loss_new = run(True)
    finally:
        for k, v in saved.items():
            setattr(cls, k, v)
        for k, v in saved_syms.items():
            setattr(mod, k, v)

    assert ns_new.opt.pml_variant == variant
    for s in opt.scales:
        k = "loss/%d" % s
        assert common.rel_err(loss_new[k].detach(), loss_ref[k].detach()) < 1e-4, k    # two fp32 evaluations at 32x64
        a, b = out_new[("disp", s)].grad, out_ref[("disp", s)].grad
        assert (a - b).norm() / b.norm() < 2e-2, "grad_disp/%d" % s                      # near-tie flips at 32x64
    assert common.rel_err(loss_new["loss"].detach(), loss_ref["loss"].detach()) < 1e-4
    for f in (-1, 1):
        a, b = out_new[("cam_T_cam", 0, f)].grad, out_ref[("cam_T_cam", 0, f)].grad
        assert (a - b).norm() / b.norm() < 2e-2
    if variant != "gru":
        w_ref, w_new = ns_ref.writers["train"], ns_new.writers["train"]
        assert set(w_new.images) == set(w_ref.images) and set(w_new.scalars) == set(w_ref.scalars)
        for name, img in w_ref.images.items():
            got = w_new.images[name]
            assert got.shape == img.shape, name
            if name.startswith("color_pred_"):
                assert (got - img).abs().max() < 2e-4, name        # outputs[("color", f, 0)], trainer.py:679-682
            elif name.startswith("automask_"):
                assert (got != img).float().mean() < 5e-3, name    # identity_selection: equal up to near-tie flips


def test_disp_heads_on_the_reference_depth_decoder(emu_lib):
    """layers.install_disp_heads on the reference's own DepthDecoder (networks/depth_decoder.py): same outputs,
    same parameter names, gradients for every parameter."""
    import numpy as np
    from ssde_b200 import layers as L
    reference_runner.load("trainer")
    from networks.depth_decoder import DepthDecoder
    torch.manual_seed(0)
    enc_ch = np.array([16, 16, 24, 32, 48])        # a narrow encoder keeps the CPU run short
    dec = DepthDecoder(num_ch_enc=enc_ch)
    feats = [torch.randn(1, int(c), 32 >> i, 64 >> i) for i, c in enumerate(enc_ch)]
    want = {k: v.detach().clone() for k, v in dec(feats).items()}
    keys = list(dec.state_dict().keys())
    L.install_disp_heads(dec)
    assert list(dec.state_dict().keys()) == keys
    got = dec(feats)
    assert set(got) == set(want)
    for k in want:
        assert (got[k] - want[k]).abs().max() < 2e-6, k
    sum(v.mean() for v in got.values()).backward()
    assert all(p.grad is not None for p in dec.parameters())
