"""Run the UNMODIFIED reference loss path on CPU (build container only).  TEST INFRASTRUCTURE.

``/root/reference`` exists only in the build container, never on the GPU box, so this module is
used solely by ``tests/golden/make_golden.py`` (fixture generation) and by
``tests/test_install_reference.py`` (skipped when the reference tree is absent).

How the reference is driven: ``Trainer.__init__`` cannot run (hard-coded
``cuda:N`` devices, KITTI on disk: trainer.py:44,67), and ``--no_cuda`` is parsed
(options.py:216) but never read, so the three loss methods are called *unbound* on a
``SimpleNamespace`` that carries exactly the attributes they touch.  The imports the trainers make
but the path never uses (tensorboardX, GPUtil, IPython, skimage) are stubbed.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import torch

REFERENCE_ROOT = os.environ.get("PML_REFERENCE_ROOT", "/root/reference")

_MODULES = {"trainer": "trainer", "fusion": "trainer_fusion", "fusion_v3": "trainer_fusion_v3",
            "gru": "trainer_gru"}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "layers.py"))


def _install_stubs():
    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    stub("tensorboardX", SummaryWriter=object)
    stub("GPUtil", showUtilization=lambda *a, **k: None)
    stub("IPython", embed=lambda *a, **k: None)
    stub("skimage")
    stub("skimage.transform", resize=None)
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    try:
        import PIL.Image as _I  # Pillow >= 10 dropped ANTIALIAS (datasets/mono_dataset.py uses it)
        if not hasattr(_I, "ANTIALIAS"):
            _I.ANTIALIAS = _I.LANCZOS
    except Exception:
        pass


def load(variant="trainer"):
    """Import /root/reference/<trainer module> and return (module, layers module)."""
    assert available(), "reference tree not present"
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    layers = importlib.import_module("layers")
    mod = importlib.import_module(_MODULES[variant])
    return mod, layers


class _RecordMin:
    """Record the index tensor of every ``torch.min(x, dim=1)`` the reference performs."""
    def __init__(self):
        self.idxs = []
        self._orig = torch.min

    def __enter__(self):
        orig = self._orig
        rec = self.idxs

        def wrapped(*a, **k):
            r = orig(*a, **k)
            if isinstance(r, tuple) or hasattr(r, "indices"):
                rec.append(r[1].detach().clone())
            return r
        torch.min = wrapped
        return self

    def __exit__(self, *exc):
        torch.min = self._orig


def run(opt, inputs, outputs, variant="trainer", noise_seed=0, dtype=torch.float32, want_grad=True):
    """Execute reference generate_images_pred + compute_losses (+ backward of losses['loss']).

    ``inputs``/``outputs`` follow the reference schema (flat 3-tuple keys; for variant="gru" the
    4-tuple sequence layout).  The global CPU generator is seeded with ``noise_seed`` right before
    the call so the tie-break noise equals ``synthetic.draw_noise(..., seed=noise_seed)``."""
    mod, layers = load(variant)
    Trainer = mod.Trainer
    opt = SimpleNamespace(**vars(opt))

    def conv(t):
        t = t.detach().cpu()
        return t.to(dtype).clone() if t.is_floating_point() else t.clone()

    inp = {k: conv(v) for k, v in inputs.items()}
    out = {k: conv(v) for k, v in outputs.items()}
    flat = [k for k in out if isinstance(k, tuple) and k[0] == "predictive_mask"]
    if flat:   # trainer.py:573 reads outputs["predictive_mask"]["disp", scale]
        out["predictive_mask"] = {("disp", k[1]): out.pop(k) for k in flat}
    n_flat = out[("disp", opt.scales[0])].shape[0]

    ns = SimpleNamespace(opt=opt, device=torch.device("cpu"), num_scales=len(opt.scales))
    ns.ssim = layers.SSIM().to(dtype)
    ns.backproject_depth, ns.project_3d = {}, {}
    for s in opt.scales:
        h, w = opt.height // 2 ** s, opt.width // 2 ** s
        ns.backproject_depth[s] = layers.BackprojectDepth(n_flat, h, w).to(dtype)
        ns.project_3d[s] = layers.Project3D(n_flat, h, w).to(dtype)
    ns.compute_reprojection_loss = lambda pred, target: Trainer.compute_reprojection_loss(ns, pred, target)

    leaves = {}
    if want_grad:
        for s in opt.scales:
            out[("disp", s)].requires_grad_(True)
            leaves["grad_disp/{}".format(s)] = out[("disp", s)]
            if "predictive_mask" in out:
                out["predictive_mask"][("disp", s)].requires_grad_(True)
                leaves["grad_mask/{}".format(s)] = out["predictive_mask"][("disp", s)]
        for f in (-1, 1):
            out[("cam_T_cam", 0, f)].requires_grad_(True)
            leaves["grad_T/{}".format(f)] = out[("cam_T_cam", 0, f)]
            if variant in ("trainer", "fusion") and getattr(opt, "pose_model_type", "") == "posecnn":   # trainer.py:490-499, trainer_fusion.py:446-456
                for k in ("axisangle", "translation"):
                    out[(k, 0, f)].requires_grad_(True)
                    leaves["grad_{}/{}".format(k, f)] = out[(k, 0, f)]

    torch.manual_seed(noise_seed)
    # trainer.py:582 builds its BCE target with torch.ones(...) in the DEFAULT dtype, so the float64
    # run of the predictive-mask branch needs the default switched (that branch draws no noise, so
    # the generator stream of the other cases is not affected)
    old_default = torch.get_default_dtype()
    if "predictive_mask" in out or (getattr(opt, "pose_model_type", "") == "posecnn" and opt.disable_automasking):
        torch.set_default_dtype(dtype)     # posecnn: layers.get_translation_matrix uses torch.zeros (layers.py:51)
    try:
        with _RecordMin() as rec:
            Trainer.generate_images_pred(ns, inp, out)
            losses = Trainer.compute_losses(ns, inp, out)
    finally:
        torch.set_default_dtype(old_default)
    res = {"loss": losses["loss"].detach()}
    for i, s in enumerate(opt.scales):
        res["loss/{}".format(s)] = losses["loss/{}".format(s)].detach()
        if i < len(rec.idxs):
            res["argmin/{}".format(s)] = rec.idxs[i]
        key = "identity_selection/{}".format(s)
        if key in out:
            res[key] = out[key].detach()
        res["depth/{}".format(s)] = out[("depth", 0, s)].detach()
        for f in (-1, 1):
            res["color/{}/{}".format(f, s)] = out[("color", f, s)].detach()
    if want_grad:
        losses["loss"].backward()
        for name, leaf in leaves.items():
            res[name] = leaf.grad.detach() if leaf.grad is not None else torch.zeros_like(leaf)
    return res


def run_depth_losses(depth_pred, depth_gt, dtype=torch.float32):
    """Execute the reference's Trainer.compute_depth_losses (trainer.py:624-652) unbound."""
    import numpy as np
    mod, _ = load("trainer")
    ns = SimpleNamespace(depth_metric_names=["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"])
    losses = {}
    mod.Trainer.compute_depth_losses(ns, {"depth_gt": depth_gt.to(dtype)}, {("depth", 0, 0): depth_pred.to(dtype)}, losses)
    return torch.tensor([float(np.asarray(losses[k])) for k in ns.depth_metric_names], dtype=torch.float64)
