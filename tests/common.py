"""Shared helpers of the test-suite: golden fixture loading, running the product path, comparing."""
from __future__ import annotations

import glob
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import ssde_b200  # noqa: E402
from ssde_b200 import synthetic, trainer_hooks  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _key(s):
    k = json.loads(s)
    return tuple(k) if isinstance(k, list) else k


def load_golden(name):
    """-> (variant, opt, inputs, outputs, ref32, ref64, noise_seed); tensors on CPU, fp32."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    opt = SimpleNamespace(**meta["opt"])
    inputs = {_key(k[3:]): torch.from_numpy(z[k]) for k in meta["input_keys"]}
    outputs = {_key(k[4:]): torch.from_numpy(z[k]) for k in meta["output_keys"]}
    ref32 = {k.split("|", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ref_f32|")}
    ref64 = {k.split("|", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ref_f64|")}
    return meta["variant"], opt, inputs, outputs, ref32, ref64, meta["noise_seed"]


def run_product(opt, inputs, outputs, variant="trainer", device="cuda", noise_seed=None, sources=(-1, 1),
                want_grad=True, extra_opt=None):
    """Run the drop-in trainer methods (the product path) on ``device`` and collect what the
    reference would expose.  ``noise_seed``: host-generated tie-break noise seeded exactly like
    the reference run (trainer.py:594); None -> in-kernel Philox."""
    opt = SimpleNamespace(**vars(opt))
    opt.pml_variant = variant
    opt.pml_sources = list(sources)
    opt.pml_noise = "host" if noise_seed is not None else "philox"
    opt.pml_emit_depth = "all"
    opt.pml_emit_warped = True
    for k, v in (extra_opt or {}).items():
        setattr(opt, k, v)
    dev = torch.device(device)
    inp = {k: v.to(dev) for k, v in inputs.items()}
    if variant == "gru":
        inp = synthetic.to_sequence_layout(inp, opt.len_sequence)
    out = {k: v.to(dev).clone() for k, v in outputs.items()}
    flat = [k for k in out if isinstance(k, tuple) and k[0] == "predictive_mask"]
    if flat:   # the trainers read outputs["predictive_mask"]["disp", s] (trainer.py:573)
        out["predictive_mask"] = {("disp", k[1]): out.pop(k) for k in flat}
    leaves = {}
    if want_grad:
        for s in opt.scales:
            out[("disp", s)].requires_grad_(True)
            leaves["grad_disp/%d" % s] = out[("disp", s)]
            if "predictive_mask" in out:
                out["predictive_mask"][("disp", s)].requires_grad_(True)
                leaves["grad_mask/%d" % s] = out["predictive_mask"][("disp", s)]
        for f in sources:
            if f != "s":
                out[("cam_T_cam", 0, f)].requires_grad_(True)
                leaves["grad_T/%s" % f] = out[("cam_T_cam", 0, f)]
                if variant in ("trainer", "fusion") and getattr(opt, "pose_model_type", "") == "posecnn":   # trainer.py:490-499
                    for k in ("axisangle", "translation"):
                        out[(k, 0, f)].requires_grad_(True)
                        leaves["grad_%s/%s" % (k, f)] = out[(k, 0, f)]
    ns = SimpleNamespace(opt=opt, device=dev, num_scales=len(opt.scales))
    if noise_seed is not None:
        torch.manual_seed(noise_seed)
    trainer_hooks.generate_images_pred(ns, inp, out)
    losses = trainer_hooks.compute_losses(ns, inp, out)
    res = {"loss": losses["loss"].detach().cpu()}
    for s in opt.scales:
        res["loss/%d" % s] = losses["loss/%d" % s].detach().cpu()
        res["argmin/%d" % s] = out[("argmin", s)].cpu()
        if ("depth", 0, s) in out:
            res["depth/%d" % s] = out[("depth", 0, s)].detach().cpu()
        k = "identity_selection/%d" % s
        if k in out:
            res[k] = out[k].cpu()
        for f in sources:
            if ("color", f, s) in out:
                res["color/%s/%d" % (f, s)] = out[("color", f, s)].detach().cpu()
    if want_grad:
        losses["loss"].backward()
        for name, leaf in leaves.items():
            res[name] = leaf.grad.detach().cpu() if leaf.grad is not None else torch.zeros_like(leaf).cpu()
    return res


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| -- the 'relative' of north_star's gradient tolerance (per tensor)."""
    a, b = a.double(), b.double()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def argmin_report(got: torch.Tensor, ref: torch.Tensor, margin: torch.Tensor = None, eps=2e-6):
    """-> (n mismatches, n mismatches on pixels whose float64 decision margin exceeds eps)."""
    mism = got.long() != ref.long()
    n = int(mism.sum())
    if margin is None:
        return n, n
    return n, int((mism & (margin > eps)).sum())
