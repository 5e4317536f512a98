"""Drop-in replacements for the three loss methods of the reference ``Trainer`` classes.

    generate_images_pred(self, inputs, outputs)      trainer.py:465-515
    compute_reprojection_loss(self, pred, target)    trainer.py:517-529
    compute_losses(self, inputs, outputs) -> dict    trainer.py:531-622
    compute_depth_losses(self, inputs, outputs, losses)   trainer.py:624-652 (monitoring metrics)

and their copies in trainer_fusion.py:421-579, trainer_fusion_v3.py:447-590 and
trainer_gru.py:864-1023 (4-tuple sequence keys).  ``install(trainer_module)`` monkey-patches a
reference trainer module in place; the reference files stay untouched.

Same ``self`` contract as the reference: ``self.opt.{scales,height,width,min_depth,max_depth,
v1_multiscale,disable_automasking,avg_reprojection,predictive_mask,no_ssim,disparity_smoothness,
pose_model_type[,len_sequence]}``, ``self.device``, ``self.num_scales``.  Same dictionary schema
in and out: ``losses["loss/{s}"]``, ``losses["loss"]``, ``outputs["identity_selection/{s}"]``,
``outputs[("depth", 0, s)]``; the warped images ``outputs[("color", f, s)]`` are only written when
``self.opt.pml_emit_warped`` is set because nothing but the tensorboard image logger
(trainer.py:679-682) reads them.

Extra, optional knobs (all default to the reference's behaviour or cheaper equivalents):
    opt.pml_sources      ordered source frames; default [-1, 1] (hard-coded at trainer.py:482,550)
    opt.pml_variant      "trainer" | "fusion" | "fusion_v3" | "gru" (set by install())
    opt.pml_noise        "philox": tie-break noise drawn in-kernel (default);
                         "host": torch.randn on the CPU generator, exactly trainer.py:594-595
    opt.pml_emit_depth   "scale0" (default; compute_depth_losses reads only that) | "all" | "none"
    opt.pml_emit_warped  False (default) | True
    opt.pml_emit_selection  True (default): outputs["identity_selection/{s}"] written by compute_losses (one
                         launch for all scales) | "lazy" (set by install(): the patched Trainer.log fills it and the
                         warped scale-0 images right before the tensorboard logger reads them) | False
    opt.pml_kernel       "sweep" (default) | "cta" (first-generation kernel, cross-check only)
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from . import functional as _F
from . import layers as _L


def _opt(opt, name, default):
    return getattr(opt, name, default)


def _gather(inputs, key, n_seq, cache):
    """trainer_gru.py:890-899,943-957 concatenates the per-timestep tensors for every scale and frame; here they
    are handed to the kernels as a list of chunks and read in place (pml_segments) -- no torch.cat."""
    if key in cache:
        return cache[key]
    if n_seq and (key + (0,)) in inputs:
        val = [inputs[key + (i,)] for i in range(n_seq)]
    else:
        val = inputs[key]
    cache[key] = val
    return val


class _Pending:
    """Result of the fused call, parked in ``outputs`` between generate_images_pred and compute_losses.  The
    reference moves every entry of ``outputs`` with ``.to(device)`` in between (trainer.py:369-371)."""
    def __init__(self, res):
        self.res = res

    def to(self, *args, **kwargs):
        return self


_PENDING_KEY = "pml/pending"


def _run_fused(self, inputs, outputs):
    opt = self.opt
    # trainer.py:571: the predicted mask is only consulted when automasking is off (`elif`)
    pmask = bool(_opt(opt, "predictive_mask", False)) and bool(opt.disable_automasking)
    variant = _opt(opt, "pml_variant", "trainer")
    sources = list(_opt(opt, "pml_sources", [-1, 1]))
    n_seq = _opt(opt, "len_sequence", 0) if variant == "gru" else 0
    H, W = opt.height, opt.width
    scales = list(opt.scales)
    # trainer.py:490-499 and trainer_fusion.py:446-456 scale the translation by the mean inverse depth
    posecnn = (variant in ("trainer", "fusion") and _opt(opt, "pose_model_type", "") == "posecnn")
    per_scale_images = (opt.v1_multiscale and variant != "fusion")
    emit_depth = _opt(opt, "pml_emit_depth", "scale0")
    emit_warped = bool(_opt(opt, "pml_emit_warped", False))
    noise_mode = _opt(opt, "pml_noise", "philox")
    automask = not opt.disable_automasking
    n_id = 0 if not automask else (1 if opt.avg_reprojection else len(sources))
    cache: Dict = {}

    # groups of scales that share images / intrinsics / poses
    if per_scale_images or posecnn:
        groups = [[s] for s in scales]
    else:
        groups = [scales]

    # tie-break noise, drawn in the reference's order (one tensor per scale, trainer.py:592-595)
    noise = {}
    if automask and noise_mode == "host":
        for s in scales:
            h, w = (H // 2 ** s, W // 2 ** s) if per_scale_images else (H, W)
            B = outputs[("disp", s)].shape[0]
            noise[s] = torch.randn([B, n_id, h, w]).to(self.device)
    seed = 0
    if automask and noise_mode != "host":
        seed = _opt(opt, "pml_seed", None)
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())

    res = {"loss": {}, "terms": {}, "argmin": {}, "n_id": n_id, "total": None, "argmin_all": None, "sources": sources}
    for group in groups:
        src_scale = group[0] if per_scale_images else 0
        target = _gather(inputs, ("color", 0, src_scale), n_seq, cache)
        K = _gather(inputs, ("K", src_scale), n_seq, cache)
        inv_K = _gather(inputs, ("inv_K", src_scale), n_seq, cache)
        srcs = [_gather(inputs, ("color", f, src_scale), n_seq, cache) for f in sources]
        disps, colors, weights, fws, bces = [], [], [], [], []
        for s in group:
            disps.append(outputs[("disp", s)])
            if pmask:
                # trainer.py:573-583: resize the mask unless v1_multiscale, weight the reprojection
                # losses with it, and push it towards 1 with 0.2 * BCE(mask, ones)
                mask = outputs["predictive_mask"][("disp", s)]
                if not opt.v1_multiscale:
                    mask = _L.interpolate_bilinear(mask, [H, W])
                fws.append(mask)
                bces.append(0.2 * _F.bce_against_ones(mask))
            # trainer.py:547 smooths against the colour at the disparity's own scale;
            # trainer_fusion.py:504 against source_scale (its disparities are full-res)
            cs = src_scale if variant == "fusion" else s
            colors.append(_gather(inputs, ("color", 0, cs), n_seq, cache))
            weights.append(opt.disparity_smoothness / (2 ** s))
        Ts = []
        for f in sources:
            if f == "s":
                T = inputs["stereo_T"]
            else:
                T = outputs[("cam_T_cam", 0, f)]
            if posecnn and f != "s":
                # trainer.py:490-499: translation scaled by the mean inverse depth of this scale
                d = outputs[("disp", group[0])]
                if not per_scale_images and tuple(d.shape[2:]) != (H, W):
                    d = F.interpolate(d, [H, W], mode="bilinear", align_corners=False)
                lo, hi = 1.0 / opt.max_depth, 1.0 / opt.min_depth
                mean_inv_depth = (lo + (hi - lo) * d).mean(3, True).mean(2, True)
                T = _L.transformation_from_parameters(
                    outputs[("axisangle", 0, f)][:, 0],
                    outputs[("translation", 0, f)][:, 0] * mean_inv_depth[:, 0], f < 0)
            Ts.append(T)
        ed = [i for i, s in enumerate(group) if emit_depth == "all" or (emit_depth == "scale0" and s == 0)]
        ew = list(range(len(group))) if emit_warped else []
        if not ew:
            # outputs[("depth", 0, s)] (trainer.py:480) from the two small layer kernels instead of the
            # sweep: the sweep then stays on its specialised instantiation without by-product stores
            # (368 vs 418 us at the headline size); nothing back-propagates through this entry
            for i in ed:
                d = disps[i].detach()
                t0 = target[0] if isinstance(target, list) else target
                if d.shape[2:] != t0.shape[2:]:
                    d = _F.upsample_bilinear(d, t0.shape[2], t0.shape[3])     # trainer.py:474-475
                outputs[("depth", 0, group[i])] = _F.depth_from_disp(d, opt.min_depth, opt.max_depth)
            ed = []
        out = _F.photometric_loss(
            target, srcs, K, inv_K, Ts, disps, colors, smooth_weights=weights,
            min_depth=opt.min_depth, max_depth=opt.max_depth, no_ssim=opt.no_ssim,
            disable_automasking=opt.disable_automasking, avg_reprojection=opt.avg_reprojection,
            noise=[noise[s] for s in group] if noise else None, seed=seed + 7919 * scales.index(group[0]),
            emit_depth=ed, emit_warped=ew, frame_weights=fws if pmask else None,
            kernel=_opt(opt, "pml_kernel", "sweep"), total_div=len(scales), seed_device=_opt(opt, "pml_seed_device", None))
        if len(groups) == 1 and not pmask:
            res["total"], res["argmin_all"], res["loss_vec"] = out["total"], out["argmin_all"], out["loss"]
        for i, s in enumerate(group):
            res["loss"][s] = out["loss"][i] + bces[i] if pmask else out["loss"][i]
            res["terms"][s] = out["terms"][i]
            res["argmin"][s] = out["argmin"][i]
            if i in out["depth"]:
                outputs[("depth", 0, s)] = out["depth"][i]
            if i in out["warped"]:
                for fi, f in enumerate(sources):
                    outputs[("color", f, s)] = out["warped"][i][fi]
                    if automask:   # trainer.py:513-515
                        outputs[("color_identity", f, s)] = torch.cat(srcs[fi], 0) if isinstance(srcs[fi], list) else srcs[fi]
    return res


def ingest_colors(inputs, frame_ids, num_scales=4, device=None, non_blocking=True):
    """Device-side replacement of the loss-side half of ``MonoDataset.preprocess``
    (datasets/mono_dataset.py:99-111) + the fp32 upload of trainer.py:233-237.

    ``inputs[("color_u8", f)]``: uint8 [B,H,W,3] scale-0 frames as the decoder / PIL delivers them
    (host, ideally pinned, or already on the device) -- or all frames stacked in ``inputs["color_u8"]``
    [F,B,H,W,3] in the order of ``frame_ids``.  Fills ``inputs[("color", f, s)]`` for every
    scale with exactly the tensors the reference's DataLoader would have produced, from a quarter of
    the bytes on the PCIe link.  All frames go through ONE pyramid launch chain."""
    if "color_u8" in inputs:          # all frames in one tensor [F,B,H,W,3] (frame order = frame_ids): no cat
        st = inputs["color_u8"]
        dev = device if device is not None else st.device
        B = st.shape[1]
        stacked = st.to(dev, non_blocking=non_blocking).reshape((-1,) + tuple(st.shape[2:]))
    else:
        frames = [inputs[("color_u8", f)] for f in frame_ids]
        dev = device if device is not None else frames[0].device
        B = frames[0].shape[0]
        stacked = torch.cat([t.to(dev, non_blocking=non_blocking) for t in frames], 0)
    levels = _F.color_pyramid(stacked, num_scales)
    for i, f in enumerate(frame_ids):
        for s in range(num_scales):
            inputs[("color", f, s)] = levels[s][i * B:(i + 1) * B]
    return inputs


def generate_images_pred(self, inputs, outputs):
    """trainer.py:465-515.  Runs the fused sweep (warp + loss + adjoint) and parks the loss terms in
    ``outputs`` for :func:`compute_losses`; writes ``outputs[("depth", 0, s)]`` (and the warped images
    when requested).  Everything put into ``outputs`` answers ``.to(device)``, so the reference's blanket
    ``outputs[key] = ipt.to(device)`` loop (trainer.py:369-371) keeps working."""
    outputs[_PENDING_KEY] = _Pending(_run_fused(self, inputs, outputs))


def compute_reprojection_loss(self, pred, target):
    """trainer.py:517-529 on the layer-level kernels (used by callers outside the fused path)."""
    l1_loss = torch.abs(target - pred).mean(1, True)
    if self.opt.no_ssim:
        return l1_loss
    ssim = getattr(self, "ssim", None)
    ssim_loss = (ssim(pred, target) if ssim is not None else _L.SSIM()(pred, target)).mean(1, True)
    return 0.85 * ssim_loss + 0.15 * l1_loss


def compute_losses(self, inputs, outputs):
    """trainer.py:531-622: returns ``{"loss/{s}": ..., "loss": ...}`` and fills
    ``outputs["identity_selection/{s}"]``."""
    pending = outputs.pop(_PENDING_KEY, None)
    res = pending.res if isinstance(pending, _Pending) else _run_fused(self, inputs, outputs)
    losses = {}
    scales = list(self.opt.scales)
    for s in scales:
        losses["loss/{}".format(s)] = res["loss"][s]
        outputs[("argmin", s)] = res["argmin"][s]
    if res["total"] is not None and self.num_scales == len(scales):
        losses["loss"] = res["total"]            # (loss_0 + loss_1 + ...) / num_scales, written by the library
    else:
        total = 0
        for s in scales:
            total = total + res["loss"][s]
        losses["loss"] = total / self.num_scales
    mode = _opt(self.opt, "pml_emit_selection", True)
    if not self.opt.disable_automasking:
        if mode == "lazy":
            outputs["pml/selection"] = _Pending((res["argmin"], res["n_id"]))
        elif mode:
            _fill_selection(outputs, res["argmin"], res["n_id"], scales, res["argmin_all"])
    return losses


def _fill_selection(outputs, argmin, n_id, scales, argmin_all=None):
    """outputs["identity_selection/{s}"] = (idxs > n_id - 1).float() (trainer.py:606-608), one launch for all scales."""
    if argmin_all is None:   # scales evaluated by separate calls (v1_multiscale: different resolutions)
        for s in scales:
            outputs["identity_selection/{}".format(s)] = _F.selection_masks(argmin[s].unsqueeze(0), n_id)[0]
        return
    masks = _F.selection_masks(argmin_all, n_id)
    for i, s in enumerate(scales):
        outputs["identity_selection/{}".format(s)] = masks[i]


def materialize_logged_outputs(self, inputs, outputs):
    """What ``Trainer.log`` reads beyond the losses (trainer.py:676-698) and the fused path does not produce by
    default: ``outputs[("color", f, 0)]`` -- the warped scale-0 images, trainer.py:679-682 -- and
    ``outputs["identity_selection/{s}"]``.  Computed on demand (a forward-only sweep of scale 0 with the
    by-product stores enabled), i.e. only on the steps that log."""
    opt = self.opt
    sel = outputs.pop("pml/selection", None)
    if isinstance(sel, _Pending):
        argmin, n_id = sel.res
        _fill_selection(outputs, argmin, n_id, list(opt.scales))
    sources = list(_opt(opt, "pml_sources", [-1, 1]))
    if all(("color", f, 0) in outputs for f in sources) or 0 not in opt.scales:
        return
    with torch.no_grad():
        o2 = type(opt)(**vars(opt)) if hasattr(opt, "__dict__") else opt
        o2.scales, o2.pml_emit_warped, o2.pml_emit_depth, o2.pml_noise = [0], True, "none", "philox"
        shim = _Shim(o2, self.device, self.num_scales)
        tmp = {k: v for k, v in outputs.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam", "axisangle", "translation")}
        if "predictive_mask" in outputs:
            tmp["predictive_mask"] = outputs["predictive_mask"]
        _run_fused(shim, inputs, tmp)
        for f in sources:
            if ("color", f, 0) in tmp:
                outputs[("color", f, 0)] = tmp[("color", f, 0)]


class _Shim:
    def __init__(self, opt, device, num_scales):
        self.opt, self.device, self.num_scales = opt, device, num_scales


class GraphedLoss:
    """Opt-in CUDA-graph execution of the drop-in pair ``generate_images_pred`` + ``compute_losses``.

    The forward of the fused loss (input pyramid when uint8 frames are given, identity / smoothness sweeps, the
    fused sweep with its adjoint, the reductions, ``outputs[("depth", 0, 0)]``) is captured once per slot on
    STATIC input tensors and replayed with one ``cudaGraphLaunch`` per step; the backward stays the single eager
    ``pml_scale_grads`` launch.  Host cost per step: one graph launch instead of ~10 kernel launches, ~15
    allocations and the Python around them.

        runner = GraphedLoss(trainer_like)                 # .opt, .device, .num_scales like the reference Trainer
        slot = runner.capture(inputs, outputs)             # these tensors become the static slot: refill them IN PLACE
        ...
        losses = slot.replay()                             # {"loss", "loss/{s}"}; differentiable w.r.t. the slot's
        losses["loss"].backward()                          #   outputs[("disp", s)] / ("cam_T_cam", 0, f) tensors
        # or, with tensors that change address every step (network outputs):
        losses = runner(inputs, outputs)                   # copies them into slot 0 (one multi-tensor copy), replays

    Restrictions: one group of scales (no v1_multiscale / posecnn / predictive mask), in-kernel tie-break noise
    (a device-side counter advances the Philox seed on every replay), fixed shapes per slot."""

    class Slot:
        def __init__(self, runner, inputs, outputs):
            self.runner, self.inputs, self.outputs = runner, inputs, outputs
            self.graph, self.node, self.vec, self.total, self.static_diff = None, None, None, None, None

        def replay(self, diff_tensors=None):
            if diff_tensors is None:
                diff_tensors = self.static_diff
            vec, total = _GraphedFn.apply(self, *diff_tensors)
            losses = {"loss/{}".format(s): vec[i] for i, s in enumerate(self.runner.scales)}
            losses["loss"] = total
            return losses

    def __init__(self, trainer_like):
        self.owner = trainer_like
        opt = trainer_like.opt
        if opt.v1_multiscale and _opt(opt, "pml_variant", "trainer") != "fusion":
            raise ValueError("GraphedLoss: v1_multiscale evaluates every scale by its own call; not supported")
        if _opt(opt, "pose_model_type", "") == "posecnn" or (_opt(opt, "predictive_mask", False) and opt.disable_automasking):
            raise ValueError("GraphedLoss: posecnn / predictive_mask need eager glue around the fused call; not supported")
        self.scales = list(opt.scales)
        self.sources = list(_opt(opt, "pml_sources", [-1, 1]))
        self.slots = []

    def _diff_keys(self):
        return [("disp", s) for s in self.scales] + [("cam_T_cam", 0, f) for f in self.sources if f != "s"]

    def capture(self, inputs, outputs):
        from types import SimpleNamespace
        own = self.owner
        dev = own.device
        o = SimpleNamespace(**vars(own.opt))
        o.pml_noise, o.pml_emit_warped, o.pml_emit_selection = "philox", False, False
        seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        o.pml_seed_device, o.pml_seed = seed_dev, int(torch.empty((), dtype=torch.int64).random_().item())
        shim = _Shim(o, dev, own.num_scales)
        slot = GraphedLoss.Slot(self, inputs, outputs)
        slot.seed_dev = seed_dev
        slot.static_diff = [outputs[k] for k in self._diff_keys()]
        frames = [0] + [f for f in self.sources]

        def body():
            inp = dict(inputs)
            if "color_u8" in inp or ("color_u8", 0) in inp:
                ingest_colors(inp, frames, own.num_scales, device=dev)
            out = dict(outputs)
            for k in self._diff_keys():
                out[k] = outputs[k].detach().requires_grad_(True)
            res = _run_fused(shim, inp, out)
            seed_dev.add_(0x9E3779B97F4A7C15 - (1 << 64))      # next replay: another Philox stream
            return res, out
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):          # warm-up: module load, launch plans, allocator
                body()
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                res, out = body()
        torch.cuda.current_stream(dev).wait_stream(side)
        if res["total"] is None:
            raise ValueError("GraphedLoss: this configuration does not run as one fused call")
        slot.graph, slot.res, slot.captured_out = g, res, out
        slot.node = res["total"].grad_fn      # the _PhotometricLoss node of the capture: owns the static gradient buffers
        slot.vec, slot.total = res["loss_vec"].detach(), res["total"].detach()
        if ("depth", 0, 0) in out:
            outputs[("depth", 0, 0)] = out[("depth", 0, 0)]
        for s in self.scales:
            outputs[("argmin", s)] = res["argmin"][s]
        self.slots.append(slot)
        return slot

    def __call__(self, inputs, outputs):
        """Drop-in form: ``inputs`` / ``outputs`` of this step are copied into slot 0 (captured on first use)."""
        if not self.slots:
            static_in = {k: v.detach().clone() for k, v in inputs.items() if isinstance(v, torch.Tensor) and v.is_cuda}
            static_out = {k: outputs[k].detach().clone() for k in self._diff_keys()}
            self.capture(static_in, static_out)
        slot = self.slots[0]
        dst = [v for k, v in slot.inputs.items()]
        src = [inputs[k] for k in slot.inputs]
        pairs = [(d, s_) for d, s_ in zip(dst, src) if d.data_ptr() != s_.data_ptr()]
        if pairs:
            torch._foreach_copy_([d for d, _ in pairs], [s_.detach() for _, s_ in pairs])
        losses = slot.replay([outputs[k] for k in self._diff_keys()])
        for k in (("depth", 0, 0),):
            if k in slot.outputs:
                outputs[k] = slot.outputs[k]
        for s in self.scales:
            outputs[("argmin", s)] = slot.outputs[("argmin", s)]
        return losses


class _GraphedFn(torch.autograd.Function):
    """Replay of a captured forward.  inputs: the slot, then the caller's disparities and poses (copied into the
    slot's static tensors unless they ARE those tensors)."""

    @staticmethod
    def forward(ctx, slot, *tensors):
        pairs = [(d, t) for d, t in zip(slot.static_diff, tensors) if d.data_ptr() != t.data_ptr()]
        if pairs:
            torch._foreach_copy_([d for d, _ in pairs], [t.detach() for _, t in pairs])
        slot.graph.replay()
        ctx.slot = slot
        ctx.leaf = [t.is_leaf for t in tensors]
        ctx.set_materialize_grads(False)
        # the static result buffers are overwritten by the next replay: hand out copies (n_pass + 1 floats)
        return slot.vec.clone(), slot.total.clone()

    @staticmethod
    def backward(ctx, g_vec, g_total):
        slot = ctx.slot
        node = slot.node
        if g_vec is None and g_total is None:
            return (None,) * (1 + len(ctx.leaf))
        runner = slot.runner
        n_pass = len(runner.scales)
        pose = [f != "s" for f in runner.sources]              # the stereo pose is an input, not a prediction
        need_in = list(ctx.needs_input_grad[1:])
        it = iter(need_in[n_pass:])
        need = need_in[:n_pass] + [bool(next(it)) if p else False for p in pose]
        grads = _F._scale_gradients(node.plan, node.dims, node.gdisps, node.small, [], g_vec, g_total, need)
        grads = grads[:n_pass] + [g for g, p in zip(grads[n_pass:], pose) if p]
        # static buffers: a leaf would adopt the buffer itself as its .grad and see it overwritten by the next replay
        grads = [g.clone() if (g is not None and leaf) else g for g, leaf in zip(grads, ctx.leaf)]
        return (None,) + tuple(grads)


DEPTH_METRIC_NAMES = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]   # trainer.py:121-122


def compute_depth_losses(self, inputs, outputs, losses):
    """trainer.py:624-652: monitoring metrics of ``outputs[("depth", 0, 0)]`` against
    ``inputs["depth_gt"]`` (resize to the ground truth's resolution, clamp, Garg/Eigen crop, median
    scaling, seven error measures), written into ``losses`` as numpy scalars like the reference."""
    import numpy as np
    names = getattr(self, "depth_metric_names", DEPTH_METRIC_NAMES)
    vals = _F.depth_metrics(outputs[("depth", 0, 0)], inputs["depth_gt"]).cpu().numpy()
    for i, metric in enumerate(names):
        losses[metric] = np.array(vals[i])


_VARIANTS = {"trainer": "trainer", "trainer_dpt": "trainer", "trainer_fusion": "fusion",
             "trainer_fusion_v3": "fusion_v3", "trainer_gru": "gru"}


def install(trainer_module, variant=None):
    """Monkey-patch ``trainer_module.Trainer`` (a reference trainer module) with the fused methods
    and its layer symbols with the libpml-backed ones.  Returns the patched class.  ``variant`` is
    inferred from the module name (trainer / trainer_fusion / trainer_fusion_v3 / trainer_gru)."""
    cls = getattr(trainer_module, "Trainer", trainer_module)
    name = getattr(trainer_module, "__name__", "trainer").split(".")[-1]
    variant = variant or _VARIANTS.get(name, "trainer")

    def _defaults(self):
        if not hasattr(self.opt, "pml_variant"):
            self.opt.pml_variant = variant
        if not hasattr(self.opt, "pml_emit_selection") and orig_log is not None:
            self.opt.pml_emit_selection = "lazy"     # the patched log() below fills it when it is read

    def _gen(self, inputs, outputs):
        _defaults(self)
        return generate_images_pred(self, inputs, outputs)

    def _loss(self, inputs, outputs):
        _defaults(self)
        return compute_losses(self, inputs, outputs)

    orig_log = getattr(cls, "log", None)

    def _log(self, mode, inputs, outputs, losses):
        _defaults(self)
        materialize_logged_outputs(self, inputs, outputs)
        return orig_log(self, mode, inputs, outputs, losses)

    cls.generate_images_pred = _gen
    cls.compute_reprojection_loss = compute_reprojection_loss
    cls.compute_losses = _loss
    if orig_log is not None and not getattr(orig_log, "_pml_wrapped", False):
        _log._pml_wrapped = True
        cls.log = _log
    if hasattr(cls, "compute_depth_losses"):
        cls.compute_depth_losses = compute_depth_losses
    for sym in ("BackprojectDepth", "Project3D", "SSIM", "disp_to_depth", "get_smooth_loss",
                "transformation_from_parameters"):
        if hasattr(trainer_module, sym):
            setattr(trainer_module, sym, getattr(_L, sym))
    return cls
