"""CPU oracle for the photometric-loss hot path.  TEST INFRASTRUCTURE ONLY.

This is a CPU restatement (torch CPU ops, dtype-generic: float32 or float64) of the reference
algorithm on the path named by BASELINE.json's north_star.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; the product package never does (it has no CPU fallback).

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference's own
tests).  This oracle is pinned instead against outputs of the reference's own code
executed in the build container: ``tests/golden/make_golden.py`` imports
``/root/reference/trainer*.py`` unmodified, runs ``Trainer.generate_images_pred`` /
``Trainer.compute_losses`` unbound in float32 and float64 and commits the results as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function here against them.

Each function cites the reference lines it restates (paths relative to /root/reference).
Arithmetic lives in ATen (third party, unpinned by the reference; torch 2.11.0 here):
``F.grid_sample`` defaults to ``align_corners=False`` on torch >= 1.3, which is what the
reference's call without that argument (trainer.py:508-511) resolves to.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SSIM_C1 = 0.01 ** 2
SSIM_C2 = 0.03 ** 2


# --------------------------------------------------------------------------------------------
# unit functions (layers.py)
# --------------------------------------------------------------------------------------------
def disp_to_depth(disp, min_depth, max_depth):
    """layers.py:16-25 -- sigmoid disparity -> (scaled disparity, depth)."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1 / scaled


def upsample_disp(disp, height, width):
    """trainer.py:474-475 -- bilinear, align_corners=False."""
    return F.interpolate(disp, [height, width], mode="bilinear", align_corners=False)


def pixel_grid(height, width, dtype, device="cpu"):
    """layers.py:149-161 -- homogeneous pixel coordinates [3, H*W], x fastest (meshgrid 'xy')."""
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype, device=device),
                            torch.arange(width, dtype=dtype, device=device), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=dtype, device=device)], 0)


def backproject(depth, inv_K):
    """layers.py:163-168 -- depth [B,1,H,W], inv_K [B,4,4] -> camera points [B,4,H*W]."""
    B, _, H, W = depth.shape
    pix = pixel_grid(H, W, depth.dtype, depth.device).unsqueeze(0).expand(B, -1, -1)
    cam = torch.matmul(inv_K[:, :3, :3], pix)
    cam = depth.reshape(B, 1, -1) * cam
    ones = torch.ones(B, 1, H * W, dtype=depth.dtype, device=depth.device)
    return torch.cat([cam, ones], 1)


def project(points, K, T, height, width, eps=1e-7):
    """layers.py:182-193 -- camera points -> normalised sampling grid [B,H,W,2] in [-1,1]."""
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + eps)
    pix = pix.view(B, 2, height, width).permute(0, 2, 3, 1)
    norm = torch.tensor([width - 1, height - 1], dtype=points.dtype, device=points.device)
    return (pix / norm - 0.5) * 2


def warp(img, grid):
    """trainer.py:508-511 -- bilinear, border padding, align_corners False (torch >= 1.3 default)."""
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="border", align_corners=False)


def ssim(x, y):
    """layers.py:234-248 -- reflection-pad(1) 3x3 mean SSIM dissimilarity, clamp((1-n/d)/2, 0, 1)."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x * x, 3, 1) - mu_x * mu_x
    sigma_y = F.avg_pool2d(y * y, 3, 1) - mu_y * mu_y
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + SSIM_C1) * (2 * sigma_xy + SSIM_C2)
    d = (mu_x * mu_x + mu_y * mu_y + SSIM_C1) * (sigma_x + sigma_y + SSIM_C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def reprojection_loss(pred, target, no_ssim=False):
    """trainer.py:517-529 -- 0.85 * mean_c SSIM + 0.15 * mean_c L1 (or L1 alone)."""
    l1 = (target - pred).abs().mean(1, True)
    if no_ssim:
        return l1
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def smooth_loss(disp, img):
    """layers.py:202-215 -- edge-aware first-order smoothness (two separately normalised means)."""
    gdx = (disp[:, :, :, :-1] - disp[:, :, :, 1:]).abs()
    gdy = (disp[:, :, :-1, :] - disp[:, :, 1:, :]).abs()
    gix = (img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True)
    giy = (img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True)
    return (gdx * torch.exp(-gix)).mean() + (gdy * torch.exp(-giy)).mean()


def normalised_smooth_loss(disp, color):
    """trainer.py:612-614 -- mean-normalise the disparity per image, then smooth_loss."""
    mean_disp = disp.mean(2, True).mean(3, True)
    return smooth_loss(disp / (mean_disp + 1e-7), color)


def rot_from_axisangle(vec):
    """layers.py:64-103 -- Rodrigues formula into a 4x4, vec [B,1,3]."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    rows = [
        [x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa],
        [x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa],
        [z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca],
    ]
    R = torch.zeros(vec.shape[0], 4, 4, dtype=vec.dtype, device=vec.device)
    top = torch.stack([torch.stack([e.reshape(-1) for e in row], -1) for row in rows], 1)
    R = torch.cat([torch.cat([top, torch.zeros_like(top[:, :, :1])], 2),
                   torch.zeros_like(R[:, :1, :])], 1)
    R[:, 3, 3] = 1
    return R


def transformation_from_parameters(axisangle, translation, invert=False):
    """layers.py:28-45 (+ get_translation_matrix :48-61)."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    Tm = torch.eye(4, dtype=t.dtype, device=t.device).repeat(t.shape[0], 1, 1)
    Tm = torch.cat([Tm[:, :, :3], torch.cat([t.reshape(-1, 3, 1), torch.ones_like(t.reshape(-1, 3, 1)[:, :1])], 1)], 2)
    return torch.matmul(R, Tm) if invert else torch.matmul(Tm, R)


# --------------------------------------------------------------------------------------------
# the path: generate_images_pred + compute_losses
# --------------------------------------------------------------------------------------------
def _gather(inputs, key, len_sequence):
    """Sequence trainer concatenates per-timestep tensors on the fly (trainer_gru.py:890-899,943-957)."""
    if len_sequence and (key + (0,)) in inputs:
        return torch.cat([inputs[key + (i,)] for i in range(len_sequence)], 0)
    return inputs[key]


def generate_images_pred(opt, inputs, outputs, sources=(-1, 1), variant="trainer"):
    """trainer.py:465-515 (trainer_fusion.py:421-472 with variant="fusion": disparities arrive
    full-res and are not interpolated; trainer_fusion_v3.py:447-482; trainer_gru.py:864-908 with
    variant="gru": 4-tuple keys).  Fills outputs[("depth",0,s)], ("sample",f,s), ("color",f,s)."""
    n_seq = opt.len_sequence if variant == "gru" else 0
    for scale in opt.scales:
        disp = outputs[("disp", scale)]
        if variant == "fusion":
            source_scale = 0
        elif opt.v1_multiscale:
            source_scale = scale
        else:
            disp = upsample_disp(disp, opt.height, opt.width)
            source_scale = 0
        _, depth = disp_to_depth(disp, opt.min_depth, opt.max_depth)
        outputs[("depth", 0, scale)] = depth
        H, W = depth.shape[2:]
        for frame_id in sources:
            if frame_id == "s":
                T = inputs["stereo_T"]
            else:
                T = outputs[("cam_T_cam", 0, frame_id)]
            if variant in ("trainer", "fusion") and opt.pose_model_type == "posecnn" and frame_id != "s":
                # trainer.py:490-499, trainer_fusion.py:446-456 (absent from trainer_fusion_v3 / trainer_gru)
                axisangle = outputs[("axisangle", 0, frame_id)]
                translation = outputs[("translation", 0, frame_id)]
                mean_inv_depth = (1 / depth).mean(3, True).mean(2, True)
                T = transformation_from_parameters(
                    axisangle[:, 0], translation[:, 0] * mean_inv_depth[:, 0], frame_id < 0)
            inv_K = _gather(inputs, ("inv_K", source_scale), n_seq)
            K = _gather(inputs, ("K", source_scale), n_seq)
            src = _gather(inputs, ("color", frame_id, source_scale), n_seq)
            grid = project(backproject(depth, inv_K), K, T, H, W)
            outputs[("sample", frame_id, scale)] = grid
            outputs[("color", frame_id, scale)] = warp(src, grid)


def compute_losses(opt, inputs, outputs, sources=(-1, 1), variant="trainer",
                   noise: Optional[Sequence[torch.Tensor]] = None, keep_maps=False, forced_argmin=None,
                   pixel_weight=None):
    """trainer.py:531-622 (trainer_fusion.py:488-579; trainer_fusion_v3.py:498-590;
    trainer_gru.py:926-1023).  ``noise`` = pre-drawn tie-break tensors, one per scale, in the
    order the reference draws them (trainer.py:592-595); None draws from the global generator.
    ``forced_argmin`` ({scale: int64 [B,H,W]}) replaces the min by a gather with the given
    selection: gradients of a candidate path are then comparable even where fp32 rounding of a
    near-tie made the device pick the other candidate (test use only).
    ``pixel_weight`` ({scale: [B,H,W]}) multiplies the per-pixel photometric term before the mean
    (test use only: measures how much of a pose gradient a set of pixels carries).

    Returns ``losses`` like the reference plus, under ``outputs``: ``identity_selection/{s}``,
    ``("argmin", s)`` (int64 [B,H,W], what the reference's ``torch.min`` returns) and, with
    keep_maps, ``("margin", s)`` (second-best minus best candidate) and ``("to_optimise", s)``."""
    n_seq = opt.len_sequence if variant == "gru" else 0
    losses = {}
    total = 0
    for si, scale in enumerate(opt.scales):
        source_scale = scale if (opt.v1_multiscale and variant != "fusion") else 0
        disp = outputs[("disp", scale)]
        color = _gather(inputs, ("color", 0, source_scale if variant == "fusion" else scale), n_seq)
        target = _gather(inputs, ("color", 0, source_scale), n_seq)

        reproj = torch.cat([reprojection_loss(outputs[("color", f, scale)], target, opt.no_ssim)
                            for f in sources], 1)
        extra = 0
        if opt.disable_automasking and getattr(opt, "predictive_mask", False):
            # trainer.py:571-583: predicted per-frame mask, resized unless v1_multiscale, weights
            # the reprojection losses; 0.2 * BCE(mask, ones) pushes it towards 1
            mask = outputs["predictive_mask"][("disp", scale)]
            if not opt.v1_multiscale:
                mask = F.interpolate(mask, [opt.height, opt.width], mode="bilinear", align_corners=False)
            reproj = reproj * mask
            extra = 0.2 * F.binary_cross_entropy(mask, torch.ones_like(mask))
        if opt.avg_reprojection:
            reproj = reproj.mean(1, keepdim=True)

        if not opt.disable_automasking:
            ident = torch.cat([reprojection_loss(_gather(inputs, ("color", f, source_scale), n_seq),
                                                 target, opt.no_ssim) for f in sources], 1)
            if opt.avg_reprojection:
                ident = ident.mean(1, keepdim=True)
            nz = noise[si] if noise is not None else torch.randn(ident.shape)
            ident = ident + nz.to(ident.dtype).to(ident.device) * 0.00001
            combined = torch.cat((ident, reproj), dim=1)
        else:
            ident = None
            combined = reproj

        if combined.shape[1] == 1:
            to_opt = combined
            idxs = torch.zeros_like(combined[:, 0], dtype=torch.int64)
        else:
            to_opt, idxs = torch.min(combined, dim=1)
            if forced_argmin is not None:
                idxs = forced_argmin[scale].to(torch.int64)
                to_opt = combined.gather(1, idxs.unsqueeze(1))[:, 0]
        outputs[("argmin", scale)] = idxs
        if ident is not None:
            outputs["identity_selection/{}".format(scale)] = (idxs > ident.shape[1] - 1).to(disp.dtype)
        if keep_maps:
            outputs[("to_optimise", scale)] = to_opt.detach().reshape(idxs.shape)
            if combined.shape[1] > 1:
                top2 = torch.topk(combined.detach(), 2, dim=1, largest=False).values
                outputs[("margin", scale)] = top2[:, 1] - top2[:, 0]
            else:
                outputs[("margin", scale)] = torch.full_like(to_opt.detach().reshape(idxs.shape), float("inf"))

        if pixel_weight is not None:
            to_opt = to_opt.reshape(idxs.shape) * pixel_weight[scale].to(to_opt.dtype)
        loss = extra + to_opt.mean()
        loss = loss + opt.disparity_smoothness * normalised_smooth_loss(disp, color) / (2 ** scale)
        total = total + loss
        losses["loss/{}".format(scale)] = loss
    losses["loss"] = total / len(opt.scales)
    return losses


def nest_predictive_mask(outputs):
    """Test dictionaries keep the predicted masks flat as ``("predictive_mask", s)`` (so they can be
    stored like every other tensor); the trainers read ``outputs["predictive_mask"][("disp", s)]``
    (trainer.py:573, written at trainer.py:294)."""
    flat = [k for k in outputs if isinstance(k, tuple) and k[0] == "predictive_mask"]
    if flat:
        outputs["predictive_mask"] = {("disp", k[1]): outputs.pop(k) for k in flat}
    return outputs


def run(opt, inputs, outputs, sources=(-1, 1), variant="trainer", noise=None, dtype=None,
        want_grad=True, keep_maps=True, forced_argmin=None, pixel_weight=None):
    """Whole path on copies of the dictionaries: fwd (+ bwd of losses["loss"]).

    Returns a dict: loss (float), loss/{s}, argmin/{s}, identity_selection/{s}, margin/{s},
    grad_disp/{s}, grad_T/{f} (for pose frames), depth/{s}."""
    def conv(t):
        t = t.detach().cpu()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        return t.clone()

    inp = {k: conv(v) for k, v in inputs.items()}
    out = nest_predictive_mask({k: conv(v) for k, v in outputs.items()})
    leaves = {}
    if want_grad:
        for s in opt.scales:
            out[("disp", s)].requires_grad_(True)
            leaves["grad_disp/{}".format(s)] = out[("disp", s)]
            if "predictive_mask" in out:
                out["predictive_mask"][("disp", s)].requires_grad_(True)
                leaves["grad_mask/{}".format(s)] = out["predictive_mask"][("disp", s)]
        for f in sources:
            if f != "s":
                out[("cam_T_cam", 0, f)].requires_grad_(True)
                leaves["grad_T/{}".format(f)] = out[("cam_T_cam", 0, f)]
                if variant in ("trainer", "fusion") and opt.pose_model_type == "posecnn":   # trainer.py:490-499
                    for k in ("axisangle", "translation"):
                        out[(k, 0, f)].requires_grad_(True)
                        leaves["grad_{}/{}".format(k, f)] = out[(k, 0, f)]
    nz = None if noise is None else [conv(n) for n in noise]
    generate_images_pred(opt, inp, out, sources, variant)
    losses = compute_losses(opt, inp, out, sources, variant, nz, keep_maps=keep_maps,
                            forced_argmin=forced_argmin, pixel_weight=pixel_weight)
    res = {"loss": losses["loss"].detach()}
    for s in opt.scales:
        res["loss/{}".format(s)] = losses["loss/{}".format(s)].detach()
        res["argmin/{}".format(s)] = out[("argmin", s)]
        res["depth/{}".format(s)] = out[("depth", 0, s)].detach()
        key = "identity_selection/{}".format(s)
        if key in out:
            res[key] = out[key].detach()
        if keep_maps:
            res["margin/{}".format(s)] = out[("margin", s)]
            res["to_optimise/{}".format(s)] = out[("to_optimise", s)]
        for f in sources:
            res["color/{}/{}".format(f, s)] = out[("color", f, s)].detach()
            res["sample/{}/{}".format(f, s)] = out[("sample", f, s)].detach()
    if want_grad:
        losses["loss"].backward()
        for name, leaf in leaves.items():
            res[name] = leaf.grad.detach() if leaf.grad is not None else torch.zeros_like(leaf)
    return res


def compute_depth_errors(gt, pred):
    """layers.py:251-269 -> (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3)."""
    thresh = torch.max(gt / pred, pred / gt)
    a1 = (thresh < 1.25).float().mean().to(gt.dtype)          # the reference averages these in fp32
    a2 = (thresh < 1.25 ** 2).float().mean().to(gt.dtype)
    a3 = (thresh < 1.25 ** 3).float().mean().to(gt.dtype)
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def compute_depth_losses(depth_pred, depth_gt, dtype=torch.float64):
    """trainer.py:624-652 restated: resize to the ground truth's resolution, clamp to [1e-3, 80],
    mask = gt > 0 inside the Garg/Eigen crop (rows 153:371, columns 44:1197 of a 375x1242 map),
    median scaling, clamp, compute_depth_errors.  -> tensor [7] in depth_metric_names order."""
    depth_pred, depth_gt = depth_pred.to(dtype), depth_gt.to(dtype)
    pred = torch.clamp(F.interpolate(depth_pred, list(depth_gt.shape[2:]), mode="bilinear", align_corners=False), 1e-3, 80)
    mask = depth_gt > 0
    crop = torch.zeros_like(mask)
    crop[:, :, 153:371, 44:1197] = 1
    mask = mask * crop
    gt = depth_gt[mask]
    pred = pred[mask]
    pred = pred * (torch.median(gt) / torch.median(pred))
    pred = torch.clamp(pred, min=1e-3, max=80)
    return torch.stack(compute_depth_errors(gt, pred))
