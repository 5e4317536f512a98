"""Synthetic KITTI-shaped batches for the photometric-loss path.

The reference has no data in this container and publishes no fixtures, so every
parity test and benchmark runs on deterministic synthetic tensors laid out with the
reference's own dictionary schema:

* ``inputs[("color", frame_id, scale)]``  [B,3,H/2^s,W/2^s]  (datasets/mono_dataset.py:116-139)
* ``inputs[("K", scale)]`` / ``inputs[("inv_K", scale)]``  [B,4,4]  (mono_dataset.py:174-183,
  KITTI intrinsics kitti_dataset.py:25-28, ``inv_K = pinv(K)`` in fp32)
* ``inputs["stereo_T"]``  [B,4,4]  (mono_dataset.py:203-209)
* ``outputs[("disp", scale)]``  [B,1,H/2^s,W/2^s]  (networks/depth_decoder.py:62-66)
* ``outputs[("cam_T_cam", 0, frame_id)]``  [B,4,4]  (trainer.py:378-442)

The sequence trainer's 4-tuple layout (datasets/kitti_dataset_seq.py:109-140,
``("color", f, s, j)``, ``("K", s, j)``) is produced by :func:`to_sequence_layout`.

Everything is generated on the CPU with a seeded ``torch.Generator`` and moved to the
requested device afterwards, so CPU oracle and CUDA path see identical bits.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

KITTI_K = np.array([[0.58, 0, 0.5, 0],
                    [0, 1.92, 0.5, 0],
                    [0, 0, 1, 0],
                    [0, 0, 0, 1]], dtype=np.float32)


def make_options(height=192, width=640, scales=(0, 1, 2, 3), batch_size=12, frame_ids=(0, -1, 1),
                 **overrides) -> SimpleNamespace:
    """Loss-affecting subset of ``MonodepthOptions`` (options.py:100-216) with its defaults."""
    opt = SimpleNamespace(
        height=height, width=width, scales=list(scales), batch_size=batch_size,
        frame_ids=list(frame_ids), min_depth=0.1, max_depth=100.0,
        disparity_smoothness=1e-3, v1_multiscale=False, avg_reprojection=False,
        disable_automasking=False, predictive_mask=False, no_ssim=False,
        pose_model_type="separate_resnet", use_stereo=False, len_sequence=1)
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


def intrinsics(batch: int, height: int, width: int, num_scales: int = 4):
    """Per-scale K and pinv(K) exactly as the dataset builds them (mono_dataset.py:174-183)."""
    out = {}
    for s in range(num_scales):
        K = KITTI_K.copy()
        K[0, :] *= width // (2 ** s)
        K[1, :] *= height // (2 ** s)
        inv_K = np.linalg.pinv(K)
        out[("K", s)] = torch.from_numpy(K).unsqueeze(0).repeat(batch, 1, 1).contiguous()
        out[("inv_K", s)] = torch.from_numpy(inv_K).unsqueeze(0).repeat(batch, 1, 1).contiguous()
    return out


def _box_blur(x: torch.Tensor, k: int) -> torch.Tensor:
    pad = k // 2
    return F.avg_pool2d(F.pad(x, (pad, pad, pad, pad), mode="replicate"), k, 1)


def _rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """Rodrigues rotation as a 4x4 (same closed form as layers.py:64-103)."""
    angle = vec.norm(dim=-1, keepdim=True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle)[..., 0], torch.sin(angle)[..., 0]
    C = 1 - ca
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    R = torch.zeros(vec.shape[0], 4, 4, dtype=vec.dtype)
    R[:, 0, 0] = x * x * C + ca
    R[:, 0, 1] = x * y * C - z * sa
    R[:, 0, 2] = z * x * C + y * sa
    R[:, 1, 0] = x * y * C + z * sa
    R[:, 1, 1] = y * y * C + ca
    R[:, 1, 2] = y * z * C - x * sa
    R[:, 2, 0] = z * x * C - y * sa
    R[:, 2, 1] = y * z * C + x * sa
    R[:, 2, 2] = z * z * C + ca
    R[:, 3, 3] = 1
    return R


def pose_matrix(axisangle: torch.Tensor, translation: torch.Tensor, invert: bool) -> torch.Tensor:
    """axis-angle + translation -> 4x4, following layers.py:28-45 (invert => R^T, -t, R@T order)."""
    R = _rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = -t
    T = torch.eye(4, dtype=axisangle.dtype).repeat(axisangle.shape[0], 1, 1)
    T[:, :3, 3] = t
    return (R @ T) if invert else (T @ R)


class _Texture:
    """Band-limited analytic RGB texture tex(u, v) in [0,1]: a sum of sinusoids with 8-120 px
    wavelengths, different per image.  Evaluable at fractional coordinates, so consistent views
    of one scene can be rendered without any resampling."""
    def __init__(self, batch, g, n_comp=10):
        wl = 8.0 * (15.0 ** torch.rand(batch, 3, n_comp, generator=g))
        ang = torch.rand(batch, 3, n_comp, generator=g) * (2 * math.pi)
        self.fx = (torch.cos(ang) / wl * 2 * math.pi)
        self.fy = (torch.sin(ang) / wl * 2 * math.pi)
        self.ph = torch.rand(batch, 3, n_comp, generator=g) * (2 * math.pi)
        a = torch.rand(batch, 3, n_comp, generator=g) + 0.2
        self.a = 0.45 * a / a.sum(-1, keepdim=True)

    def __call__(self, u, v):
        # u, v: [B,H,W] -> [B,3,H,W]
        out = torch.full((u.shape[0], 3) + tuple(u.shape[1:]), 0.5)
        for k in range(self.a.shape[-1]):
            arg = (self.fx[:, :, k, None, None] * u[:, None] + self.fy[:, :, k, None, None] * v[:, None]
                   + self.ph[:, :, k, None, None])
            out += self.a[:, :, k, None, None] * torch.sin(arg)
        return out


def _project_pixels(disp, K, inv_K, T, xs, ys, min_depth=0.1, max_depth=100.0):
    """Pixel positions of the target grid seen from the source camera (data generation only)."""
    depth = 1.0 / (1.0 / max_depth + (1.0 / min_depth - 1.0 / max_depth) * disp[:, 0])
    pix = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(1, 3, -1)
    cam = (inv_K[:, :3, :3] @ pix) * depth.reshape(depth.shape[0], 1, -1)
    cam = torch.cat([cam, torch.ones_like(cam[:, :1])], 1)
    p = ((K @ T)[:, :3] @ cam)
    u = (p[:, 0] / (p[:, 2] + 1e-7)).reshape(depth.shape)
    v = (p[:, 1] / (p[:, 2] + 1e-7)).reshape(depth.shape)
    return u, v


def _invert_flow(fu, fv, xs, ys, iters=10):
    """Solve p + flow(p) = x for p on the pixel grid (fixed point, bilinear flow lookup)."""
    B, H, W = fu.shape
    flow = torch.stack([fu, fv], 1)
    pu, pv = xs.expand(B, -1, -1).clone(), ys.expand(B, -1, -1).clone()
    for _ in range(iters):
        grid = torch.stack([pu / (W - 1) * 2 - 1, pv / (H - 1) * 2 - 1], -1)
        f = F.grid_sample(flow, grid, mode="bilinear", padding_mode="border", align_corners=True)
        pu, pv = xs - f[:, 0], ys - f[:, 1]
    return pu, pv


def make_batch(batch=12, height=192, width=640, scales=(0, 1, 2, 3), sources=(-1, 1), seed=0,
               style="kitti", full_res_disp=False, device="cpu", dtype=torch.float32, predictive_mask=False):
    """Build (inputs, outputs) dictionaries in the reference schema.

    style:
      "kitti"    low-passed colours, network-like disparities in [0.03, 0.6], small poses
                 (5-15 px parallax, <10 % border-clipped) -- the benchmark distribution
      "uniform"  iid U(0,1) colours and disparities, larger poses (stress: SSIM variance, clipping)
      "static"   source frames identical to the target (identity loss exactly 0 -> only the
                 tie-break noise orders the candidates, mono_dataset.py:165-170)
      "constant" constant-colour images (SSIM denominators collapse to C1*C2)
      "oof"      poses that throw most samples out of the frustum (border clamp path)
    full_res_disp: every ``("disp", s)`` is emitted at H x W (trainer_fusion.py:427-433).
    predictive_mask: also emit ``outputs[("predictive_mask", s)]`` [B,S,h_s,w_s], a sigmoid of smooth
      noise like the mask decoder's output (trainer.py:294, networks/depth_decoder.py:62-66); the
      trainers read it as ``outputs["predictive_mask"][("disp", s)]`` (trainer.py:573).
    """
    g = torch.Generator().manual_seed(seed)
    num_scales = max(scales) + 1
    frames = [0] + [f for f in sources]
    inputs: Dict = {}
    outputs: Dict = {}

    kitti = style in ("kitti", "static")
    if kitti:
        tex = _Texture(batch, g)
        ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32),
                                torch.arange(width, dtype=torch.float32), indexing="ij")
    for fi, f in enumerate(frames):
        if style == "constant":
            img = torch.rand(batch, 3, 1, 1, generator=g).expand(batch, 3, height, width).contiguous()
        elif kitti:
            img = None  # filled below, once disparity and poses exist
        else:
            img = torch.rand(batch, 3, height, width, generator=g)
        if img is not None:
            for s in range(num_scales):
                inputs[("color", f, s)] = (img if s == 0 else F.avg_pool2d(img, 2 ** s)).contiguous()

    inputs.update(intrinsics(batch, height, width, num_scales))

    if style == "uniform":
        for s in scales:
            h, w = (height, width) if full_res_disp else (height // 2 ** s, width // 2 ** s)
            outputs[("disp", s)] = torch.rand(batch, 1, h, w, generator=g)
    else:
        k = max(3, (min(height, width) // 8) | 1)
        n = _box_blur(_box_blur(torch.randn(batch, 1, height, width, generator=g), k), k)
        n = n / (n.std() + 1e-6)
        true_disp = 0.03 + 0.57 * torch.sigmoid(1.5 * n)          # "ground truth", full resolution
        for s in scales:
            # each decoder head predicts the truth up to a small, smooth error
            d = true_disp if (full_res_disp or s == 0) else F.avg_pool2d(true_disp, 2 ** s)
            err = _box_blur(torch.randn(d.shape, generator=g), 5) * 0.05
            outputs[("disp", s)] = (d * (1 + err)).clamp(0.01, 0.99).contiguous()

    for f in sources:
        if f == "s":
            T = torch.eye(4).repeat(batch, 1, 1)
            T[:, 0, 3] = 0.1 * (1 - 2 * (torch.arange(batch) % 2)).float()
            inputs["stereo_T"] = T.contiguous()
            continue
        if style == "oof":
            aa_std, t_std, tz_std = 0.6, 2.0, 2.0
        elif style == "uniform":
            aa_std, t_std, tz_std = 0.02, 0.05, 0.05
        else:
            aa_std, t_std, tz_std = 0.005, 0.02, 0.03
        aa = torch.randn(batch, 1, 3, generator=g) * aa_std
        tr = torch.randn(batch, 1, 3, generator=g) * t_std
        tr[..., 2] += torch.randn(batch, 1, generator=g) * tz_std
        outputs[("axisangle", 0, f)] = aa.unsqueeze(1)
        outputs[("translation", 0, f)] = tr.unsqueeze(1)
        outputs[("cam_T_cam", 0, f)] = pose_matrix(aa[:, 0], tr[:, 0], invert=(f < 0)).contiguous()

    if kitti:
        # Render target and sources from one analytic texture so that warping a source with the
        # (near-true) disparity and pose reproduces the target: source(x') = tex(p) where
        # p + flow(p) = x' and flow = projected position minus pixel position under the true
        # disparity (fixed-point inverse of the warp).
        inputs[("color", 0, 0)] = tex(xs.expand(batch, -1, -1), ys.expand(batch, -1, -1))
        for f in sources:
            if style == "static":
                img = inputs[("color", 0, 0)].clone()
            else:
                T = inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]
                u, v = _project_pixels(true_disp, inputs[("K", 0)], inputs[("inv_K", 0)], T, xs, ys)
                # grid_sample(align_corners=False) reads source index u*W/(W-1)-0.5
                u = u * (width / (width - 1.0)) - 0.5
                v = v * (height / (height - 1.0)) - 0.5
                img = tex(*_invert_flow(u - xs, v - ys, xs, ys))
            inputs[("color", f, 0)] = img
        for f in frames:
            img = inputs[("color", f, 0)].clamp(0, 1).contiguous()
            for s in range(num_scales):
                inputs[("color", f, s)] = (img if s == 0 else F.avg_pool2d(img, 2 ** s)).contiguous()

    if predictive_mask:
        for s in scales:
            h, w = height // 2 ** s, width // 2 ** s
            n = _box_blur(torch.randn(batch, len(sources), h, w, generator=g), 5)
            outputs[("predictive_mask", s)] = torch.sigmoid(4.0 * n + 1.0).contiguous()

    def cvt(t):
        return t.to(device=device, dtype=dtype) if t.is_floating_point() else t.to(device)

    inputs = {k: cvt(v) for k, v in inputs.items()}
    outputs = {k: cvt(v) for k, v in outputs.items()}
    return inputs, outputs


def draw_noise(batch, height, width, scales, n_identity, seed=0, v1_multiscale=False,
               dtype=torch.float32) -> List[torch.Tensor]:
    """Pre-draw the tie-break noise exactly as the reference consumes the global CPU generator:
    one ``torch.randn([B, n_identity, h, w])`` per scale, in scale order (trainer.py:594-595)."""
    torch.manual_seed(seed)
    out = []
    for s in scales:
        h, w = (height // 2 ** s, width // 2 ** s) if v1_multiscale else (height, width)
        out.append(torch.randn([batch, n_identity, h, w]).to(dtype))
    return out


def to_sequence_layout(inputs: Dict, len_sequence: int) -> Dict:
    """Split a flat batch of B = bs*n images into the sequence trainer's 4-tuple keys
    (kitti_dataset_seq.py:109-140): chunk i of the flat batch becomes time index i, so that
    ``torch.cat([inputs[(k, f, s, i)] for i in range(n)], 0)`` (trainer_gru.py:890-899) restores it."""
    n = len_sequence
    out = {}
    for key, val in inputs.items():
        if isinstance(key, tuple):
            chunks = torch.chunk(val, n, dim=0)
            assert len(chunks) == n
            for i, c in enumerate(chunks):
                out[key + (i,)] = c.contiguous()
        else:
            out[key] = val
    return out


def algorithmic_bytes(batch, height, width, n_sources, n_scales=4) -> int:
    """SURVEY.md §8(d) / BASELINE.md §4 algorithmic bytes per fwd+bwd step:
    N * [16 + 16 S + n (26 + 28 S) + 12 Q + 24 Q'] with Q = sum_s 4^-s, Q' = sum_{s>=1} 4^-s."""
    N = batch * height * width
    Q = sum(4.0 ** -s for s in range(n_scales))
    Qp = Q - 1.0
    per_px = 16 + 16 * n_sources + n_scales * (26 + 28 * n_sources) + 12 * Q + 24 * Qp
    return int(round(N * per_px))
