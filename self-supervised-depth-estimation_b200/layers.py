"""Drop-in replacements for the hot-path symbols of the reference's ``layers.py``.

Same names, constructor arguments, forward signatures and return shapes as
/root/reference/layers.py (:16-25 disp_to_depth, :28-45 transformation_from_parameters,
:139-168 BackprojectDepth, :171-193 Project3D, :202-215 get_smooth_loss, :218-248 SSIM), each
backed by one sm_100a kernel forward and one backward through libpml.so.  The modules are
stateless: unlike the reference they register no parameters/buffers (the reference's
``id_coords`` / ``ones`` / ``pix_coords`` grids, layers.py:149-161, are recomputed in registers),
so nothing new appears in any state_dict and ``.to(device)`` is a no-op that still works.

The reference's unfused trainer code (``generate_images_pred`` written against these layers plus
``F.grid_sample``) therefore runs unchanged on top of this module; the fused fast path is
``trainer_hooks``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as _F


def disp_to_depth(disp, min_depth, max_depth):
    """layers.py:16-25 -> (scaled_disp, depth)."""
    return _F._DispToDepth.apply(disp, min_depth, max_depth)


def transformation_from_parameters(axisangle, translation, invert=False):
    """layers.py:28-45: axisangle, translation [B,1,3] (or [B,3]) -> [B,4,4]."""
    return _F._Pose.apply(axisangle, translation, bool(invert))


class BackprojectDepth(nn.Module):
    """layers.py:139-168.  ``batch_size`` is accepted for signature compatibility; the batch is
    read from ``inv_K`` so a short last batch works (the reference needs drop_last=True)."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size = batch_size
        self.height = height
        self.width = width

    def forward(self, depth, inv_K):
        return _F._Backproject.apply(depth, inv_K, self.height, self.width)


class Project3D(nn.Module):
    """layers.py:171-193: points [B,4,H*W], K, T [B,4,4] -> sampling grid [B,H,W,2]."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size = batch_size
        self.height = height
        self.width = width
        self.eps = eps

    def forward(self, points, K, T):
        return _F._Project.apply(points, K, T, self.height, self.width, self.eps)


class SSIM(nn.Module):
    """layers.py:218-248: reflection-padded 3x3 SSIM dissimilarity, [B,C,H,W] -> [B,C,H,W]."""

    def __init__(self):
        super().__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y):
        return _F._SSIM.apply(x, y)


def get_smooth_loss(disp, img):
    """layers.py:202-215: edge-aware smoothness of ``disp`` [B,1,H,W] under ``img`` [B,C,H,W]."""
    return _F._SmoothLoss.apply(disp, img)


def interpolate_bilinear(x, size):
    """The reference's ``F.interpolate(x, [H, W], mode="bilinear", align_corners=False)`` calls
    (trainer.py:474-475 on disparities, :574-576 on the predictive mask): [B,C,h,w] -> [B,C,H,W]."""
    return _F.upsample_bilinear(x, size[0], size[1])


class DispHead(nn.Module):
    """``sigmoid(Conv3x3(x))`` -- one disparity head of the reference's ``DepthDecoder``
    (networks/depth_decoder.py:46-47 ``("dispconv", s)``, :62-66) as one fused kernel.

    Wraps the reference's own ``Conv3x3`` module (layers.py:121-136) and keeps using ITS parameters:
    ``self.conv`` is the original ``nn.Conv2d(C_in, 1, 3)``, so state_dict keys, optimiser state and
    checkpoints are unchanged.  ``forward`` returns the disparity (the sigmoid is inside)."""

    def __init__(self, conv3x3):
        super().__init__()
        conv = getattr(conv3x3, "conv", conv3x3)
        if not isinstance(conv, nn.Conv2d) or conv.out_channels != 1 or tuple(conv.kernel_size) != (3, 3):
            raise TypeError("DispHead wraps a Conv3x3 / nn.Conv2d(C_in, 1, 3)")
        pad = getattr(conv3x3, "pad", None)
        if pad is not None and not isinstance(pad, nn.ReflectionPad2d):
            raise TypeError("DispHead implements the reflection-padded Conv3x3 (use_refl=True, layers.py:127-128)")
        self.pad = pad if pad is not None else nn.ReflectionPad2d(1)
        self.conv = conv

    def forward(self, x):
        return _F.disp_head(x, self.conv.weight, self.conv.bias)


def install_disp_heads(depth_decoder):
    """Replace the ``("dispconv", s)`` + ``sigmoid`` pairs of a reference ``DepthDecoder`` instance
    (networks/depth_decoder.py:46-49,62-66) by :class:`DispHead`.  The decoder's ``forward`` is untouched: it
    still evaluates ``self.sigmoid(self.convs[("dispconv", i)](x))`` -- the head now returns the disparity and
    ``self.sigmoid`` becomes the identity.  Parameter names stay the same.  Returns the decoder."""
    replaced = {}
    for key, mod in list(depth_decoder.convs.items()):
        if isinstance(key, tuple) and key[0] == "dispconv" and not isinstance(mod, DispHead):
            head = DispHead(mod)
            replaced[id(mod)] = head
            depth_decoder.convs[key] = head
    if hasattr(depth_decoder, "decoder"):     # the nn.ModuleList that registers the parameters
        for i, mod in enumerate(depth_decoder.decoder):
            if id(mod) in replaced:
                depth_decoder.decoder[i] = replaced[id(mod)]
    if replaced:
        depth_decoder.sigmoid = nn.Identity()
    return depth_decoder
