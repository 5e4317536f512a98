"""DepthDecoder disparity heads: oracle against the reference's own modules (golden),
kernels against the float64 oracle, and the DepthDecoder drop-in."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import disp_head_oracle as dho
from ssde_b200 import functional as Fn, layers as L

GOLDEN = os.path.join(common.GOLDEN_DIR, "aux", "disp_head.npz")


def _cases():
    z = np.load(GOLDEN)
    for s in range(4):
        yield s, {k.split("|")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith("s%d|" % s)}


def test_oracle_matches_the_reference_heads():
    for s, c in _cases():
        r = dho.run(c["x"], c["weight"], c["bias"], c["g_disp"])
        for k in ("disp", "g_x", "g_weight", "g_bias"):
            assert common.rel_err(r[k], c[k]) < 1e-12, (s, k)


def _check_kernels(device):
    for s, c in _cases():
        x = c["x"].to(device).requires_grad_(True)
        w = c["weight"].to(device).requires_grad_(True)
        b = c["bias"].to(device).requires_grad_(True)
        disp = Fn.disp_head(x, w, b)
        disp.backward(c["g_disp"].to(device))
        assert (disp.detach().double().cpu() - c["disp"]).abs().max() < 2e-6, s      # a sigmoid of an fp32 sum of 9 C terms
        for got, k in ((x.grad, "g_x"), (w.grad, "g_weight"), (b.grad, "g_bias")):
            assert common.rel_err(got.cpu(), c[k]) < 2e-5, (s, k, common.rel_err(got.cpu(), c[k]))
    # sizes that are not multiples of the 30-column strips / 16-row chunks, weights only (no g_x)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 16, 37, 95, generator=g)
    w, b = torch.randn(1, 16, 3, 3, generator=g) * 0.2, torch.randn(1, generator=g)
    gd = torch.randn(3, 1, 37, 95, generator=g)
    ref = dho.run(x, w, b, gd)
    wd, bd = w.to(device).requires_grad_(True), b.to(device).requires_grad_(True)
    disp = Fn.disp_head(x.to(device), wd, bd)
    disp.backward(gd.to(device))
    assert (disp.detach().double().cpu() - ref["disp"]).abs().max() < 2e-6
    assert common.rel_err(wd.grad.cpu(), ref["g_weight"]) < 2e-5 and common.rel_err(bd.grad.cpu(), ref["g_bias"]) < 2e-5


    # enough 32 x 32 tiles for the tiled kernels of the large heads (pml_head_tiled: B * tiles >= 296), sizes that
    # are not multiples of the tile, a channel count that is not a multiple of the 8-channel march; with and without g_x
    # (9, 5, 70, 150): register-tile forward (w % 4 != 0); (20, 6, 70, 152): forward staged through shared memory (cp.async)
    for (Bt, Ct, ht, wt_) in ((9, 5, 70, 150), (20, 6, 70, 152)):
        x = torch.randn(Bt, Ct, ht, wt_, generator=g)
        w, b = torch.randn(1, Ct, 3, 3, generator=g) * 0.2, torch.randn(1, generator=g)
        gd = torch.randn(Bt, 1, ht, wt_, generator=g)
        ref = dho.run(x, w, b, gd)
        for with_gx in (True, False):
            xd = x.clone().to(device).requires_grad_(with_gx)
            wd, bd = w.clone().to(device).requires_grad_(True), b.clone().to(device).requires_grad_(True)
            disp = Fn.disp_head(xd, wd, bd)
            disp.backward(gd.to(device))
            assert (disp.detach().double().cpu() - ref["disp"]).abs().max() < 2e-6
            assert common.rel_err(wd.grad.cpu(), ref["g_weight"]) < 2e-5 and common.rel_err(bd.grad.cpu(), ref["g_bias"]) < 2e-5
            if with_gx:
                assert common.rel_err(xd.grad.cpu(), ref["g_x"]) < 2e-5


def test_kernels_on_the_emulator(emu_lib):
    _check_kernels("cpu")


@pytest.mark.gpu
def test_kernels_on_the_gpu(cuda_lib):
    _check_kernels("cuda")


def _decoder_dropin(device):
    """A DepthDecoder-shaped module (same attribute layout as networks/depth_decoder.py:30-49) keeps its
    parameters and its forward code; only the head modules are swapped."""
    import torch.nn as nn

    class Conv3x3(nn.Module):            # layers.py:121-136
        def __init__(self, cin, cout):
            super().__init__()
            self.pad = nn.ReflectionPad2d(1)
            self.conv = nn.Conv2d(cin, cout, 3)

        def forward(self, x):
            return self.conv(self.pad(x))

    class Dec(nn.Module):
        def __init__(self):
            super().__init__()
            self.convs = {("dispconv", s): Conv3x3(16 * 2 ** s, 1) for s in range(4)}
            self.decoder = nn.ModuleList(list(self.convs.values()))
            self.sigmoid = nn.Sigmoid()

        def forward(self, feats):
            return {("disp", s): self.sigmoid(self.convs[("dispconv", s)](feats[s])) for s in range(4)}
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False     # the stock convolution would otherwise run in TF32 (1e-4 off)
    dec = Dec().to(device)
    feats = [torch.randn(2, 16 * 2 ** s, 32 >> s, 64 >> s, device=device) for s in range(4)]
    want = {k: v.detach() for k, v in dec(feats).items()}
    keys = list(dec.state_dict().keys())
    L.install_disp_heads(dec)
    assert list(dec.state_dict().keys()) == keys
    got = dec(feats)
    for k in want:
        assert (got[k] - want[k]).abs().max() < 2e-6, k
    sum(v.mean() for v in got.values()).backward()
    assert all(p.grad is not None for p in dec.parameters())


def test_decoder_dropin_on_the_emulator(emu_lib):
    _decoder_dropin("cpu")


@pytest.mark.gpu
def test_decoder_dropin_on_the_gpu(cuda_lib):
    _decoder_dropin("cuda")
