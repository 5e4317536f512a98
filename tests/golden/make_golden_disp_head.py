"""tests/golden/aux/disp_head.npz: outputs and gradients of the reference's OWN disparity heads
(networks/depth_decoder.py:46-47,62-66 over layers.Conv3x3) in float64 (build container only).

    python tests/golden/make_golden_disp_head.py      # needs /root/reference
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import reference_runner  # noqa: E402


def main():
    assert reference_runner.available()
    reference_runner.load("trainer")          # puts /root/reference on sys.path, stubs the unused imports
    from networks.depth_decoder import DepthDecoder
    torch.manual_seed(0)
    dec = DepthDecoder(num_ch_enc=np.array([64, 64, 128, 256, 512])).double()
    blob = {}
    g = torch.Generator().manual_seed(1)
    # (scale, B, h, w): the head's input has num_ch_dec[s] = 16 / 32 / 64 / 128 channels
    for s, B, h, w in ((0, 2, 16, 40), (1, 1, 17, 33), (2, 1, 8, 36), (3, 1, 6, 5)):
        head = dec.convs[("dispconv", s)]
        C = head.conv.in_channels
        # inputs drawn in fp32 (stored as such) so that the fp32 kernels see exactly the float64 reference's values
        x = torch.randn(B, C, h, w, generator=g).double().requires_grad_(True)
        gd = torch.randn(B, 1, h, w, generator=g).double()
        with torch.no_grad():
            for p_ in head.parameters():
                p_.copy_(p_.float().double())
        for p in head.parameters():
            p.grad = None
        disp = dec.sigmoid(head(x))
        disp.backward(gd)
        for k, v in (("x", x), ("weight", head.conv.weight), ("bias", head.conv.bias), ("g_disp", gd), ("disp", disp),
                     ("g_x", x.grad), ("g_weight", head.conv.weight.grad), ("g_bias", head.conv.bias.grad)):
            a = v.detach().numpy().copy()
            blob["s%d|%s" % (s, k)] = a.astype(np.float32) if k in ("x", "weight", "bias", "g_disp") else a
    path = os.path.join(HERE, "aux", "disp_head.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
