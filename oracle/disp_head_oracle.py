"""CPU restatement of the DepthDecoder disparity head.  TEST INFRASTRUCTURE ONLY (see photometric_oracle.py).

    networks/depth_decoder.py:46-47   self.convs[("dispconv", s)] = Conv3x3(num_ch_dec[s], 1)
    networks/depth_decoder.py:62-66   outputs[("disp", i)] = self.sigmoid(self.convs[("dispconv", i)](x))
    layers.py:121-136                 Conv3x3 = ReflectionPad2d(1) + Conv2d(in, out, 3)

Pinned by tests/golden/aux/disp_head.npz, produced by the reference's own ``DepthDecoder`` modules
(tests/golden/make_golden_disp_head.py)."""
import torch
import torch.nn.functional as F


def disp_head(x, weight, bias):
    """x [B,C,h,w], weight [1,C,3,3], bias [1] -> sigmoid(conv3x3(reflection_pad1(x))) [B,1,h,w]."""
    return torch.sigmoid(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), weight, bias))


def run(x, weight, bias, g_disp, dtype=torch.float64):
    """-> dict(disp, g_x, g_weight, g_bias) for the upstream gradient g_disp."""
    x = x.detach().to(dtype).clone().requires_grad_(True)
    weight = weight.detach().to(dtype).clone().requires_grad_(True)
    bias = bias.detach().to(dtype).clone().requires_grad_(True)
    d = disp_head(x, weight, bias)
    d.backward(g_disp.to(dtype))
    return {"disp": d.detach(), "g_x": x.grad, "g_weight": weight.grad, "g_bias": bias.grad}
