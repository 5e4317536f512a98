"""DepthDecoder disparity heads: fused libpml kernels vs the stock PyTorch recipe
(ReflectionPad2d + Conv2d + Sigmoid, networks/depth_decoder.py:62-66) at the headline sizes.
    python tools/disp_head_bench.py [B=12 H=192 W=640]
Prints ms and achieved GB/s against the algorithmic bytes (forward: x read once + disp written;
backward: x read + g_x written + disp / g_disp read)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from ssde_b200 import functional as Fn

B, H, W = int(os.environ.get("B", 12)), int(os.environ.get("H", 192)), int(os.environ.get("W", 640))
dev = torch.device("cuda", 0)


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


tot = {"fused_fwd": 0, "fused_bwd": 0, "torch_fwd": 0, "torch_bwd": 0}
for s in range(4):
    C, h, w = 16 * 2 ** s, H >> s, W >> s
    x = torch.randn(B, C, h, w, device=dev, requires_grad=True)
    conv = nn.Conv2d(C, 1, 3).to(dev)
    pad, sig = nn.ReflectionPad2d(1), nn.Sigmoid()
    gd = torch.randn(B, 1, h, w, device=dev)
    d_f = Fn.disp_head(x, conv.weight, conv.bias)
    d_t = sig(conv(pad(x)))
    err = (d_f - d_t).abs().max().item()
    t_ff = timeit(lambda: Fn.disp_head(x, conv.weight, conv.bias))
    t_tf = timeit(lambda: sig(conv(pad(x))))
    def bw(d):
        return lambda: torch.autograd.grad(d, [x, conv.weight, conv.bias], gd, retain_graph=True)
    t_fb, t_tb = timeit(bw(d_f)), timeit(bw(d_t))
    fb = (B * C * h * w + B * h * w) * 4
    bb = (2 * B * C * h * w + 2 * B * h * w) * 4
    print("scale %d C=%3d %3dx%3d | fused fwd %.4f ms (%.0f GB/s) bwd %.4f ms (%.0f GB/s) | torch fwd %.4f bwd %.4f ms | max |disp diff| %.1e"
          % (s, C, h, w, t_ff, fb / t_ff / 1e6, t_fb, bb / t_fb / 1e6, t_tf, t_tb, err))
    tot["fused_fwd"] += t_ff; tot["fused_bwd"] += t_fb; tot["torch_fwd"] += t_tf; tot["torch_bwd"] += t_tb
print("all four heads: fused fwd+bwd %.3f ms, torch fwd+bwd %.3f ms" % (tot["fused_fwd"] + tot["fused_bwd"], tot["torch_fwd"] + tot["torch_bwd"]))
