"""Kernel-level timing of the fused sweep (CUDA events around the sweep launch, eager), fwd+bwd and
forward only, for A/B experiments:  [B=12 H=192 W=640 S=2] python tools/exp_time.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ssde_b200 import functional as Fn

class A: pass
a = A(); a.batch, a.height, a.width, a.sources = int(os.environ.get("B", 12)), int(os.environ.get("H", 192)), int(os.environ.get("W", 640)), int(os.environ.get("S", 2))
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
opt, srcs, sets = bench.make_sets(a, 4, 0)
pe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
pe[0].record(); pe[1].record(); torch.cuda.synchronize()
steps = [bench.fused_step_fn(opt, srcs, i, o, dev, prof_events=pe) for (i, o) in sets]

def fwd_fn(inputs, outputs):
    target = inputs[("color", 0, 0)].to(dev)
    sources = [inputs[("color", f, 0)].to(dev) for f in srcs]
    K, inv_K = inputs[("K", 0)].to(dev), inputs[("inv_K", 0)].to(dev)
    Ts = [outputs[("cam_T_cam", 0, f)].to(dev) for f in srcs]
    disps = [outputs[("disp", s)].to(dev) for s in opt.scales]
    colors = [inputs[("color", 0, s)].to(dev) for s in opt.scales]
    weights = [opt.disparity_smoothness / 2 ** s for s in opt.scales]
    return lambda: Fn.photometric_loss(target, sources, K, inv_K, Ts, disps, colors, smooth_weights=weights,
                                       min_depth=opt.min_depth, max_depth=opt.max_depth, seed=1, prof_events=pe)
fsteps = [fwd_fn(i, o) for (i, o) in sets]
for name, fns in (("fwd+bwd", steps), ("fwd only", fsteps)):
    durs, tot = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(5 + reps):
        e0.record()
        fns[i % len(fns)]()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            durs.append(pe[0].elapsed_time(pe[1])); tot.append(e0.elapsed_time(e1))
    durs.sort(); tot.sort()
    print("%-9s sweep kernel median %.4f ms  min %.4f ms   (eager step median %.3f ms)" % (name, durs[len(durs) // 2], durs[0], tot[len(tot) // 2]))
