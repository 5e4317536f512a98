import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssde_b200 import synthetic, functional as Fn, trainer_hooks
dev = torch.device("cuda")
B, H, W = 2, 96, 320
frames = [0, -1, 1]
i, o = synthetic.make_batch(B, H, W, seed=1)
u8 = torch.stack([(i[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8) for f in frames], 0).contiguous().to(dev)
inp = {k: v.to(dev) for k, v in i.items() if not (isinstance(k, tuple) and k[0] == "color")}
inp["color_u8"] = u8
trainer_hooks.ingest_colors(inp, frames, 4, device=dev)
for trial in range(3):
    Ts = [o[("cam_T_cam", 0, f)].to(dev).clone().requires_grad_(True) for f in (-1, 1)]
    disps = [o[("disp", s)].to(dev).clone().requires_grad_(True) for s in range(4)]
    depth = Fn.depth_from_disp(disps[0].detach(), 0.1, 100.0)
    out = Fn.photometric_loss(inp[("color", 0, 0)], [inp[("color", f, 0)] for f in (-1, 1)], inp[("K", 0)], inp[("inv_K", 0)], Ts, disps,
                              [inp[("color", 0, s)] for s in range(4)], smooth_weights=[1e-3 / 2 ** s for s in range(4)], seed=5, total_div=4)
    node = out["total"].grad_fn
    torch.cuda.synchronize()
    small = node.small
    so = node.plan.small_off
    print(trial, "after fwd: small NaN?", bool(torch.isnan(small).any()), "gT NaN?", bool(torch.isnan(small[so["gT"]:]).any()), "gconst", small[so["gconst"]:so["gconst"] + 8].tolist())
    masks = Fn.selection_masks(out["argmin_all"], 2)
    torch.cuda.synchronize()
    print("   after masks: gT NaN?", bool(torch.isnan(small[so["gT"]:]).any()), "small ptr", hex(small.data_ptr()), "masks ptr", hex(masks.data_ptr()), "depth ptr", hex(depth.data_ptr()),
          "argmin ptr", hex(out["argmin_all"].data_ptr()), "garena ptr", hex(node.garena.data_ptr()))
    out["total"].backward()
    torch.cuda.synchronize()
    print("   T grads:", ["NaN" if torch.isnan(t.grad).any() else "%.3e" % t.grad.abs().max().item() for t in Ts])
