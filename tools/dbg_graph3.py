import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from ssde_b200 import synthetic, hostio, functional as Fn, trainer_hooks
dev = torch.device("cuda")
B, H, W = 2, 96, 320
opt = synthetic.make_options(H, W, batch_size=B)
opt.pml_sources, opt.pml_variant, opt.pml_emit_depth = [-1, 1], "trainer", "scale0"
frames = [0, -1, 1]
i, o = synthetic.make_batch(B, H, W, seed=1)
u8 = torch.stack([(i[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8) for f in frames], 0).contiguous().to(dev)
def rep(tag, Ts, loss):
    torch.cuda.synchronize()
    print(tag, "loss %.6f" % loss.item(), ["NaN" if torch.isnan(t.grad).any() else "%.3e" % t.grad.abs().max().item() for t in Ts])
def hooks(tag, inp, emit_depth="scale0", sel=True):
    out = {k: v.to(dev).clone().requires_grad_(True) for k, v in o.items() if k[0] in ("disp", "cam_T_cam")}
    o2 = SimpleNamespace(**vars(opt)); o2.pml_emit_depth = emit_depth; o2.pml_emit_selection = sel
    ns = SimpleNamespace(opt=o2, device=dev, num_scales=4)
    trainer_hooks.generate_images_pred(ns, inp, out)
    losses = trainer_hooks.compute_losses(ns, inp, out)
    losses["loss"].backward()
    rep(tag, [out[("cam_T_cam", 0, f)] for f in (-1, 1)], losses["loss"])
base = {k: v.to(dev) for k, v in i.items()}
hooks("hooks fp32 colours", dict(base))
inp = {k: v for k, v in base.items() if not (isinstance(k, tuple) and k[0] == "color")}
inp["color_u8"] = u8
trainer_hooks.ingest_colors(inp, frames, 4, device=dev)
hooks("hooks ingest colours", dict(inp))
hooks("hooks ingest colours, no depth", dict(inp), emit_depth="none")
hooks("hooks ingest colours, no selection", dict(inp), sel=False)
Ts = [o[("cam_T_cam", 0, f)].to(dev).clone().requires_grad_(True) for f in (-1, 1)]
disps = [o[("disp", s)].to(dev).clone().requires_grad_(True) for s in range(4)]
out = Fn.photometric_loss(inp[("color", 0, 0)], [inp[("color", f, 0)] for f in (-1, 1)], inp[("K", 0)], inp[("inv_K", 0)], Ts, disps,
                          [inp[("color", 0, s)] for s in range(4)], smooth_weights=[1e-3 / 2 ** s for s in range(4)], seed=5)
out["total"].backward()
rep("functional ingest colours", Ts, out["total"])
