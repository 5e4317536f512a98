import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from ssde_b200 import synthetic, trainer_hooks, hostio
dev = torch.device("cuda")
B, H, W = 2, 96, 320
opt = synthetic.make_options(H, W, batch_size=B)
opt.pml_sources, opt.pml_variant, opt.pml_emit_depth = [-1, 1], "trainer", "scale0"
ns = SimpleNamespace(opt=opt, device=dev, num_scales=4)
frames = [0, -1, 1]
batches = []
for seed in (1, 2, 3):
    i, o = synthetic.make_batch(B, H, W, seed=seed)
    hb = {"color_u8": torch.stack([(i[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8) for f in frames], 0).contiguous()}
    hb.update({k: v for k, v in i.items() if not (isinstance(k, tuple) and k[0] == "color")})
    hb.update({k: v for k, v in o.items() if k[0] in ("disp", "cam_T_cam")})
    batches.append(hostio.PinnedBatch(hb))

def eager(hb, tag):
    d2, ar = hb.upload(dev)
    inp2 = {k: v for k, v in d2.items() if not (isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam"))}
    out2 = {k: v.requires_grad_(True) for k, v in d2.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")}
    trainer_hooks.ingest_colors(inp2, frames, 4, device=dev)
    o2 = SimpleNamespace(**vars(opt)); o2.pml_noise = "philox"
    ns2 = SimpleNamespace(opt=o2, device=dev, num_scales=4)
    trainer_hooks.generate_images_pred(ns2, inp2, out2)
    want = trainer_hooks.compute_losses(ns2, inp2, out2)
    want["loss"].backward()
    torch.cuda.synchronize()
    print(tag, "loss %.6f" % want["loss"].item(), {str(k): ("NaN" if torch.isnan(v.grad).any() else "%.3e" % v.grad.abs().max().item()) for k, v in out2.items() if k[0] in ("disp", "cam_T_cam")})

eager(batches[0], "eager before any graph")
eager(batches[1], "eager again")
if len(sys.argv) > 1:
    sys.exit(0)
d, arena = batches[0].upload(dev)
inp = {k: v for k, v in d.items() if not (isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam"))}
out = {k: v.requires_grad_(True) for k, v in d.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")}
runner = trainer_hooks.GraphedLoss(ns)
slot = runner.capture(inp, out)
for n, hb in enumerate(batches):
    hb.upload_into(arena)
    for k, v in out.items():
        if k[0] in ("disp", "cam_T_cam"): v.grad = None
    losses = slot.replay()
    losses["loss"].backward()
    torch.cuda.synchronize()
    print("graph step", n, "loss %.6f" % losses["loss"].item(), {str(k): ("NaN" if torch.isnan(v.grad).any() else "%.3e" % v.grad.abs().max().item()) for k, v in out.items() if k[0] in ("disp", "cam_T_cam")})
    eager(hb, "  eager after graph step %d" % n)
