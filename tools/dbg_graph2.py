import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssde_b200 import synthetic, hostio, functional as Fn
dev = torch.device("cuda")
B, H, W = 2, 96, 320
opt = synthetic.make_options(H, W, batch_size=B)
i, o = synthetic.make_batch(B, H, W, seed=1)
hb = {k: v for k, v in i.items()}
hb.update({k: v for k, v in o.items() if k[0] in ("disp", "cam_T_cam")})
pb = hostio.PinnedBatch(hb)
d, arena = pb.upload(dev)
def run(tag, clone_keys):
    g = lambda k: (d[k].detach().clone() if k[0] in clone_keys else d[k].detach())
    Ts = [g(("cam_T_cam", 0, f)).requires_grad_(True) for f in (-1, 1)]
    disps = [g(("disp", s)).requires_grad_(True) for s in range(4)]
    out = Fn.photometric_loss(g(("color", 0, 0)), [g(("color", f, 0)) for f in (-1, 1)], g(("K", 0)), g(("inv_K", 0)), Ts, disps,
                              [g(("color", 0, s)) for s in range(4)], smooth_weights=[1e-3 / 2 ** s for s in range(4)], seed=5)
    out["total"].backward()
    torch.cuda.synchronize()
    print(tag, "loss %.6f" % out["total"].item(), ["NaN" if torch.isnan(t.grad).any() else "%.3e" % t.grad.abs().max().item() for t in Ts],
          "ptr%256:", [t.data_ptr() % 256 for t in Ts], "K ptr%256", g(("K", 0)).data_ptr() % 256)
run("all views", ())
run("clone T", ("cam_T_cam",))
run("clone K", ("K", "inv_K"))
run("clone color", ("color",))
run("clone disp", ("disp",))
run("clone all", ("cam_T_cam", "K", "inv_K", "color", "disp"))
