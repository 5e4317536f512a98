"""Full training step around the fused loss, BASELINE.json configs[2] (SURVEY.md section 8 e-ii):
mono + stereo (frame_ids 0 -1 1 s), 320x1024, batch 8 per GPU, stock PyTorch/cuDNN networks under
DistributedDataParallel (NCCL), the photometric loss through the Trainer drop-ins.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_step_bench.py [--loss fused|torch|none]

The networks are random-init stand-ins of the standard Monodepth2 architecture (torchvision ResNet-18 encoder,
skip-connection depth decoder with four sigmoid disparity heads, ResNet-18 pose network on frame pairs); they
stay stock PyTorch, only the loss path differs between the arms:
  fused : ssde_b200.trainer_hooks (libpml.so)
  torch : the same ATen op recipe the reference trainer issues on the GPU (oracle port executed on CUDA tensors;
          a BASELINE being measured, never part of the product path)
  none  : networks only (loss = mean disparity), to read off the loss path's share of the step
Synthetic KITTI-shaped frames, one fixed batch per rank resident on the device.
"""
import argparse, json, os, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.distributed as dist
import torchvision


class ConvBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1, padding_mode="reflect")
    def forward(self, x):
        return F.elu(self.conv(x), inplace=True)


class Encoder(nn.Module):
    def __init__(self, in_images=1):
        super().__init__()
        r = torchvision.models.resnet18(weights=None)
        r.fc = nn.Identity()     # unused classifier head: DDP requires every parameter to receive a gradient
        if in_images > 1:
            r.conv1 = nn.Conv2d(3 * in_images, 64, 7, 2, 3, bias=False)
        self.r = r
        self.ch = [64, 64, 128, 256, 512]
    def forward(self, x):
        r = self.r
        f0 = r.relu(r.bn1(r.conv1((x - 0.45) / 0.225)))
        f1 = r.layer1(r.maxpool(f0)); f2 = r.layer2(f1); f3 = r.layer3(f2); f4 = r.layer4(f3)
        return [f0, f1, f2, f3, f4]


class DepthDecoder(nn.Module):
    def __init__(self, enc_ch, scales=(0, 1, 2, 3)):
        super().__init__()
        dec = [16, 32, 64, 128, 256]
        self.scales = scales
        self.up0, self.up1, self.disp = nn.ModuleDict(), nn.ModuleDict(), nn.ModuleDict()
        for i in range(4, -1, -1):
            cin = enc_ch[-1] if i == 4 else dec[i + 1]
            self.up0[str(i)] = ConvBlock(cin, dec[i])
            self.up1[str(i)] = ConvBlock(dec[i] + (enc_ch[i - 1] if i > 0 else 0), dec[i])
        for s in scales:
            self.disp[str(s)] = nn.Conv2d(dec[s], 1, 3, padding=1, padding_mode="reflect")
    def forward(self, feats):
        out, x = {}, feats[-1]
        for i in range(4, -1, -1):
            x = F.interpolate(self.up0[str(i)](x), scale_factor=2, mode="nearest")
            if i > 0:
                x = torch.cat([x, feats[i - 1]], 1)
            x = self.up1[str(i)](x)
            if i in self.scales:
                out[("disp", i)] = torch.sigmoid(self.disp[str(i)](x))
        return out


class PoseDecoder(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.sq = nn.Conv2d(cin, 256, 1)
        self.c0, self.c1, self.c2 = nn.Conv2d(256, 256, 3, 1, 1), nn.Conv2d(256, 256, 3, 1, 1), nn.Conv2d(256, 6, 1)
    def forward(self, f):
        x = F.relu(self.sq(f))
        x = self.c2(F.relu(self.c1(F.relu(self.c0(x))))).mean(3).mean(2)
        x = 0.01 * x.view(-1, 1, 1, 6)
        return x[..., :3], x[..., 3:]


class Nets(nn.Module):
    def __init__(self):
        super().__init__()
        self.enc, self.pose_enc = Encoder(1), Encoder(2)
        self.dec, self.pose_dec = DepthDecoder(self.enc.ch), PoseDecoder(512)
    def forward(self, inputs):
        outputs = self.dec(self.enc(inputs[("color", 0, 0)]))
        for f in (-1, 1):
            pair = [inputs[("color", f, 0)], inputs[("color", 0, 0)]] if f < 0 else [inputs[("color", 0, 0)], inputs[("color", f, 0)]]
            aa, tr = self.pose_dec(self.pose_enc(torch.cat(pair, 1))[-1])
            outputs[("axisangle", 0, f)], outputs[("translation", 0, f)] = aa, tr
        return outputs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8); ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=1024); ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5); ap.add_argument("--loss", default="fused", choices=["fused", "torch", "none"])
    ap.add_argument("--cpu", action="store_true", help="debug the DDP plumbing on CPU / gloo (loss arms none | torch only)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cpu") if args.cpu else torch.device("cuda", local)
    if not args.cpu:
        torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        if args.cpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=dev)
    from ssde_b200 import synthetic, trainer_hooks, layers as L
    sources = [-1, 1, "s"]
    opt = synthetic.make_options(args.height, args.width, batch_size=args.batch)
    opt.pml_sources, opt.pml_variant, opt.pml_noise, opt.pml_emit_depth = sources, "trainer", "philox", "scale0"
    inputs, _ = synthetic.make_batch(args.batch, args.height, args.width, sources=sources, seed=100 + rank)
    inputs = {k: v.to(dev) for k, v in inputs.items()}
    torch.manual_seed(0)
    nets = Nets().to(dev)
    n_params = sum(p.numel() for p in nets.parameters())
    model = nn.parallel.DistributedDataParallel(nets, device_ids=None if args.cpu else [local]) if world > 1 else nets
    optim = torch.optim.Adam(model.parameters(), 1e-4)
    ns = SimpleNamespace(opt=opt, device=dev, num_scales=4)
    if args.loss == "torch":
        from oracle import photometric_oracle as po     # baseline arm only

    def step():
        outputs = model(inputs)
        if args.loss == "none":
            # same autograd extent as the real loss: every disparity head and both pose outputs get a gradient
            loss = sum(outputs[("disp", s)].mean() for s in opt.scales) + \
                sum(outputs[(k, 0, f)].sum() for k in ("axisangle", "translation") for f in (-1, 1))
        else:
            for f in (-1, 1):
                tfp = L.transformation_from_parameters if args.loss == "fused" else po.transformation_from_parameters
                outputs[("cam_T_cam", 0, f)] = tfp(outputs[("axisangle", 0, f)][:, 0], outputs[("translation", 0, f)][:, 0], f < 0)
            if args.loss == "fused":
                trainer_hooks.generate_images_pred(ns, inputs, outputs)
                loss = trainer_hooks.compute_losses(ns, inputs, outputs)["loss"]
            else:
                po.generate_images_pred(opt, inputs, outputs, sources, "trainer")
                loss = po.compute_losses(opt, inputs, outputs, sources, "trainer")["loss"]
        optim.zero_grad(set_to_none=True)
        loss.backward()
        optim.step()
        return loss

    for _ in range(args.warmup):
        step()
    if args.cpu:
        import time
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss = step()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3])
    else:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    if rank == 0:
        print(json.dumps({"bench": "ddp_step C3", "loss_path": args.loss, "n_gpus": world, "batch_per_gpu": args.batch,
                          "height": args.height, "width": args.width, "sources": [str(s) for s in sources], "params_M": round(n_params / 1e6, 2),
                          "ms_per_step": round(ms, 3), "images_per_s": round(world * args.batch / ms * 1e3, 1),
                          "loss": float(loss.item())}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
