"""N>1 path on CPU: two gloo ranks each run the loss on their shard of the batch (emulated
kernels) -- the path has no data-path collective, so the only exchange is the
DDP-style averaging of losses / gradients and the max-over-ranks timing of bench.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from ssde_b200 import synthetic


def _worker(rank, world, port, emu_path, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssde_b200 import _cabi
    _cabi.set_library_for_testing(_cabi.Library(emu_path, emulator=True))
    torch.set_num_threads(1)
    B, H, W = 2, 32, 64
    opt = synthetic.make_options(H, W, batch_size=B)
    inputs, outputs = synthetic.make_batch(B, H, W, seed=40)
    noise_seed = 6
    # shard = one image per rank; the tie-break noise is drawn for the whole batch and sliced
    full_noise = synthetic.draw_noise(B, H, W, opt.scales, 2, seed=noise_seed)
    sl = slice(rank, rank + 1)
    from ssde_b200 import functional as Fn
    disps = [outputs[("disp", s)][sl].clone().requires_grad_(True) for s in opt.scales]
    Ts = [outputs[("cam_T_cam", 0, f)][sl].clone().requires_grad_(True) for f in (-1, 1)]
    out = Fn.photometric_loss(
        inputs[("color", 0, 0)][sl], [inputs[("color", f, 0)][sl] for f in (-1, 1)],
        inputs[("K", 0)][sl], inputs[("inv_K", 0)][sl], Ts, disps,
        [inputs[("color", 0, s)][sl] for s in opt.scales],
        smooth_weights=[opt.disparity_smoothness / 2 ** s for s in opt.scales],
        noise=[n[sl] for n in full_noise])
    loss = out["loss"].mean()
    loss.backward()
    # what DDP does: average across ranks
    l = loss.detach().clone()
    dist.all_reduce(l)
    l /= world
    # bench.py's timing rule: max over ranks
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros_like(disps[0].grad) for _ in range(world)]
    dist.all_gather(gathered, disps[0].grad)
    # input pyramid: every rank ingests the uint8 frames of its own images
    frames = torch.randint(0, 256, (B, H, W, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    lvl = Fn.color_pyramid(frames[sl], 3)[2]
    pyr = [torch.zeros_like(lvl) for _ in range(world)]
    dist.all_gather(pyr, lvl)
    if rank == 0:
        ret["loss"] = l.item()
        ret["tmax"] = t.item()
        ret["grad0"] = torch.cat(gathered, 0) / world
        ret["pyr2"] = torch.cat(pyr, 0)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(emu_lib):
    import __graft_entry__ as ge
    emu_path = ge.build_emu()
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, emu_path, ret), nprocs=world, join=True)
    B, H, W = 2, 32, 64
    opt = synthetic.make_options(H, W, batch_size=B)
    inputs, outputs = synthetic.make_batch(B, H, W, seed=40)
    full = common.run_product(opt, inputs, outputs, device="cpu", noise_seed=6)
    assert abs(ret["loss"] - full["loss"].item()) / abs(full["loss"].item()) < 1e-6
    assert ret["tmax"] == 2.0
    assert common.rel_err(ret["grad0"], full["grad_disp/0"]) < 1e-5
    # the sharded pyramid equals the single-process one bit for bit (and the numpy oracle)
    from ssde_b200 import functional as Fn
    from oracle import pyramid_oracle as pyo
    frames = torch.randint(0, 256, (B, H, W, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    assert torch.equal(ret["pyr2"], Fn.color_pyramid(frames, 3)[2])
    assert (ret["pyr2"].numpy() == pyo.pyramid(frames.numpy(), 3)[0][2]).all()
