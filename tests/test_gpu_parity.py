"""Parity of the CUDA path (through the C ABI, the autograd Functions and the Trainer drop-ins)
against the golden fixtures of the reference and the float64 CPU oracle.  Run with -m gpu."""
import pytest
import torch

import common
import parity
from ssde_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", common.golden_names())
def test_golden_fixture(cuda_lib, name):
    variant, opt, inputs, outputs, r32, r64, seed = common.load_golden(name)
    got = common.run_product(opt, inputs, outputs, variant, device="cuda", noise_seed=seed)
    rep = parity.check(got, opt, variant, inputs, outputs, seed, r32, r64, degenerate=(name == "trainer_constant"))
    print(name, {k: v for k, v in rep.items()})


CONFIGS = {
    # BASELINE.json configs at sizes the float64 oracle finishes in seconds
    "c2_headline_b2": dict(B=2, H=192, W=640, sources=(-1, 1), variant="trainer", style="kitti"),
    "c3_stereo_320x1024": dict(B=1, H=320, W=1024, sources=(-1, 1, "s"), variant="trainer", style="kitti"),
    "c4_gru_seq5": dict(B=5, H=192, W=640, sources=(-1, 1), variant="gru", style="kitti", len_sequence=5),
    "c5_small_96x320_s1": dict(B=3, H=96, W=320, sources=(1,), variant="trainer", style="kitti"),
    "c5_s4": dict(B=1, H=96, W=320, sources=(-1, 1, -2, 2), variant="trainer", style="kitti"),
    "c5_s8": dict(B=1, H=96, W=320, sources=(-1, 1, -2, 2, -3, 3, -4, 4), variant="trainer", style="kitti"),
    "uniform_stress": dict(B=2, H=96, W=320, sources=(-1, 1), variant="trainer", style="uniform"),
    "out_of_frustum": dict(B=2, H=96, W=320, sources=(-1, 1), variant="trainer", style="oof"),
    "tanh_range_disp": dict(B=2, H=64, W=160, sources=(-1, 1), variant="fusion", style="kitti", neg_disp=True),
    # sizes that are not multiples of the 28-column strips / of the row chunks, odd low-res extents
    "odd_sizes": dict(B=3, H=72, W=200, sources=(-1, 1), variant="trainer", style="kitti", scales=[0, 1, 2, 3]),
    "tiny": dict(B=1, H=16, W=32, sources=(-1, 1), variant="trainer", style="uniform", scales=[0, 1]),
    "single_scale3": dict(B=2, H=64, W=96, sources=(1,), variant="trainer", style="kitti", scales=[3]),
    "four_sources_avg": dict(B=1, H=64, W=160, sources=(-1, 1, -2, 2), variant="trainer", style="kitti",
                             opt=dict(avg_reprojection=True)),
    # --predictive_mask ablation (trainer.py:571-583): two frames, three frames with stereo, v1_multiscale
    "predictive_mask": dict(B=2, H=96, W=320, sources=(-1, 1), variant="trainer", style="kitti", pmask=True, seed=33,
                            opt=dict(disable_automasking=True, predictive_mask=True)),
    "predictive_mask_stereo": dict(B=1, H=96, W=320, sources=(-1, 1, "s"), variant="trainer", style="kitti", pmask=True,
                                   opt=dict(disable_automasking=True, predictive_mask=True)),
    "predictive_mask_v1": dict(B=2, H=64, W=160, sources=(-1, 1), variant="trainer", style="kitti", pmask=True,
                               opt=dict(disable_automasking=True, predictive_mask=True, v1_multiscale=True)),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_against_oracle(cuda_lib, name):
    c = dict(CONFIGS[name])
    B, H, W, sources, variant = c["B"], c["H"], c["W"], c["sources"], c["variant"]
    kw = dict(c.get("opt", {}))
    if "scales" in c:
        kw["scales"] = c["scales"]
    opt = synthetic.make_options(H, W, batch_size=B, len_sequence=c.get("len_sequence", 1), **kw)
    inputs, outputs = synthetic.make_batch(B, H, W, sources=sources, seed=c.get("seed", 31), style=c["style"],
                                           full_res_disp=(variant == "fusion"), scales=opt.scales,
                                           predictive_mask=c.get("pmask", False))
    if c.get("neg_disp"):
        # Fusion's UpscalePS ends in tanh (fusion_v2.py:234-235): disparities are not confined to
        # [0,1]; keep sigma > 0 so that depth stays finite in the oracle too
        for s in opt.scales:
            outputs[("disp", s)] = outputs[("disp", s)] * 1.6 - 0.0005
    got = common.run_product(opt, inputs, outputs, variant, device="cuda", noise_seed=5, sources=sources)
    rep = parity.check(got, opt, variant, inputs, outputs, 5, sources=sources)
    print(name, rep)


PHILOX_CASES = {
    # (sources, B, H, W): mode 0 at the headline resolution, modes 1/3/2 with three, four and eight frames,
    # stereo-only training (one source frame: the scalar single-frame instantiation)
    "s1_stereo_96x320": (("s",), 2, 96, 320),
    "s5_64x160": ((-1, 1, -2, 2, "s"), 2, 64, 160),
    "s2_192x640": ((-1, 1), 2, 192, 640),
    "s2_odd_72x200": ((-1, 1), 3, 72, 200),
    "s3_stereo_96x320": ((-1, 1, "s"), 2, 96, 320),
    "s4_64x160": ((-1, 1, -2, 2), 2, 64, 160),
    "s8_64x160": ((-1, 1, -2, 2, -3, 3, -4, 4), 2, 64, 160),
}


@pytest.mark.parametrize("name", list(PHILOX_CASES))
def test_default_training_instantiation(cuda_lib, name):
    """The instantiations bench.py times and trainer_hooks runs by default -- in-kernel Philox noise, no
    by-product stores: sweep_kernel<GRAD,SSIM,MODE,EMIT=false,COMMON=true> -- against the zero-noise float64
    oracle: loss, selection wherever the float64 margin exceeds the noise bound, gradients (forced selection)."""
    sources, B, H, W = PHILOX_CASES[name]
    opt = synthetic.make_options(H, W, batch_size=B)
    inputs, outputs = synthetic.make_batch(B, H, W, sources=sources, seed=11)
    torch.manual_seed(5)
    got = common.run_product(opt, inputs, outputs, "trainer", device="cuda", noise_seed=None, sources=sources,
                             extra_opt=dict(pml_emit_warped=False, pml_emit_depth="scale0"))
    assert not any(k.startswith("color/") for k in got)
    rep = parity.check(got, opt, "trainer", inputs, outputs, 0, sources=sources, philox=True)
    print(name, rep)


@pytest.mark.parametrize("size", [(32, 64), (64, 160), (96, 320)])
@pytest.mark.parametrize("seed", range(12))
def test_seed_sweep(cuda_lib, seed, size):
    """12 fresh seeds x 3 sizes, image styles alternating: parity must not depend on the seed."""
    H, W = size
    style = "kitti" if seed % 2 == 0 else "uniform"
    opt = synthetic.make_options(H, W, batch_size=2)
    inputs, outputs = synthetic.make_batch(2, H, W, seed=100 + seed, style=style)
    got = common.run_product(opt, inputs, outputs, "trainer", device="cuda", noise_seed=seed + 4)
    parity.check(got, opt, "trainer", inputs, outputs, seed + 4,
                 loss_tol=parity.LOSS_TOL if 2 * H * W >= 20000 else parity.LOSS_TOL_SMALL)


@pytest.mark.parametrize("H,W,seed,style", [(64, 160, 3, "kitti"), (96, 320, 100, "kitti"), (96, 320, 101, "uniform"),
                                            (192, 640, 7, "kitti")])
def test_pose_gradient_with_kink_pixels_weighed_out(cuda_lib, H, W, seed, style):
    """Sharp gradient test: the derivative-kink pixels are removed from the loss on both sides (per-pixel frame
    weights in the library, pixel weights in the float64 oracle) instead of being covered by an allowance."""
    opt = synthetic.make_options(H, W, batch_size=2)
    inputs, outputs = synthetic.make_batch(2, H, W, seed=seed, style=style)
    rep = parity.pose_gradient_check("cuda", opt, inputs, outputs)
    print(rep)


def test_graphed_loss_matches_the_eager_dropins(cuda_lib):
    """trainer_hooks.GraphedLoss (CUDA-graph replay of generate_images_pred + compute_losses on static slots)
    gives the eager drop-ins' losses, selection and gradients, step after step, with uint8 ingest in the graph."""
    from types import SimpleNamespace
    from ssde_b200 import trainer_hooks, hostio
    dev = torch.device("cuda")
    B, H, W = 2, 96, 320
    opt = synthetic.make_options(H, W, batch_size=B)
    opt.pml_sources, opt.pml_variant, opt.pml_emit_depth = [-1, 1], "trainer", "scale0"
    ns = SimpleNamespace(opt=opt, device=dev, num_scales=4)
    frames = [0, -1, 1]
    batches = []
    for seed in (1, 2, 3):
        i, o = synthetic.make_batch(B, H, W, seed=seed)
        hb = {"color_u8": torch.stack([(i[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8)
                                       for f in frames], 0).contiguous()}
        hb.update({k: v for k, v in i.items() if not (isinstance(k, tuple) and k[0] == "color")})
        hb.update({k: v for k, v in o.items() if k[0] in ("disp", "cam_T_cam")})
        batches.append(hostio.PinnedBatch(hb))
    d, arena = batches[0].upload(dev)
    inp = {k: v for k, v in d.items() if not (isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam"))}
    out = {k: v.requires_grad_(True) for k, v in d.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")}
    runner = trainer_hooks.GraphedLoss(ns)
    slot = runner.capture(inp, out)
    for hb in batches:
        hb.upload_into(arena)
        for k, v in out.items():
            if k[0] in ("disp", "cam_T_cam"):
                v.grad = None
        losses = slot.replay()
        losses["loss"].backward()
        got = {k: v.detach().clone() for k, v in losses.items()}
        got_g = {k: v.grad.clone() for k, v in out.items() if k[0] in ("disp", "cam_T_cam")}
        got_am = {s: out[("argmin", s)].clone() for s in opt.scales} if ("argmin", 0) in out else None
        # eager reference on copies of the same device tensors
        d2, _ = hb.upload(dev)
        inp2 = {k: v for k, v in d2.items() if not (isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam"))}
        out2 = {k: v.requires_grad_(True) for k, v in d2.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")}
        trainer_hooks.ingest_colors(inp2, frames, 4, device=dev)
        o2 = SimpleNamespace(**vars(opt)); o2.pml_noise = "philox"
        ns2 = SimpleNamespace(opt=o2, device=dev, num_scales=4)
        trainer_hooks.generate_images_pred(ns2, inp2, out2)
        want = trainer_hooks.compute_losses(ns2, inp2, out2)
        want["loss"].backward()
        for k in want:
            assert common.rel_err(got[k].cpu(), want[k].detach().cpu()) < 2e-6, k     # another Philox stream on near-ties only
        for k in got_g:     # the two runs draw different tie-break noise: a handful of near-tie pixels select differently
            assert parity.l2_err(got_g[k].cpu(), out2[k].grad.cpu()) < 2e-2, (k, parity.l2_err(got_g[k].cpu(), out2[k].grad.cpu()))
        if got_am is not None:
            for s in opt.scales:
                assert (got_am[s] != out2[("argmin", s)]).float().mean().item() < 1e-3


def test_graphed_loss_dropin_call_with_network_outputs(cuda_lib):
    """GraphedLoss.__call__: this step's tensors live at new addresses (a DataLoader batch, network outputs with an
    autograd history); they are copied into slot 0, the graph is replayed and the gradients flow back into the
    'network' parameters like through the eager drop-ins."""
    from types import SimpleNamespace
    from ssde_b200 import trainer_hooks
    dev = torch.device("cuda")
    B, H, W = 2, 64, 160
    opt = synthetic.make_options(H, W, batch_size=B)
    opt.pml_sources, opt.pml_variant, opt.pml_emit_depth, opt.pml_noise = [-1, 1], "trainer", "scale0", "philox"
    ns = SimpleNamespace(opt=opt, device=dev, num_scales=4)
    runner = trainer_hooks.GraphedLoss(ns)
    for seed in (4, 5, 6):
        i, o = synthetic.make_batch(B, H, W, seed=seed)
        inp = {k: v.to(dev) for k, v in i.items()}
        res = {}
        for mode in ("graph", "eager"):
            # a stand-in network: disparity = sigmoid(logit parameter), pose = parameter
            logits = {s: torch.logit(o[("disp", s)].to(dev).clamp(1e-3, 1 - 1e-3)).requires_grad_(True) for s in opt.scales}
            poses = {f: o[("cam_T_cam", 0, f)].to(dev).clone().requires_grad_(True) for f in (-1, 1)}
            out = {("disp", s): torch.sigmoid(logits[s]) for s in opt.scales}
            out.update({("cam_T_cam", 0, f): poses[f] * 1.0 for f in (-1, 1)})
            if mode == "graph":
                losses = runner(dict(inp), out)
            else:
                trainer_hooks.generate_images_pred(ns, dict(inp), out)
                losses = trainer_hooks.compute_losses(ns, dict(inp), out)
            losses["loss"].backward()
            res[mode] = (losses["loss"].detach().clone(), {s: logits[s].grad.clone() for s in opt.scales},
                         {f: poses[f].grad.clone() for f in (-1, 1)}, out[("depth", 0, 0)].detach().clone())
        g, e = res["graph"], res["eager"]
        assert common.rel_err(g[0].cpu(), e[0].cpu()) < 2e-6
        for s in opt.scales:
            assert parity.l2_err(g[1][s].cpu(), e[1][s].cpu()) < 2e-2, s      # different Philox streams on near-ties
        for f in (-1, 1):
            assert parity.l2_err(g[2][f].cpu(), e[2][f].cpu()) < 2e-2, f
        assert common.rel_err(g[3].cpu(), e[3].cpu()) < 1e-6


def test_tensors_on_a_non_current_device(cuda_lib):
    """The reference trainer keeps its tensors on cuda:1 / cuda:3 without ever calling set_device
    (trainer.py:44,67): every libpml launch must follow its tensors' device, not the current one."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    assert torch.cuda.current_device() == 0
    B, H, W = 2, 64, 160
    opt = synthetic.make_options(H, W, batch_size=B)
    inputs, outputs = synthetic.make_batch(B, H, W, seed=3)
    got = common.run_product(opt, inputs, outputs, "trainer", device="cuda:1", noise_seed=7)
    assert torch.cuda.current_device() == 0
    parity.check(got, opt, "trainer", inputs, outputs, 7)


@pytest.mark.parametrize("kernel", ["cta", "sweep"])
def test_kernel_generations_agree(cuda_lib, kernel):
    """The CTA-strip kernel and the warp-strip sweep (pair sweeps for S > 2) are interchangeable:
    each passes the same parity check."""
    for sources in ((-1, 1), (-1, 1, "s")):
        B, H, W = 2, 96, 320
        opt = synthetic.make_options(H, W, batch_size=B)
        inputs, outputs = synthetic.make_batch(B, H, W, sources=sources, seed=41)
        got = common.run_product(opt, inputs, outputs, "trainer", device="cuda", noise_seed=6, sources=sources,
                                 extra_opt=dict(pml_kernel=kernel))
        parity.check(got, opt, "trainer", inputs, outputs, 6, sources=sources)


def _full_size(B=12, H=192, W=640, seed=0):
    opt = synthetic.make_options(H, W, batch_size=B)
    inputs, outputs = synthetic.make_batch(B, H, W, seed=seed)
    return opt, inputs, outputs


def test_full_size_determinism_and_forward_only(cuda_lib):
    opt, inputs, outputs = _full_size()
    a = common.run_product(opt, inputs, outputs, device="cuda", noise_seed=1)
    b = common.run_product(opt, inputs, outputs, device="cuda", noise_seed=1)
    c = common.run_product(opt, inputs, outputs, device="cuda", noise_seed=1, want_grad=False)
    for s in opt.scales:
        assert torch.equal(a["argmin/%d" % s], b["argmin/%d" % s])
        assert torch.equal(a["argmin/%d" % s], c["argmin/%d" % s])
        assert a["loss/%d" % s].item() == b["loss/%d" % s].item()
        # the forward-only call uses shorter row chunks (more resident warps): same pixels, another summation order
        assert abs(c["loss/%d" % s].item() - a["loss/%d" % s].item()) <= 1e-6 * abs(a["loss/%d" % s].item())
        assert torch.equal(a["grad_disp/%d" % s], b["grad_disp/%d" % s]) or \
            common.rel_err(a["grad_disp/%d" % s], b["grad_disp/%d" % s]) < 1e-6   # 4-way corner atomics
    assert torch.equal(a["grad_T/1"], b["grad_T/1"])


def test_full_size_batch_decomposition(cuda_lib):
    """The loss is a mean over images (trainer.py:610): B=12 equals the mean of twelve B=1 runs
    (same per-image noise), and each image's disparity gradient is 1/12 of its stand-alone one."""
    opt, inputs, outputs = _full_size()
    B = 12
    seed = 3
    full = common.run_product(opt, inputs, outputs, device="cuda", noise_seed=seed)
    noise = synthetic.draw_noise(B, opt.height, opt.width, opt.scales, 2, seed=seed)
    acc = {s: 0.0 for s in opt.scales}
    from ssde_b200 import functional as Fn
    dev = torch.device("cuda")
    for b in (0, 5, 11):
        sl = slice(b, b + 1)
        disps = [outputs[("disp", s)][sl].to(dev).requires_grad_(True) for s in opt.scales]
        out = Fn.photometric_loss(
            inputs[("color", 0, 0)][sl].to(dev), [inputs[("color", f, 0)][sl].to(dev) for f in (-1, 1)],
            inputs[("K", 0)][sl].to(dev), inputs[("inv_K", 0)][sl].to(dev),
            [outputs[("cam_T_cam", 0, f)][sl].to(dev) for f in (-1, 1)], disps,
            [inputs[("color", 0, s)][sl].to(dev) for s in opt.scales],
            smooth_weights=[1e-3 / 2 ** s for s in opt.scales], noise=[n[sl].to(dev) for n in noise])
        (out["loss"].sum() / len(opt.scales)).backward()
        for i, s in enumerate(opt.scales):
            assert torch.equal(out["argmin"][i][0].cpu(), full["argmin/%d" % s][b])
            # photometric part scales by 1/B exactly; the smoothness normaliser B*h*(w-1) too
            assert common.rel_err(disps[i].grad.cpu()[0] / B, full["grad_disp/%d" % s][b]) < 2e-5


def test_full_size_gradient_linearity(cuda_lib):
    opt, inputs, outputs = _full_size(B=4)
    from types import SimpleNamespace
    from ssde_b200 import trainer_hooks
    dev = torch.device("cuda")
    grads = []
    for scale in (1.0, 2.5):
        o = SimpleNamespace(**vars(opt)); o.pml_noise = "host"
        ns = SimpleNamespace(opt=o, device=dev, num_scales=4)
        inp = {k: v.to(dev) for k, v in inputs.items()}
        out = {k: v.to(dev).clone() for k, v in outputs.items()}
        out[("disp", 0)].requires_grad_(True)
        out[("cam_T_cam", 0, 1)].requires_grad_(True)
        torch.manual_seed(0)
        trainer_hooks.generate_images_pred(ns, inp, out)
        losses = trainer_hooks.compute_losses(ns, inp, out)
        (losses["loss"] * scale).backward()
        grads.append((out[("disp", 0)].grad.clone(), out[("cam_T_cam", 0, 1)].grad.clone()))
    assert common.rel_err(grads[1][0].cpu(), grads[0][0].cpu() * 2.5) < 1e-6
    assert common.rel_err(grads[1][1].cpu(), grads[0][1].cpu() * 2.5) < 1e-6


def test_depth_metrics(cuda_lib):
    import layer_checks
    layer_checks.depth_metrics("cuda")


def test_layer_dropins(cuda_lib):
    import layer_checks
    layer_checks.run("cuda")
    layer_checks.run("cuda", B=3, H=40, W=300)


def test_reference_unfused_code_path_on_dropin_layers(cuda_lib):
    """The reference's own unfused recipe (BackprojectDepth -> Project3D -> F.grid_sample ->
    SSIM/L1, trainer.py:501-529) written against the drop-in layers reproduces the oracle."""
    import torch.nn.functional as F
    from oracle import photometric_oracle as po
    from ssde_b200 import layers as L
    B, H, W = 2, 64, 160
    opt = synthetic.make_options(H, W, batch_size=B)
    inputs, outputs = synthetic.make_batch(B, H, W, seed=8)
    dev = torch.device("cuda")
    disp = outputs[("disp", 0)].to(dev)
    _, depth = L.disp_to_depth(disp, opt.min_depth, opt.max_depth)
    cam = L.BackprojectDepth(B, H, W).to(dev)(depth, inputs[("inv_K", 0)].to(dev))
    grid = L.Project3D(B, H, W).to(dev)(cam, inputs[("K", 0)].to(dev), outputs[("cam_T_cam", 0, 1)].to(dev))
    pred = F.grid_sample(inputs[("color", 1, 0)].to(dev), grid, padding_mode="border", align_corners=False)
    tgt = inputs[("color", 0, 0)].to(dev)
    rp = 0.85 * L.SSIM()(pred, tgt).mean(1, True) + 0.15 * (tgt - pred).abs().mean(1, True)
    o = {k: v.double() for k, v in outputs.items()}
    i = {k: v.double() for k, v in inputs.items()}
    po.generate_images_pred(opt, i, o, sources=(1,))
    want = po.reprojection_loss(o[("color", 1, 0)], i[("color", 0, 0)])
    assert (rp.double().cpu() - want).abs().max().item() < 5e-4
    assert abs(rp.mean().item() - want.mean().item()) / want.mean().item() < 1e-5
