"""Memory-safety sweep of the kernel sources: the host-emulator build compiled with AddressSanitizer
(tests/emu/run_asan.sh) executes the fused paths, the pyramid kernels and the layer kernels on shapes
that exercise strip / chunk / tile borders.  compute-sanitizer is not available on the GPU pool, so
this is the out-of-bounds check of the kernel indexing.  TEST INFRASTRUCTURE (optional, a few minutes)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
from ssde_b200 import _cabi, synthetic, functional as Fn
lib = _cabi.Library(os.environ["PML_EMU_ASAN_LIB"], emulator=True); _cabi.set_library_for_testing(lib)
import common, parity
# fused path: golden default, predmask, gru, wide (multi strip / chunk), 5 sources with stereo, S=8, odd sizes
for name in ("trainer_default", "trainer_predmask", "trainer_wide", "gru_seq3", "trainer_v1multiscale"):
    variant, opt, inputs, outputs, r32, r64, seed = common.load_golden(name)
    got = common.run_product(opt, inputs, outputs, variant, device="cpu", noise_seed=seed)
    print(name, "ok", float(got["loss"]))
for sources, (B, H, W) in (((-1, 1, -2, 2, "s"), (1, 32, 64)), ((-1, 1, -2, 2, -3, 3, -4, 4), (1, 32, 64)), ((-1, 1), (3, 40, 104)), ((1,), (2, 24, 40))):
    opt = synthetic.make_options(H, W, batch_size=B, scales=[0, 1, 2] if H % 8 else [0, 1, 2, 3])
    inputs, outputs = synthetic.make_batch(B, H, W, sources=sources, seed=6, scales=opt.scales)
    got = common.run_product(opt, inputs, outputs, "trainer", device="cpu", noise_seed=None, sources=sources)
    print(sources, (B, H, W), "ok", float(got["loss"]))
# default training configuration (in-kernel noise, no by-product stores): selection pre-pass + COMMON modes 1/3/2, the scalar
# single-frame instantiation (3 and 5 sources, stereo-only), odd sizes
for sources, (B, H, W) in (((-1, 1, "s"), (2, 40, 104)), ((-1, 1, -2, 2), (1, 32, 64)), ((-1, 1, -2, 2, "s"), (1, 24, 40)), (("s",), (2, 40, 72)),
                           ((-1, 1, -2, 2, -3, 3, "s"), (1, 24, 40))):
    opt = synthetic.make_options(H, W, batch_size=B, scales=[0, 1, 2])
    inputs, outputs = synthetic.make_batch(B, H, W, sources=sources, seed=7, scales=opt.scales)
    got = common.run_product(opt, inputs, outputs, "trainer", device="cpu", noise_seed=None, sources=sources,
                             extra_opt=dict(pml_emit_warped=False, pml_emit_depth="scale0"))
    print("default configuration", sources, (B, H, W), "ok", float(got["loss"]))
# pyramid: packed + scalar kernels, border tiles, non-multiple tile sizes
rng = np.random.default_rng(0)
for (N, H, W, n) in ((2, 96, 160, 4), (1, 32, 64, 4), (1, 8, 16, 4), (1, 40, 72, 2), (3, 192, 640, 4)):
    fr = torch.from_numpy(rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8))
    out = Fn.color_pyramid(fr, n)
    print("pyramid", (N, H, W, n), "ok", float(out[-1].mean()))
# disparity heads: forward / backward on sizes off the 30-column strips and 16-row chunks, channel split over warps
for (B, C, h, w) in ((2, 16, 37, 95), (1, 128, 6, 5), (1, 64, 17, 31), (3, 32, 16, 30), (9, 5, 70, 150), (40, 3, 33, 65), (40, 6, 33, 68)):   # the last three: tiled kernels (register tile, staged)
    x = torch.randn(B, C, h, w, requires_grad=True)
    wt, bs = (torch.randn(1, C, 3, 3) * 0.1).requires_grad_(True), torch.zeros(1, requires_grad=True)
    d = Fn.disp_head(x, wt, bs)
    d.backward(torch.randn_like(d))
    print("disp head", (B, C, h, w), "ok", float(d.mean()))
# selection masks: pixel counts off the 16-pixel vector path
for n in ((2, 3, 5, 7), (4, 1, 8, 16), (1, 2, 33, 65)):
    am = torch.randint(0, 4, n, dtype=torch.uint8)
    m = Fn.selection_masks(am, 2)
    assert torch.equal(m, (am > 1).float())
print("selection masks ok")
# cross-check kernel (CTA strips) on a multi-strip shape
opt = synthetic.make_options(40, 104, batch_size=2, scales=[0, 1, 2])
inputs, outputs = synthetic.make_batch(2, 40, 104, seed=6, scales=opt.scales)
got = common.run_product(opt, inputs, outputs, "trainer", device="cpu", noise_seed=3, extra_opt=dict(pml_kernel="cta"))
print("cta kernel ok", float(got["loss"]))
import layer_checks
layer_checks.run("cpu"); layer_checks.depth_metrics("cpu")
print("layers ok")
