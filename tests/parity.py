"""Parity criteria shared by the emulator tests (CPU) and the GPU tests.

Tolerances (north_star): per-scale loss within 1e-5 relative, gradients within 1e-4 relative, both
stated against the float64 CPU oracle; selection indices identical.  What makes those statements
testable for an fp32 implementation of a piecewise-smooth function:

* selection (``torch.min`` over the candidates): where the two best candidates differ by less than
  ``TIE_EPS`` in float64, fp32 evaluation -- the reference's own included -- may pick either.
  Indices must be identical everywhere else (``n_far == 0``), the near-tie flips must stay below
  ``FLIP_SHARE`` of the pixels (or half of the near-tie pixels, or twice what the fp32 reference flips
  against its own float64 run),
  and the report carries the mismatch counts against BOTH the float64 and the fp32 reference.
* gradients are compared against the float64 oracle evaluated *with the device's selection*
  (oracle ``forced_argmin``); the fp32 floor quoted beside each error is the fp32 oracle with the
  SAME forced selection, so neither contains arg-min flips.
* derivative kinks.  The loss has three kinds of points where the derivative jumps; a pixel whose
  float64 value sits within fp32 round-off of one may legitimately take either branch.  They are
  identified explicitly, counted, and masked -- nothing else is:
    - ``cell``:  a bilinear sampling coordinate within fp32 round-off of an integer or of the clip limits
                 (grid_sample's slope changes from one cell to the next / is switched off, trainer.py:508);
    - ``l1``:    the sign of (warped - target) in some channel flips when the sampling coordinate moves by
                 its fp32 round-off (sign(x - y), trainer.py:520);
    - ``clamp``: an SSIM value within CLAMP_EPS of 1 before the clamp (layers.py:248).  The clamp at 0 is
                 not a kink of the function: n/d <= 1 in exact arithmetic, only fp32 round-off on windows
                 with x ~ y gets there; the fused kernels follow float64 and keep the gradient.
  A kink at loss pixel p touches the disparity gradient of the 3x3 pixels around p (the SSIM window)
  and, at a lower scale, the low-resolution cells under them; those entries are excluded from the
  per-pixel comparison.  The pose gradient sums over all pixels: its allowance is twice the share of
  the float64 pose gradient that the kink pixels carry (measured with the oracle, ``pixel_weight``).
  The share of masked pixels per kind is asserted to stay below ``KINK_SHARE``.
* everything that is not masked must meet GRAD_TOL in the max-norm and in L2, plus three times the fp32
  oracle's own (forced, masked) deviation from float64: two fp32 evaluations of E[x^2] - mu^2 with
  independent rounding, compared by their worst entry out of 1e4 .. 1e6.  On most fixtures the library is
  closer to float64 than the ATen fp32 run; its worst entry has been seen at 3.7x the ATen run's worst.
* loss: LOSS_TOL plus twice the deviation of the fp32 oracle itself.  fp32 SSIM evaluates
  E[x^2] - mu^2 with ~5e-8 absolute noise against C2 = 9e-4, i.e. ~1e-5 noise per pixel on a small
  dissimilarity; averaged over few pixels that noise does not vanish (32x64 images: up to 6e-5
  relative for the ATen fp32 run and for this library alike; >= 64x160: below 1e-5 + floor).
"""
import torch
import torch.nn.functional as F

import common
from oracle import photometric_oracle as po
from ssde_b200 import synthetic

TIE_EPS = 5e-5
LOSS_TOL = 1e-5
LOSS_TOL_SMALL = 1e-4   # images with fewer than ~20k pixels: the per-pixel fp32 SSIM noise does not average out (see above)
GRAD_TOL = 1e-4
FLIP_SHARE = 5e-4      # near-tie selection flips allowed (observed: 1e-4 .. 3e-4 on the goldens)
KINK_SHARE = {"cell": 5e-2, "l1": 5e-2, "clamp": 1e-3}   # caps on the share of loss pixels per kink kind.  Observed on the
# KITTI-like views: cell <= 8e-4 (expected 2.4e-4 per coordinate), l1 <= 3e-3, clamp 0; the caps leave room for the iid-random stress
# images, whose warp fields are chaotic (fp32 coordinate errors of 1e-4 px): cell up to 2e-2 there.  Every report carries the shares.
L1_EPS = 2e-6           # floor of the |warped - target| threshold (the per-pixel threshold is 4x the fp32 oracle's own colour error)
CLAMP_EPS = 2e-5        # fp32 error of an SSIM value near the clamp (E[x^2] - mu^2 cancellation against C2 = 9e-4)
PHILOX_BOUND = 1.4e-4  # Box-Muller on 32-bit uniforms: |noise| <= 6.66, two candidates, x 1e-5


def l2_err(a, b):
    a, b = a.double(), b.double()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def _ssim_raw(x, y):
    """layers.py:234-248 before the clamp: (1 - n/d) / 2, float64."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sx = F.avg_pool2d(x * x, 3, 1) - mu_x * mu_x
    sy = F.avg_pool2d(y * y, 3, 1) - mu_y * mu_y
    sxy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + po.SSIM_C1) * (2 * sxy + po.SSIM_C2)
    d = (mu_x * mu_x + mu_y * mu_y + po.SSIM_C1) * (sx + sy + po.SSIM_C2)
    return (1 - n / d) / 2


def kink_pixels(o64, opt, variant, inputs, sources, scale, argmin, o32=None):
    """-> dict of [B,H,W] bool maps (cell, l1, clamp) of loss pixels on a derivative kink (float64).
    Only the candidate that ``argmin`` selects at a pixel carries gradient there, so only its kinks count.

    The yardstick is the fp32 error of a sampling coordinate: ``tol`` = 16 ulps, or 4x the deviation of the
    fp32 oracle's own coordinate from float64 in the 3x3 neighbourhood where the geometry is ill-conditioned
    (large parallax, near the epipole; ``o32``).  ``cell``: an integer (or the clip limits 0, n-1) lies within
    ``tol`` of the coordinate.  ``l1``: re-sampling the source at the coordinate shifted by +-tol changes the
    sign of (warped - target) in some channel, or |warped - target| < L1_EPS."""
    n_seq = opt.len_sequence if variant == "gru" else 0
    src_scale = scale if (opt.v1_multiscale and variant != "fusion") else 0
    inp = synthetic.to_sequence_layout(inputs, opt.len_sequence) if variant == "gru" else inputs
    target = po._gather(inp, ("color", 0, src_scale), n_seq).double()
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else len(sources))
    cell = torch.zeros(argmin.shape, dtype=torch.bool)
    l1, clamp = cell.clone(), cell.clone()
    for fi, f in enumerate(sources):
        sel = (argmin == n_id) if opt.avg_reprojection else (argmin == n_id + fi)
        # the warped value of pixel q enters the SSIM windows of the 3x3 pixels around it: its sampling
        # slopes matter wherever one of those windows selected this frame
        sel_near = F.max_pool2d(sel[:, None].float(), 3, 1, 1)[:, 0] > 0
        g = o64["sample/%s/%d" % (f, scale)]
        g32 = o32["sample/%s/%d" % (f, scale)].double() if o32 is not None else g
        Hs, Ws = g.shape[1], g.shape[2]
        tols = []
        for c, n in ((0, Ws), (1, Hs)):
            ix = ((g[..., c] + 1) * n - 1) / 2
            dev = (((g32[..., c] + 1) * n - 1) / 2 - ix).abs()
            dev = F.max_pool2d(dev[:, None], 3, 1, 1)[:, 0]
            tol = torch.maximum(1e-6 * (ix.abs() + 16), 4 * dev)
            tols.append(tol)
            # including 0 and n-1, where the clip of grid_sample's border mode switches the slope off;
            # coordinates well outside are not ambiguous
            cell |= sel_near & ((ix - ix.round()).abs() < tol) & (ix > -tol) & (ix < n - 1 + tol)
        x = o64["color/%s/%d" % (f, scale)].double()
        src = po._gather(inp, ("color", f, src_scale), n_seq).double()
        sgn = torch.sign(x - target)
        amb = ((x - target).abs() < L1_EPS).any(1)
        for sx in (-1.0, 1.0):
            for sy in (-1.0, 1.0):
                gp = torch.stack([g[..., 0] + sx * 2 * tols[0] / Ws, g[..., 1] + sy * 2 * tols[1] / Hs], -1)
                amb |= (torch.sign(po.warp(src, gp) - target) != sgn).any(1)
        l1 |= sel & amb
        if not opt.no_ssim:
            raw = _ssim_raw(x, target)
            clamp |= sel & ((raw - 1).abs() < CLAMP_EPS).any(1)
    return {"cell": cell, "l1": l1, "clamp": clamp}


def oracle_pair(opt, variant, inputs, outputs, seed, sources=(-1, 1), forced=None, dtype=torch.float64,
                pixel_weight=None, zero_noise=False):
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else len(sources))
    v1 = opt.v1_multiscale and variant != "fusion"
    B = outputs[("disp", opt.scales[0])].shape[0]
    noise = synthetic.draw_noise(B, opt.height, opt.width, opt.scales, max(n_id, 1), seed=seed or 0, v1_multiscale=v1)
    if zero_noise:
        noise = [torch.zeros_like(n) for n in noise]
    inp = synthetic.to_sequence_layout(inputs, opt.len_sequence) if variant == "gru" else inputs
    return po.run(opt, inp, outputs, sources=sources, variant=variant, noise=noise if n_id else None,
                  dtype=dtype, forced_argmin=forced, pixel_weight=pixel_weight)


def check(got, opt, variant, inputs, outputs, seed, ref32=None, ref64=None, sources=(-1, 1),
          degenerate=False, report=None, philox=False, loss_tol=LOSS_TOL):
    """Assert parity of one product run ``got`` (common.run_product output).  ``ref32``/``ref64``:
    golden reference outputs when available (else the oracle stands in, pinned to them elsewhere).
    ``philox``: the product drew its tie-break noise in-kernel; the oracle then runs with zero noise
    and the selection is only compared where the float64 margin exceeds the noise bound."""
    o64 = oracle_pair(opt, variant, inputs, outputs, seed, sources, zero_noise=philox)
    if ref64 is None or philox:
        ref64 = o64
    if ref32 is None or philox:
        ref32 = oracle_pair(opt, variant, inputs, outputs, seed, sources, dtype=torch.float32, zero_noise=philox)
    rep = report if report is not None else {}
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else len(sources))
    tie_eps = PHILOX_BOUND if philox else TIE_EPS
    # ---- losses
    # fp32 floor of the total: the per-scale deviations of the fp32 reference can cancel in its own
    # total by luck (min over many frames -> tiny losses dominated by SSIM's variance cancellation),
    # so the total is judged against the magnitude-weighted per-scale floors as well
    floor_total = sum(abs(float(ref32["loss/%d" % s]) - float(ref64["loss/%d" % s])) for s in opt.scales) / \
        max(sum(abs(float(ref64["loss/%d" % s])) for s in opt.scales), 1e-30)
    for k in ["loss"] + ["loss/%d" % s for s in opt.scales]:
        e = common.rel_err(got[k], ref64[k])
        floor = common.rel_err(ref32[k], ref64[k])
        if k == "loss":
            floor = max(floor, floor_total)
        allow = loss_tol + 2 * floor
        if philox and k != "loss":
            # the reference adds noise * 1e-5 to the identity candidates it minimises over (trainer.py:594-597);
            # the zero-noise oracle does not: at most 6.66e-5 per pixel on which an identity candidate won
            s = int(k.split("/")[1])
            id_share = (got["argmin/%d" % s].long() < n_id).float().mean().item()
            allow += 6.66e-5 * id_share / max(abs(float(ref64[k])), 1e-30) * 0.05   # gated generator: runs on < 5 % of the rows
        rep[k] = (e, floor)
        assert e <= allow, "%s: rel err %.3e (fp32 reference itself %.3e)" % (k, e, floor)
    # ---- selection
    forced = {}
    for s in opt.scales:
        a = got["argmin/%d" % s]
        forced[s] = a.long()
        if not degenerate:
            # a single candidate (no automask + avg_reprojection): the reference takes no min at all
            r64 = ref64.get("argmin/%d" % s, o64["argmin/%d" % s])
            n, n_far = common.argmin_report(a, r64, o64["margin/%d" % s], eps=tie_eps)
            n32 = n_ref = None
            if "argmin/%d" % s in ref32:
                n32 = int((a.long() != ref32["argmin/%d" % s].long()).sum())
                n_ref = int((ref32["argmin/%d" % s].long() != r64.long()).sum())
            n_near = int((o64["margin/%d" % s] <= tie_eps).sum())
            rep["argmin/%d" % s] = {"vs_f64": n, "beyond_near_tie": n_far, "vs_f32_ref": n32, "f32_ref_vs_f64": n_ref,
                                    "near_tie_pixels": n_near, "pixels": a.numel()}
            assert n_far == 0, "scale %d: %d selection mismatches beyond near-ties" % (s, n_far)
            if not philox:
                # ... and rare even among the near-ties: fp32 noise (~1e-5) against TIE_EPS flips a minority of them
                assert n <= max(4, int(a.numel() * FLIP_SHARE), n_near // 2, 2 * (n_ref or 0)), \
                    "scale %d: %d near-tie flips of %d (%d near-tie pixels, fp32 reference: %s)" % (s, n, a.numel(), n_near, n_ref)
        k = "identity_selection/%d" % s
        if k in got:
            assert torch.equal(got[k].float(), (a.long() > n_id - 1).float()), k
    # ---- gradients, conditional on the device's own selection
    if any(k.startswith("grad_") for k in got):
        of = oracle_pair(opt, variant, inputs, outputs, seed, sources, forced=forced, zero_noise=philox)
        of32 = oracle_pair(opt, variant, inputs, outputs, seed, sources, forced=forced, dtype=torch.float32,
                           zero_noise=philox)     # forward quantities (sampling grid, warped colours) do not depend on `forced`
        masks, keep_w, keep_w3, shares = {}, {}, {}, {}
        for s in opt.scales:
            kinks = kink_pixels(o64, opt, variant, inputs, sources, s, forced[s], of32)
            any_k = kinks["cell"] | kinks["l1"] | kinks["clamp"]
            shares[s] = {k: round(v.float().mean().item(), 6) for k, v in kinks.items()}
            shares[s]["any"] = round(any_k.float().mean().item(), 6)
            for kind, cap in KINK_SHARE.items():
                assert shares[s][kind] <= cap or degenerate, "scale %d: %.2e of the pixels sit on a '%s' kink" % (s, shares[s][kind], kind)
            keep_w[s] = (~any_k).to(torch.float64)
            m = F.max_pool2d(any_k[:, None].float(), 3, 1, 1)         # the 3x3 SSIM window around a kink pixel
            keep_w3[s] = 1.0 - m[:, 0].to(torch.float64)
            hd = got["grad_disp/%d" % s].shape[2]
            k = m.shape[2] // hd
            if k > 1:
                m = F.max_pool2d(F.max_pool2d(m, k), 3, 1, 1)            # low-res cells under them (bilinear footprint)
            masks["grad_disp/%d" % s] = m > 0
        rep["kink_share"] = shares
        # share of every pose-type gradient carried by the kink pixels (float64, forced selection)
        # ... and by their 3x3 neighbourhoods: where float64 sits on the switched-off side of a clip limit the kink
        # pixel itself carries nothing, its neighbours show what the other branch would carry
        ok = oracle_pair(opt, variant, inputs, outputs, seed, sources, forced=forced, pixel_weight=keep_w, zero_noise=philox)
        ok3 = oracle_pair(opt, variant, inputs, outputs, seed, sources, forced=forced, pixel_weight=keep_w3, zero_noise=philox)
        for k in sorted(of):
            if not k.startswith("grad_"):
                continue
            a, b, b32 = got[k].double(), of[k].double(), of32[k].double()
            allow = allow2 = 0.0
            if k in masks:
                keep = (~masks[k]).double()
                a, b, b32 = a * keep, b * keep, b32 * keep
            elif not k.startswith("grad_mask"):
                d, d3 = (of[k].double() - ok[k].double()), (of[k].double() - ok3[k].double())
                allow = 2 * (d.abs().max().item() + d3.abs().max().item()) / max(of[k].abs().max().item(), 1e-300)
                allow2 = 2 * (d.norm().item() + d3.norm().item()) / max(of[k].double().norm().item(), 1e-300)
            e, e2 = common.rel_err(a, b), l2_err(a, b)
            f, f2 = common.rel_err(b32, b), l2_err(b32, b)
            rep[k] = {"max": e, "l2": e2, "fp32_ref_max": f, "fp32_ref_l2": f2, "kink_allowance": allow}
            assert e <= GRAD_TOL + 3 * f + allow, "%s: max-norm rel err %.3e (fp32 reference itself %.3e, kink allowance %.1e)" % (k, e, f, allow)
            assert e2 <= GRAD_TOL + 3 * f2 + allow2, "%s: L2 rel err %.3e (fp32 reference itself %.3e)" % (k, e2, f2)
    # ---- by-products (trainer.py:480, :508)
    for s in opt.scales:
        k = "depth/%d" % s
        if k in got and k in o64:
            assert common.rel_err(got[k], o64[k]) < 1e-5, k
        for f in sources:
            k = "color/%s/%d" % (f, s)
            if k in got and k in o64:
                # fp32 sampling coordinates (|error| ~ 1e-5 .. 1e-4 px) times the image slope (up to 1 / px on iid-random images)
                floor = (ref32[k].double() - o64[k]).abs().max().item() if k in ref32 else 0.0
                e = (got[k].double() - o64[k]).abs().max().item()
                assert e < 2e-4 + 2 * floor, "%s: max abs err %.2e (fp32 reference itself %.2e)" % (k, e, floor)
    return rep


def pose_gradient_check(device, opt, inputs, outputs, sources=(-1, 1)):
    """A sharp test of the pose (and disparity) gradients: the kink pixels are taken OUT of the loss on both
    sides instead of being bounded by an allowance.  The library runs the ``disable_automasking`` configuration
    with per-pixel frame weights (the predictive-mask input of trainer.py:571-579, here 0 on the 3x3
    neighbourhoods of kink pixels and 1 elsewhere); the float64 oracle weighs the same pixels out
    (``pixel_weight``).  What is left is smooth, so GRAD_TOL + 2 x (fp32 oracle's own deviation) must hold for
    d loss / d T without any allowance.  Returns the report."""
    from types import SimpleNamespace
    from ssde_b200 import functional as Fn
    o = SimpleNamespace(**vars(opt))
    o.disable_automasking = True
    dev = torch.device(device)
    scales = list(o.scales)
    H, W = o.height, o.width

    def tensors(req):
        tgt = inputs[("color", 0, 0)].to(dev)
        srcs = [inputs[("color", f, 0)].to(dev) for f in sources]
        K, iK = inputs[("K", 0)].to(dev), inputs[("inv_K", 0)].to(dev)
        Ts = [(inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]).to(dev).clone().requires_grad_(req) for f in sources]
        disps = [outputs[("disp", s)].to(dev).clone().requires_grad_(req) for s in scales]
        cols = [inputs[("color", 0, s)].to(dev) for s in scales]
        return tgt, srcs, K, iK, Ts, disps, cols
    kw = dict(smooth_weights=[o.disparity_smoothness / 2 ** s for s in scales], min_depth=o.min_depth, max_depth=o.max_depth,
              disable_automasking=True)
    tgt, srcs, K, iK, Ts, disps, cols = tensors(False)
    first = Fn.photometric_loss(tgt, srcs, K, iK, Ts, disps, cols, **kw)
    forced = {s: first["argmin"][i].long().cpu() for i, s in enumerate(scales)}
    o64 = oracle_pair(o, "trainer", inputs, outputs, 0, sources)
    o32 = oracle_pair(o, "trainer", inputs, outputs, 0, sources, dtype=torch.float32)
    keep = {}
    for s in scales:
        kinks = kink_pixels(o64, o, "trainer", inputs, sources, s, forced[s], o32)
        any_k = kinks["cell"] | kinks["l1"] | kinks["clamp"]
        # with weight 0 every candidate of a pixel is 0 and the selection there is moot; pixels whose selection
        # is a near-tie are weighed out too, so that both sides agree on the selection everywhere else
        any_k |= o64["margin/%d" % s] <= TIE_EPS
        keep[s] = 1.0 - F.max_pool2d(any_k[:, None].float(), 3, 1, 1)[:, 0]
    tgt, srcs, K, iK, Ts, disps, cols = tensors(True)
    fws = [keep[s][:, None].expand(-1, len(sources), -1, -1).contiguous().to(dev) for s in scales]
    out = Fn.photometric_loss(tgt, srcs, K, iK, Ts, disps, cols, frame_weights=fws, **kw)
    out["total"].backward()
    kw64 = {s: keep[s].double() for s in scales}
    sel = {s: out["argmin"][i].long().cpu() for i, s in enumerate(scales)}
    of = oracle_pair(o, "trainer", inputs, outputs, 0, sources, forced=sel, pixel_weight=kw64)
    of32 = oracle_pair(o, "trainer", inputs, outputs, 0, sources, forced=sel, pixel_weight=kw64, dtype=torch.float32)
    rep = {"kept_share": {s: round(keep[s].mean().item(), 5) for s in scales}}
    e = common.rel_err(out["total"].detach().cpu(), of["loss"])
    rep["loss"] = e
    assert e <= 1e-4, "loss with kink pixels weighed out: rel err %.3e" % e
    for fi, f in enumerate(sources):
        if f == "s":
            continue
        k = "grad_T/%s" % f
        a, b, b32 = Ts[fi].grad.double().cpu(), of[k].double(), of32[k].double()
        e, e2, fl, fl2 = common.rel_err(a, b), l2_err(a, b), common.rel_err(b32, b), l2_err(b32, b)
        rep[k] = {"max": e, "l2": e2, "fp32_ref_max": fl, "fp32_ref_l2": fl2}
        # d loss / d T is a cancelling sum over all pixels: both fp32 evaluations sit well above 1e-4 of float64 on it
        # (ATen: 5e-5 .. 5e-4), the library within 3x of the ATen run's own deviation (accumulating the per-pixel
        # terms in double changes its result in the 4th digit only: the noise is in the terms, not in the sum)
        assert e <= GRAD_TOL + 3 * fl, "%s (kink pixels weighed out): max-norm rel err %.3e (fp32 reference itself %.3e)" % (k, e, fl)
        assert e2 <= GRAD_TOL + 3 * fl2, "%s (kink pixels weighed out): L2 rel err %.3e (fp32 reference itself %.3e)" % (k, e2, fl2)
    for i, s in enumerate(scales):
        k = "grad_disp/%d" % s
        a, b, b32 = disps[i].grad.double().cpu(), of[k].double(), of32[k].double()
        e, fl = common.rel_err(a, b), common.rel_err(b32, b)
        rep[k] = {"max": e, "fp32_ref_max": fl}
        assert e <= GRAD_TOL + 3 * fl, "%s (kink pixels weighed out): max-norm rel err %.3e (fp32 reference itself %.3e)" % (k, e, fl)
    return rep
