"""Parity criteria shared by the emulator tests (CPU) and the GPU tests.

Tolerances (north_star): per-scale loss within 1e-5 relative, gradients within 1e-4 relative,
both stated against the float64 CPU oracle; selection indices identical.  Two refinements make
those statements testable for an fp32 implementation:

* fp32 rounding of near-ties: where the two best candidates differ by less than ``TIE_EPS`` in
  float64, fp32 evaluation (the reference's own included -- its golden fp32 run disagrees with its
  float64 run on the same pixels) may pick either.  Indices must be identical everywhere else, and
  the number of such pixels must stay tiny.
* gradients are compared against the float64 oracle evaluated *with the device's selection*
  (oracle ``forced_argmin``), so that a legitimately different near-tie choice does not mask or
  fake a gradient error; every pixel is compared, none is excluded.
* bilinear sampling has a discontinuous derivative where a sampling coordinate is an integer; a
  pixel whose float64 coordinate lies within a few fp32 ulps of one may legitimately take the
  neighbouring cell's slope.  Such pixels (typically 0-3 per test image) are masked out of the
  per-pixel disparity-gradient comparison, and the pose gradient (a sum over all pixels) gets an
  allowance of their share n_ambiguous / N.
* where the fp32 reference itself is further than the tolerance from float64 (sign of |x-y| at
  near-equality, SSIM variance cancellation on constant images), twice its own deviation is added.
"""
import torch

import common
from oracle import photometric_oracle as po
from ssde_b200 import synthetic

TIE_EPS = 5e-5
LOSS_TOL = 1e-5
GRAD_TOL = 1e-4


def l2_err(a, b):
    a, b = a.double(), b.double()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def ambiguous_cells(o64, opt, sources, scale):
    """[B,H,W] bool: some source's float64 sampling coordinate is within ~16 fp32 ulps of an integer."""
    amb = None
    for f in sources:
        g = o64["sample/%s/%d" % (f, scale)]
        Hs, Ws = g.shape[1], g.shape[2]
        for c, n in ((0, Ws), (1, Hs)):
            ix = ((g[..., c] + 1) * n - 1) / 2
            ix = ix.clamp(0, n - 1)
            a = (ix - ix.round()).abs() < 1e-6 * (ix.abs() + 16)
            # exactly clipped coordinates are not ambiguous (slope is zeroed on both sides)
            a &= (ix > 0) & (ix < n - 1)
            amb = a if amb is None else (amb | a)
    return amb


def oracle_pair(opt, variant, inputs, outputs, seed, sources=(-1, 1), forced=None, dtype=torch.float64):
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else len(sources))
    v1 = opt.v1_multiscale and variant != "fusion"
    B = outputs[("disp", opt.scales[0])].shape[0]
    noise = synthetic.draw_noise(B, opt.height, opt.width, opt.scales, max(n_id, 1), seed=seed, v1_multiscale=v1)
    inp = synthetic.to_sequence_layout(inputs, opt.len_sequence) if variant == "gru" else inputs
    return po.run(opt, inp, outputs, sources=sources, variant=variant, noise=noise if n_id else None,
                  dtype=dtype, forced_argmin=forced)


def check(got, opt, variant, inputs, outputs, seed, ref32=None, ref64=None, sources=(-1, 1),
          degenerate=False, report=None):
    """Assert parity of one product run ``got`` (common.run_product output).  ``ref32``/``ref64``:
    golden reference outputs when available (else the oracle stands in, pinned to them elsewhere)."""
    o64 = oracle_pair(opt, variant, inputs, outputs, seed, sources)
    if ref64 is None:
        ref64 = o64
    if ref32 is None:
        ref32 = oracle_pair(opt, variant, inputs, outputs, seed, sources, dtype=torch.float32)
    rep = report if report is not None else {}
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else len(sources))
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
        rep[k] = e
        assert e <= LOSS_TOL + 2 * floor, "%s: rel err %.3e (fp32 reference itself %.3e)" % (k, e, floor)
    # ---- selection
    forced = {}
    for s in opt.scales:
        a = got["argmin/%d" % s]
        forced[s] = a.long()
        if not degenerate:
            # a single candidate (no automask + avg_reprojection): the reference takes no min at all
            n, n_far = common.argmin_report(a, ref64.get("argmin/%d" % s, o64["argmin/%d" % s]), o64["margin/%d" % s], eps=TIE_EPS)
            rep["argmin/%d" % s] = (n, n_far)
            assert n_far == 0, "scale %d: %d selection mismatches beyond near-ties" % (s, n_far)
            # ... and rare: at most 0.5 % of the pixels, or twice what the fp32 reference itself flips
            # against its float64 run (many candidates with tiny losses sit inside fp32 round-off)
            n_ref = 0
            if "argmin/%d" % s in ref32 and "argmin/%d" % s in ref64:
                n_ref = int((ref32["argmin/%d" % s].long() != ref64["argmin/%d" % s].long()).sum())
            assert n <= max(4, a.numel() // 200, 2 * n_ref), "scale %d: %d near-tie flips of %d (fp32 reference: %d)" % (s, n, a.numel(), n_ref)
        k = "identity_selection/%d" % s
        if k in got:
            assert torch.equal(got[k], (a.long() > n_id - 1).float()), k
    # ---- gradients, conditional on the device's own selection
    if any(k.startswith("grad_") for k in got):
        of = oracle_pair(opt, variant, inputs, outputs, seed, sources, forced=forced)
        # fp32 noise floor of the reference itself, per gradient family (pose / disparity): the
        # pose gradient is a cancelling sum over all pixels, its fp32 noise varies frame to frame
        fam = {}
        for k in of:
            if k.startswith("grad_"):
                key = k.split("/")[0]
                f0, f1 = fam.get(key, (0.0, 0.0))
                fam[key] = (max(f0, common.rel_err(ref32[k], ref64[k])), max(f1, l2_err(ref32[k], ref64[k])))
        amb_share = 0.0
        masks = {}
        for s in opt.scales:
            amb = ambiguous_cells(o64, opt, sources, s)
            amb_share = max(amb_share, amb.float().mean().item())
            m = amb[:, None].float()
            hd = got["grad_disp/%d" % s].shape[2]
            k = m.shape[2] // hd
            if k > 1:
                m = torch.nn.functional.max_pool2d(m, k)
            masks["grad_disp/%d" % s] = torch.nn.functional.max_pool2d(m, 3, 1, 1) > 0
        rep["ambiguous_share"] = amb_share
        for k in sorted(of):
            if not k.startswith("grad_"):
                continue
            a, b = got[k].double(), of[k].double()
            if k in masks:
                keep = (~masks[k]).double()
                a, b = a * keep, b * keep
            e, e2 = common.rel_err(a, b), l2_err(a, b)
            f, f2 = fam[k.split("/")[0]]
            if k.startswith(("grad_T", "grad_axisangle", "grad_translation")):   # sums over all pixels (posecnn: through T)
                f, f2 = f + amb_share, f2 + amb_share
            rep[k] = (e, e2)
            assert e <= GRAD_TOL + 2 * f, "%s: max-norm rel err %.3e (fp32 reference itself %.3e)" % (k, e, f)
            assert e2 <= GRAD_TOL + 2 * f2, "%s: L2 rel err %.3e (fp32 reference itself %.3e)" % (k, e2, f2)
    # ---- by-products (trainer.py:480, :508)
    for s in opt.scales:
        k = "depth/%d" % s
        if k in got and k in o64:
            assert common.rel_err(got[k], o64[k]) < 1e-5, k
        for f in sources:
            k = "color/%s/%d" % (f, s)
            if k in got and k in o64:
                assert (got[k].double() - o64[k]).abs().max().item() < 2e-4, k
    return rep
