"""torch.autograd.Function wrappers over the C ABI (include/pml.h).

``photometric_loss`` is the fused path: one call = ``Trainer.generate_images_pred`` +
``Trainer.compute_losses`` (trainer.py:465-622) for a group of scales that share images.  The
CUDA sweep evaluates the loss and, when any input requires grad, its adjoint in the same pass;
``backward`` only scales the stored gradients by the incoming ones (pml_scale_grads).

PyTorch is plumbing here (device memory, current stream, autograd graph); every byte of arithmetic
happens inside libpml.so.  Tensors must be CUDA fp32; anything else raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import torch

from . import _cabi
from ._cabi import (PML_FLAG_AVG_REPROJ, PML_FLAG_KERNEL_CTA, PML_FLAG_NO_AUTOMASK, PML_FLAG_NO_SSIM, PML_MAX_PASSES,
                    PML_MAX_SOURCES, PmlProblem, get_library)


def _stream_ptr(t: torch.Tensor):
    if t.is_cuda:
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return ctypes.c_void_p(0)


def _check(t: torch.Tensor, name: str, lib) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s); the sm_100a kernels compute in fp32" % (name, t.dtype))
    if not t.is_cuda and not lib.emulator:
        raise RuntimeError("%s is on %s: libpml has no CPU path, move it to a CUDA device" % (name, t.device))
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _Plan:
    """Launch state of one problem configuration that survives from step to step: the pml_problem with its
    static fields filled, the workspace (re-used: every call is stream-ordered on the stream the plan was
    made for), the ctypes arrays of the backward.  Keyed by shapes, flags and stream in ``_PLANS``."""
    __slots__ = ("prob", "ws", "ws_bytes", "hd", "wd", "gptrs", "shapes", "gdisp_off", "gdisp_total", "small_off",
                 "small_total", "n_id", "segs")


_PLANS: Dict = {}
_PLANS_MAX = 64


def _tensor_list(x):
    return list(x) if isinstance(x, (list, tuple)) else None


def _make_plan(lib, cfg, target0, disps, dev, n_seg, seg_size):
    n_pass, S = cfg["n_pass"], cfg["S"]
    _, _, H, W = target0.shape
    B = disps[0].shape[0]
    pl = _Plan()
    prob = PmlProblem()
    prob.B, prob.H, prob.W, prob.S, prob.n_pass = B, H, W, S, n_pass
    prob.flags = cfg["flags"]
    prob.min_depth, prob.max_depth, prob.eps = cfg["min_depth"], cfg["max_depth"], cfg.get("eps", 1e-7)
    prob.loss_total_div = float(cfg.get("total_div") or n_pass)
    automask = not (cfg["flags"] & PML_FLAG_NO_AUTOMASK)
    pl.n_id = 0 if not automask else (1 if cfg["flags"] & PML_FLAG_AVG_REPROJ else S)
    pl.shapes = [(d.shape[2], d.shape[3]) for d in disps]
    off = 0
    pl.gdisp_off = []
    for i, (hd, wd) in enumerate(pl.shapes):
        ps = prob.passes[i]
        ps.hd, ps.wd = hd, wd
        ps.smooth_weight = cfg["smooth_weights"][i]
        pl.gdisp_off.append(off)
        off += (B * hd * wd + 3) // 4 * 4          # keep every slice 16-byte aligned
    pl.gdisp_total = off
    # one fp32 arena per call: terms [n_pass,4] | loss vector [n_pass] | total [1] | pad | grad_const [n_pass,B] | grad_T [n_pass,S,B,16]
    so = {"terms": 0, "vec": 4 * n_pass, "total": 5 * n_pass}
    o = (5 * n_pass + 1 + 3) // 4 * 4
    so["gconst"] = o
    o = (o + n_pass * B + 3) // 4 * 4
    so["gT"] = o
    pl.small_off, pl.small_total = so, o + n_pass * S * B * 16
    pl.segs = None
    if n_seg:
        pl.segs = _cabi.PmlSegments()
        pl.segs.n_seg, pl.segs.seg_size = n_seg, seg_size
        prob.segments = ctypes.pointer(pl.segs)
    pl.hd = (ctypes.c_int32 * n_pass)(*[s_[0] for s_ in pl.shapes])
    pl.wd = (ctypes.c_int32 * n_pass)(*[s_[1] for s_ in pl.shapes])
    pl.gptrs = (ctypes.c_void_p * n_pass)()
    pl.prob = prob
    pl.ws = None
    pl.ws_bytes = -1
    return pl


class _PhotometricLoss(torch.autograd.Function):
    """inputs: cfg (carries the tensors that never receive gradients: images, intrinsics, noise), then the
    differentiable tensors laid out as  disps[n_pass] | Ts[S] | frame_weights[n_pass or 0]"""

    @staticmethod
    def forward(ctx, cfg: Dict, *tensors):
        lib = get_library()
        n_pass, S = cfg["n_pass"], cfg["S"]
        has_fw = bool(cfg.get("has_fw"))
        tensors = [_check(t, "tensor[%d]" % i, lib) for i, t in enumerate(tensors)]
        disps = tensors[:n_pass]
        Ts = tensors[n_pass:n_pass + S]
        fws = tensors[n_pass + S:] if has_fw else [None] * n_pass
        img = cfg["images"]
        n_seg = img["n_seg"]

        def chk(x, name):
            if n_seg:
                return [_check(t, name, lib) for t in x]
            return _check(x, name, lib)
        target, K, inv_K = chk(img["target"], "target"), chk(img["K"], "K"), chk(img["inv_K"], "inv_K")
        sources = [chk(x, "source") for x in img["sources"]]
        colors = [chk(x, "smooth colour") for x in img["colors"]]
        noise = [_check(t, "noise", lib) for t in img["noise"]] if cfg["has_noise"] else [None] * n_pass
        target0 = target[0] if n_seg else target
        seg_size = target0.shape[0] if n_seg else 0
        B = disps[0].shape[0]
        _, C, H, W = target0.shape
        if C != 3:
            raise ValueError("target must be [B,3,H,W]")
        if (seg_size * n_seg if n_seg else target0.shape[0]) != B:
            raise ValueError("images and disparities disagree on the batch size")
        want_grad = any(ctx.needs_input_grad[1:])
        dev = target0.device
        stream = torch.cuda.current_stream(dev).cuda_stream if target0.is_cuda else 0

        key = (dev, stream, B, H, W, S, n_pass, cfg["flags"], tuple(d.shape[2:] for d in disps), cfg["min_depth"],
               cfg["max_depth"], tuple(cfg["smooth_weights"]), cfg.get("total_div"), n_seg, seg_size, has_fw)
        pl = _PLANS.get(key)
        if pl is None:
            if len(_PLANS) >= _PLANS_MAX:
                _PLANS.clear()
            pl = _PLANS[key] = _make_plan(lib, cfg, target0, disps, dev, n_seg, seg_size)
        prob = pl.prob
        n_id = pl.n_id
        img_shape = (seg_size if n_seg else B, 3, H, W)
        kshape = (seg_size if n_seg else B, 4, 4)

        prob.seed = cfg.get("seed", 0)
        sd = cfg.get("seed_device")
        prob.seed_device = sd.data_ptr() if sd is not None else None
        if n_seg:
            sg = pl.segs
            for j in range(n_seg):
                if target[j].shape != img_shape or K[j].shape != kshape or inv_K[j].shape != kshape:
                    raise ValueError("chunk %d of the batch has the wrong shape" % j)
                sg.target[j], sg.K[j], sg.inv_K[j] = target[j].data_ptr(), K[j].data_ptr(), inv_K[j].data_ptr()
                for f in range(S):
                    if sources[f][j].shape != img_shape:
                        raise ValueError("source %d, chunk %d has the wrong shape" % (f, j))
                    sg.sources[f][j] = sources[f][j].data_ptr()
            prob.target, prob.K, prob.inv_K = target[0].data_ptr(), K[0].data_ptr(), inv_K[0].data_ptr()
            for f in range(S):
                prob.sources[f] = sources[f][0].data_ptr()
        else:
            if K.shape != kshape or inv_K.shape != kshape:
                raise ValueError("K / inv_K must be [B,4,4]")
            prob.target, prob.K, prob.inv_K = target.data_ptr(), K.data_ptr(), inv_K.data_ptr()
            for f in range(S):
                if sources[f].shape != img_shape:
                    raise ValueError("source %d has the wrong shape" % f)
                prob.sources[f] = sources[f].data_ptr()
        for f in range(S):
            if Ts[f].shape != (B, 4, 4):
                raise ValueError("pose %d must be [B,4,4]" % f)
            prob.T[f] = Ts[f].data_ptr()

        f32 = dict(device=dev, dtype=torch.float32)
        emit_depth, emit_warped = cfg.get("emit_depth", ()), cfg.get("emit_warped", ())
        argmin_all = torch.empty((n_pass, B, H, W), device=dev, dtype=torch.uint8)
        small = torch.empty(pl.small_total if want_grad else pl.small_off["gconst"], **f32)
        garena = torch.empty(pl.gdisp_total, **f32) if want_grad else None
        so = pl.small_off
        depths, warpeds, gdisps, gfws = [], [], [], []
        am_ptr, n_img = argmin_all.data_ptr(), B * H * W
        for i in range(n_pass):
            d = disps[i]
            if d.dim() != 4 or d.shape[0] != B or d.shape[1] != 1:
                raise ValueError("disp %d must be [B,1,h,w]" % i)
            hd, wd = pl.shapes[i]
            ps = prob.passes[i]
            cshape = (seg_size if n_seg else B, 3, hd, wd)
            if n_seg:
                for j in range(n_seg):
                    if colors[i][j].shape != cshape:
                        raise ValueError("smoothness colour %d, chunk %d must match its disparity" % (i, j))
                    pl.segs.smooth_color[i][j] = colors[i][j].data_ptr()
                ps.smooth_color = colors[i][0].data_ptr()
            else:
                if colors[i].shape != cshape:
                    raise ValueError("smoothness colour %d must match its disparity: %s vs %s"
                                     % (i, tuple(colors[i].shape), tuple(d.shape)))
                ps.smooth_color = colors[i].data_ptr()
            ps.disp = d.data_ptr()
            if noise[i] is not None:
                if tuple(noise[i].shape) != (B, n_id, H, W):
                    raise ValueError("noise %d must be [B,%d,H,W]" % (i, n_id))
                ps.noise = noise[i].data_ptr()
            else:
                ps.noise = None
            ps.argmin = am_ptr + i * n_img
            dp = torch.empty((B, 1, H, W), **f32) if i in emit_depth else None
            wp = torch.empty((S, B, 3, H, W), **f32) if i in emit_warped else None
            ps.depth, ps.warped = _ptr(dp), _ptr(wp)
            depths.append(dp)
            warpeds.append(wp)
            if want_grad:
                g = garena[pl.gdisp_off[i]:pl.gdisp_off[i] + B * hd * wd].view(B, 1, hd, wd)
                ps.grad_disp = g.data_ptr()
                gdisps.append(g)
            else:
                ps.grad_disp = None
            if fws[i] is not None:
                if tuple(fws[i].shape) != (B, S, H, W):
                    raise ValueError("frame weights %d must be [B,%d,H,W] (the predictive mask at full resolution)" % (i, S))
                ps.frame_weight = fws[i].data_ptr()
                if want_grad:
                    gw = torch.empty_like(fws[i])
                    ps.grad_frame_weight = gw.data_ptr()
                    gfws.append(gw)
                else:
                    ps.grad_frame_weight = None
            else:
                ps.frame_weight = ps.grad_frame_weight = None
        sp = small.data_ptr()
        prob.losses, prob.loss_vector, prob.loss_total = sp + 4 * so["terms"], sp + 4 * so["vec"], sp + 4 * so["total"]
        if want_grad:
            prob.grad_disp_const, prob.grad_T = sp + 4 * so["gconst"], sp + 4 * so["gT"]
        else:
            prob.grad_disp_const = prob.grad_T = None

        prof = cfg.get("prof_events")
        if prof is not None:
            prob.prof_start, prob.prof_stop = prof[0].cuda_event, prof[1].cuda_event
        else:
            prob.prof_start = prob.prof_stop = None
        if pl.ws_bytes < 0:
            pl.ws_bytes = lib.pml_workspace_bytes(ctypes.byref(prob))
            if pl.ws_bytes == 0:
                del _PLANS[key]
                raise _cabi.PmlError("pml_workspace_bytes rejected the problem (unsupported shape: H/hd must be a "
                                     "power of two shared by both axes, S <= %d, scales <= %d)" % (PML_MAX_SOURCES, PML_MAX_PASSES))
            pl.ws = torch.empty((pl.ws_bytes + 15) // 16 * 16, device=dev, dtype=torch.uint8)
        fn = lib.pml_loss_forward_backward if want_grad else lib.pml_loss_forward
        lib.check(fn(ctypes.byref(prob), _ptr(pl.ws), pl.ws_bytes, ctypes.c_void_p(stream)),
                  "pml_loss_forward_backward" if want_grad else "pml_loss_forward")

        ctx.want_grad = want_grad
        ctx.set_materialize_grads(False)   # no zero-filled grads for the non-differentiable by-products
        if want_grad:
            ctx.plan, ctx.gdisps, ctx.small, ctx.gfws, ctx.garena = pl, gdisps, small, gfws, garena
            ctx.dims = (n_pass, S, B, has_fw, stream)
            ctx.consumed = False
        loss_vec = small[so["vec"]:so["vec"] + n_pass]
        total = small[so["total"]:so["total"] + 1].view(())
        terms = small[:4 * n_pass].view(n_pass, 4)
        outs = [loss_vec, total, terms, argmin_all] + [t for t in depths if t is not None] + [t for t in warpeds if t is not None]
        ctx.mark_non_differentiable(*outs[2:])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_vec, g_total, *unused):
        n_in = len(ctx.needs_input_grad)
        if not ctx.want_grad or (g_vec is None and g_total is None):
            return (None,) * n_in
        if ctx.consumed:
            raise RuntimeError("libpml photometric-loss gradients were already consumed in place by a previous "
                               "backward(); re-run the forward pass instead of retain_graph")
        ctx.consumed = True
        grads = _scale_gradients(ctx.plan, ctx.dims, ctx.gdisps, ctx.small, ctx.gfws, g_vec, g_total, ctx.needs_input_grad[1:])
        # the gradient buffers are handed over without keeping a reference: AccumulateGrad then adopts them as
        # the leaves' .grad instead of cloning them (one 5.9 MB copy kernel per scale-0 disparity otherwise)
        ctx.gdisps = ctx.gfws = ctx.small = ctx.garena = None
        return (None,) + tuple(grads) + (None,) * (n_in - 1 - len(grads))


def _scale_gradients(pl, dims, gdisps, small, gfws, g_vec, g_total, need):
    """Backward of the fused node: the sweep already left d loss_s / d (disp_s, T_f) in ``gdisps`` / ``small``;
    one launch (pml_scale_grads) multiplies them by the incoming gradients in place.  -> grads for
    disps | Ts | frame weights."""
    lib = get_library()
    n_pass, S, B, has_fw, _ = dims
    # autograd runs a node's backward on the stream of its forward; a replayed graph runs on the current stream
    stream = torch.cuda.current_stream(small.device).cuda_stream if small.is_cuda else 0
    if g_vec is not None:
        g_vec = g_vec.contiguous().to(torch.float32)
    if g_total is not None:
        g_total = g_total.contiguous().to(torch.float32)
    so = pl.small_off
    gT_out = torch.empty((S, B, 4, 4), device=small.device, dtype=torch.float32)
    for i in range(n_pass):
        pl.gptrs[i] = gdisps[i].data_ptr()
    sp = small.data_ptr()
    lib.check(lib.pml_scale_grads(n_pass, B, S, pl.hd, pl.wd, pl.gptrs, sp + 4 * so["gconst"], sp + 4 * so["gT"],
                                  _ptr(g_vec), _ptr(g_total), pl.prob.loss_total_div, _ptr(gT_out),
                                  ctypes.c_void_p(stream)), "pml_scale_grads")
    grads: List[Optional[torch.Tensor]] = []
    for i in range(n_pass):
        grads.append(gdisps[i] if need[i] else None)
    for f in range(S):
        grads.append(gT_out[f] if need[n_pass + f] else None)
    if gfws:   # d loss_s / d mask_s, scaled by the incoming gradient
        up = (g_vec if g_vec is not None else 0) + (g_total / pl.prob.loss_total_div if g_total is not None else 0)
        for i, gw in enumerate(gfws):
            grads.append(gw.mul_(up[i] if up.dim() else up) if need[n_pass + S + i] else None)
    return grads


def photometric_loss(target, sources: Sequence, K, inv_K, Ts: Sequence[torch.Tensor],
                     disps: Sequence[torch.Tensor], smooth_colors: Sequence, *,
                     smooth_weights: Sequence[float], min_depth=0.1, max_depth=100.0,
                     no_ssim=False, disable_automasking=False, avg_reprojection=False,
                     noise: Optional[Sequence[torch.Tensor]] = None, seed: int = 0,
                     emit_depth: Sequence[int] = (), emit_warped: Sequence[int] = (), prof_events=None,
                     frame_weights: Optional[Sequence[torch.Tensor]] = None, kernel: str = "sweep",
                     total_div: Optional[float] = None, seed_device: Optional[torch.Tensor] = None):
    """Fused view synthesis + photometric loss for ``len(disps)`` scales sharing one image set.

    Returns a dict: ``loss`` [n_pass] (differentiable w.r.t. ``disps`` and ``Ts``; element s is
    the reference's ``losses["loss/s"]``, trainer.py:618), ``total`` 0-dim = sum(loss) / total_div
    (``losses["loss"]``, trainer.py:621, when total_div is the number of scales -- the default),
    ``terms`` [n_pass,4] (loss, photometric mean, smoothness, 0), ``argmin`` list of uint8 [B,H,W]
    (torch.min index, trainer.py:604), ``depth`` {pass: [B,1,H,W]} and ``warped`` {pass: [S,B,3,H,W]}
    for the requested passes.

    ``target``, every ``sources[f]``, ``K``, ``inv_K`` and every ``smooth_colors[i]`` are either one tensor
    over the whole batch or -- the sequence trainer's layout (trainer_gru.py:890-899,943-957) -- a list of
    equally sized chunks along the batch dimension, which the kernels then read in place (no torch.cat).

    ``frame_weights`` (one [B,S,H,W] tensor per scale, differentiable) is the ``--predictive_mask``
    ablation (trainer.py:571-579): the mask, already resized to H x W, multiplies each frame's
    reprojection loss before the mean / min; like the reference it needs ``disable_automasking``.
    ``kernel="cta"`` runs the first-generation CTA-strip kernel (testing: an independent cross-check).
    ``seed_device``: int64 tensor [1] on the device whose value is xor'ed into ``seed`` by the kernels (a captured
    CUDA graph then draws fresh tie-break noise on every replay)."""
    n_pass, S = len(disps), len(sources)
    if not (1 <= n_pass <= PML_MAX_PASSES):
        raise ValueError("1..%d scales per call" % PML_MAX_PASSES)
    if not (1 <= S <= PML_MAX_SOURCES):
        raise ValueError("1..%d source frames" % PML_MAX_SOURCES)
    if len(Ts) != S or len(smooth_colors) != n_pass or len(smooth_weights) != n_pass:
        raise ValueError("inconsistent argument lengths")
    flags = (PML_FLAG_NO_SSIM if no_ssim else 0) | (PML_FLAG_NO_AUTOMASK if disable_automasking else 0) | \
            (PML_FLAG_AVG_REPROJ if avg_reprojection else 0) | (PML_FLAG_KERNEL_CTA if kernel == "cta" else 0)
    use_noise = noise is not None and not disable_automasking
    if frame_weights is not None:
        if not disable_automasking:
            raise ValueError("frame_weights (predictive mask) are only used with disable_automasking (trainer.py:556,571)")
        if len(frame_weights) != n_pass:
            raise ValueError("one frame-weight tensor per scale")
    # chunked batch?
    tl = _tensor_list(target)
    n_seg = 0
    if tl is not None:
        n_seg = len(tl)
        if not (1 <= n_seg <= _cabi.PML_MAX_SEGMENTS):
            raise ValueError("1..%d chunks per batch" % _cabi.PML_MAX_SEGMENTS)
        if kernel == "cta":
            raise ValueError("the CTA-strip cross-check kernel takes whole-batch tensors only")
        groups = [tl, _tensor_list(K), _tensor_list(inv_K)] + [_tensor_list(x) for x in sources] + \
                 [_tensor_list(x) for x in smooth_colors]
        if any(g is None or len(g) != n_seg for g in groups):
            raise ValueError("target, sources, K, inv_K and smooth_colors must all be lists of %d chunks" % n_seg)
    elif any(_tensor_list(x) is not None for x in [K, inv_K] + list(sources) + list(smooth_colors)):
        raise ValueError("target, sources, K, inv_K and smooth_colors must all be tensors or all be lists of chunks")
    nondiff = [target, K, inv_K] + list(sources) + list(smooth_colors) + (list(noise) if use_noise else [])
    for x in nondiff:
        for t in (x if isinstance(x, (list, tuple)) else (x,)):
            if isinstance(t, torch.Tensor) and t.requires_grad:
                raise NotImplementedError(
                    "gradients with respect to images / intrinsics / noise are not produced by libpml "
                    "(the reference never requests them: colour inputs carry no grad, trainer.py:233-237)")
    images = dict(target=target, K=K, inv_K=inv_K, sources=list(sources), colors=list(smooth_colors),
                  noise=list(noise) if use_noise else None, n_seg=n_seg)
    cfg = dict(n_pass=n_pass, S=S, flags=flags, min_depth=float(min_depth), max_depth=float(max_depth),
               smooth_weights=[float(w) for w in smooth_weights], has_noise=use_noise, seed=int(seed) & (2 ** 64 - 1),
               emit_depth=tuple(emit_depth), emit_warped=tuple(emit_warped), prof_events=prof_events,
               has_fw=frame_weights is not None, images=images,
               total_div=float(total_div) if total_div else None, seed_device=seed_device)
    tensors = list(disps) + list(Ts) + (list(frame_weights) if frame_weights is not None else [])
    outs = _PhotometricLoss.apply(cfg, *tensors)
    am = outs[3]
    res = {"loss": outs[0], "total": outs[1], "terms": outs[2], "argmin": [am[i] for i in range(n_pass)],
           "argmin_all": am, "depth": {}, "warped": {}}
    k = 4
    for i in sorted(set(emit_depth)):
        res["depth"][i] = outs[k]
        k += 1
    for i in sorted(set(emit_warped)):
        res["warped"][i] = outs[k]
        k += 1
    return res


def selection_masks(argmin_all: torch.Tensor, n_id: int) -> torch.Tensor:
    """``outputs["identity_selection/{s}"]`` (trainer.py:606-608) for all scales in one launch:
    uint8 selection indices [n_pass,B,H,W] -> float [n_pass,B,H,W], 1 where a reprojection candidate won."""
    lib = get_library()
    if argmin_all.dtype != torch.uint8 or argmin_all.dim() != 4:
        raise TypeError("argmin must be uint8 [n_pass,B,H,W]")
    if not argmin_all.is_cuda and not lib.emulator:
        raise RuntimeError("argmin is on %s: libpml has no CPU path" % argmin_all.device)
    argmin_all = argmin_all.contiguous()
    n_pass = argmin_all.shape[0]
    n_pix = argmin_all[0].numel()
    out = torch.empty(argmin_all.shape, device=argmin_all.device, dtype=torch.float32)
    a = (ctypes.c_void_p * n_pass)(*[argmin_all.data_ptr() + i * n_pix for i in range(n_pass)])
    o = (ctypes.c_void_p * n_pass)(*[out.data_ptr() + 4 * i * n_pix for i in range(n_pass)])
    lib.check(lib.pml_selection_masks(n_pass, n_pix, a, int(n_id), o, _stream_ptr(argmin_all)), "pml_selection_masks")
    return out


# ------------------------------------------------------------------------------------------------
# layer-level functions (layers.py signatures)
# ------------------------------------------------------------------------------------------------
class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, min_depth, max_depth):
        lib = get_library()
        disp = _check(disp, "disp", lib)
        scaled, depth = torch.empty_like(disp), torch.empty_like(disp)
        lib.check(lib.pml_disp_to_depth_fwd(_ptr(disp), _ptr(scaled), _ptr(depth), disp.numel(),
                                            float(min_depth), float(max_depth), _stream_ptr(disp)), "pml_disp_to_depth_fwd")
        ctx.save_for_backward(disp)
        ctx.rng = (float(min_depth), float(max_depth))
        return scaled, depth

    @staticmethod
    def backward(ctx, g_scaled, g_depth):
        lib = get_library()
        (disp,) = ctx.saved_tensors
        g = torch.empty_like(disp)
        gs = g_scaled.contiguous() if g_scaled is not None else None
        gd = g_depth.contiguous() if g_depth is not None else None
        lib.check(lib.pml_disp_to_depth_bwd(_ptr(disp), _ptr(gs), _ptr(gd), _ptr(g), disp.numel(),
                                            ctx.rng[0], ctx.rng[1], _stream_ptr(disp)), "pml_disp_to_depth_bwd")
        return g, None, None


def depth_from_disp(disp: torch.Tensor, min_depth: float, max_depth: float) -> torch.Tensor:
    """``disp_to_depth(disp)[1]`` (layers.py:16-25) without an autograd node: the ``outputs[("depth", 0, s)]``
    by-product of trainer.py:480, which nothing back-propagates through."""
    lib = get_library()
    disp = _check(disp.detach(), "disp", lib)
    depth = torch.empty_like(disp)
    lib.check(lib.pml_disp_to_depth_fwd(_ptr(disp), None, _ptr(depth), disp.numel(), float(min_depth), float(max_depth),
                                        _stream_ptr(disp)), "pml_disp_to_depth_fwd")
    return depth


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, H, W):
        lib = get_library()
        depth, inv_K = _check(depth, "depth", lib), _check(inv_K, "inv_K", lib)
        B = inv_K.shape[0]
        if depth.numel() != B * H * W:
            raise ValueError("depth has %d elements, expected %d" % (depth.numel(), B * H * W))
        cam = torch.empty((B, 4, H * W), device=depth.device, dtype=torch.float32)
        lib.check(lib.pml_backproject_fwd(_ptr(depth), _ptr(inv_K), _ptr(cam), B, H, W, _stream_ptr(depth)),
                  "pml_backproject_fwd")
        ctx.save_for_backward(inv_K)
        ctx.dims = (B, H, W, depth.shape)
        return cam

    @staticmethod
    def backward(ctx, g_cam):
        lib = get_library()
        (inv_K,) = ctx.saved_tensors
        B, H, W, shape = ctx.dims
        g_cam = g_cam.contiguous()
        g = torch.empty(shape, device=g_cam.device, dtype=torch.float32)
        lib.check(lib.pml_backproject_bwd(_ptr(g_cam), _ptr(inv_K), _ptr(g), B, H, W, _stream_ptr(g_cam)),
                  "pml_backproject_bwd")
        return g, None, None, None


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W, eps):
        lib = get_library()
        points, K, T = _check(points, "points", lib), _check(K, "K", lib), _check(T, "T", lib)
        B = points.shape[0]
        if points.shape != (B, 4, H * W):
            raise ValueError("points must be [B,4,H*W]")
        grid = torch.empty((B, H, W, 2), device=points.device, dtype=torch.float32)
        lib.check(lib.pml_project_fwd(_ptr(points), _ptr(K), _ptr(T), _ptr(grid), B, H, W, float(eps),
                                      _stream_ptr(points)), "pml_project_fwd")
        ctx.save_for_backward(points, K, T)
        ctx.dims = (B, H, W, float(eps))
        return grid

    @staticmethod
    def backward(ctx, g_grid):
        lib = get_library()
        points, K, T = ctx.saved_tensors
        B, H, W, eps = ctx.dims
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("no gradient for the intrinsics K (never trained in the reference)")
        g_grid = g_grid.contiguous()
        g_pts = torch.empty_like(points)
        g_T = torch.empty((B, 4, 4), device=points.device, dtype=torch.float32)
        nb = lib.pml_project_bwd_workspace_bytes(B, H, W)
        ws = torch.empty(nb, device=points.device, dtype=torch.uint8)
        lib.check(lib.pml_project_bwd(_ptr(points), _ptr(K), _ptr(T), _ptr(g_grid), _ptr(g_pts), _ptr(g_T),
                                      _ptr(ws), nb, B, H, W, eps, _stream_ptr(points)), "pml_project_bwd")
        return g_pts, None, g_T, None, None, None


class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        lib = get_library()
        x, y = _check(x, "x", lib), _check(y, "y", lib)
        if x.shape != y.shape or x.dim() != 4:
            raise ValueError("SSIM expects two [B,C,H,W] tensors of equal shape")
        B, C, H, W = x.shape
        out = torch.empty_like(x)
        lib.check(lib.pml_ssim_fwd(_ptr(x), _ptr(y), _ptr(out), B * C, H, W, _stream_ptr(x)), "pml_ssim_fwd")
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = get_library()
        x, y = ctx.saved_tensors
        B, C, H, W = x.shape
        g = g.contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        lib.check(lib.pml_ssim_bwd(_ptr(x), _ptr(y), _ptr(g), _ptr(gx), _ptr(gy), B * C, H, W, _stream_ptr(x)),
                  "pml_ssim_bwd")
        return gx, gy


class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        lib = get_library()
        disp, img = _check(disp, "disp", lib), _check(img, "img", lib)
        B, C, H, W = img.shape
        if disp.shape != (B, 1, H, W):
            raise ValueError("disp must be [B,1,H,W] matching img")
        out = torch.empty((), device=disp.device, dtype=torch.float32)
        nb = lib.pml_smooth_workspace_bytes(B, H, W)
        ws = torch.empty(nb, device=disp.device, dtype=torch.uint8)
        lib.check(lib.pml_smooth_fwd(_ptr(disp), _ptr(img), _ptr(out), _ptr(ws), nb, B, C, H, W, _stream_ptr(disp)),
                  "pml_smooth_fwd")
        ctx.save_for_backward(disp, img)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = get_library()
        disp, img = ctx.saved_tensors
        B, C, H, W = img.shape
        g = g.contiguous().to(torch.float32)
        gd = torch.empty_like(disp) if ctx.needs_input_grad[0] else None
        gi = torch.empty_like(img) if ctx.needs_input_grad[1] else None
        lib.check(lib.pml_smooth_bwd(_ptr(disp), _ptr(img), _ptr(g), _ptr(gd), _ptr(gi), B, C, H, W, _stream_ptr(disp)),
                  "pml_smooth_bwd")
        return gd, gi


class _Pose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        lib = get_library()
        aa = _check(axisangle, "axisangle", lib).reshape(-1, 3).contiguous()
        tr = _check(translation, "translation", lib).reshape(-1, 3).contiguous()
        B = aa.shape[0]
        if tr.shape[0] != B:
            raise ValueError("axisangle / translation batch mismatch")
        T = torch.empty((B, 4, 4), device=aa.device, dtype=torch.float32)
        lib.check(lib.pml_pose_fwd(_ptr(aa), _ptr(tr), _ptr(T), B, 1 if invert else 0, _stream_ptr(aa)), "pml_pose_fwd")
        ctx.save_for_backward(aa, tr)
        ctx.meta = (bool(invert), axisangle.shape, translation.shape)
        return T

    @staticmethod
    def backward(ctx, gT):
        lib = get_library()
        aa, tr = ctx.saved_tensors
        invert, sa, st = ctx.meta
        gT = gT.contiguous()
        ga, gt = torch.empty_like(aa), torch.empty_like(tr)
        lib.check(lib.pml_pose_bwd(_ptr(aa), _ptr(tr), _ptr(gT), _ptr(ga), _ptr(gt), aa.shape[0], 1 if invert else 0,
                                   _stream_ptr(aa)), "pml_pose_bwd")
        return ga.reshape(sa), gt.reshape(st), None


class _Upsample(torch.autograd.Function):
    """F.interpolate(x, [H, W], mode="bilinear", align_corners=False) (trainer.py:474-475, :574-576)."""
    @staticmethod
    def forward(ctx, x, H, W):
        lib = get_library()
        x = _check(x, "x", lib)
        if x.dim() != 4:
            raise ValueError("x must be [B,C,h,w]")
        B, C, h, w = x.shape
        out = torch.empty((B, C, H, W), device=x.device, dtype=torch.float32)
        lib.check(lib.pml_upsample_fwd(_ptr(x), _ptr(out), B * C, h, w, H, W, _stream_ptr(x)), "pml_upsample_fwd")
        ctx.dims = (B, C, h, w, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = get_library()
        B, C, h, w, H, W = ctx.dims
        g = g.contiguous()
        gx = torch.empty((B, C, h, w), device=g.device, dtype=torch.float32)
        lib.check(lib.pml_upsample_bwd(_ptr(g), _ptr(gx), B * C, h, w, H, W, _stream_ptr(g)), "pml_upsample_bwd")
        return gx, None, None


class _BceOnes(torch.autograd.Function):
    """nn.BCELoss()(mask, torch.ones_like(mask)) (trainer.py:582) -> 0-dim tensor."""
    @staticmethod
    def forward(ctx, mask):
        lib = get_library()
        mask = _check(mask, "mask", lib)
        out = torch.empty((), device=mask.device, dtype=torch.float32)
        nb = lib.pml_bce_workspace_bytes()
        ws = torch.empty(nb, device=mask.device, dtype=torch.uint8)
        lib.check(lib.pml_bce_ones_fwd(_ptr(mask), mask.numel(), _ptr(out), _ptr(ws), nb, _stream_ptr(mask)), "pml_bce_ones_fwd")
        ctx.save_for_backward(mask)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = get_library()
        (mask,) = ctx.saved_tensors
        g = g.contiguous().to(torch.float32)
        gm = torch.empty_like(mask)
        lib.check(lib.pml_bce_ones_bwd(_ptr(mask), _ptr(g), _ptr(gm), mask.numel(), _stream_ptr(mask)), "pml_bce_ones_bwd")
        return gm


class _DispHead(torch.autograd.Function):
    """sigmoid(Conv3x3(x)) of the DepthDecoder disparity heads (networks/depth_decoder.py:46-47,62-66;
    layers.py:121-136): reflection pad + 3x3 convolution to one channel + sigmoid, one kernel each way."""
    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = get_library()
        x, weight, bias = _check(x, "x", lib), _check(weight, "weight", lib), _check(bias, "bias", lib)
        if x.dim() != 4 or weight.shape != (1, x.shape[1], 3, 3) or bias.numel() != 1:
            raise ValueError("x [B,C,h,w], weight [1,C,3,3], bias [1] expected")
        B, C, h, w = x.shape
        disp = torch.empty((B, 1, h, w), device=x.device, dtype=torch.float32)
        lib.check(lib.pml_disp_head_fwd(_ptr(x), _ptr(weight), _ptr(bias), _ptr(disp), B, C, h, w, _stream_ptr(x)),
                  "pml_disp_head_fwd")
        ctx.save_for_backward(x, weight, disp)
        return disp

    @staticmethod
    def backward(ctx, g):
        lib = get_library()
        x, weight, disp = ctx.saved_tensors
        B, C, h, w = x.shape
        g = g.contiguous().to(torch.float32)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(weight)
        gb = torch.empty((1,), device=x.device, dtype=torch.float32)
        nb = lib.pml_disp_head_bwd_workspace_bytes(B, C, h, w)
        ws = torch.empty(nb, device=x.device, dtype=torch.uint8)
        lib.check(lib.pml_disp_head_bwd(_ptr(x), _ptr(weight), _ptr(disp), _ptr(g), _ptr(gx), _ptr(gw), _ptr(gb), _ptr(ws), nb,
                                        B, C, h, w, _stream_ptr(x)), "pml_disp_head_bwd")
        return gx, gw, gb


def disp_head(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    return _DispHead.apply(x, weight, bias)


def upsample_bilinear(x: torch.Tensor, height: int, width: int) -> torch.Tensor:
    return _Upsample.apply(x, int(height), int(width))


def bce_against_ones(mask: torch.Tensor) -> torch.Tensor:
    return _BceOnes.apply(mask)


# ------------------------------------------------------------------------------------------------
# input colour pyramid (datasets/mono_dataset.py:84-111)
# ------------------------------------------------------------------------------------------------
def color_pyramid(frames_u8: torch.Tensor, n_scales: int = 4) -> List[torch.Tensor]:
    """uint8 scale-0 frames [N,H,W,3] (PIL / numpy layout, on the device) -> list over scales of
    fp32 [N,3,H>>s,W>>s]: what ``MonoDataset.preprocess`` + ``ToTensor`` produce on the CPU for
    ``("color", f, s)``, bit for bit (Pillow 8-bit Lanczos chain, value / 255)."""
    lib = get_library()
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
        raise TypeError("frames must be uint8 [N,H,W,3]")
    if not frames_u8.is_cuda and not lib.emulator:
        raise RuntimeError("frames are on %s: libpml has no CPU path, move them to a CUDA device" % frames_u8.device)
    frames_u8 = frames_u8.contiguous()
    N, H, W, _ = frames_u8.shape
    outs = [torch.empty((N, 3, H >> s, W >> s), device=frames_u8.device, dtype=torch.float32) for s in range(n_scales)]
    nb = lib.pml_pyramid_workspace_bytes(N, H, W, n_scales)
    ws = torch.empty(nb, device=frames_u8.device, dtype=torch.uint8)
    ptrs = (ctypes.c_void_p * n_scales)(*[o.data_ptr() for o in outs])
    lib.check(lib.pml_pyramid_u8(_ptr(frames_u8), N, H, W, n_scales, ptrs, _ptr(ws), nb, _stream_ptr(frames_u8)),
              "pml_pyramid_u8")
    return outs


# ------------------------------------------------------------------------------------------------
# monitoring metrics (trainer.py:624-652, layers.py:251-269)
# ------------------------------------------------------------------------------------------------
GARG_CROP_375x1242 = (153, 371, 44, 1197)   # trainer.py:641: crop_mask[:, :, 153:371, 44:1197] = 1


def depth_metrics(depth_pred: torch.Tensor, depth_gt: torch.Tensor, crop=GARG_CROP_375x1242,
                  min_depth: float = 1e-3, max_depth: float = 80.0) -> torch.Tensor:
    """``Trainer.compute_depth_losses`` without its Python glue: depth_pred [B,1,H,W] (any
    resolution), depth_gt [B,1,Hg,Wg] (0 = no measurement) -> 7 floats on the device in the order of
    ``depth_metric_names``: abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3.

    Resize + clamp + mask + crop run in one kernel, the two medians are found by radix selection (three
    histogram passes, no sort), median scaling + the seven error sums in a last one.  No host synchronisation."""
    lib = get_library()
    depth_pred = _check(depth_pred.detach(), "depth_pred", lib)
    depth_gt = _check(depth_gt, "depth_gt", lib)
    if depth_pred.dim() != 4 or depth_gt.dim() != 4 or depth_pred.shape[1] != 1 or depth_gt.shape[1] != 1 \
            or depth_pred.shape[0] != depth_gt.shape[0]:
        raise ValueError("depth_pred [B,1,H,W] and depth_gt [B,1,Hg,Wg] expected")
    B, _, H, W = depth_pred.shape
    _, _, Hg, Wg = depth_gt.shape
    dev = depth_pred.device
    n = B * Hg * Wg
    pred = torch.empty(n, device=dev, dtype=torch.float32)
    gt = torch.empty(n, device=dev, dtype=torch.float32)
    count = torch.empty(1, device=dev, dtype=torch.int32)
    cy0, cy1, cx0, cx1 = crop
    st = _stream_ptr(depth_pred)
    lib.check(lib.pml_depth_metrics_prepare(_ptr(depth_pred), _ptr(depth_gt), _ptr(pred), _ptr(gt), _ptr(count),
                                            B, H, W, Hg, Wg, cy0, cy1, cx0, cx1, float(min_depth), float(max_depth), st),
              "pml_depth_metrics_prepare")
    # torch.median == lower middle element == sorted[(n_valid - 1) // 2]; masked entries are +inf: radix selection on the device
    ratio = torch.empty(1, device=dev, dtype=torch.float32)
    out = torch.empty(7, device=dev, dtype=torch.float32)
    nb = lib.pml_depth_metrics_workspace_bytes()
    ws = torch.empty(nb, device=dev, dtype=torch.uint8)
    lib.check(lib.pml_depth_metrics_median_ratio(_ptr(pred), _ptr(gt), n, _ptr(count), _ptr(ratio), _ptr(ws), nb, st),
              "pml_depth_metrics_median_ratio")                                                          # trainer.py:645
    lib.check(lib.pml_depth_metrics_reduce(_ptr(pred), _ptr(gt), n, _ptr(ratio), _ptr(count), float(min_depth),
                                           float(max_depth), _ptr(out), _ptr(ws), nb, st), "pml_depth_metrics_reduce")
    return out
