"""The CPU oracle against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import pytest
import torch

import common
from oracle import photometric_oracle as po
from ssde_b200 import synthetic


def _oracle_for(name, dtype):
    variant, opt, inputs, outputs, r32, r64, seed = common.load_golden(name)
    n_id = 0 if opt.disable_automasking else (1 if opt.avg_reprojection else 2)
    v1 = opt.v1_multiscale and variant != "fusion"
    B = outputs[("disp", opt.scales[0])].shape[0]
    noise = synthetic.draw_noise(B, opt.height, opt.width, opt.scales, max(n_id, 1), seed=seed, v1_multiscale=v1)
    inp = synthetic.to_sequence_layout(inputs, opt.len_sequence) if variant == "gru" else inputs
    got = po.run(opt, inp, outputs, variant=variant, noise=noise if n_id else None, dtype=dtype)
    return got, (r32 if dtype == torch.float32 else r64), opt


@pytest.mark.parametrize("name", common.golden_names())
def test_oracle_float64_matches_reference(name):
    got, ref, opt = _oracle_for(name, torch.float64)
    for k, v in ref.items():
        if k.startswith("argmin/") or k.startswith("identity_selection/"):
            assert torch.equal(got[k].long(), v.long()), k
        else:
            assert common.rel_err(got[k], v) < 1e-8, (k, common.rel_err(got[k], v))  # float64 round-off only


@pytest.mark.parametrize("name", common.golden_names())
def test_oracle_float32_matches_reference(name):
    got, ref, opt = _oracle_for(name, torch.float32)
    for k, v in ref.items():
        if k.startswith("argmin/") or k.startswith("identity_selection/"):
            assert torch.equal(got[k].long(), v.long()), k
        elif k.startswith("loss"):
            assert common.rel_err(got[k], v) < 1e-6, (k, common.rel_err(got[k], v))
        else:
            # same ATen ops in a different association (e.g. the division by (W-1, H-1) is one
            # broadcast op here, two in-place ops in layers.py:190-191): fp32 round-off only
            assert common.rel_err(got[k], v) < 2e-5, (k, common.rel_err(got[k], v))


def test_golden_covers_every_variant_and_flag():
    names = set(common.golden_names())
    for need in ("trainer_default", "trainer_avg", "trainer_noautomask", "trainer_nossim", "trainer_v1multiscale",
                 "fusion_default", "fusion_v3_default", "gru_seq3", "trainer_static", "trainer_constant", "trainer_oof",
                 "trainer_predmask", "trainer_predmask_avg", "trainer_posecnn"):
        assert need in names


def test_unit_functions_closed_forms():
    """Closed forms the kernels restate: upsample coordinates, grid_sample
    unnormalisation, SSIM of identical inputs, smoothness of a constant disparity."""
    g = torch.Generator().manual_seed(0)
    d = torch.rand(1, 1, 4, 6, generator=g, dtype=torch.float64)
    up = po.upsample_disp(d, 8, 12)
    # pixel (x=5, y=3): src = (5.5/2 - .5, 3.5/2 - .5) = (2.25, 1.25)
    want = (0.75 * (0.75 * d[0, 0, 1, 2] + 0.25 * d[0, 0, 1, 3]) + 0.25 * (0.75 * d[0, 0, 2, 2] + 0.25 * d[0, 0, 2, 3]))
    assert abs(up[0, 0, 3, 5] - want) < 1e-15
    x = torch.rand(1, 3, 8, 8, generator=g, dtype=torch.float64)
    assert po.ssim(x, x).abs().max() == 0
    assert po.smooth_loss(torch.ones(1, 1, 8, 8), torch.rand(1, 3, 8, 8, generator=g)).item() == 0
    s, dep = po.disp_to_depth(torch.tensor([0.0, 1.0]), 0.1, 100.0)
    assert torch.allclose(dep, torch.tensor([100.0, 0.1]))


def test_depth_metrics_oracle_matches_reference():
    """oracle.compute_depth_losses restates trainer.py:624-652 / layers.py:251-269; the fixture holds
    the outputs of the unmodified reference (tests/golden/make_golden_depth.py)."""
    import os
    import numpy as np
    from oracle import photometric_oracle as po
    z = np.load(os.path.join(common.GOLDEN_DIR, "aux", "depth_metrics.npz"))
    pred, gt = torch.from_numpy(z["pred"]), torch.from_numpy(z["gt"])
    got64 = po.compute_depth_losses(pred, gt, torch.float64)
    got32 = po.compute_depth_losses(pred, gt, torch.float32)
    assert torch.allclose(got64, torch.from_numpy(z["ref_f64"]), rtol=1e-10, atol=0)
    assert torch.allclose(got32.double(), torch.from_numpy(z["ref_f32"]), rtol=1e-6, atol=0)
