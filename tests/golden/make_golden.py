"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py [case ...]  # needs /root/reference; no argument = every case

For every case this script builds a small synthetic batch (package ``synthetic`` module), runs
``/root/reference/trainer*.py``'s ``Trainer.generate_images_pred`` + ``Trainer.compute_losses``
unbound (oracle/reference_runner.py) in float32 and in float64, and stores inputs + reference
outputs.  The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and are the ground truth
of the GPU parity tests (tests/test_gpu_parity.py).  Nothing at test time reads /root/reference.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import ssde_b200  # noqa: E402
from ssde_b200 import synthetic  # noqa: E402
from oracle import reference_runner  # noqa: E402

# name -> (variant, opt overrides, batch kwargs)
CASES = {
    "trainer_default":      ("trainer", {}, dict(style="kitti", seed=11)),
    "trainer_uniform":      ("trainer", {}, dict(style="uniform", seed=12)),
    "trainer_static":       ("trainer", {}, dict(style="static", seed=13)),
    "trainer_constant":     ("trainer", {}, dict(style="constant", seed=14)),
    "trainer_oof":          ("trainer", {}, dict(style="oof", seed=15)),
    "trainer_avg":          ("trainer", dict(avg_reprojection=True), dict(style="kitti", seed=16)),
    "trainer_noautomask":   ("trainer", dict(disable_automasking=True), dict(style="kitti", seed=17)),
    "trainer_nossim":       ("trainer", dict(no_ssim=True), dict(style="kitti", seed=18)),
    "trainer_v1multiscale": ("trainer", dict(v1_multiscale=True), dict(style="kitti", seed=19)),
    "trainer_scales02":     ("trainer", dict(scales=[0, 2]), dict(style="kitti", seed=20)),
    "fusion_default":       ("fusion", {}, dict(style="kitti", seed=21, full_res_disp=True)),
    "fusion_v3_default":    ("fusion_v3", {}, dict(style="kitti", seed=22)),
    "gru_seq3":             ("gru", dict(len_sequence=3, batch_size=1), dict(style="kitti", seed=23, batch=3)),
    "trainer_predmask":     ("trainer", dict(disable_automasking=True, predictive_mask=True), dict(style="kitti", seed=25, predictive_mask=True)),
    "trainer_predmask_avg": ("trainer", dict(disable_automasking=True, predictive_mask=True, avg_reprojection=True),
                             dict(style="kitti", seed=26, predictive_mask=True)),
    # posecnn builds T inside generate_images_pred (trainer.py:490-499) through layers.get_translation_matrix, whose
    # torch.zeros takes the DEFAULT dtype: the float64 run needs the default switched, which would change the
    # torch.randn stream of the automask noise -- hence this case runs without automasking (no noise drawn)
    "trainer_posecnn":      ("trainer", dict(pose_model_type="posecnn", disable_automasking=True), dict(style="kitti", seed=27)),
    "trainer_wide":         ("trainer", {}, dict(style="kitti", seed=24, batch=1, height=64, width=160)),
    # the same translation scaling exists in trainer_fusion.py:446-456 (full-resolution disparities, no resize)
    "fusion_posecnn":       ("fusion", dict(pose_model_type="posecnn", disable_automasking=True),
                             dict(style="kitti", seed=28, full_res_disp=True)),
}
NOISE_SEED = 1234


def build_case(name):
    variant, overrides, bkw = CASES[name]
    bkw = dict(bkw)
    B = bkw.pop("batch", 2)
    H = bkw.pop("height", 32)
    W = bkw.pop("width", 64)
    kw = dict(batch_size=B)
    kw.update(overrides)
    opt = synthetic.make_options(H, W, **kw)
    inputs, outputs = synthetic.make_batch(B, H, W, scales=opt.scales, **bkw)
    return variant, opt, inputs, outputs


def main():
    assert reference_runner.available(), "needs the reference tree"
    for name in (sys.argv[1:] or CASES):
        variant, opt, inputs, outputs = build_case(name)
        ref_in = synthetic.to_sequence_layout(inputs, opt.len_sequence) if variant == "gru" else inputs
        blob = {}
        meta = {"variant": variant, "opt": {k: v for k, v in vars(opt).items()},
                "noise_seed": NOISE_SEED, "torch": torch.__version__,
                "input_keys": [], "output_keys": []}
        for k, v in inputs.items():
            key = "in|" + json.dumps(k)
            blob[key] = v.numpy()
            meta["input_keys"].append(key)
        for k, v in outputs.items():
            key = "out|" + json.dumps(k)
            blob[key] = v.numpy()
            meta["output_keys"].append(key)
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            res = reference_runner.run(opt, ref_in, outputs, variant, noise_seed=NOISE_SEED, dtype=dt)
            for k, v in res.items():
                if k.startswith("color/") or k.startswith("depth/"):
                    continue
                a = v.numpy()
                if k.startswith("argmin/"):
                    a = a.astype(np.uint8)
                if k.startswith("identity_selection/"):
                    a = a.astype(np.uint8)
                blob["ref_%s|%s" % (tag, k)] = a
        blob["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **blob)
        print("%-22s %7.1f KB  loss32=%.8f loss64=%.12f" % (
            name, os.path.getsize(path) / 1024, float(blob["ref_f32|loss"]), float(blob["ref_f64|loss"])))


if __name__ == "__main__":
    main()
