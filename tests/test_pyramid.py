"""Input colour pyramid (datasets/mono_dataset.py:84-111): the numpy oracle is pinned
against fixtures produced by the reference's own MonoDataset.preprocess (real Pillow + torchvision)
and against the installed Pillow; the kernels (emulator here, CUDA under -m gpu) must match bit for bit."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import pyramid_oracle as pyo
from ssde_b200 import functional as Fn, trainer_hooks

FIX = os.path.join(common.GOLDEN_DIR, "aux", "pyramid.npz")


def _cases():
    z = np.load(FIX)
    for name in ("small", "odd_tiles", "tiny"):
        frames = z[name + "|frames"]
        ref = [[z["%s|ref|%d|%d" % (name, n, s)] for s in range(4)] for n in range(frames.shape[0])]
        yield name, frames, ref


def test_oracle_matches_reference_fixture():
    for name, frames, ref in _cases():
        levels_f, _ = pyo.pyramid(frames, 4)
        for n in range(frames.shape[0]):
            for s in range(4):
                assert np.array_equal(levels_f[s][n], ref[n][s]), (name, n, s)   # bit-exact fp32


def test_oracle_matches_installed_pillow_general_ratios():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for (h, w, oh, ow) in [(375, 1242, 192, 640), (30, 20, 45, 31), (17, 33, 17, 16), (64, 14, 32, 7)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.LANCZOS))
        assert np.array_equal(pyo.resize_lanczos(img, oh, ow), ref), (h, w, oh, ow)


def check_kernels(device):
    """Shared by the emulator test and the GPU test: kernel output == reference fixture, bit for bit."""
    for name, frames, ref in _cases():
        got = Fn.color_pyramid(torch.from_numpy(frames).to(device), 4)
        for s in range(4):
            g = got[s].cpu().numpy()
            for n in range(frames.shape[0]):
                assert np.array_equal(g[n], ref[n][s]), (name, n, s, np.abs(g[n] - ref[n][s]).max())
    # fewer scales, one scale, and the drop-in ingest (three frames through one launch chain)
    name, frames, ref = next(_cases())
    got = Fn.color_pyramid(torch.from_numpy(frames).to(device), 2)
    assert len(got) == 2 and np.array_equal(got[1].cpu().numpy()[0], ref[0][1])
    got = Fn.color_pyramid(torch.from_numpy(frames).to(device), 1)
    assert np.array_equal(got[0].cpu().numpy()[1], ref[1][0])
    inputs = {("color_u8", f): torch.from_numpy(frames[i:i + 1]) for i, f in enumerate((0, -1, 1))}
    trainer_hooks.ingest_colors(inputs, (0, -1, 1), 4, device=device)
    for i, f in enumerate((0, -1, 1)):
        for s in range(4):
            assert np.array_equal(inputs[("color", f, s)].cpu().numpy()[0], ref[i][s])
    # a size the 2x chain cannot serve is refused, not approximated
    with pytest.raises(Exception):
        Fn.color_pyramid(torch.zeros(1, 36, 64, 3, dtype=torch.uint8, device=device), 4)


def test_emulated_kernels_bit_exact(emu_lib):
    check_kernels("cpu")


def test_full_size_matches_oracle_emulated(emu_lib):
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (1, 192, 640, 3), dtype=np.uint8)
    want, _ = pyo.pyramid(frames, 4)
    got = Fn.color_pyramid(torch.from_numpy(frames), 4)
    for s in range(4):
        assert np.array_equal(got[s].numpy(), want[s]), s


@pytest.mark.gpu
def test_cuda_kernels_bit_exact(cuda_lib):
    check_kernels("cuda")
    rng = np.random.default_rng(6)
    frames = rng.integers(0, 256, (4, 192, 640, 3), dtype=np.uint8)
    want, _ = pyo.pyramid(frames, 4)
    got = Fn.color_pyramid(torch.from_numpy(frames).cuda(), 4)
    for s in range(4):
        assert np.array_equal(got[s].cpu().numpy(), want[s]), s
