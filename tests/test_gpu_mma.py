"""Parity of the FP64 mma.sync 1-D convolution kernels (conv_mma.cu) against the oracle and against the DFMA
streaming kernels (conv1d.cu), with the switch forced on -- and off -- whatever the process default is.

Bar: <= 1 ULP of the float Quantum against the reference arithmetic (BASELINE.json north_star), >= 99.9 % bit-identical
(the matrix path accumulates in FP64 like every other path: it differs only in association)."""
import numpy as np
import pytest

import util
from util import P, make_image, max_ulp, oracle

pytestmark = pytest.mark.gpu

im = pytest.importorskip("imagemagick_b200")


def _dev(a):
    import torch
    return im.Image(torch.from_numpy(a).cuda())


def _host(img):
    return img.pixels.cpu().numpy() if img.on_device else img.pixels


def orc(fn, src, *args):
    h, w, ch = src.shape
    dst = np.empty_like(src)
    assert getattr(oracle(), fn)(P(src), P(dst), w, h, ch, *args) == 0
    return dst


def with_mma(fn, on=1):
    util.set_option("conv_mma", on)
    n0 = util.get_option("conv_mma_launches")
    out = fn()
    return out, util.get_option("conv_mma_launches") - n0


SIZES = [(67, 45), (256, 131), (1, 1), (5, 300), (300, 3), (8, 8), (9, 33), (640, 130), (31, 1100)]


@pytest.mark.parametrize("kind", ["noise", "alpha_blocks", "hdr"])
@pytest.mark.parametrize("radius,sigma", [(0.0, 0.5), (0.0, 1.0), (0.0, 2.0), (0.0, 3.0), (0.0, 4.0), (3.0, 1.5), (0.0, 2.4),
                                          (0.0, 3.6), (13.0, 3.0)])
def test_blur_mma_path_against_the_oracle(kind, radius, sigma):
    for (w, h) in SIZES:
        src = make_image(w, h, 4, seed=w * 5 + h, kind=kind)
        want = orc("orc_blur", src, radius, sigma)
        got, n = with_mma(lambda: _host(im.BlurImage(_dev(src), radius, sigma)))
        assert n == 2, "both passes must run on the mma.sync kernels"
        assert max_ulp(got, want) <= 1, (w, h, kind, radius, sigma, max_ulp(got, want))


def test_blur_mma_1024_mostly_bit_identical_and_agrees_with_the_dfma_kernels():
    src = make_image(1024, 1024, 4, seed=42)
    for sigma in (2.0, 4.0):
        want = orc("orc_blur", src, 0.0, sigma)
        got, n = with_mma(lambda: _host(im.BlurImage(_dev(src), 0.0, sigma)))
        assert n == 2
        d = util.ulp_distance(got, want)
        assert d.max() <= 1 and (d == 0).mean() > 0.9999
        ref, n = with_mma(lambda: _host(im.BlurImage(_dev(src), 0.0, sigma)), on=0)
        assert n == 0
        d = util.ulp_distance(got, ref)
        assert d.max() <= 1 and (d == 0).mean() > 0.9999


def test_mma_strip_seams_and_image_edges():
    """A tall and a wide image: several strips per pass, ragged last blocks, edge replication on all four sides."""
    for (w, h) in ((40, 1500), (1500, 40)):
        src = make_image(w, h, 4, seed=9, kind="alpha_blocks")
        want = orc("orc_blur", src, 0.0, 4.0)
        got, n = with_mma(lambda: _host(im.BlurImage(_dev(src), 0.0, 4.0)))
        assert n == 2 and max_ulp(got, want) <= 1


@pytest.mark.parametrize("radius,sigma", [(0.0, 2.4), (0.0, 3.6), (13.0, 3.0), (2.0, 1.0), (0.0, 4.0), (0.0, 2.0)])
def test_mma_non_finite_samples_stay_local(radius, sigma):
    """Every Toeplitz tile multiplies samples outside an output's window by zero taps: inf / NaN pixels must poison
    exactly the outputs they poison in the reference (the flagged blocks take the scalar path)."""
    src = make_image(150, 110, 4, seed=21)
    src[30, 40, 0] = np.inf
    src[31, 90, 3] = -np.inf
    src[80, 20, 1] = np.nan
    src[100, 140, 0] = np.inf
    src[0, 0, 2] = np.inf
    src[109, 149, 3] = np.nan
    want = orc("orc_blur", src, radius, sigma)
    got, n = with_mma(lambda: _host(im.BlurImage(_dev(src), radius, sigma)))
    assert n == 2
    assert np.isfinite(want).mean() > 0.3
    assert np.array_equal(np.isnan(got), np.isnan(want))
    inf = np.isinf(want)
    assert np.array_equal(np.isinf(got), inf) and np.array_equal(got[inf], want[inf])
    ok = np.isfinite(want)
    d = util.ulp_distance(np.where(ok, got, np.float32(0)), np.where(ok, want, np.float32(0)))
    assert d.max() <= 1


@pytest.mark.parametrize("sigma", [1.0, 2.0, 4.0, 2.6])
def test_mma_reciprocal_clamp_around_the_threshold(sigma):
    rng = np.random.default_rng(11)
    h, w = 96, 160
    src = (rng.random((h, w, 4), dtype=np.float32) * np.float32(65535)).astype(np.float32)
    alpha = np.zeros((h, w), np.float32)
    ys, xs = np.mgrid[4:h:9, 4:w:11]
    alpha[ys, xs] = ((6.5535e-8 / 0.2) * np.float32(2.0) ** rng.integers(-6, 7, size=ys.shape)).astype(np.float32)
    alpha[:8, :8] = 65535.0
    alpha[40:44, 100:104] = np.float32(1e-30)
    src[..., 3] = alpha
    want = orc("orc_blur", src, 0.0, sigma)
    got, n = with_mma(lambda: _host(im.BlurImage(_dev(src), 0.0, sigma)))
    assert n == 2 and np.isfinite(want).all()
    assert max_ulp(got, want) <= 1


@pytest.mark.parametrize("kind", ["noise", "alpha_blocks", "hdr"])
@pytest.mark.parametrize("sigma", [0.5, 1.0, 2.0, 4.0])
def test_mma_rank1_gaussian_keeps_the_double_intermediate(kind, sigma):
    for (w, h) in ((97, 61), (300, 5), (3, 200), (640, 130)):
        src = make_image(w, h, 4, seed=w + 3, kind=kind)
        want = orc("orc_gaussian_blur", src, 0.0, sigma)
        got, n = with_mma(lambda: _host(im.GaussianBlurImage(_dev(src), 0.0, sigma)))
        assert n == 2
        assert max_ulp(got, want) <= 1, (w, h, kind, sigma)


@pytest.mark.parametrize("args", [(0.0, 4.0, 1.5, 0.02), (0.0, 2.0, 0.8, 0.0), (0.0, 1.0, 2.0, 0.3)])
def test_mma_unsharp_epilogue(args):
    src = make_image(333, 217, 4, seed=5, kind="alpha_blocks")
    want = orc("orc_unsharp", src, *args)
    fused, n = with_mma(lambda: _host(im.UnsharpMaskImage(_dev(src), *args)))
    assert n == 2
    assert max_ulp(fused, want) <= 1
    util.set_option("no_fused_unsharp", 1)
    unfused, _ = with_mma(lambda: _host(im.UnsharpMaskImage(_dev(src), *args)))
    assert np.array_equal(fused, unfused)
    # the pass-through branch returns the input itself
    blur = orc("orc_blur", src, args[0], args[1])
    passthrough = np.abs(2.0 * (src.astype(np.float64) - blur.astype(np.float64))) < 65535.0 * args[3]
    assert np.array_equal(fused[passthrough], src[passthrough])
