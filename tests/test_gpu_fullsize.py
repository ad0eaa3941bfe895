"""BASELINE.json's configurations at their FULL sizes on the GPU.

The oracle cannot process 8192^2 / 16384^2 images in test time, so the full-size runs are checked
 (a) window by window: an output window of a neighbourhood operator depends only on the input window
     plus the operator's halo, so the oracle is run on crops (corners -- edge replication --, the centre,
     and windows that straddle the kernels' strip / chunk boundaries) and compared with the same window of
     the full-size GPU result at the operator's own parity bar;
 (b) through size-independent properties: a constant image stays constant, point operators are checked on
     sampled rows, two independent GPU implementations (streaming vs gather resize) agree, and the result of
     the device-resident path equals the host-buffer path.
"""
import numpy as np
import pytest

import util
from util import P, max_ulp, oracle

pytestmark = pytest.mark.gpu

im = pytest.importorskip("imagemagick_b200")
torch = pytest.importorskip("torch")


def big_image(w, h, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.rand((h, w, 4), device="cuda", generator=g) * 65535.0
    t[: h // 3, : w // 5, 3] = 0.0                     # a fully transparent region (PerceptibleReciprocal clamp)
    t[h // 2:, w // 2:, 3] = 65535.0                   # an opaque one
    return t


def windows(w, h, size, extra=()):
    """Corners, centre and caller-chosen offsets; clipped to the image."""
    pts = [(0, 0), (w - size, 0), (0, h - size), (w - size, h - size), ((w - size) // 2, (h - size) // 2)]
    pts += list(extra)
    return [(max(0, min(x, w - size)), max(0, min(y, h - size))) for x, y in pts]


def crop_with_halo(t, x0, y0, size, halo):
    """Host crop [y0-halo, y0+size+halo) x [x0-halo, ...) clipped to the image + the offset of the window in it."""
    h, w = t.shape[0], t.shape[1]
    xa, ya = max(0, x0 - halo), max(0, y0 - halo)
    xb, yb = min(w, x0 + size + halo), min(h, y0 + size + halo)
    return np.ascontiguousarray(t[ya:yb, xa:xb].cpu().numpy()), x0 - xa, y0 - ya


def test_config2_blur_8192_sigma4_windows_and_constant():
    W = H = 8192
    src = big_image(W, H, 2)
    out = im.BlurImage(im.Image(src), 0.0, 4.0).pixels
    size, halo = 320, 16
    # strips of the column pass are 529 rows, of the row pass 529 columns; chunks of the row loader 8 pixels
    extra = [(529 - 100, 529 - 100), (3 * 529 - 160, 5 * 529 - 160), (4000, 15 * 529 - 200)]
    for x0, y0 in windows(W, H, size, extra):
        crop, ox, oy = crop_with_halo(src, x0, y0, size, halo)
        want = np.empty_like(crop)
        assert oracle().orc_blur(P(crop), P(want), crop.shape[1], crop.shape[0], 4, 0.0, 4.0) == 0
        got = out[y0:y0 + size, x0:x0 + size].cpu().numpy()
        d = util.ulp_distance(got, want[oy:oy + size, ox:ox + size])
        assert d.max() <= 1, (x0, y0, int(d.max()))
        assert (d == 0).mean() > 0.9999, (x0, y0, float((d == 0).mean()))
    del out
    const = torch.full((2048, W, 4), 12345.678, device="cuda")
    const[..., 3] = 65535.0
    res = im.BlurImage(im.Image(const), 0.0, 4.0).pixels
    assert max_ulp(res.cpu().numpy(), const.cpu().numpy()) <= 1


def test_config3_resize_16384_to_8192_lanczos(monkeypatch):
    W = H = 16384
    src = big_image(W, H, 3)
    out = im.ResizeImage(im.Image(src), W // 2, H // 2, im.LanczosFilter).pixels
    # (a) top-left window against the oracle: the crop starts at the origin, so the contribution lists
    #     (bisect, start, weights) are the ones the full-size image uses
    n = 384
    crop = np.ascontiguousarray(src[: 2 * n + 16, : 2 * n + 16].cpu().numpy())
    want = np.empty((n + 8, n + 8, 4), np.float32)
    assert oracle().orc_resize(P(crop), 2 * n + 16, 2 * n + 16, 4, P(want), n + 8, n + 8, 22) == 0
    got = out[:n, :n].cpu().numpy()
    assert max_ulp(got, want[:n, :n]) <= 1
    # (b) the streaming kernels against the independent gather kernels on the whole image
    util.set_option("no_resize_stream", 1)
    ref = im.ResizeImage(im.Image(src), W // 2, H // 2, im.LanczosFilter).pixels
    util.set_option("no_resize_stream", 0)
    a = out.view(torch.int32).to(torch.int64)
    b = ref.view(torch.int32).to(torch.int64)
    a = torch.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = torch.where(b < 0, -(b & 0x7FFFFFFF), b)
    d = (a - b).abs()
    print("config3 stream vs gather: max ULP", int(d.max()), "fraction identical", float((d == 0).double().mean()))
    assert int(d.max()) <= 1
    assert float((d == 0).double().mean()) > 0.99999
    del ref, a, b, d
    # (c) a constant opaque image stays constant (the weights sum to 1 within rounding)
    const = torch.full((4096, W, 4), 40000.25, device="cuda")
    const[..., 3] = 65535.0
    res = im.ResizeImage(im.Image(const), W // 2, 2048, im.LanczosFilter).pixels
    exp = np.empty((2048, W // 2, 4), np.float32)
    exp[..., :3] = 40000.25
    exp[..., 3] = 65535.0
    assert max_ulp(res.cpu().numpy(), exp) <= 1


def test_config4_lab_then_dilate_8192():
    W = H = 8192
    src = big_image(W, H, 4)
    src[100, :64, :3] = 0.0                                  # black pixels: L must be exactly 0
    img = im.Image(src.clone())
    im.TransformImageColorspace(img, im.LabColorspace)
    lab = img.pixels
    for y in (0, 100, 4097, H - 1):                          # point operator: sampled rows
        row = np.ascontiguousarray(src[y:y + 1].cpu().numpy())
        want = row.copy()
        assert oracle().orc_colorspace(P(want), W, 1, 4, 23, 11) == 0
        assert max_ulp(lab[y:y + 1].cpu().numpy(), want) <= 1, y
    out = im.MorphologyImage(im.Image(lab), im.DilateMorphology, 1, "Disk:3").pixels
    k = util.orc_kernel("disk", 3, 1, 0, 0)
    size, halo = 256, 3
    for x0, y0 in windows(W, H, size, [(26 * 30 - 100, 1000), (5000, 2048 - 128)]):
        crop, ox, oy = crop_with_halo(lab, x0, y0, size, halo)
        want = util.orc_morphology(crop, 4, 1, [k])
        got = out[y0:y0 + size, x0:x0 + size].cpu().numpy()
        assert max_ulp(got, want[oy:oy + size, ox:ox + size]) == 0, (x0, y0)   # bit exact


def test_config5_pipeline_4096_device_equals_host_path():
    W = H = 4096
    src = big_image(W, H, 5)
    dev = im.ResizeImage(im.BlurImage(im.Image(src), 0.0, 4.0), W // 2, H // 2, im.LanczosFilter).pixels.cpu().numpy()
    host_src = src.cpu().numpy()
    host = im.ResizeImage(im.BlurImage(im.Image(host_src), 0.0, 4.0), W // 2, H // 2, im.LanczosFilter).pixels
    assert np.array_equal(dev.view(np.int32), host.view(np.int32))       # same kernels, same bits
    n = 256                                                              # top-left window of the pipeline vs the oracle
    crop = np.ascontiguousarray(host_src[: 2 * n + 64, : 2 * n + 64])
    b = np.empty_like(crop)
    assert oracle().orc_blur(P(crop), P(b), crop.shape[1], crop.shape[0], 4, 0.0, 4.0) == 0
    b = np.ascontiguousarray(b[: 2 * n + 16, : 2 * n + 16])
    want = np.empty((n + 8, n + 8, 4), np.float32)
    assert oracle().orc_resize(P(b), 2 * n + 16, 2 * n + 16, 4, P(want), n + 8, n + 8, 22) == 0
    # blur within 1 ULP, then a resize of a 1-ULP-different input: compare with a matching tolerance
    got = dev[:n, :n]
    err = np.abs(got.astype(np.float64) - want[:n, :n].astype(np.float64))
    assert err.max() <= 3 * 0.00390625, err.max()                        # 3 float ULPs at the top of the Quantum range
