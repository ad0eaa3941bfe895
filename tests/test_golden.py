"""Pins the oracle (CPU restatement) to golden vectors produced by the REAL reference
(tests/golden/make_golden.py ran ImageMagick 7.1.1-45 compiled from source).  Bit-exact."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import util
from util import P, oracle

G = np.load(Path(__file__).resolve().parent / "golden" / "hotpath_golden.npz")
TAGS = sorted({k.split("/")[0] for k in G.files if k.startswith("c")})
KERNEL_ARGS = {"Disk:3": ("disk", 3, 1, 0, 0), "Diamond:2": ("diamond", 2, 1, 0, 0),
               "Rectangle:5x3+1+2": ("rectangle", 5, 3, 1, 2)}
METHODS = {"erode": 3, "dilate": 4, "open": 8, "close": 9, "smooth": 12, "edgein": 13, "edgeout": 14, "edge": 15,
           "tophat": 16, "bottomhat": 17}
FILTERS = {"lanczos": 22, "mitchell": 12, "undefined": 0, "triangle": 3, "point": 1}
SPACES = {"srgb": 23, "lab": 11, "xyz": 26, "rgb": 21, "ohta": 18, "rec601ycbcr": 19, "rec709ycbcr": 20, "yiq": 30,
          "yuv": 32, "cmy": 1, "ycbcr": 27}


def golden_cases(tag):
    return [k for k in G.files if k.startswith(tag + "/") and not k.endswith("/src") and not k.endswith("/clamp_src")]


def run_oracle(src, name, clamp_src=None):
    h, w, ch = src.shape
    dst = np.empty_like(src)
    o = oracle()
    if name.startswith("blur_"):
        r, s = name[5:].split("x")
        assert o.orc_blur(P(src), P(dst), w, h, ch, float(r), float(s)) == 0
    elif name.startswith("gaussian_"):
        r, s = name[9:].split("x")
        assert o.orc_gaussian_blur(P(src), P(dst), w, h, ch, float(r), float(s)) == 0
    elif name.startswith("unsharp_"):
        rs, gain, thr = name[8:].split("_")
        r, s = rs.split("x")
        assert o.orc_unsharp(P(src), P(dst), w, h, ch, float(r), float(s), float(gain), float(thr)) == 0
    elif name.startswith("sharpen_"):
        r, s = name[8:].split("x")
        assert o.orc_sharpen(P(src), P(dst), w, h, ch, float(r), float(s)) == 0
    elif name.startswith("edge_") and name[5:].replace(".", "").isdigit():
        assert o.orc_edge(P(src), P(dst), w, h, ch, float(name[5:])) == 0
    elif name.split("_")[0] in METHODS:
        m, kname = name.split("_", 1)
        k = util.orc_kernel(*KERNEL_ARGS[kname])
        dst = util.orc_morphology(src, METHODS[m], 1, [k])
    elif name == "convolve_user3x3":
        vals = np.array([[1.0, 2.0, 0.5], [0.0, -1.0, np.nan], [3.0, 0.25, -2.0]])
        dst = util.orc_morphology(src, 1, 1, [util.orc_kernel_from_array(vals, 1, 1)])
    elif name.startswith("resize_"):
        _, f, size = name.split("_")
        ow, oh = map(int, size.split("x"))
        dst = np.empty((oh, ow, ch), np.float32)
        assert o.orc_resize(P(src), w, h, ch, P(dst), ow, oh, FILTERS[f]) == 0
    elif name.startswith("sample_"):
        ow, oh = map(int, name[7:].split("x"))
        dst = np.empty((oh, ow, ch), np.float32)
        assert o.orc_sample(P(src), w, h, ch, P(dst), ow, oh) == 0
    elif name.startswith("threshold_"):
        _, op, *rest = name.split("_")
        if op == "bilevel":
            dst = util.orc_threshold(src, 0, [float(rest[0])])
        elif op == "clamp":
            dst = util.orc_threshold(clamp_src, 3, [0.0])
        else:
            dst = util.orc_threshold(src, 1 if op == "black" else 2, util.parse_thresholds(rest[0]))
    elif name.startswith("colorspace_"):
        _, a, b = name.split("_")
        dst = src.copy()
        assert o.orc_colorspace(P(dst), w, h, ch, SPACES[a], SPACES[b]) == 0
    else:
        raise AssertionError(name)
    return dst


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_golden_bit_exact(tag):
    src = np.ascontiguousarray(G[tag + "/src"])
    cases = golden_cases(tag)
    assert len(cases) > 20
    for key in cases:
        got = run_oracle(src, key.split("/", 1)[1], np.ascontiguousarray(G[tag + "/clamp_src"]))
        want = G[key]
        assert got.shape == want.shape, key
        assert util.max_ulp(got, want) == 0, key


def test_validate_c_known_answer_rgb_to_lab():
    """tests/validate.c:244-259: sRGB (0.545877,0.966567,0.463759) -> L*a*b* 88.456154,-54.671483,51.662818
    (the reference's own tolerance is 1% of QuantumRange, validate.c:67)."""
    buf = np.ascontiguousarray(G["kat/srgb"]).copy()
    assert oracle().orc_colorspace(P(buf), 1, 1, 3, 23, 11) == 0
    assert util.max_ulp(buf, G["kat/lab"]) == 0
    L = buf[0, 0, 0] / 65535.0 * 100.0
    a = (buf[0, 0, 1] / 65535.0 - 0.5) * 255.0
    b = (buf[0, 0, 2] / 65535.0 - 0.5) * 255.0
    assert abs(L - 88.456154) < 5e-3 and abs(a + 54.671483) < 5e-3 and abs(b - 51.662818) < 5e-3


def test_probed_kernel_sizes():
    """SURVEY appendix B probes: -blur 0x2 -> 17 taps, 0x4 -> 33; gaussian 0x4 -> 29x29; Disk:3 -> 29 cells."""
    o = oracle()
    assert o.orc_optimal_kernel_width_1d(0.0, 2.0) == 17
    assert o.orc_optimal_kernel_width_1d(0.0, 4.0) == 33
    assert o.orc_optimal_kernel_width_2d(0.0, 4.0) == 29
    k = util.orc_kernel("disk", 3, 1, 0, 0).array()
    assert k.shape == (7, 7) and int(np.sum(~np.isnan(k))) == 29


# ---- the CUDA path against the same golden vectors (outputs of the real reference) ---------------------------
def run_cuda(im, src, name, clamp_src):
    import torch
    dev = lambda a: im.Image(torch.from_numpy(np.ascontiguousarray(a)).cuda())
    host = lambda img: img.pixels.cpu().numpy()
    if name.startswith("blur_"):
        r, s = name[5:].split("x")
        return host(im.BlurImage(dev(src), float(r), float(s))), 1
    if name.startswith("gaussian_"):
        r, s = name[9:].split("x")
        return host(im.GaussianBlurImage(dev(src), float(r), float(s))), 1
    if name.startswith("unsharp_"):
        rs, gain, thr = name[8:].split("_")
        r, s = rs.split("x")
        return host(im.UnsharpMaskImage(dev(src), float(r), float(s), float(gain), float(thr))), 1
    if name.startswith("sharpen_"):
        r, s = name[8:].split("x")
        return host(im.SharpenImage(dev(src), float(r), float(s))), 1
    if name.startswith("edge_") and name[5:].replace(".", "").isdigit():
        return host(im.EdgeImage(dev(src), float(name[5:]))), 1
    if name.split("_")[0] in METHODS:
        m, kname = name.split("_", 1)
        return host(im.MorphologyImage(dev(src), METHODS[m], 1, kname)), 0
    if name == "convolve_user3x3":
        return host(im.ConvolveImage(dev(src), "3x3: 1,2,0.5 0,-1,nan 3,0.25,-2")), 1
    if name.startswith("resize_"):
        _, f, size = name.split("_")
        ow, oh = map(int, size.split("x"))
        return host(im.ResizeImage(dev(src), ow, oh, FILTERS[f])), 1
    if name.startswith("sample_"):
        ow, oh = map(int, name[7:].split("x"))
        return host(im.SampleImage(dev(src), ow, oh)), 0
    if name.startswith("threshold_"):
        _, op, *rest = name.split("_")
        img = dev(clamp_src if op == "clamp" else src)
        if op == "bilevel":
            im.BilevelImage(img, float(rest[0]))
        elif op == "clamp":
            im.ClampImage(img)
        elif op == "black":
            im.BlackThresholdImage(img, rest[0])
        else:
            im.WhiteThresholdImage(img, rest[0])
        return host(img), 0
    if name.startswith("colorspace_"):
        _, a, b = name.split("_")
        img = dev(src)
        img.colorspace = SPACES[a]
        im.TransformImageColorspace(img, SPACES[b])
        return host(img), 1
    raise AssertionError(name)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_cuda_path_matches_reference_golden(tag):
    """The product (through the C-ABI) against the REAL reference's outputs: bit exact for erode/dilate, the
    difference methods and the threshold operators, <= 1 ULP for convolution / resize / colourspace
    (UnsharpMask included: its point pass runs on the float-rounded blur like the reference's)."""
    im = pytest.importorskip("imagemagick_b200")
    src = np.ascontiguousarray(G[tag + "/src"])
    clamp_src = np.ascontiguousarray(G[tag + "/clamp_src"])
    for key in golden_cases(tag):
        got, bar = run_cuda(im, src, key.split("/", 1)[1], clamp_src)
        want = G[key]
        assert got.shape == want.shape, key
        assert util.max_ulp(got, want) <= bar, (key, util.max_ulp(got, want))


# ---- second golden file: operators added late in round 2 (tests/golden/make_golden_r02b.py) ----------------------
G2 = np.load(Path(__file__).resolve().parent / "golden" / "r02b_golden.npz")
FILTERS2 = {"jinc": 13, "kaiser": 16}
HEXCONE2 = {"hcl": 4, "hclp": 5, "hsb": 6, "hsi": 7, "hsl": 8, "hsv": 9, "hwb": 10, "srgb": 23,
            "lch": 12, "lchab": 13, "lchuv": 14, "oklab": 38, "oklch": 39, "jzazbz": 34, "lms": 16, "luv": 17, "xyy": 25, "displayp3": 35, "adobe98": 36, "prophoto": 37, "cat02lms": 40}
SMOOTH2 = ("hsi", "jzazbz", "lch", "lchab", "lchuv", "oklab", "oklch", "lms", "luv", "xyy", "displayp3", "adobe98", "prophoto", "cat02lms")    # <= 1 ULP; the others bit exact


STATISTICS2 = {"mode": 6, "nonpeak": 7}


def _poster(src):
    return (np.round(src / 16384.0) * 16384.0).astype(np.float32)


def _r02b_cases(tag):
    return [k for k in G2.files if k.startswith(tag + "/") and not k.endswith("/src")]


def test_kernel_strings_expand_like_the_reference():
    """AcquireKernelInfo (morphology.c:485): every named kernel of KernelInfoType, the '@' / '>' / '<' expansions (eighth
    and quarter turns, mirror images) of named and user kernels, the distance-kernel scale flags -- bit-identical values
    and origins for every list element; strings the reference rejects are rejected."""
    im = pytest.importorskip("imagemagick_b200")
    strings = [str(x) for x in G2["kernel/strings"]]
    assert len(strings) > 90
    for n, name in enumerate(strings):
        count = int(G2[f"kernel/{n}/count"][0])
        try:
            mine = im.AcquireKernelInfo(name).arrays()
        except im.MagickB200Error:
            mine = []
        assert len(mine) == count, (name, len(mine), count)
        for idx, (vals, x, y) in enumerate(mine):
            want = G2[f"kernel/{n}/{idx}/values"]
            assert vals.shape == want.shape, (name, idx)
            assert np.array_equal(np.isnan(vals), np.isnan(want)), (name, idx)
            assert np.array_equal(np.nan_to_num(vals), np.nan_to_num(want)), (name, idx)
            assert [x, y] == list(G2[f"kernel/{n}/{idx}/origin"]), (name, idx)


@pytest.mark.parametrize("tag", ["c3", "c4"])
def test_oracle_matches_reference_golden_r02b(tag):
    """Jinc / Kaiser ResizeImage and the hue / saturation colourspaces: the oracle against arrays the real reference
    produced (bit exact)."""
    o = oracle()
    src = np.ascontiguousarray(G2[tag + "/src"])
    h, w, ch = src.shape
    for key in _r02b_cases(tag):
        name = key.split("/", 1)[1]
        want = G2[key]
        if name.startswith("resize_"):
            _, f, size = name.split("_")
            ow, oh = map(int, size.split("x"))
            got = np.empty((oh, ow, ch), np.float32)
            assert o.orc_resize(P(src), w, h, ch, P(got), ow, oh, FILTERS2[f]) == 0
        elif name.startswith("statistic_"):
            _, stat, size, which = name.split("_")
            ww, wh = map(int, size.split("x"))
            image = src if which == "src" else _poster(src)
            got = np.empty_like(image)
            assert o.orc_statistic(P(image), P(got), w, h, ch, STATISTICS2[stat], ww, wh) == 0
        else:
            _, a, b = name.split("_")
            got = src.copy()
            assert o.orc_colorspace(P(got), w, h, ch, HEXCONE2[a], HEXCONE2[b]) == 0
        assert util.max_ulp(got, want) == 0, key


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c3", "c4"])
def test_cuda_path_matches_reference_golden_r02b(tag):
    im = pytest.importorskip("imagemagick_b200")
    import torch
    src = np.ascontiguousarray(G2[tag + "/src"])
    for key in _r02b_cases(tag):
        name = key.split("/", 1)[1]
        want = G2[key]
        img = im.Image(torch.from_numpy(src.copy()).cuda())
        if name.startswith("resize_"):
            _, f, size = name.split("_")
            ow, oh = map(int, size.split("x"))
            got, bar = im.ResizeImage(img, ow, oh, FILTERS2[f]).pixels.cpu().numpy(), 1
        elif name.startswith("statistic_"):
            _, stat, size, which = name.split("_")
            ww, wh = map(int, size.split("x"))
            if which == "poster":
                img = im.Image(torch.from_numpy(_poster(src)).cuda())
            got, bar = im.StatisticImage(img, STATISTICS2[stat], ww, wh).pixels.cpu().numpy(), 0
        else:
            _, a, b = name.split("_")
            img.colorspace = HEXCONE2[a]
            im.TransformImageColorspace(img, HEXCONE2[b])
            got, bar = img.pixels.cpu().numpy(), (1 if (a in SMOOTH2 or b in SMOOTH2) else 0)
            if b in ("lch", "lchab", "lchuv"):            # achromatic pixels: the reference's hue is rounding noise
                got, want = got.copy(), want.copy()
                grey = np.abs(want[..., 1].astype(np.float64) - 32767.5) < 1.0e-4
                got[..., 2][grey] = 0
                want[..., 2][grey] = 0
        d = util.ulp_or_noise(got, want) if bar else util.ulp_distance(got, want)
        assert int(d.max()) <= bar, (key, int(d.max()))


def _defines_to_dict(text):
    return dict(kv.split("=", 1) for kv in text.split(";"))


@pytest.mark.gpu
def test_cuda_resize_with_defines_matches_reference_golden():
    """ResizeImage under -define filter:* : the product (python mirror parsing the strings like the shim) against arrays
    the real reference produced with the same artifacts set."""
    im = pytest.importorskip("imagemagick_b200")
    import torch
    src = np.ascontiguousarray(G2["c4/src"])
    for n, entry in enumerate(str(x) for x in G2["resizedef/defines"]):
        filt, defines = entry.split("|", 1)
        for size in ("20x15", "82x62"):
            want = G2[f"resizedef/{n}/{size}"]
            ow, oh = map(int, size.split("x"))
            got = im.ResizeImage(im.Image(torch.from_numpy(src.copy()).cuda()), ow, oh, int(filt),
                                 artifacts=_defines_to_dict(defines)).pixels.cpu().numpy()
            assert util.max_ulp(got, want) <= 1, (entry, size, util.max_ulp(got, want))


def _settings_of(defines):
    """-> (dict for the python mirror, ColorspaceOptions for the oracle)"""
    d = _defines_to_dict(defines) if defines else {}
    values = {}
    if "color:illuminant" in d:
        values["illuminant"] = d["color:illuminant"]
    for key, field in (("white-luminance", "white_luminance"), ("film-gamma", "film_gamma"), ("reference-black", "reference_black"),
                       ("reference-white", "reference_white")):
        if key in d:
            values[field] = float(d[key])
    return d, util.ColorspaceOptions.of(**values)


def _colordef_cases():
    return [tuple(str(x).split("|", 2)) for x in G2["colordef/cases"]]


def test_oracle_colorspace_settings_match_reference_golden():
    """Illuminants, Jzazbz white luminance, the Log film settings, Log and YCC: the oracle against arrays the real
    reference produced with the artifacts / properties set (bit exact)."""
    import ctypes as C
    src = np.ascontiguousarray(G2["c4/src"])
    h, w, ch = src.shape
    for n, (frm, to, defines) in enumerate(_colordef_cases()):
        got = src.copy()
        _, opts = _settings_of(defines)
        assert oracle().orc_colorspace_ex(P(got), w, h, ch, int(frm), int(to), C.byref(opts)) == 0
        want = G2[f"colordef/{n}"]
        assert np.array_equal(np.isnan(got), np.isnan(want)), (frm, to, defines)
        assert util.max_ulp(np.nan_to_num(got), np.nan_to_num(want)) == 0, (frm, to, defines)


@pytest.mark.gpu
def test_cuda_colorspace_settings_match_reference_golden():
    """... and the CUDA path (python mirror parsing the strings like the shim): YCC and the forward Log gather bit exact,
    everything else <= 1 ULP."""
    im = pytest.importorskip("imagemagick_b200")
    import torch
    src = np.ascontiguousarray(G2["c4/src"])
    for n, (frm, to, defines) in enumerate(_colordef_cases()):
        frm, to = int(frm), int(to)
        settings, _ = _settings_of(defines)
        img = im.Image(torch.from_numpy(src.copy()).cuda())
        img.colorspace = frm
        im.TransformImageColorspace(img, to, settings=settings)
        got, want = img.pixels.cpu().numpy().copy(), G2[f"colordef/{n}"].copy()
        exact = (frm, to) in ((23, 28), (28, 23), (23, 15))
        if to in (12, 13, 14):
            grey = np.abs(want[..., 1].astype(np.float64) - 32767.5) < 1.0e-4
            got[..., 2][grey] = 0
            want[..., 2][grey] = 0
        ok = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ok), (frm, to, defines)
        got, want = np.where(ok, got, np.float32(0)), np.where(ok, want, np.float32(0))
        d = util.ulp_distance(got, want) if exact else util.ulp_or_noise(got, want)
        if frm == 21 and to == 28:
            continue          # the linear -> sRGB leg is a <= 1 ULP operator in front of a quantiser: pinned in test_gpu_parity
        assert int(d.max()) <= (0 if exact else 1), (frm, to, defines, int(d.max()))
