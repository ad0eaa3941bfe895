"""Generates tests/golden/r02b_golden.npz from the REAL reference (oracle/_ref/libmagickref.so): the operators added
late in round 2 -- ResizeImage with the Jinc and Kaiser filters, the hue / saturation colourspaces (HCL, HCLp, HSB, HSI,
HSL, HSV, HWB) and the XYZ-derived ones (LMS, CAT02LMS, Luv, xyY, DisplayP3, Adobe98, ProPhoto) in both directions.  Run in the authoring container only:   python tests/golden/make_golden_r02b.py
tests/test_golden.py pins the oracle (CPU) and the CUDA path (-m gpu) to these arrays."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import util  # noqa: E402

OUT = Path(__file__).resolve().parent / "r02b_golden.npz"
W, H = 41, 31
HEXCONE = {"hcl": 4, "hclp": 5, "hsb": 6, "hsi": 7, "hsl": 8, "hsv": 9, "hwb": 10,
           "lch": 12, "lchab": 13, "lchuv": 14, "oklab": 38, "oklch": 39, "lms": 16, "luv": 17, "xyy": 25, "displayp3": 35, "adobe98": 36, "prophoto": 37, "cat02lms": 40}


def source(ch):
    src = util.make_image(W, H, ch, seed=300 + ch, kind="alpha_blocks" if ch == 4 else "noise")
    special = np.array([[0, 0, 0], [65535, 65535, 65535], [32768, 32768, 32768], [65535, 0, 0], [0, 65535, 0],
                        [0, 0, 65535], [65535, 65535, 0], [40000, 40000, 100], [100, 40000, 40000], [10922.5, 0, 0],
                        [43690, 65535, 65535], [0, 65535, 32767.5], [65535, 32768, 16384], [70000, -20, 300]], np.float32)
    src[0, :len(special), :3] = special
    return src


def main():
    r, P = util.ref(), util.P
    out = {}
    for ch in (3, 4):
        src = source(ch)
        out[f"c{ch}/src"] = src
        for filt, name in ((13, "jinc"), (16, "kaiser")):
            for (ow, oh) in ((20, 15), (82, 62), (41, 13)):
                dst = np.empty((oh, ow, ch), np.float32)
                assert r.ref_resize(P(src), W, H, ch, P(dst), ow, oh, filt) == 0
                out[f"c{ch}/resize_{name}_{ow}x{oh}"] = dst
        for name, cs in HEXCONE.items():
            for frm, to, tag in ((23, cs, f"srgb_{name}"), (cs, 23, f"{name}_srgb")):
                buf = src.copy()
                assert r.ref_colorspace(P(buf), W, H, ch, frm, to) == 0
                out[f"c{ch}/colorspace_{tag}"] = buf
    np.savez_compressed(OUT, **out)
    print(OUT, len(out), "arrays", OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
