"""Developer diagnostic: where do the XYZ-derived colourspaces differ from the oracle by more than 1 ULP?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import util
import imagemagick_b200 as im
rng = np.random.default_rng(1)
for kind in ("noise", "hdr"):
    src = util.make_image(131, 67, 4, seed=7, kind=kind)
    for cs in (16, 17, 25, 35, 36, 37, 40, 26, 11):
        for frm, to in ((23, cs), (cs, 23)):
            want = src.copy()
            assert util.oracle().orc_colorspace(util.P(want), 131, 67, 4, frm, to) == 0
            img = im.Image(torch.from_numpy(src.copy()).cuda()); img.colorspace = frm
            im.TransformImageColorspace(img, to)
            got = img.pixels.cpu().numpy()
            ok = np.isfinite(want) & np.isfinite(got)
            d = util.ulp_distance(np.where(ok, got, np.float32(0)), np.where(ok, want, np.float32(0)))
            bad = d > 1
            msg = f"{kind} {frm}->{to}: >1ULP {int(bad.sum())} of {bad.size}, nonfinite mismatch {int((np.isfinite(want) != np.isfinite(got)).sum())}"
            if bad.any():
                msg += f", max|want| among them {np.abs(want[bad]).max():.3e}, max abs diff {np.abs(want[bad].astype(np.float64) - got[bad]).max():.3e}, max ulp {int(d.max())}"
                idx = np.argwhere(bad)[:3]
                for (y, x, c) in idx:
                    msg += f"\n      px {src[y, x, :3]} ch{c}: want {want[y, x, c]!r} got {got[y, x, c]!r}"
            print(msg, flush=True)
