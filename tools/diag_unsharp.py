"""Developer diagnostic: fused vs separate UnsharpMask point pass on the mma.sync kernels."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import util
import imagemagick_b200 as im
src = util.make_image(333, 217, 4, seed=5, kind="alpha_blocks")
for mma in (1, 0):
    util.set_option("conv_mma", mma)
    for args in ((0.0, 2.0, 0.8, 0.0), (0.0, 4.0, 0.8, 0.0), (0.0, 2.0, 1.5, 0.0), (0.0, 2.0, 0.8, 0.02)):
        util.set_option("no_fused_unsharp", 0)
        f = im.UnsharpMaskImage(im.Image(torch.from_numpy(src).cuda()), *args).pixels.cpu().numpy()
        util.set_option("no_fused_unsharp", 1)
        u = im.UnsharpMaskImage(im.Image(torch.from_numpy(src).cuda()), *args).pixels.cpu().numpy()
        want = np.empty_like(src)
        util.oracle().orc_unsharp(util.P(src), util.P(want), 333, 217, 4, *[float(a) for a in args])
        d = util.ulp_distance(f, u)
        idx = np.argwhere(d > 0)
        print("mma", mma, args, "mismatches", len(idx), "max ulp", int(d.max()), "fused vs oracle", util.max_ulp(f, want),
              "unfused vs oracle", util.max_ulp(u, want), "nan", int(np.isnan(f).sum()), int(np.isnan(u).sum()))
        for (y, x, c) in idx[:6]:
            print("   ", y, x, c, "src", src[y, x, c], "fused", f[y, x, c], "unfused", u[y, x, c], "want", want[y, x, c])
