"""Developer micro-benchmark: device-resident timings of the hot-path operators (CUDA events).
usage: python tools/devbench.py [blur|resize|lab|dilate|gauss|conv2d|stencils|all] [size]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import imagemagick_b200 as im

which = sys.argv[1] if len(sys.argv) > 1 else "all"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
PEAK = 6576.1


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, npix, bytes_per_px):
    gbs = npix * bytes_per_px / ms / 1e6
    print(f"{name:34s} {ms:9.3f} ms  {npix / ms / 1e3:10.1f} Mpix/s  {gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f}% of measured HBM peak", flush=True)


torch.manual_seed(0)
if which in ("blur", "all"):
    x = im.Image(torch.rand(size, size, 4, device="cuda") * 65535)
    for sigma in (4.0, 2.0):
        ms = timeit(lambda: im.BlurImage(x, 0.0, sigma))
        report(f"BlurImage {size}^2 RGBA sigma={sigma}", ms, size * size, 64)
    k = im.AcquireKernelInfo("blur:0x4")
    ms = timeit(lambda: im.ConvolveImage(x, k)); report("  row pass only (33 taps)", ms, size * size, 32)
    k = im.AcquireKernelInfo("blur:0x4+90")
    ms = timeit(lambda: im.ConvolveImage(x, k)); report("  column pass only (33 taps)", ms, size * size, 32)
    del x
if which == "taps":          # DFMA streaming kernels against the FP64 mma.sync kernels, per window length
    from imagemagick_b200 import _lib
    x = im.Image(torch.rand(size, size, 4, device="cuda") * 65535)
    for mma in (0, 1):
        _lib.check(_lib.load().mb200_set_option(b"conv_mma", mma))
        for sigma in (1.0, 2.0, 3.0, 4.0):
            ms = timeit(lambda: im.BlurImage(x, 0.0, sigma))
            report(f"mma={mma} BlurImage sigma={sigma}", ms, size * size, 64)
        ms = timeit(lambda: im.GaussianBlurImage(x, 0.0, 4.0)); report(f"mma={mma} GaussianBlurImage(0,4) rank-1", ms, size * size, 32)
        ms = timeit(lambda: im.GaussianBlurImage(x, 0.0, 2.0)); report(f"mma={mma} GaussianBlurImage(0,2) rank-1", ms, size * size, 32)
        ms = timeit(lambda: im.UnsharpMaskImage(x, 0.0, 4.0, 1.5, 0.02)); report(f"mma={mma} UnsharpMaskImage(0,4)", ms, size * size, 80)
        ms = timeit(lambda: im.UnsharpMaskImage(x, 0.0, 2.0, 1.5, 0.02)); report(f"mma={mma} UnsharpMaskImage(0,2)", ms, size * size, 80)
    del x
if which in ("resize", "all"):
    s2 = size * 2 if size <= 8192 else size
    x = im.Image(torch.rand(s2, s2, 4, device="cuda") * 65535)
    ms = timeit(lambda: im.ResizeImage(x, s2 // 2, s2 // 2, im.LanczosFilter))
    report(f"ResizeImage {s2}^2->{s2 // 2}^2 Lanczos", ms, s2 * s2, 36)
    del x
if which in ("lab", "all"):
    x = im.Image(torch.rand(size, size, 4, device="cuda") * 65535)
    def f():
        x.colorspace = im.sRGBColorspace
        im.TransformImageColorspace(x, im.LabColorspace)
    ms = timeit(f); report(f"sRGB->Lab {size}^2", ms, size * size, 32)
    del x
if which in ("dilate", "all"):
    x = im.Image(torch.rand(size, size, 4, device="cuda") * 65535)
    k = im.AcquireKernelInfo("Disk:3")
    ms = timeit(lambda: im.MorphologyImage(x, im.DilateMorphology, 1, k)); report(f"Dilate Disk:3 {size}^2", ms, size * size, 32)
    del x
if which in ("gauss", "all"):
    s = min(size, 4096)
    x = im.Image(torch.rand(s, s, 4, device="cuda") * 65535)
    ms = timeit(lambda: im.GaussianBlurImage(x, 0.0, 4.0), iters=3, warm=1); report(f"GaussianBlurImage 2-D 29x29 {s}^2", ms, s * s, 32)

if which in ("conv2d", "all"):
    hd = im.Image(torch.rand(1080, 1920, 4, device="cuda") * 65535)
    ms = timeit(lambda: im.SharpenImage(hd, 5.0, 2.0), iters=9); report("SharpenImage 5x2 (11x11) 1920x1080", ms, 1920 * 1080, 32)
    x = im.Image(torch.rand(size, size, 4, device="cuda") * 65535)
    ms = timeit(lambda: im.SharpenImage(x, 5.0, 2.0)); report(f"SharpenImage 5x2 (11x11) {size}^2", ms, size * size, 32)
    ms = timeit(lambda: im.EdgeImage(x, 1.0)); report(f"EdgeImage(1) (3x3) {size}^2", ms, size * size, 32)
    ms = timeit(lambda: im.EmbossImage(x, 0.0, 1.0), iters=3); report(f"EmbossImage(0,1) {size}^2 (conv + equalize)", ms, size * size, 64)
    ms = timeit(lambda: im.ConvolveImage(x, "LoG:0x2")); report(f"ConvolveImage LoG:0x2 {size}^2", ms, size * size, 32)
    del x, hd
if which in ("stencils", "all"):
    s = min(size, 4096)
    x = im.Image(torch.rand(s, s, 4, device="cuda") * 65535)
    ms = timeit(lambda: im.StatisticImage(x, im.MedianStatistic, 3, 3), iters=3); report(f"StatisticImage Median 3x3 {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.StatisticImage(x, im.MeanStatistic, 5, 5), iters=3); report(f"StatisticImage Mean 5x5 {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.BilateralBlurImage(x, 7, 7, 20.0, 2.0), iters=3); report(f"BilateralBlurImage 7x7 {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.SelectiveBlurImage(x, 0.0, 1.5, 6553.5), iters=3); report(f"SelectiveBlurImage 0x1.5 t=10% {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.AdaptiveBlurImage(x, 0.0, 1.5), iters=3); report(f"AdaptiveBlurImage 0x1.5 {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.RotationalBlurImage(x, 5.0), iters=3); report(f"RotationalBlurImage(5) {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.MotionBlurImage(x, 0.0, 4.0, 30.0), iters=3); report(f"MotionBlurImage(0,4,30) {s}^2", ms, s * s, 32)
    ms = timeit(lambda: im.EqualizeImage(x), iters=3); report(f"EqualizeImage {s}^2", ms, s * s, 32)
