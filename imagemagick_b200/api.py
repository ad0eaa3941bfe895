"""Host-side mirror of the MagickCore entry points of the hot path.

Same names, argument meaning and error behaviour as the reference operators
(MagickCore/effect.c, morphology.c, resize.c, colorspace.c), on top of the C-ABI in
include/magick_b200.h:

    BlurImage, GaussianBlurImage, ConvolveImage, UnsharpMaskImage   effect.c:765/1709/1170/4256
    SharpenImage, EdgeImage, EmbossImage, MotionBlurImage           effect.c:3991/1520/1600/2347
    EqualizeImage                                                   enhance.c:2040
    BilateralBlurImage, RotationalBlurImage                         effect.c:821/3129
    StatisticImage                                                  statistic.c:2918
    MorphologyImage, AcquireKernelInfo                              morphology.c:4129/485
    ResizeImage, SampleImage, ScaleImage, ThumbnailImage (pixel)    resize.c:3761/3907/4106/4591
    TransformImageColorspace                                        colorspace.c:1751
    BilevelImage, BlackThresholdImage, WhiteThresholdImage, ClampImage  threshold.c:805/927/2518/1087

An `Image` wraps the pixel cache: an (rows, columns, channels) float32 array of raw
Quantum values (0..65535), either a NumPy array (host; every call stages through
HBM and back -- the end-to-end path) or a CUDA torch tensor (device-resident; the
kernels run on torch's current stream and the result stays in HBM).

Operators that return `Image *` in the reference return a NEW Image here and never
modify their input; TransformImageColorspace works in place and returns True, like
the reference.  Failures raise MagickB200Error (the reference returns NULL / MagickFalse
and fills an ExceptionInfo); MB200_EUNSUPPORTED is the "decline" signal on which the
MagickCore shim falls back to the stock CPU path.  Nothing here computes pixels on
the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np

from . import _lib
from ._lib import KernelPtr, MagickB200Error, check

# MagickCore/morphology.h:69-99
UndefinedMorphology, ConvolveMorphology, CorrelateMorphology, ErodeMorphology, DilateMorphology = 0, 1, 2, 3, 4
OpenMorphology, CloseMorphology, SmoothMorphology = 8, 9, 12
EdgeInMorphology, EdgeOutMorphology, EdgeMorphology, TopHatMorphology, BottomHatMorphology = 13, 14, 15, 16, 17
ErodeIntensityMorphology, DilateIntensityMorphology, IterativeDistanceMorphology = 5, 6, 7
OpenIntensityMorphology, CloseIntensityMorphology = 10, 11
HitAndMissMorphology, ThinningMorphology, ThickenMorphology = 18, 19, 20

# MagickCore/resample.h:32-69
(UndefinedFilter, PointFilter, BoxFilter, TriangleFilter, HermiteFilter, HannFilter, HammingFilter,
 BlackmanFilter, GaussianFilter, QuadraticFilter, CubicFilter, CatromFilter, MitchellFilter, JincFilter,
 SincFilter, SincFastFilter, KaiserFilter, WelchFilter, ParzenFilter, BohmanFilter, BartlettFilter,
 LagrangeFilter, LanczosFilter, LanczosSharpFilter, Lanczos2Filter, Lanczos2SharpFilter, RobidouxFilter,
 RobidouxSharpFilter, CosineFilter, SplineFilter, LanczosRadiusFilter, CubicSplineFilter,
 MagicKernelSharp2013Filter, MagicKernelSharp2021Filter, SentinelFilter) = range(35)

# MagickCore/colorspace.h:27-67
LabColorspace, RGBColorspace, sRGBColorspace, XYZColorspace = 11, 21, 23, 26
CMYColorspace, OHTAColorspace, Rec601YCbCrColorspace, Rec709YCbCrColorspace = 1, 18, 19, 20
YCbCrColorspace, YDbDrColorspace, YIQColorspace, YPbPrColorspace, YUVColorspace = 27, 29, 30, 31, 32
LCHColorspace, LCHabColorspace, LCHuvColorspace, OklabColorspace, OklchColorspace, JzazbzColorspace = 12, 13, 14, 38, 39, 34
LogColorspace, YCCColorspace = 15, 28
LMSColorspace, LuvColorspace, xyYColorspace, DisplayP3Colorspace, Adobe98Colorspace, ProPhotoColorspace, CAT02LMSColorspace = 16, 17, 25, 35, 36, 37, 40
HCLColorspace, HCLpColorspace, HSBColorspace, HSIColorspace, HSLColorspace, HSVColorspace, HWBColorspace = 4, 5, 6, 7, 8, 9, 10

# kernel types of include/magick_b200.h
(UserDefinedKernel, BlurKernel, GaussianKernel, DiskKernel, SquareKernel, DiamondKernel, OctagonKernel,
 PlusKernel, CrossKernel, RectangleKernel, UnityKernel, DoGKernel, LoGKernel, BinomialKernel) = range(14)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class Image:
    """The pixel cache of one image: rows x columns x channels float32 Quantum."""

    def __init__(self, pixels, colorspace: int = sRGBColorspace):
        if _is_torch(pixels):
            import torch
            if not pixels.is_cuda:
                raise ValueError("torch pixel caches must live on a CUDA device (use NumPy for host buffers)")
            if pixels.dtype != torch.float32 or pixels.dim() != 3:
                raise ValueError("pixels must be float32 of shape (rows, columns, channels)")
            pixels = pixels.contiguous()
        else:
            pixels = np.ascontiguousarray(pixels, dtype=np.float32)
            if pixels.ndim != 3:
                raise ValueError("pixels must have shape (rows, columns, channels)")
        if not 1 <= pixels.shape[2] <= 4:
            raise ValueError("1..4 channels (Gray, Gray+Alpha, RGB, RGBA)")
        self.pixels = pixels
        self.colorspace = colorspace

    rows = property(lambda self: int(self.pixels.shape[0]))
    columns = property(lambda self: int(self.pixels.shape[1]))
    channels = property(lambda self: int(self.pixels.shape[2]))
    on_device = property(lambda self: _is_torch(self.pixels))

    def _ptr(self) -> int:
        return self.pixels.data_ptr() if self.on_device else self.pixels.ctypes.data

    def _new_like(self, rows: Optional[int] = None, columns: Optional[int] = None) -> "Image":
        shape = (rows or self.rows, columns or self.columns, self.channels)
        if self.on_device:
            import torch
            out = torch.empty(shape, dtype=torch.float32, device=self.pixels.device)
        else:
            out = np.empty(shape, dtype=np.float32)
        return Image(out, self.colorspace)


def _stream(image: Image):
    import torch
    handle = torch.cuda.current_stream(image.pixels.device).cuda_stream
    # torch's default stream is the legacy NULL stream; NULL means "library stream" in the
    # C-ABI, so name the legacy stream explicitly (cudaStreamLegacy == 0x1).
    return C.c_void_p(handle if handle else 1)


def _activate(image: Image) -> None:
    if image.on_device:
        import torch
        dev = image.pixels.device.index
        if dev is None:
            dev = torch.cuda.current_device()
        check(_lib.load().mb200_set_device(dev))


class KernelInfo:
    """Owning handle of a KernelInfo list (MagickCore/morphology.h:102-130)."""

    def __init__(self, ptr: KernelPtr):
        if not ptr:
            raise MagickB200Error(_lib.EINVAL, "kernel could not be parsed / built")
        self._ptr = ptr

    def __del__(self):
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr:
            try:
                _lib.load().mb200_destroy_kernel_info(ptr)
            except Exception:
                pass

    def __iter__(self):
        p = self._ptr
        while p:
            yield p.contents
            p = p.contents.next

    def arrays(self):
        """[(values(h,w) float64, x, y), ...] for every kernel of the list."""
        out = []
        for k in self:
            n = k.width * k.height
            vals = np.ctypeslib.as_array(k.values, shape=(n,)).copy().reshape(k.height, k.width)
            out.append((vals, int(k.x), int(k.y)))
        return out


def AcquireKernelInfo(kernel_string: str) -> KernelInfo:
    """MagickCore/morphology.c:485."""
    return KernelInfo(_lib.load().mb200_acquire_kernel_info(kernel_string.encode()))


def AcquireKernelBuiltIn(kernel_type: int, rho: float = 0.0, sigma: float = 0.0, xi: float = 0.0,
                         psi: float = 0.0) -> KernelInfo:
    """MagickCore/morphology.c:950 (GeometryInfo rho, sigma, xi, psi)."""
    return KernelInfo(_lib.load().mb200_acquire_kernel_builtin(kernel_type, rho, sigma, xi, psi))


def _as_kernel(kernel: Union[str, KernelInfo]) -> KernelInfo:
    return AcquireKernelInfo(kernel) if isinstance(kernel, str) else kernel


def _same_size_op(image: Image, dev_fn: str, host_fn: str, *args) -> Image:
    lib = _lib.load()
    out = image._new_like()
    if image.on_device:
        _activate(image)
        check(getattr(lib, dev_fn)(image._ptr(), out._ptr(), image.columns, image.rows, image.channels, *args,
                                   _stream(image)))
    else:
        check(getattr(lib, host_fn)(image._ptr(), out._ptr(), image.columns, image.rows, image.channels, *args))
    return out


def BlurImage(image: Image, radius: float, sigma: float) -> Image:
    """MagickCore/effect.c:765 -- separable Gaussian ("blur:RxS;blur:RxS+90")."""
    return _same_size_op(image, "mb200_blur_image_dev", "mb200_blur_image", float(radius), float(sigma))


def GaussianBlurImage(image: Image, radius: float, sigma: float) -> Image:
    """MagickCore/effect.c:1709 -- true 2-D "gaussian:RxS" kernel."""
    return _same_size_op(image, "mb200_gaussian_blur_image_dev", "mb200_gaussian_blur_image", float(radius),
                         float(sigma))


def ConvolveImage(image: Image, kernel_info: Union[str, KernelInfo]) -> Image:
    """MagickCore/effect.c:1170."""
    k = _as_kernel(kernel_info)
    return _same_size_op(image, "mb200_convolve_image_dev", "mb200_convolve_image", k._ptr)


def MorphologyImage(image: Image, method: int, iterations: int, kernel: Union[str, KernelInfo],
                    bias: float = 0.0) -> Image:
    """MagickCore/morphology.c:4129 (bias == the "convolve:bias" artifact)."""
    k = _as_kernel(kernel)
    return _same_size_op(image, "mb200_morphology_image_dev", "mb200_morphology_image", int(method),
                         int(iterations), k._ptr, float(bias))


def UnsharpMaskImage(image: Image, radius: float, sigma: float, gain: float, threshold: float) -> Image:
    """MagickCore/effect.c:4256."""
    return _same_size_op(image, "mb200_unsharp_mask_image_dev", "mb200_unsharp_mask_image", float(radius),
                         float(sigma), float(gain), float(threshold))


def SharpenImage(image: Image, radius: float, sigma: float) -> Image:
    """MagickCore/effect.c:3991 -- ConvolveImage with the inline sharpening kernel."""
    return _same_size_op(image, "mb200_sharpen_image_dev", "mb200_sharpen_image", float(radius), float(sigma))


def EdgeImage(image: Image, radius: float) -> Image:
    """MagickCore/effect.c:1520 -- ConvolveImage with the all -1 / centre n-1 kernel."""
    return _same_size_op(image, "mb200_edge_image_dev", "mb200_edge_image", float(radius))


def EmbossImage(image: Image, radius: float, sigma: float) -> Image:
    """MagickCore/effect.c:1600 -- ConvolveImage with the anti-diagonal emboss kernel, then EqualizeImage."""
    return _same_size_op(image, "mb200_emboss_image_dev", "mb200_emboss_image", float(radius), float(sigma))


def EqualizeImage(image: Image, sync_channels: bool = True) -> bool:
    """MagickCore/enhance.c:2040 -- in place.  sync_channels: the channel mask carries SyncChannels (the default)."""
    return _in_place(image, "mb200_equalize_image_dev", "mb200_equalize_image", int(bool(sync_channels)))


# MagickCore/statistic.h:141-151
(UndefinedStatistic, GradientStatistic, MaximumStatistic, MeanStatistic, MedianStatistic, MinimumStatistic, ModeStatistic,
 NonpeakStatistic, RootMeanSquareStatistic, StandardDeviationStatistic, ContrastStatistic) = range(11)


def StatisticImage(image: Image, statistic_type: int, width: int, height: int) -> Image:
    """MagickCore/statistic.c:2918."""
    return _same_size_op(image, "mb200_statistic_image_dev", "mb200_statistic_image", int(statistic_type), int(width),
                         int(height))


def RotationalBlurImage(image: Image, angle: float) -> Image:
    """MagickCore/effect.c:3129."""
    return _same_size_op(image, "mb200_rotational_blur_image_dev", "mb200_rotational_blur_image", float(angle))


def BilateralBlurImage(image: Image, width: int, height: int, intensity_sigma: float, spatial_sigma: float) -> Image:
    """MagickCore/effect.c:821."""
    return _same_size_op(image, "mb200_bilateral_blur_image_dev", "mb200_bilateral_blur_image", int(width), int(height),
                         float(intensity_sigma), float(spatial_sigma))


def AdaptiveBlurImage(image: Image, radius: float, sigma: float) -> Image:
    """MagickCore/effect.c:128."""
    return _same_size_op(image, "mb200_adaptive_blur_image_dev", "mb200_adaptive_blur_image", float(radius), float(sigma))


def AdaptiveSharpenImage(image: Image, radius: float, sigma: float) -> Image:
    """MagickCore/effect.c:447."""
    return _same_size_op(image, "mb200_adaptive_sharpen_image_dev", "mb200_adaptive_sharpen_image", float(radius), float(sigma))


def SelectiveBlurImage(image: Image, radius: float, sigma: float, threshold: float) -> Image:
    """MagickCore/effect.c:3406 (threshold in quantum units)."""
    return _same_size_op(image, "mb200_selective_blur_image_dev", "mb200_selective_blur_image", float(radius), float(sigma),
                         float(threshold))


def MotionBlurImage(image: Image, radius: float, sigma: float, angle: float) -> Image:
    """MagickCore/effect.c:2347."""
    return _same_size_op(image, "mb200_motion_blur_image_dev", "mb200_motion_blur_image", float(radius), float(sigma),
                         float(angle))


class FilterOptions(C.Structure):
    """mb200_filter_options: the "filter:*" expert settings of AcquireResizeFilter (MagickCore/resize.c:999-1226) as values."""
    _fields_ = [("set", C.c_uint), ("window", C.c_int), ("keep_filter", C.c_int), ("lobes", C.c_long), ("sigma", C.c_double),
                ("kaiser_beta", C.c_double), ("blur", C.c_double), ("support", C.c_double), ("win_support", C.c_double),
                ("b", C.c_double), ("c", C.c_double)]


_FILTER_NAMES = {n[:-len("Filter")].lower(): v for n, v in list(globals().items())
                 if n.endswith("Filter") and isinstance(v, int)}


def filter_options_from_artifacts(artifacts) -> "FilterOptions | None":
    """What the shim does in C (b200_filter_options): read the -define filter:* strings the way the reference does."""
    if not artifacts:
        return None
    o = FilterOptions()
    truthy = str(artifacts.get("filter:filter", "")).strip().lower() in ("true", "yes", "on", "1")
    w = artifacts.get("filter:window")
    if w is not None and str(w).strip().lower() in _FILTER_NAMES and _FILTER_NAMES[str(w).strip().lower()] > 0:
        o.window, o.keep_filter, o.set = _FILTER_NAMES[str(w).strip().lower()], int(truthy), o.set | 1
    for key, field, bit in (("filter:sigma", "sigma", 2), ("filter:alpha", "kaiser_beta", 4), ("filter:kaiser-beta", "kaiser_beta", 4),
                            ("filter:blur", "blur", 16), ("filter:support", "support", 32), ("filter:win-support", "win_support", 64),
                            ("filter:b", "b", 128), ("filter:c", "c", 256)):
        if key in artifacts:
            setattr(o, field, float(artifacts[key]))
            o.set |= bit
    if "filter:kaiser-alpha" in artifacts:
        o.kaiser_beta, o.set = float(artifacts["filter:kaiser-alpha"]) * 3.14159265358979323846264338327950288419716939937510, o.set | 4
    if "filter:lobes" in artifacts:
        o.lobes, o.set = int(float(artifacts["filter:lobes"])), o.set | 8
    return o if o.set else None


def ResizeImage(image: Image, columns: int, rows: int, filter: int = UndefinedFilter, artifacts=None) -> Image:
    """MagickCore/resize.c:3761.  `artifacts`: the image's "-define filter:*" settings, e.g. {"filter:blur": "0.8"}."""
    if columns <= 0 or rows <= 0:
        raise MagickB200Error(_lib.EINVAL, "NegativeOrZeroImageSize")
    lib = _lib.load()
    out = image._new_like(rows=rows, columns=columns)
    opts = filter_options_from_artifacts(artifacts)
    ref = C.byref(opts) if opts is not None else None
    if image.on_device:
        _activate(image)
        check(lib.mb200_resize_image_ex_dev(image._ptr(), image.columns, image.rows, image.channels, out._ptr(),
                                            columns, rows, int(filter), ref, _stream(image)))
    else:
        check(lib.mb200_resize_image_ex(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns,
                                        rows, int(filter), ref))
    return out


def SampleImage(image: Image, columns: int, rows: int) -> Image:
    """MagickCore/resize.c:3907 -- nearest-sample gather."""
    if columns <= 0 or rows <= 0:
        raise MagickB200Error(_lib.EINVAL, "NegativeOrZeroImageSize")
    lib = _lib.load()
    out = image._new_like(rows=rows, columns=columns)
    if image.on_device:
        _activate(image)
        check(lib.mb200_sample_image_dev(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns,
                                         rows, _stream(image)))
    else:
        check(lib.mb200_sample_image(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns, rows))
    return out


def ScaleImage(image: Image, columns: int, rows: int) -> Image:
    """MagickCore/resize.c:4106 -- box scaling."""
    if columns <= 0 or rows <= 0:
        raise MagickB200Error(_lib.EINVAL, "NegativeOrZeroImageSize")
    lib = _lib.load()
    out = image._new_like(rows=rows, columns=columns)
    if image.on_device:
        _activate(image)
        check(lib.mb200_scale_image_dev(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns, rows,
                                        _stream(image)))
    else:
        check(lib.mb200_scale_image(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns, rows))
    return out


def ThumbnailImage(image: Image, columns: int, rows: int, filter: int = UndefinedFilter) -> Image:
    """MagickCore/resize.c:4591 -- pixel path (sample / box / LanczosSharp cascade); `filter` is image->filter."""
    if columns <= 0 or rows <= 0:
        raise MagickB200Error(_lib.EINVAL, "NegativeOrZeroImageSize")
    lib = _lib.load()
    out = image._new_like(rows=rows, columns=columns)
    if image.on_device:
        _activate(image)
        check(lib.mb200_thumbnail_image_dev(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns,
                                            rows, int(filter), _stream(image)))
    else:
        check(lib.mb200_thumbnail_image(image._ptr(), image.columns, image.rows, image.channels, out._ptr(), columns,
                                        rows, int(filter)))
    return out


class ColorspaceOptions(C.Structure):
    """mb200_colorspace_options: the image settings sRGBTransformImage / TransformsRGBImage read (MagickCore/colorspace.c:761,
    :996, :1085-1095) as values."""
    _fields_ = [("set", C.c_uint), ("illuminant", C.c_int), ("white_luminance", C.c_double), ("film_gamma", C.c_double),
                ("reference_black", C.c_double), ("reference_white", C.c_double)]


_ILLUMINANTS = {"a": 0, "b": 1, "c": 2, "d50": 3, "d55": 4, "d65": 5, "d75": 6, "e": 7, "f2": 8, "f7": 9, "f11": 10}


def colorspace_options_from_settings(settings) -> "ColorspaceOptions | None":
    """What the shim does in C (b200_colorspace_options): the "color:illuminant" artifact and the "white-luminance",
    "film-gamma", "reference-black", "reference-white" properties.  An unparsable illuminant selects D65 like the
    reference's UndefinedIlluminant (MagickCore/color.h:42)."""
    if not settings:
        return None
    o = ColorspaceOptions()
    if "color:illuminant" in settings:
        o.illuminant, o.set = _ILLUMINANTS.get(str(settings["color:illuminant"]).strip().lower(), 5), o.set | 1
    for key, field, bit in (("white-luminance", "white_luminance", 2), ("film-gamma", "film_gamma", 4),
                            ("reference-black", "reference_black", 8), ("reference-white", "reference_white", 16)):
        if key in settings:
            setattr(o, field, float(settings[key]))
            o.set |= bit
    return o if o.set else None


def TransformImageColorspace(image: Image, colorspace: int, settings=None) -> bool:
    """MagickCore/colorspace.c:1751 -- in place; updates image.colorspace.  `settings`: the image's artifacts / properties
    the transform reads, e.g. {"color:illuminant": "D50"} or {"reference-white": "700"}."""
    lib = _lib.load()
    if image.colorspace == colorspace:
        return True
    opts = colorspace_options_from_settings(settings)
    ref = C.byref(opts) if opts is not None else None
    if image.on_device:
        _activate(image)
        check(lib.mb200_transform_colorspace_ex_dev(image._ptr(), image.columns, image.rows, image.channels,
                                                    image.colorspace, colorspace, ref, _stream(image)))
    else:
        check(lib.mb200_transform_colorspace_ex(image._ptr(), image.columns, image.rows, image.channels,
                                                image.colorspace, colorspace, ref))
    image.colorspace = colorspace
    return True


def _in_place(image: Image, dev_fn: str, host_fn: str, *args) -> bool:
    lib = _lib.load()
    if image.on_device:
        _activate(image)
        check(getattr(lib, dev_fn)(image._ptr(), image.columns, image.rows, image.channels, *args, _stream(image)))
    else:
        check(getattr(lib, host_fn)(image._ptr(), image.columns, image.rows, image.channels, *args))
    return True


def BilevelImage(image: Image, threshold: float) -> bool:
    """MagickCore/threshold.c:805 -- in place; a non-gray image is re-tagged sRGB (:827)."""
    ok = _in_place(image, "mb200_bilevel_image_dev", "mb200_bilevel_image", float(threshold))
    if image.channels >= 3:
        image.colorspace = sRGBColorspace
    return ok


def BlackThresholdImage(image: Image, thresholds: str) -> bool:
    """MagickCore/threshold.c:927 -- in place."""
    return _in_place(image, "mb200_black_threshold_image_dev", "mb200_black_threshold_image", int(image.colorspace),
                     thresholds.encode())


def WhiteThresholdImage(image: Image, thresholds: str) -> bool:
    """MagickCore/threshold.c:2518 -- in place."""
    return _in_place(image, "mb200_white_threshold_image_dev", "mb200_white_threshold_image", int(image.colorspace),
                     thresholds.encode())


def ClampImage(image: Image) -> bool:
    """MagickCore/threshold.c:1087 -- in place."""
    return _in_place(image, "mb200_clamp_image_dev", "mb200_clamp_image")


def MorphologyPrimitive(image: Image, method: int, kernel: Union[str, KernelInfo], bias: float = 0.0):
    """One MorphologyPrimitive pass (MagickCore/morphology.c:2566): returns (Image, changed).
    Device-resident images only (the primitive has no host-buffer entry point)."""
    if not image.on_device:
        raise ValueError("MorphologyPrimitive needs a device-resident Image")
    k = _as_kernel(kernel)
    out = image._new_like()
    changed = C.c_longlong(0)
    _activate(image)
    check(_lib.load().mb200_morphology_primitive_dev(image._ptr(), out._ptr(), image.columns, image.rows,
                                                     image.channels, int(method), k._ptr, float(bias),
                                                     C.byref(changed), _stream(image)))
    return out, int(changed.value)


def launch_count() -> int:
    return int(_lib.load().mb200_launch_count())


def device_count() -> int:
    return int(_lib.load().mb200_device_count())
