/*
  magick_b200.h -- C-ABI of libmagickb200.so: ImageMagick's per-pixel hot path
  (separable / 2-D convolution, erode/dilate morphology, filtered resize,
  sRGB<->Lab/XYZ/linear colourspace) as hand-written sm_100a CUDA kernels.

  Plain pointers and sizes only.  No torch / CUDA types in the signatures (a
  CUDA stream is passed as void*).  Pixel buffers are the reference's own pixel
  cache layout for its default Q16-HDRI build (MagickCore/cache.c:5142,
  MagickCore/pixel.c:6158): tightly packed, row-major, channel-interleaved
  float32 Quantum, values 0..65535, `channels` floats per pixel:

      channels 1 = Gray, 2 = Gray+Alpha, 3 = RGB, 4 = RGBA (alpha last).

  Images with alpha give their colour channels the Blend trait exactly like
  MagickCore/pixel.c:6356-6381 (alpha-weighted convolution / resize).

  Every entry point names the reference interface it replaces.  Return value:
  0 (MB200_OK) on success, a negative MB200_E* code otherwise; the message is
  available from mb200_last_error() (thread-local).  There is NO CPU fallback:
  when no sm_100 device is usable the calls fail with MB200_ENODEVICE -- the
  MagickCore shim (imagemagick_b200/shim) then returns NULL so that the caller's
  stock CPU path runs, which is the accelerate hook contract of
  MagickCore/effect.c:783-787 and MagickCore/resize.c:3818-3826.
*/
#ifndef MAGICK_B200_H
#define MAGICK_B200_H

#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

#if defined(_WIN32)
# define MB200_API
#else
# define MB200_API __attribute__((visibility("default")))
#endif

enum {
  MB200_OK = 0,
  MB200_EINVAL = -1,        /* bad argument */
  MB200_ENODEVICE = -2,     /* no usable CUDA device / wrong architecture */
  MB200_ECUDA = -3,         /* CUDA runtime error (see mb200_last_error) */
  MB200_ENOMEM = -4,
  MB200_EUNSUPPORTED = -5   /* valid in the reference but not implemented here:
                               the shim must fall back to the CPU path */
};

/* MagickCore/morphology.h:69-99 MorphologyMethod -- same numeric values */
typedef enum {
  MB200_UndefinedMorphology = 0,
  MB200_ConvolveMorphology = 1,
  MB200_CorrelateMorphology = 2,
  MB200_ErodeMorphology = 3,
  MB200_DilateMorphology = 4,
  MB200_ErodeIntensityMorphology = 5,
  MB200_DilateIntensityMorphology = 6,
  MB200_IterativeDistanceMorphology = 7,
  MB200_OpenMorphology = 8,
  MB200_CloseMorphology = 9,
  MB200_OpenIntensityMorphology = 10,
  MB200_CloseIntensityMorphology = 11,
  MB200_SmoothMorphology = 12,
  MB200_EdgeInMorphology = 13,
  MB200_EdgeOutMorphology = 14,
  MB200_EdgeMorphology = 15,
  MB200_TopHatMorphology = 16,
  MB200_BottomHatMorphology = 17,
  MB200_HitAndMissMorphology = 18,
  MB200_ThinningMorphology = 19,
  MB200_ThickenMorphology = 20
} mb200_morphology_method;

/* MagickCore/resample.h:32-69 FilterType -- same numeric values */
typedef enum {
  MB200_UndefinedFilter = 0, MB200_PointFilter, MB200_BoxFilter, MB200_TriangleFilter,
  MB200_HermiteFilter, MB200_HannFilter, MB200_HammingFilter, MB200_BlackmanFilter,
  MB200_GaussianFilter, MB200_QuadraticFilter, MB200_CubicFilter, MB200_CatromFilter,
  MB200_MitchellFilter, MB200_JincFilter, MB200_SincFilter, MB200_SincFastFilter,
  MB200_KaiserFilter, MB200_WelchFilter, MB200_ParzenFilter, MB200_BohmanFilter,
  MB200_BartlettFilter, MB200_LagrangeFilter, MB200_LanczosFilter, MB200_LanczosSharpFilter,
  MB200_Lanczos2Filter, MB200_Lanczos2SharpFilter, MB200_RobidouxFilter,
  MB200_RobidouxSharpFilter, MB200_CosineFilter, MB200_SplineFilter,
  MB200_LanczosRadiusFilter, MB200_CubicSplineFilter, MB200_MagicKernelSharp2013Filter,
  MB200_MagicKernelSharp2021Filter, MB200_SentinelFilter
} mb200_filter_type;

/* MagickCore/colorspace.h:27-67 ColorspaceType -- same numeric values */
typedef enum {
  MB200_CMYColorspace = 1,
  MB200_HCLColorspace = 4,            /* hue / saturation family of the generic branch (colorspace-private.h:149-529, */
  MB200_HCLpColorspace = 5,           /* :801-1064): bit exact, HSI <= 1 ULP                                        */
  MB200_HSBColorspace = 6,
  MB200_HSIColorspace = 7,
  MB200_HSLColorspace = 8,
  MB200_HSVColorspace = 9,
  MB200_HWBColorspace = 10,
  MB200_LabColorspace = 11,
  MB200_LCHColorspace = 12,           /* polar Lab (alias of LCHab) and Luv: the hue of an achromatic pixel is */
  MB200_LCHabColorspace = 13,         /* rounding noise in the reference as well                               */
  MB200_LCHuvColorspace = 14,
  MB200_LogColorspace = 15,           /* table gather, colorspace.c:1055-1163 / :2391-2500 */
  MB200_LMSColorspace = 16,           /* XYZ-derived spaces of the generic branch: <= 1 ULP */
  MB200_LuvColorspace = 17,
  MB200_OHTAColorspace = 18,          /* LUT branch, colorspace.c:1229-1494 */
  MB200_Rec601YCbCrColorspace = 19,   /* LUT branch */
  MB200_Rec709YCbCrColorspace = 20,   /* LUT branch */
  MB200_RGBColorspace = 21,           /* linear RGB */
  MB200_sRGBColorspace = 23,
  MB200_xyYColorspace = 25,
  MB200_XYZColorspace = 26,
  MB200_YCbCrColorspace = 27,
  MB200_YCCColorspace = 28,           /* PhotoYCC, LUT branch :1347-1389 / :2681-2711, bit exact */
  MB200_YDbDrColorspace = 29,
  MB200_YIQColorspace = 30,
  MB200_YPbPrColorspace = 31,
  MB200_YUVColorspace = 32,
  MB200_JzazbzColorspace = 34,
  MB200_DisplayP3Colorspace = 35,
  MB200_Adobe98Colorspace = 36,
  MB200_ProPhotoColorspace = 37,
  MB200_OklabColorspace = 38,
  MB200_OklchColorspace = 39,
  MB200_CAT02LMSColorspace = 40
} mb200_colorspace;

/* Mirror of KernelInfo (MagickCore/morphology.h:102-130): a singly linked list
   of width x height double arrays, NaN == "not part of the neighbourhood",
   (x,y) the origin.  `type` is an mb200_kernel_type (needed only because the
   reference's 180-degree RotateKernelInfo is a no-op for some built-ins,
   MagickCore/morphology.c:4281-4305). */
typedef enum {
  MB200_UserDefinedKernel = 0, MB200_BlurKernel, MB200_GaussianKernel, MB200_DiskKernel,
  MB200_SquareKernel, MB200_DiamondKernel, MB200_OctagonKernel, MB200_PlusKernel,
  MB200_CrossKernel, MB200_RectangleKernel, MB200_UnityKernel, MB200_DoGKernel,
  MB200_LoGKernel, MB200_BinomialKernel,
  /* the rest of KernelInfoType (MagickCore/morphology.h:27-68): named convolution kernels, hit-and-miss lists, distances */
  MB200_CometKernel, MB200_LaplacianKernel, MB200_SobelKernel, MB200_FreiChenKernel, MB200_RobertsKernel,
  MB200_PrewittKernel, MB200_CompassKernel, MB200_KirschKernel, MB200_RingKernel, MB200_PeaksKernel, MB200_EdgesKernel,
  MB200_CornersKernel, MB200_DiagonalsKernel, MB200_LineEndsKernel, MB200_LineJunctionsKernel, MB200_RidgesKernel,
  MB200_ConvexHullKernel, MB200_ThinSEKernel, MB200_SkeletonKernel, MB200_ChebyshevKernel, MB200_ManhattanKernel,
  MB200_OctagonalKernel, MB200_EuclideanKernel
} mb200_kernel_type;

typedef struct mb200_kernel_info {
  int type;
  size_t width, height;
  long x, y;
  double *values;
  double minimum, maximum, negative_range, positive_range, angle;
  struct mb200_kernel_info *next;
} mb200_kernel_info;

/* ------------------------------------------------------------- runtime ---- */

/* Number of usable sm_100 devices (0 when there is no GPU / driver). */
MB200_API int mb200_device_count(void);
/* Bind the calling thread (and the library's per-device state) to `device`.
   One process per GPU is the intended deployment (imagemagick_b200.dist). */
MB200_API int mb200_set_device(int device);
MB200_API const char *mb200_last_error(void);
MB200_API const char *mb200_version(void);
/* Number of kernel launches issued by this library since process start
   (bench.py's "gpu_launches"). */
MB200_API unsigned long long mb200_launch_count(void);
/* Block until all work queued by this library on `stream` (NULL = the library's
   own stream for the current device) has finished. */
MB200_API int mb200_synchronize(void *stream);

/* Device pixel-cache staging (the CUDA analogue of AcquireMagickCLCacheInfo /
   GetAuthenticOpenCLBuffer, MagickCore/opencl.c:528-553, MagickCore/cache.c:1259). */
MB200_API int mb200_malloc(void **dev_ptr, size_t bytes);
MB200_API int mb200_free(void *dev_ptr);
MB200_API int mb200_malloc_host(void **host_ptr, size_t bytes);   /* pinned */
MB200_API int mb200_free_host(void *host_ptr);
/* Asynchronous on `stream` when the host memory is pinned; pageable memory is moved through the bounce ring
   (mb200_upload returns once the host buffer has been consumed, mb200_download when it holds the data). */
MB200_API int mb200_upload(void *dev_dst, const void *host_src, size_t bytes, void *stream);
MB200_API int mb200_download(void *host_dst, const void *dev_src, size_t bytes, void *stream);

/* Gives the memory cached in the library's private stream-ordered pool (operator temporaries of the
   current device) back to the driver, keeping at most `keep_bytes`.  The host application's default pool is
   never touched. */
MB200_API int mb200_trim(size_t keep_bytes);
/* Measures the FP64 FMA issue rate of the current device (FMA/s; 16 independent DFMA chains per thread):
   the co-limit of the FP64-accumulating convolution kernels (bench.py's second roofline entry). */
MB200_API int mb200_probe_fp64_fma_rate(double *fma_per_second);
/* Test / developer hook: force the generic kernels ("no_rank1", "no_morph_stream", "no_resize_stream",
   "resize_regular_h", "no_fused_unsharp", "resize_fused", "conv_mma" = the FP64 mma.sync kernels of conv_mma.cu for
   RGBA 1-D passes: 1 whenever possible, 0 never, -1 automatic = float-in / float-out passes; initialised from the MB200_<NAME> environment variables).  mb200_get_option reads a switch back;
   "conv_mma_launches" counts the passes the mma.sync kernels have served since process start. */
MB200_API int mb200_set_option(const char *name, int value);
MB200_API int mb200_get_option(const char *name, int *value);

/* ------------------------------------------ pixel cache staged into HBM ---- */
/* Residency of HOST pixel caches in HBM -- the CUDA analogue of the reference's OpenCL cache plumbing:
     mb200_cache_attach        AcquireMagickCLCacheInfo      MagickCore/opencl.c:528-553
     (operators below)         GetAuthenticOpenCLBuffer      MagickCore/cache.c:1259-1292
     mb200_cache_sync          CopyOpenCLBuffer              MagickCore/cache.c:5341-5364 (sites :1710, :2771, :4079)
     mb200_cache_detach        RelinquishMagickCLCacheInfo   MagickCore/cache.c:978-984
   An attached buffer keeps one HBM copy for its lifetime.  The host-buffer operators find it by the buffer's
   base address: inputs whose HBM copy is current are not uploaded again, results are written to the HBM copy.
   EAGER mode (default): a result is copied to its host buffer before the operator returns, and an input's HBM
   copy is trusted only while it is the sole current copy -- always safe, no cooperation needed.
   LAZY mode (mb200_cache_set_lazy(1)): results stay in HBM until mb200_cache_sync(host) and resident inputs
   are not re-uploaded; the caller must call mb200_cache_sync before the host READS a buffer and
   mb200_cache_host_written after the host WRITES one (the three CopyOpenCLBuffer sites of cache.c are exactly
   those places; INTEGRATION.md shows the patch).  Unattached buffers are staged per call in either mode.
   Host memory that is pinned (mb200_malloc_host, cudaHostRegister, MB200_CACHE_REGISTER) moves at PCIe speed
   (55 GB/s); pageable memory goes through a threaded pinned bounce ring (49 / 44 GB/s up / down instead of
   cudaMemcpy's 9 / 19 GB/s). */
enum { MB200_CACHE_REGISTER = 1 };   /* pin the buffer with cudaHostRegister (67-450 ms per GiB, once) */
MB200_API int mb200_cache_attach(void *host_pixels, size_t bytes, int flags);
MB200_API int mb200_cache_detach(void *host_pixels);            /* drops the HBM copy WITHOUT syncing */
MB200_API int mb200_cache_sync(void *host_pixels);              /* HBM -> host if the HBM copy is newer */
MB200_API int mb200_cache_host_written(void *host_pixels);      /* the HBM copy is stale */
MB200_API int mb200_cache_resident(const void *host_pixels);    /* -1 unknown; bit 0: HBM copy current, bit 1: host copy current */
MB200_API int mb200_cache_set_lazy(int on);                     /* returns the previous mode */
/* uploads, upload bytes, downloads, download bytes, resident-input hits, bytes moved through the bounce ring */
MB200_API void mb200_cache_stats(unsigned long long out[6]);
MB200_API int mb200_copy_threads(void);

/* ------------------------------------------------------- kernel builders ---- */

/* AcquireKernelInfo (MagickCore/morphology.c:485): parses the reference's kernel
   strings -- "blur:RxS[+angle]", "gaussian:RxS", "dog:", "log:", "disk:R[,scale]",
   "square:", "diamond:", "octagon:", "plus:", "cross:", "rectangle:WxH+X+Y",
   "unity", "binomial:", user arrays "WxH+X+Y:v,v,..." and old-style "v,v,v,..." --
   and ';'-separated lists.  Returns NULL on a parse error / unsupported name. */
MB200_API mb200_kernel_info *mb200_acquire_kernel_info(const char *kernel_string);
/* AcquireKernelBuiltIn (MagickCore/morphology.c:950) with GeometryInfo rho,sigma,xi,psi. */
MB200_API mb200_kernel_info *mb200_acquire_kernel_builtin(int type, double rho, double sigma,
                                                         double xi, double psi);
MB200_API mb200_kernel_info *mb200_clone_kernel_info(const mb200_kernel_info *kernel);
MB200_API mb200_kernel_info *mb200_destroy_kernel_info(mb200_kernel_info *kernel);
/* ScaleKernelInfo (MagickCore/morphology.c:4571); flags: 1 = NormalizeValue,
   2 = CorrelateNormalizeValue. */
MB200_API void mb200_scale_kernel_info(mb200_kernel_info *kernel, double scaling_factor, int flags);
/* GetOptimalKernelWidth1D/2D (MagickCore/gem.c:262, :302). */
MB200_API size_t mb200_optimal_kernel_width_1d(double radius, double sigma);
MB200_API size_t mb200_optimal_kernel_width_2d(double radius, double sigma);

/* The kernels SharpenImage (MagickCore/effect.c:3991-4063) and EdgeImage (:1520-1570) build inline
   before calling ConvolveImage. */
MB200_API mb200_kernel_info *mb200_sharpen_kernel(double radius, double sigma);
MB200_API mb200_kernel_info *mb200_edge_kernel(double radius);
/* The anti-diagonal kernel EmbossImage (MagickCore/effect.c:1600-1665) builds inline. */
MB200_API mb200_kernel_info *mb200_emboss_kernel(double radius, double sigma);

/* MotionBlurImage's taps (GetMotionBlurKernel, MagickCore/effect.c:2316-2345) and integer offsets along
   `angle` (:2390-2398).  Returns the tap count; pass NULL arrays to query it. */
MB200_API long mb200_motion_blur_kernel(double radius, double sigma, double angle, double *taps,
    long *offset_x, long *offset_y, size_t max_taps);

/* Resize contribution table of one axis: exactly the start/stop/weights that
   HorizontalFilter / VerticalFilter (MagickCore/resize.c:3398-3443, :3614-3657)
   compute per output column/row.  weights is out_n * max_taps doubles (row o at
   weights[o*max_taps]); returns max_taps (>0) or a negative error.  Pass
   start/count/weights == NULL to only query max_taps. */
MB200_API long mb200_resize_contributions(int filter, size_t in_n, size_t out_n, double factor,
                                         long *start, int *count, double *weights,
                                         size_t max_taps);
/* The "filter:*" expert settings AcquireResizeFilter reads from the image artifacts (MagickCore/resize.c:999-1226), as
   values: the caller (the shim, with the reference's own StringToDouble / ParseCommandOption) does the string parsing.
   `set` says which fields are meaningful.  window + keep_filter restate :999-1043: a "filter:window" alone turns the
   weighting function into SincFast; with a truthy "filter:filter" the requested filter keeps its weighting function. */
enum { MB200_FO_WINDOW = 1, MB200_FO_SIGMA = 2, MB200_FO_KAISER_BETA = 4, MB200_FO_LOBES = 8, MB200_FO_BLUR = 16,
       MB200_FO_SUPPORT = 32, MB200_FO_WIN_SUPPORT = 64, MB200_FO_B = 128, MB200_FO_C = 256 };
typedef struct mb200_filter_options {
  unsigned set;
  int window;              /* FilterType of "filter:window" */
  int keep_filter;         /* "filter:filter" was a truthy string (:1000) */
  long lobes;              /* "filter:lobes" */
  double sigma;            /* "filter:sigma" (Gaussian) */
  double kaiser_beta;      /* "filter:alpha" / "filter:kaiser-beta" / pi * "filter:kaiser-alpha", last one wins (:1104-1117) */
  double blur, support, win_support, b, c;
} mb200_filter_options;
/* ... with expert settings (options == NULL: none) */
MB200_API long mb200_resize_contributions_ex(int filter, const mb200_filter_options *options, size_t in_n, size_t out_n,
                                            double factor, long *start, int *count, double *weights, size_t max_taps);
MB200_API double mb200_resize_filter_weight_ex(int filter, const mb200_filter_options *options, double x);
MB200_API double mb200_resize_filter_support_ex(int filter, const mb200_filter_options *options);
/* GetResizeFilterWeight / GetResizeFilterSupport (MagickCore/resize.c:1690, :1656). */
MB200_API double mb200_resize_filter_weight(int filter, double x);
MB200_API double mb200_resize_filter_support(int filter);

/* --------------------------------------- device-resident operators (HBM) ---- */
/* src/dst are DEVICE pointers (from mb200_malloc or any CUDA allocation, e.g. a
   torch tensor's data_ptr()); they must not alias unless stated.  `stream` is a
   cudaStream_t passed as void* (NULL = library stream).  Calls are asynchronous
   with respect to the host unless a `changed` result is requested. */

/* MorphologyPrimitive (MagickCore/morphology.c:2566-3227) for ONE kernel:
   method in {Convolve, Erode, Dilate}.  *changed (host, may be NULL) receives the
   reference's return value (pixels changed).  Width-1 Convolve kernels take the
   reference's column path semantics (:2654-2807). */
MB200_API int mb200_morphology_primitive_dev(const float *src, float *dst, size_t width,
    size_t height, int channels, int method, const mb200_kernel_info *kernel, double bias,
    long long *changed, void *stream);

/* MorphologyImage / MorphologyApply (MagickCore/morphology.c:4129, :3634) with the
   default compose (re-iterate kernel lists): iterations (<0 = until unchanged),
   compound methods Open / Close / Smooth, Correlate, and -- for single kernels -- the
   "difference" methods EdgeIn / EdgeOut / Edge / TopHat / BottomHat, whose final
   CompositeImage(..., DifferenceCompositeOp, ...) step (:3995-4012) runs as one point
   kernel (bit exact).  HitAndMiss / Thinning / Thicken / Distance / Voronoi / the
   *Intensity methods return MB200_EUNSUPPORTED. */
MB200_API int mb200_morphology_image_dev(const float *src, float *dst, size_t width,
    size_t height, int channels, int method, long iterations,
    const mb200_kernel_info *kernel, double bias, void *stream);

/* ConvolveImage (MagickCore/effect.c:1170) */
MB200_API int mb200_convolve_image_dev(const float *src, float *dst, size_t width, size_t height,
    int channels, const mb200_kernel_info *kernel, void *stream);
/* BlurImage (MagickCore/effect.c:765) == AccelerateBlurImage (accelerate-private.h:36) */
MB200_API int mb200_blur_image_dev(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, double sigma, void *stream);
/* GaussianBlurImage (MagickCore/effect.c:1709) */
MB200_API int mb200_gaussian_blur_image_dev(const float *src, float *dst, size_t width,
    size_t height, int channels, double radius, double sigma, void *stream);
/* UnsharpMaskImage (MagickCore/effect.c:4256) == AccelerateUnsharpMaskImage (:46) */
MB200_API int mb200_unsharp_mask_image_dev(const float *src, float *dst, size_t width,
    size_t height, int channels, double radius, double sigma, double gain, double threshold,
    void *stream);
/* SharpenImage (MagickCore/effect.c:3991) and EdgeImage (:1520): ConvolveImage with the kernel above */
MB200_API int mb200_sharpen_image_dev(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, double sigma, void *stream);
MB200_API int mb200_edge_image_dev(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, void *stream);
/* EqualizeImage (MagickCore/enhance.c:2040; the reference's hook is AccelerateEqualizeImage, accelerate-private.h), in place:
   histogram on the device, the 65 536-entry map on the host in the reference's order, table lookup on the device --
   bit exact.  sync_channels != 0: the channel mask carries SyncChannels (the default mask does): one histogram of the
   pixel intensity for all channels; 0: one histogram per channel.  Synchronises `stream`. */
MB200_API int mb200_equalize_image_dev(float *buf, size_t width, size_t height, int channels, int sync_channels,
    void *stream);
/* EmbossImage (MagickCore/effect.c:1600): ConvolveImage with mb200_emboss_kernel, then EqualizeImage. */
MB200_API int mb200_emboss_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma, void *stream);
/* StatisticImage (MagickCore/statistic.c:2918): `type` is a StatisticType (statistic.h:141-151: 1 Gradient, 2 Maximum,
   3 Mean, 4 Median, 5 Minimum, 6 Mode, 7 Nonpeak, 8 RootMeanSquare, 9 StandardDeviation, 10 Contrast) over a
   window_width x window_height neighbourhood, bit exact. */
MB200_API int mb200_statistic_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    int type, size_t window_width, size_t window_height, void *stream);
/* RotationalBlurImage (MagickCore/effect.c:3129) == AccelerateRotationalBlurImage (accelerate-private.h), bit exact. */
MB200_API int mb200_rotational_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    double angle, void *stream);
/* BilateralBlurImage (MagickCore/effect.c:821), odd window sizes (even ones return MB200_EUNSUPPORTED), bit exact. */
MB200_API int mb200_bilateral_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    size_t window_width, size_t window_height, double intensity_sigma, double spatial_sigma, void *stream);
/* AdaptiveBlurImage (MagickCore/effect.c:128) / AdaptiveSharpenImage (:447): EdgeImage -> AutoLevelImage -> BlurImage ->
   AutoLevelImage gives the edge map that selects a kernel size per pixel; every stage in the reference's operation
   order, bit exact. */
MB200_API int mb200_adaptive_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma, void *stream);
MB200_API int mb200_adaptive_sharpen_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma, void *stream);
/* SelectiveBlurImage (MagickCore/effect.c:3406): contrast-gated Gaussian (threshold in quantum units), bit exact. */
MB200_API int mb200_selective_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma, double threshold, void *stream);
/* MotionBlurImage (MagickCore/effect.c:2347) == AccelerateMotionBlurImage (accelerate-private.h). */
MB200_API int mb200_motion_blur_image_dev(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, double sigma, double angle, void *stream);
/* ResizeImage (MagickCore/resize.c:3761) == AccelerateResizeImage (:43).  filter
   UndefinedFilter applies the reference's own default choice (:3806-3816). */
MB200_API int mb200_resize_image_ex_dev(const float *src, size_t width, size_t height, int channels, float *dst,
                                       size_t out_width, size_t out_height, int filter,
                                       const mb200_filter_options *options, void *stream);
MB200_API int mb200_resize_image_dev(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t out_width, size_t out_height, int filter, void *stream);
/* SampleImage (MagickCore/resize.c:3907): nearest-sample gather with the default sampling offset
   (0.5 - MagickEpsilon; the "sample:offset" artifact is the shim's decline), bit exact. */
MB200_API int mb200_sample_image_dev(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t out_width, size_t out_height, void *stream);
/* ScaleImage (MagickCore/resize.c:4106): box scaling.  The reference's sequential (span, scale) state machines are
   simulated on the host (mb200_scale_contributions) into per-output term lists; the gather kernel accumulates them in
   the reference's order with unfused double operations: bit exact, reductions and enlargements alike. */
MB200_API int mb200_scale_image_dev(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t out_width, size_t out_height, void *stream);
MB200_API long mb200_scale_contributions(int axis, size_t in_n, size_t out_n, long *offsets, int *index,
    double *weight, size_t max_terms);
/* ThumbnailImage (MagickCore/resize.c:4591-4650), pixel path only: SampleImage to 4x the target when both
   integer reduction factors exceed 4, ResizeImage(BoxFilter) to 2x when they exceed 2, then
   ResizeImage(`filter` = image->filter; UndefinedFilter selects LanczosSharp like the reference).  The
   profile stripping and Thumb::* properties the reference adds afterwards are left to the caller. */
MB200_API int mb200_thumbnail_image_dev(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t columns, size_t rows, int filter, void *stream);
/* TransformImageColorspace (MagickCore/colorspace.c:1751), in place on `buf`. */
MB200_API int mb200_transform_colorspace_dev(float *buf, size_t width, size_t height,
    int channels, int from_colorspace, int to_colorspace, void *stream);
/* ... with the image settings the reference reads inside sRGBTransformImage / TransformsRGBImage, as values (the shim
   parses them with the reference's own functions): the "color:illuminant" artifact (colorspace.c:761-773, :2093-2105) =
   the reference white of Lab, LCH, LCHab, LCHuv and Luv; the "white-luminance" property (:996, :2331) of Jzazbz; the
   "film-gamma", "reference-black" and "reference-white" properties (:1085-1095, :2416-2426) of Log.  (Log's "gamma"
   property, :1081, cannot be set on an image: SetImageProperty diverts that key to image->gamma, property.c:4583.)
   options == NULL or a clear `set` bit: the reference's default. */
typedef enum {                          /* MagickCore/color.h:40-54 IlluminantType -- same numeric values */
  MB200_AIlluminant = 0, MB200_BIlluminant, MB200_CIlluminant, MB200_D50Illuminant, MB200_D55Illuminant,
  MB200_D65Illuminant, MB200_D75Illuminant, MB200_EIlluminant, MB200_F2Illuminant, MB200_F7Illuminant, MB200_F11Illuminant
} mb200_illuminant;
enum { MB200_CO_ILLUMINANT = 1, MB200_CO_WHITE_LUMINANCE = 2, MB200_CO_FILM_GAMMA = 4, MB200_CO_REFERENCE_BLACK = 8,
       MB200_CO_REFERENCE_WHITE = 16 };
typedef struct mb200_colorspace_options {
  unsigned set;
  int illuminant;                      /* mb200_illuminant */
  double white_luminance;
  double film_gamma, reference_black, reference_white;    /* reference_black / _white within 0..1024 */
} mb200_colorspace_options;
MB200_API int mb200_transform_colorspace_ex_dev(float *buf, size_t width, size_t height, int channels,
    int from_colorspace, int to_colorspace, const mb200_colorspace_options *options, void *stream);
/* The host-built tables behind the Log and YCC legs (no device needed): the 65536-entry `logmap` of colorspace.c:1100-1106
   (forward != 0) / :2431-2440 (inverse) for the given film settings, and the 1389-entry PhotoYCC table of :1828. */
MB200_API int mb200_log_colorspace_table(int forward, const mb200_colorspace_options *options, float *table65536);
MB200_API int mb200_ycc_table(float *table1389);

/* Threshold point operators of MagickCore/threshold.c, in place on `buf`, bit exact.
   BilevelImage (:805): every channel (alpha included) := intensity <= threshold ? 0 : QuantumRange,
   intensity = GetPixelIntensity (pixel.c:2356, Rec709Luma; the gray sample for Gray / Gray+Alpha).
   A non-gray image is re-tagged sRGB by the reference (:827); the caller owns that tag. */
MB200_API int mb200_bilevel_image_dev(float *buf, size_t width, size_t height, int channels,
    double threshold, void *stream);
/* BlackThresholdImage (:927) / WhiteThresholdImage (:2518).  `thresholds` is the reference's
   geometry string "v[,v[,v[,v]]][%]" = red[,green[,blue[,alpha]]] (missing green/blue default to
   red, alpha to 100, '%' scales all by QuantumRange/100, :955-985).  Gray images (which the
   reference first promotes to sRGB, :949), linear-RGB images (`colorspace` == MB200_RGBColorspace:
   the intensity then needs EncodePixelGamma) and other geometry syntax return MB200_EUNSUPPORTED. */
MB200_API int mb200_black_threshold_image_dev(float *buf, size_t width, size_t height, int channels,
    int colorspace, const char *thresholds, void *stream);
MB200_API int mb200_white_threshold_image_dev(float *buf, size_t width, size_t height, int channels,
    int colorspace, const char *thresholds, void *stream);
/* ClampImage (:1087): ClampPixel on every channel (HDRI: below 0 -> 0, >= QuantumRange -> QuantumRange). */
MB200_API int mb200_clamp_image_dev(float *buf, size_t width, size_t height, int channels, void *stream);

/* Copy-trait channels.  With a `-channel` selection the reference hands the unselected channels through from the
   operator's source (MagickCore/morphology.c:2733-2737, effect.c:4346-4350); ResizeImage takes the nearest source sample
   of each pass (resize.c:3697-3707).  The operators above compute every channel; these point passes put the Copy channels
   back into `dst` (bit c of update_mask set = channel c is updated by the operator).  Valid for the operators whose
   selected channels do not depend on the selection: Blur, GaussianBlur, Convolve, the non-difference morphology methods,
   UnsharpMask, Sharpen, Edge and Resize. */
MB200_API int mb200_restore_channels_dev(float *dst, const float *src, size_t width, size_t height, int channels,
    unsigned update_mask, void *stream);
MB200_API int mb200_resize_copy_channels_dev(const float *src, size_t width, size_t height, int channels, float *dst,
    size_t out_width, size_t out_height, int filter, unsigned update_mask, void *stream);

/* ---------------------------------------------- host-buffer operators ---- */
/* Same operators on HOST buffers: stage into HBM, run, copy back, synchronise.
   These are what the MagickCore shim calls with the pixel-cache pointers
   (GetVirtualPixels / GetAuthenticPixels, MagickCore/cache.c:3257, :1490). */
MB200_API int mb200_blur_image(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, double sigma);
MB200_API int mb200_gaussian_blur_image(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, double sigma);
MB200_API int mb200_convolve_image(const float *src, float *dst, size_t width, size_t height,
    int channels, const mb200_kernel_info *kernel);
MB200_API int mb200_morphology_image(const float *src, float *dst, size_t width, size_t height,
    int channels, int method, long iterations, const mb200_kernel_info *kernel, double bias);
MB200_API int mb200_unsharp_mask_image(const float *src, float *dst, size_t width, size_t height,
    int channels, double radius, double sigma, double gain, double threshold);
MB200_API int mb200_sharpen_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma);
MB200_API int mb200_edge_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius);
MB200_API int mb200_motion_blur_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma, double angle);
MB200_API int mb200_restore_channels(float *dst, const float *src, size_t width, size_t height, int channels,
    unsigned update_mask);
MB200_API int mb200_resize_copy_channels(const float *src, size_t width, size_t height, int channels, float *dst,
    size_t out_width, size_t out_height, int filter, unsigned update_mask);
MB200_API int mb200_statistic_image(const float *src, float *dst, size_t width, size_t height, int channels, int type,
    size_t window_width, size_t window_height);
MB200_API int mb200_rotational_blur_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double angle);
MB200_API int mb200_bilateral_blur_image(const float *src, float *dst, size_t width, size_t height, int channels,
    size_t window_width, size_t window_height, double intensity_sigma, double spatial_sigma);
MB200_API int mb200_adaptive_blur_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma);
MB200_API int mb200_adaptive_sharpen_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma);
MB200_API int mb200_selective_blur_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma, double threshold);
MB200_API int mb200_equalize_image(float *buf, size_t width, size_t height, int channels, int sync_channels);
MB200_API int mb200_emboss_image(const float *src, float *dst, size_t width, size_t height, int channels,
    double radius, double sigma);
MB200_API int mb200_resize_image_ex(const float *src, size_t width, size_t height, int channels, float *dst,
                                   size_t out_width, size_t out_height, int filter, const mb200_filter_options *options);
MB200_API int mb200_resize_image(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t out_width, size_t out_height, int filter);
MB200_API int mb200_sample_image(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t out_width, size_t out_height);
MB200_API int mb200_scale_image(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t out_width, size_t out_height);
MB200_API int mb200_thumbnail_image(const float *src, size_t width, size_t height, int channels,
    float *dst, size_t columns, size_t rows, int filter);
MB200_API int mb200_transform_colorspace(float *buf, size_t width, size_t height, int channels,
    int from_colorspace, int to_colorspace);
MB200_API int mb200_transform_colorspace_ex(float *buf, size_t width, size_t height, int channels,
    int from_colorspace, int to_colorspace, const mb200_colorspace_options *options);
MB200_API int mb200_bilevel_image(float *buf, size_t width, size_t height, int channels, double threshold);
MB200_API int mb200_black_threshold_image(float *buf, size_t width, size_t height, int channels,
    int colorspace, const char *thresholds);
MB200_API int mb200_white_threshold_image(float *buf, size_t width, size_t height, int channels,
    int colorspace, const char *thresholds);
MB200_API int mb200_clamp_image(float *buf, size_t width, size_t height, int channels);

#if defined(__cplusplus)
}
#endif
#endif /* MAGICK_B200_H */
