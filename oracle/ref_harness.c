/*
  oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.

  Thin driver around the UNMODIFIED reference MagickCore (compiled from
  /root/reference by oracle/Makefile into oracle/_ref/libmagickref.so).  It lets
  tests and bench.py's cpu_baseline leg run the reference's own BlurImage,
  GaussianBlurImage, ConvolveImage, MorphologyImage, UnsharpMaskImage,
  ResizeImage and TransformImageColorspace on raw, tightly packed float buffers
  (Quantum == float, values in [0,65535], channels interleaved) with no file I/O
  and no scaling, so results can be compared bit-for-bit.

  channels: 1 = Gray, 2 = Gray+Alpha, 3 = RGB, 4 = RGBA  (the reference's own
  channel order, pixel.c ResetPixelChannelMap).
*/
#include "MagickCore/studio.h"
#include "MagickCore/MagickCore.h"
#include <string.h>

static int g_init = 0;

static void ensure_init(void)
{
  if (!g_init) { MagickCoreGenesis("magickref", MagickFalse); g_init = 1; }
}

__attribute__((visibility("default")))
void ref_set_threads(int n)
{
  ensure_init();
  (void) SetMagickResourceLimit(ThreadResource, (MagickSizeType) n);
#ifdef _OPENMP
  omp_set_num_threads(n);
#endif
}

__attribute__((visibility("default")))
int ref_get_threads(void)
{
  ensure_init();
  return (int) GetMagickResourceLimit(ThreadResource);
}

static Image *make_image(const float *src, size_t w, size_t h, int ch,
                         int colorspace, ExceptionInfo *ex)
{
  ImageInfo *info;
  Image *im;
  Quantum *q;

  ensure_init();
  info = AcquireImageInfo();
  im = AcquireImage(info, ex);
  info = DestroyImageInfo(info);
  if (im == (Image *) NULL) return im;
  if (SetImageExtent(im, w, h, ex) == MagickFalse) return DestroyImage(im);
  if (ch == 1 || ch == 2)
    (void) SetImageColorspace(im, GRAYColorspace, ex);
  if (ch == 2 || ch == 4)
    im->alpha_trait = BlendPixelTrait;
  if (colorspace >= 0 && ch >= 3)
    im->colorspace = (ColorspaceType) colorspace;   /* tag only, no transform */
  (void) SetImageStorageClass(im, DirectClass, ex);
  /* re-derive the channel map after the alpha_trait change (pixel.c:6132) */
  (void) SetImageColorspace(im, im->colorspace, ex);
  if ((int) GetPixelChannels(im) != ch) return DestroyImage(im);
  q = GetAuthenticPixels(im, 0, 0, w, h, ex);
  if (q == (Quantum *) NULL) return DestroyImage(im);
  memcpy(q, src, w * h * (size_t) ch * sizeof(float));
  (void) SyncAuthenticPixels(im, ex);
  return im;
}

static int export_image(const Image *im, float *dst, size_t w, size_t h, int ch,
                        ExceptionInfo *ex)
{
  const Quantum *p;
  if (im == (const Image *) NULL) return -1;
  if (im->columns != w || im->rows != h || (int) GetPixelChannels(im) != ch) return -2;
  p = GetVirtualPixels(im, 0, 0, w, h, ex);
  if (p == (const Quantum *) NULL) return -3;
  memcpy(dst, p, w * h * (size_t) ch * sizeof(float));
  return 0;
}

#define BEGIN ExceptionInfo *ex = AcquireExceptionInfo(); Image *im = NULL, *out = NULL; int rc = -1;
#define END   if (out) out = DestroyImage(out); if (im) im = DestroyImage(im); ex = DestroyExceptionInfo(ex); return rc;

__attribute__((visibility("default")))
int ref_blur(const float *src, float *dst, size_t w, size_t h, int ch,
             double radius, double sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = BlurImage(im, radius, sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_gaussian_blur(const float *src, float *dst, size_t w, size_t h, int ch,
                      double radius, double sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = GaussianBlurImage(im, radius, sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_unsharp(const float *src, float *dst, size_t w, size_t h, int ch,
                double radius, double sigma, double gain, double threshold)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = UnsharpMaskImage(im, radius, sigma, gain, threshold, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_sharpen(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = SharpenImage(im, radius, sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_edge(const float *src, float *dst, size_t w, size_t h, int ch, double radius)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = EdgeImage(im, radius, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_adaptive_blur(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = AdaptiveBlurImage(im, radius, sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_adaptive_sharpen(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = AdaptiveSharpenImage(im, radius, sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_selective_blur(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma, double threshold)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = SelectiveBlurImage(im, radius, sigma, threshold, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_emboss(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = EmbossImage(im, radius, sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

/* EqualizeImage in place; sync == 0 clears SyncChannels from the channel mask (per-channel histograms) */
__attribute__((visibility("default")))
int ref_equalize(float *buf, size_t w, size_t h, int ch, int sync)
{
  BEGIN
  im = make_image(buf, w, h, ch, -1, ex);
  if (im)
    {
      if (!sync) (void) SetPixelChannelMask(im, (ChannelType) (AllChannels & ~SyncChannels));
      if (EqualizeImage(im, ex) != MagickFalse) rc = export_image(im, buf, w, h, ch, ex);
    }
  END
}

__attribute__((visibility("default")))
int ref_motion_blur(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma, double angle)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = MotionBlurImage(im, radius, sigma, angle, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_bilateral_blur(const float *src, float *dst, size_t w, size_t h, int ch, size_t width, size_t height,
                       double intensity_sigma, double spatial_sigma)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = BilateralBlurImage(im, width, height, intensity_sigma, spatial_sigma, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_rotational_blur(const float *src, float *dst, size_t w, size_t h, int ch, double angle)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = RotationalBlurImage(im, angle, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_statistic(const float *src, float *dst, size_t w, size_t h, int ch, int type, size_t width, size_t height)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = StatisticImage(im, (StatisticType) type, width, height, ex); rc = export_image(out, dst, w, h, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_convolve(const float *src, float *dst, size_t w, size_t h, int ch,
                 const char *kernel)
{
  BEGIN
  KernelInfo *k;
  im = make_image(src, w, h, ch, -1, ex);
  k = AcquireKernelInfo(kernel, ex);
  if (im && k) { out = ConvolveImage(im, k, ex); rc = export_image(out, dst, w, h, ch, ex); }
  if (k) k = DestroyKernelInfo(k);
  END
}

/* method: MorphologyMethod enum value (morphology.h:69-99) */
__attribute__((visibility("default")))
int ref_morphology(const float *src, float *dst, size_t w, size_t h, int ch,
                   int method, long iterations, const char *kernel)
{
  BEGIN
  KernelInfo *k;
  im = make_image(src, w, h, ch, -1, ex);
  k = AcquireKernelInfo(kernel, ex);
  if (im && k) { out = MorphologyImage(im, (MorphologyMethod) method, (ssize_t) iterations, k, ex);
                 rc = export_image(out, dst, w, h, ch, ex); }
  if (k) k = DestroyKernelInfo(k);
  END
}

/* ResizeImage with "-define" artifacts: defines = "filter:blur=0.8;filter:lobes=2" (set with SetImageArtifact, like the CLI) */
__attribute__((visibility("default")))
int ref_resize_defines(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh, int filter,
                       const char *defines)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) {
    char *copy = AcquireString(defines), *p = copy;
    while (p != (char *) NULL && *p != '\0') {
      char *end = strchr(p, ';'), *eq;
      if (end != (char *) NULL) *end = '\0';
      eq = strchr(p, '=');
      if (eq != (char *) NULL) { *eq = '\0'; (void) SetImageArtifact(im, p, eq + 1); }
      p = end != (char *) NULL ? end + 1 : (char *) NULL;
    }
    copy = DestroyString(copy);
    out = ResizeImage(im, ow, oh, (FilterType) filter, ex);
    rc = export_image(out, dst, ow, oh, ch, ex);
  }
  END
}

/* filter: FilterType enum value (resample.h:32-69) */
__attribute__((visibility("default")))
int ref_resize(const float *src, size_t w, size_t h, int ch,
               float *dst, size_t ow, size_t oh, int filter)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = ResizeImage(im, ow, oh, (FilterType) filter, ex); rc = export_image(out, dst, ow, oh, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_sample(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = SampleImage(im, ow, oh, ex); rc = export_image(out, dst, ow, oh, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_scale(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = ScaleImage(im, ow, oh, ex); rc = export_image(out, dst, ow, oh, ch, ex); }
  END
}

__attribute__((visibility("default")))
int ref_thumbnail(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh)
{
  BEGIN
  im = make_image(src, w, h, ch, -1, ex);
  if (im) { out = ThumbnailImage(im, ow, oh, ex); rc = export_image(out, dst, ow, oh, ch, ex); }
  END
}

/* in place; from/to: ColorspaceType enum values (colorspace.h:27-67) */
__attribute__((visibility("default")))
int ref_colorspace(float *buf, size_t w, size_t h, int ch, int from, int to)
{
  BEGIN
  im = make_image(buf, w, h, ch, from, ex);
  if (im && TransformImageColorspace(im, (ColorspaceType) to, ex) != MagickFalse)
    rc = export_image(im, buf, w, h, ch, ex);
  END
}

/* ... with image settings: "key=value;key=value"; keys with a "color:" prefix are set as artifacts (-define), the
   others as properties (-set), which is how TransformImageColorspace looks them up. */
__attribute__((visibility("default")))
int ref_colorspace_defines(float *buf, size_t w, size_t h, int ch, int from, int to, const char *defines)
{
  BEGIN
  im = make_image(buf, w, h, ch, from, ex);
  if (im) {
    char *copy = AcquireString(defines), *p = copy;
    while (p != (char *) NULL && *p != '\0') {
      char *end = strchr(p, ';'), *eq;
      if (end != (char *) NULL) *end = '\0';
      eq = strchr(p, '=');
      if (eq != (char *) NULL) {
        *eq = '\0';
        if (strncmp(p, "color:", 6) == 0) (void) SetImageArtifact(im, p, eq + 1);
        else (void) SetImageProperty(im, p, eq + 1, ex);
      }
      p = end != (char *) NULL ? end + 1 : (char *) NULL;
    }
    copy = DestroyString(copy);
    if (TransformImageColorspace(im, (ColorspaceType) to, ex) != MagickFalse)
      rc = export_image(im, buf, w, h, ch, ex);
  }
  END
}

/* threshold.c point operators, in place.  op: 0 BilevelImage(threshold), 1 BlackThresholdImage(thresholds),
   2 WhiteThresholdImage(thresholds), 3 ClampImage.  The channel count must survive the call (gray
   images are promoted to sRGB by the black/white operators: rc = -2 then). */
__attribute__((visibility("default")))
int ref_threshold(float *buf, size_t w, size_t h, int ch, int op, double threshold, const char *thresholds)
{
  BEGIN
  MagickBooleanType ok = MagickFalse;
  im = make_image(buf, w, h, ch, -1, ex);
  if (im)
    {
      switch (op)
      {
        case 0: ok = BilevelImage(im, threshold, ex); break;
        case 1: ok = BlackThresholdImage(im, thresholds, ex); break;
        case 2: ok = WhiteThresholdImage(im, thresholds, ex); break;
        case 3: ok = ClampImage(im, ex); break;
        default: break;
      }
      if (ok != MagickFalse) rc = export_image(im, buf, w, h, ch, ex);
    }
  END
}

/* Kernel generation probe: returns the idx-th kernel of the list the string
   parses to (after the same normalisation ConvolveImage's callers apply is NOT
   applied here -- this is AcquireKernelInfo's raw output). */
__attribute__((visibility("default")))
int ref_kernel(const char *kernel, int idx, double *values, size_t max_values,
               size_t *kw, size_t *kh, long *kx, long *ky)
{
  ExceptionInfo *ex = AcquireExceptionInfo();
  KernelInfo *k, *p;
  int rc = -1, i;
  ensure_init();
  k = AcquireKernelInfo(kernel, ex);
  for (p = k, i = 0; p && i < idx; i++) p = p->next;
  if (p && p->width * p->height <= max_values)
    {
      size_t j;
      *kw = p->width; *kh = p->height; *kx = (long) p->x; *ky = (long) p->y;
      for (j = 0; j < p->width * p->height; j++) values[j] = (double) p->values[j];
      rc = 0;
    }
  if (k) k = DestroyKernelInfo(k);
  ex = DestroyExceptionInfo(ex);
  return rc;
}

__attribute__((visibility("default")))
const char *ref_version(void)
{
  size_t v;
  ensure_init();
  return GetMagickVersion(&v);
}

__attribute__((visibility("default")))
const char *ref_features(void)
{
  ensure_init();
  return GetMagickFeatures();
}
