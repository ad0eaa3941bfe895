/*
  shim_harness.c -- end-to-end check of the drop-in boundary (test infrastructure).

  Linked with the UNMODIFIED reference MagickCore (oracle/_ref/libMagickCoreRef.a), the shim and
  libmagickb200 using ld --wrap: every call below enters through ImageMagick's own exported entry
  point, is served by the GPU, and is compared against __real_X (the stock CPU path) on the same
  Image.  Exit code 0 == all within the parity bar and every operator actually hit the GPU path.
*/
#include "MagickCore/studio.h"
#include "MagickCore/MagickCore.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

extern Image *__real_BlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_GaussianBlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_UnsharpMaskImage(const Image *, const double, const double, const double, const double, ExceptionInfo *);
extern Image *__real_MorphologyImage(const Image *, const MorphologyMethod, const ssize_t, const KernelInfo *, ExceptionInfo *);
extern Image *__real_ResizeImage(const Image *, const size_t, const size_t, const FilterType, ExceptionInfo *);
extern MagickBooleanType __real_TransformImageColorspace(Image *, const ColorspaceType, ExceptionInfo *);
extern Image *__real_SampleImage(const Image *, const size_t, const size_t, ExceptionInfo *);
extern Image *__real_ScaleImage(const Image *, const size_t, const size_t, ExceptionInfo *);
extern Image *__real_ThumbnailImage(const Image *, const size_t, const size_t, ExceptionInfo *);
extern Image *__real_MinifyImage(const Image *, ExceptionInfo *);
extern Image *__real_MotionBlurImage(const Image *, const double, const double, const double, ExceptionInfo *);
extern Image *__real_ConvolveImage(const Image *, const KernelInfo *, ExceptionInfo *);
extern Image *__real_ResampleImage(const Image *, const double, const double, const FilterType, ExceptionInfo *);
extern int mb200_device_count(void);
extern Image *__real_SharpenImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_EmbossImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_StatisticImage(const Image *, const StatisticType, const size_t, const size_t, ExceptionInfo *);
extern Image *__real_RotationalBlurImage(const Image *, const double, ExceptionInfo *);
extern Image *__real_AdaptiveBlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_AdaptiveSharpenImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_SelectiveBlurImage(const Image *, const double, const double, const double, ExceptionInfo *);
extern Image *__real_BilateralBlurImage(const Image *, const size_t, const size_t, const double, const double, ExceptionInfo *);
extern MagickBooleanType __real_EqualizeImage(Image *, ExceptionInfo *);
extern Image *__real_EdgeImage(const Image *, const double, ExceptionInfo *);
extern MagickBooleanType __real_BilevelImage(Image *, const double, ExceptionInfo *);
extern MagickBooleanType __real_BlackThresholdImage(Image *, const char *, ExceptionInfo *);
extern MagickBooleanType __real_WhiteThresholdImage(Image *, const char *, ExceptionInfo *);
extern MagickBooleanType __real_ClampImage(Image *, ExceptionInfo *);
extern long B200ShimHits(void), B200ShimFallbacks(void);
extern void B200ShimEnable(int);

static long ulp(float a, float b)
{
  int ia, ib;
  memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
  if (ia < 0) ia = -(ia & 0x7fffffff);
  if (ib < 0) ib = -(ib & 0x7fffffff);
  return labs((long) ia - (long) ib);
}

static long compare(const Image *a, const Image *b, ExceptionInfo *ex)
{
  const Quantum *p, *q;
  size_t i, n;
  long worst = 0;
  if (!a || !b || a->columns != b->columns || a->rows != b->rows) return 1L << 40;
  n = a->columns * a->rows * GetPixelChannels(a);
  p = GetVirtualPixels(a, 0, 0, a->columns, a->rows, ex);
  q = GetVirtualPixels(b, 0, 0, b->columns, b->rows, ex);
  for (i = 0; i < n; i++) { long d = ulp((float) p[i], (float) q[i]); if (d > worst) worst = d; }
  return worst;
}

static Image *noise_image(size_t w, size_t h, MagickBooleanType alpha, ExceptionInfo *ex)
{
  ImageInfo *info = AcquireImageInfo();
  Image *im = AcquireImage(info, ex);
  Quantum *q;
  size_t i, n;
  unsigned long long s = 88172645463325252ULL;
  info = DestroyImageInfo(info);
  (void) SetImageExtent(im, w, h, ex);
  if (alpha) im->alpha_trait = BlendPixelTrait;
  (void) SetImageStorageClass(im, DirectClass, ex);
  (void) SetImageColorspace(im, sRGBColorspace, ex);
  q = GetAuthenticPixels(im, 0, 0, w, h, ex);
  n = w * h * GetPixelChannels(im);
  for (i = 0; i < n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; q[i] = (Quantum) ((s >> 40) * (65535.0 / 16777215.0)); }
  (void) SyncAuthenticPixels(im, ex);
  return im;
}

#define CPU(expr) (B200ShimEnable(0), cpu_tmp = (expr), B200ShimEnable(1), cpu_tmp)
#define CHECK(name, bar, gpu, cpu) do { Image *g_ = (gpu); Image *c_ = (cpu); long d_ = compare(g_, c_, ex); \
  printf("%-34s max ULP %ld (bar %d)%s\n", name, d_, bar, d_ <= bar ? "" : "  FAIL"); if (d_ > bar) failures++; \
  if (g_) DestroyImage(g_); if (c_) DestroyImage(c_); } while (0)

int main(void)
{
  ExceptionInfo *ex;
  Image *rgba, *rgb, *a, *b, *cpu_tmp;
  KernelInfo *k;
  int failures = 0;
  MagickCoreGenesis("shim_harness", MagickFalse);
  ex = AcquireExceptionInfo();
  rgba = noise_image(517, 389, MagickTrue, ex);
  rgb = noise_image(300, 200, MagickFalse, ex);

  CHECK("BlurImage(0,4) RGBA", 1, BlurImage(rgba, 0.0, 4.0, ex), CPU(__real_BlurImage(rgba, 0.0, 4.0, ex)));
  CHECK("BlurImage(0,2) RGB", 1, BlurImage(rgb, 0.0, 2.0, ex), CPU(__real_BlurImage(rgb, 0.0, 2.0, ex)));
  CHECK("GaussianBlurImage(0,1.5) RGBA", 1, GaussianBlurImage(rgba, 0.0, 1.5, ex), CPU(__real_GaussianBlurImage(rgba, 0.0, 1.5, ex)));
  CHECK("UnsharpMaskImage RGBA", 1, UnsharpMaskImage(rgba, 0.0, 2.0, 1.5, 0.02, ex), CPU(__real_UnsharpMaskImage(rgba, 0.0, 2.0, 1.5, 0.02, ex)));
  CHECK("ResizeImage Lanczos 2x down RGBA", 1, ResizeImage(rgba, 258, 194, LanczosFilter, ex), CPU(__real_ResizeImage(rgba, 258, 194, LanczosFilter, ex)));
  CHECK("ResizeImage default up RGB", 1, ResizeImage(rgb, 450, 300, UndefinedFilter, ex), CPU(__real_ResizeImage(rgb, 450, 300, UndefinedFilter, ex)));
  CHECK("MotionBlurImage(0,3,30) RGBA", 1, MotionBlurImage(rgba, 0.0, 3.0, 30.0, ex), CPU(__real_MotionBlurImage(rgba, 0.0, 3.0, 30.0, ex)));
  CHECK("MinifyImage RGBA (Spline 2x)", 1, MinifyImage(rgba, ex), CPU(__real_MinifyImage(rgba, ex)));
  CHECK("ResampleImage 36 dpi RGB (Lanczos)", 1, ResampleImage(rgb, 36.0, 36.0, LanczosFilter, ex), CPU(__real_ResampleImage(rgb, 36.0, 36.0, LanczosFilter, ex)));
  CHECK("SampleImage 517x389 -> 100x77 RGBA", 0, SampleImage(rgba, 100, 77, ex), CPU(__real_SampleImage(rgba, 100, 77, ex)));
  CHECK("ScaleImage 517x389 -> 200x150 RGBA", 0, ScaleImage(rgba, 200, 150, ex), CPU(__real_ScaleImage(rgba, 200, 150, ex)));
  CHECK("ScaleImage 300x200 -> 450x333 RGB", 0, ScaleImage(rgb, 450, 333, ex), CPU(__real_ScaleImage(rgb, 450, 333, ex)));
  CHECK("SharpenImage(0,1) RGBA", 1, SharpenImage(rgba, 0.0, 1.0, ex), CPU(__real_SharpenImage(rgba, 0.0, 1.0, ex)));
  CHECK("EdgeImage(1) RGB", 1, EdgeImage(rgb, 1.0, ex), CPU(__real_EdgeImage(rgb, 1.0, ex)));
  CHECK("StatisticImage Median 3x3 RGBA", 0, StatisticImage(rgba, MedianStatistic, 3, 3, ex), CPU(__real_StatisticImage(rgba, MedianStatistic, 3, 3, ex)));
  CHECK("StatisticImage Mode 3x3 RGB", 0, StatisticImage(rgb, ModeStatistic, 3, 3, ex), CPU(__real_StatisticImage(rgb, ModeStatistic, 3, 3, ex)));
  CHECK("StatisticImage Nonpeak 5x5 RGBA", 0, StatisticImage(rgba, NonpeakStatistic, 5, 5, ex), CPU(__real_StatisticImage(rgba, NonpeakStatistic, 5, 5, ex)));
  CHECK("StatisticImage StdDev 5x3 RGB", 0, StatisticImage(rgb, StandardDeviationStatistic, 5, 3, ex), CPU(__real_StatisticImage(rgb, StandardDeviationStatistic, 5, 3, ex)));
  CHECK("RotationalBlurImage(7) RGBA", 0, RotationalBlurImage(rgba, 7.0, ex), CPU(__real_RotationalBlurImage(rgba, 7.0, ex)));
  CHECK("AdaptiveBlurImage 0x1.5 RGBA", 0, AdaptiveBlurImage(rgba, 0.0, 1.5, ex), CPU(__real_AdaptiveBlurImage(rgba, 0.0, 1.5, ex)));
  CHECK("AdaptiveSharpenImage 0x1 RGB", 0, AdaptiveSharpenImage(rgb, 0.0, 1.0, ex), CPU(__real_AdaptiveSharpenImage(rgb, 0.0, 1.0, ex)));
  CHECK("SelectiveBlurImage 0x1.5 t=10% RGBA", 0, SelectiveBlurImage(rgba, 0.0, 1.5, 6553.5, ex),
        CPU(__real_SelectiveBlurImage(rgba, 0.0, 1.5, 6553.5, ex)));
  CHECK("BilateralBlurImage 5x5 RGB", 0, BilateralBlurImage(rgb, 5, 5, 20.0, 2.0, ex), CPU(__real_BilateralBlurImage(rgb, 5, 5, 20.0, 2.0, ex)));
  k = AcquireKernelInfo("Disk:3", ex);
  CHECK("MorphologyImage Dilate Disk:3", 0, MorphologyImage(rgba, DilateMorphology, 1, k, ex), CPU(__real_MorphologyImage(rgba, DilateMorphology, 1, k, ex)));
  CHECK("MorphologyImage Erode x2 Disk:3", 0, MorphologyImage(rgb, ErodeMorphology, 2, k, ex), CPU(__real_MorphologyImage(rgb, ErodeMorphology, 2, k, ex)));
  CHECK("MorphologyImage Edge Disk:3 RGBA", 0, MorphologyImage(rgba, EdgeMorphology, 1, k, ex), CPU(__real_MorphologyImage(rgba, EdgeMorphology, 1, k, ex)));
  CHECK("MorphologyImage TopHat Disk:3 RGB", 0, MorphologyImage(rgb, TopHatMorphology, 1, k, ex), CPU(__real_MorphologyImage(rgb, TopHatMorphology, 1, k, ex)));
  k = DestroyKernelInfo(k);
  k = AcquireKernelInfo("Corners", ex);
  CHECK("MorphologyImage HitAndMiss Corners", 0, MorphologyImage(rgb, HitAndMissMorphology, 1, k, ex), CPU(__real_MorphologyImage(rgb, HitAndMissMorphology, 1, k, ex)));
  k = DestroyKernelInfo(k);
  k = AcquireKernelInfo("Skeleton", ex);
  CHECK("MorphologyImage Thinning x3 Skeleton", 0, MorphologyImage(rgba, ThinningMorphology, 3, k, ex), CPU(__real_MorphologyImage(rgba, ThinningMorphology, 3, k, ex)));
  k = DestroyKernelInfo(k);
  k = AcquireKernelInfo("Disk:2", ex);
  CHECK("MorphologyImage OpenIntensity Disk:2", 0, MorphologyImage(rgba, OpenIntensityMorphology, 1, k, ex), CPU(__real_MorphologyImage(rgba, OpenIntensityMorphology, 1, k, ex)));
  k = DestroyKernelInfo(k);
  k = AcquireKernelInfo("Euclidean:2", ex);
  CHECK("MorphologyImage IterativeDistance x4", 0, MorphologyImage(rgb, IterativeDistanceMorphology, 4, k, ex), CPU(__real_MorphologyImage(rgb, IterativeDistanceMorphology, 4, k, ex)));
  k = DestroyKernelInfo(k);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (TransformImageColorspace(a, LabColorspace, ex) == MagickFalse || a->colorspace != LabColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, LabColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace sRGB->Lab", 1, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (TransformImageColorspace(a, HSLColorspace, ex) == MagickFalse || a->colorspace != HSLColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, HSLColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace sRGB->HSL", 0, a, b);
  a = CloneImage(rgb, 0, 0, MagickTrue, ex); b = CloneImage(rgb, 0, 0, MagickTrue, ex);
  (void) SetImageColorspace(a, HWBColorspace, ex); (void) SetImageColorspace(b, HWBColorspace, ex);
  if (TransformImageColorspace(a, HSVColorspace, ex) == MagickFalse || a->colorspace != HSVColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, HSVColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace HWB->HSV", 0, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (TransformImageColorspace(a, LuvColorspace, ex) == MagickFalse || a->colorspace != LuvColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, LuvColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace sRGB->Luv", 1, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  (void) SetImageArtifact(a, "color:illuminant", "D50"); (void) SetImageArtifact(b, "color:illuminant", "D50");
  if (TransformImageColorspace(a, LabColorspace, ex) == MagickFalse || a->colorspace != LabColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, LabColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace sRGB->Lab (D50)", 1, a, b);
  a = CloneImage(rgb, 0, 0, MagickTrue, ex); b = CloneImage(rgb, 0, 0, MagickTrue, ex);
  (void) SetImageProperty(a, "reference-white", "700", ex); (void) SetImageProperty(b, "reference-white", "700", ex);
  if (TransformImageColorspace(a, LogColorspace, ex) == MagickFalse || a->colorspace != LogColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, LogColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace sRGB->Log (reference-white 700)", 1, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (TransformImageColorspace(a, YCCColorspace, ex) == MagickFalse || a->colorspace != YCCColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, YCCColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace sRGB->YCC", 0, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  (void) SetImageColorspace(a, YCCColorspace, ex); (void) SetImageColorspace(b, YCCColorspace, ex);
  if (TransformImageColorspace(a, sRGBColorspace, ex) == MagickFalse || a->colorspace != sRGBColorspace) failures++;
  B200ShimEnable(0); (void) __real_TransformImageColorspace(b, sRGBColorspace, ex); B200ShimEnable(1);
  CHECK("TransformImageColorspace YCC->sRGB", 0, a, b);
  CHECK("ResizeImage Jinc 50% RGBA", 1, ResizeImage(rgba, rgba->columns / 2, rgba->rows / 2, JincFilter, ex),
        CPU(__real_ResizeImage(rgba, rgba->columns / 2, rgba->rows / 2, JincFilter, ex)));
  (void) SetImageArtifact(rgba, "filter:blur", "0.85"); (void) SetImageArtifact(rgba, "filter:lobes", "2");
  CHECK("ResizeImage Lanczos 50% blur=.85 lobes=2", 1, ResizeImage(rgba, rgba->columns / 2, rgba->rows / 2, LanczosFilter, ex),
        CPU(__real_ResizeImage(rgba, rgba->columns / 2, rgba->rows / 2, LanczosFilter, ex)));
  (void) DeleteImageArtifact(rgba, "filter:blur"); (void) DeleteImageArtifact(rgba, "filter:lobes");
  CHECK("ResizeImage Kaiser 150% RGB", 1, ResizeImage(rgb, rgb->columns * 3 / 2, rgb->rows * 3 / 2, KaiserFilter, ex),
        CPU(__real_ResizeImage(rgb, rgb->columns * 3 / 2, rgb->rows * 3 / 2, KaiserFilter, ex)));
  /* threshold.c point operators: in place, bit exact */
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (BilevelImage(a, 30000.0, ex) == MagickFalse) failures++;
  B200ShimEnable(0); (void) __real_BilevelImage(b, 30000.0, ex); B200ShimEnable(1);
  CHECK("BilevelImage(30000) RGBA", 0, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (BlackThresholdImage(a, "40%,50%,60%", ex) == MagickFalse) failures++;
  B200ShimEnable(0); (void) __real_BlackThresholdImage(b, "40%,50%,60%", ex); B200ShimEnable(1);
  CHECK("BlackThresholdImage RGBA", 0, a, b);
  a = CloneImage(rgb, 0, 0, MagickTrue, ex); b = CloneImage(rgb, 0, 0, MagickTrue, ex);
  if (WhiteThresholdImage(a, "45000", ex) == MagickFalse) failures++;
  B200ShimEnable(0); (void) __real_WhiteThresholdImage(b, "45000", ex); B200ShimEnable(1);
  CHECK("WhiteThresholdImage RGB", 0, a, b);
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (EqualizeImage(a, ex) == MagickFalse) failures++;
  B200ShimEnable(0); (void) __real_EqualizeImage(b, ex); B200ShimEnable(1);
  CHECK("EqualizeImage RGBA", 0, a, b);
  {
    /* EmbossImage ends in an equalisation of the convolved image: a discontinuous map, so a 1-ULP difference of one
       convolved sample may move many outputs a little; compared in absolute terms like the thumbnail cascade */
    Image *g = EmbossImage(rgb, 0.0, 1.0, ex), *c = CPU(__real_EmbossImage(rgb, 0.0, 1.0, ex));
    double worst = 0.0;
    if (!g || !c) { printf("EmbossImage: FAIL\n"); failures++; }
    else {
      const Quantum *p = GetVirtualPixels(g, 0, 0, g->columns, g->rows, ex), *q = GetVirtualPixels(c, 0, 0, c->columns, c->rows, ex);
      size_t i, n = g->columns * g->rows * GetPixelChannels(g);
      for (i = 0; i < n; i++) { double d = fabs((double) p[i] - (double) q[i]); if (d > worst) worst = d; }
      printf("%-34s max |diff| %.5f Quantum (bar 2.0)%s\n", "EmbossImage(0,1) RGB", worst, worst <= 2.0 ? "" : "  FAIL");
      if (worst > 2.0) failures++;
    }
    if (g) DestroyImage(g);
    if (c) DestroyImage(c);
  }
  a = CloneImage(rgba, 0, 0, MagickTrue, ex); b = CloneImage(rgba, 0, 0, MagickTrue, ex);
  if (ClampImage(a, ex) == MagickFalse) failures++;
  B200ShimEnable(0); (void) __real_ClampImage(b, ex); B200ShimEnable(1);
  CHECK("ClampImage RGBA", 0, a, b);
  {
    /* ThumbnailImage: the cascade is re-issued through the wrapped stages, the metadata comes from the real function.
       A cascade of <= 1 ULP stages is compared in absolute terms (a 1-ULP difference of a bright input sample is many
       ULPs of a dark output sample). */
    const char *const props[] = { "Thumb::URI", "Thumb::Image::Width", "Thumb::Image::Height", "Thumb::Document::Pages",
                                  "Thumb::Size", "software", (const char *) NULL };
    size_t sizes[3][2] = { { 100, 75 }, { 200, 150 }, { 400, 300 } };
    int s, pi;
    for (s = 0; s < 3; s++) {
      Image *g = ThumbnailImage(rgba, sizes[s][0], sizes[s][1], ex);
      Image *c = CPU(__real_ThumbnailImage(rgba, sizes[s][0], sizes[s][1], ex));
      double worst = 0.0;
      if (!g || !c || g->columns != c->columns || g->rows != c->rows) { printf("ThumbnailImage: geometry FAIL\n"); failures++; }
      else {
        const Quantum *p = GetVirtualPixels(g, 0, 0, g->columns, g->rows, ex), *q = GetVirtualPixels(c, 0, 0, c->columns, c->rows, ex);
        size_t i, n = g->columns * g->rows * GetPixelChannels(g);
        for (i = 0; i < n; i++) { double d = fabs((double) p[i] - (double) q[i]); if (d > worst) worst = d; }
        printf("%-34s max |diff| %.5f Quantum (bar 0.02)%s\n", "ThumbnailImage RGBA", worst, worst <= 0.02 ? "" : "  FAIL");
        if (worst > 0.02) failures++;
        if (g->depth != c->depth || g->page.width != c->page.width || g->interlace != c->interlace) { printf("ThumbnailImage: attributes FAIL\n"); failures++; }
        for (pi = 0; props[pi] != (const char *) NULL; pi++) {
          const char *a1 = GetImageProperty(g, props[pi], ex), *b1 = GetImageProperty(c, props[pi], ex);
          if ((a1 == NULL) != (b1 == NULL) || (a1 != NULL && strcmp(a1, b1) != 0)) { printf("ThumbnailImage: property %s FAIL\n", props[pi]); failures++; }
        }
      }
      if (g) DestroyImage(g);
      if (c) DestroyImage(c);
    }
  }
  {
    /* -channel selections: unselected channels carry the Copy trait and are handed through (nearest sample for resize) */
    const long hits0 = B200ShimHits();
    KernelInfo *uk = AcquireKernelInfo("3x3: 1,2,1, 2,4,2, 1,2,1", ex);
    (void) SetPixelChannelMask(rgba, (ChannelType) (RedChannel | BlueChannel | AlphaChannel));
    CHECK("BlurImage(0,2) -channel RBA RGBA", 1, BlurImage(rgba, 0.0, 2.0, ex), CPU(__real_BlurImage(rgba, 0.0, 2.0, ex)));
    CHECK("ResizeImage Lanczos -channel RBA RGBA", 1, ResizeImage(rgba, 258, 194, LanczosFilter, ex), CPU(__real_ResizeImage(rgba, 258, 194, LanczosFilter, ex)));
    CHECK("UnsharpMaskImage -channel RBA RGBA", 1, UnsharpMaskImage(rgba, 0.0, 2.0, 1.5, 0.02, ex), CPU(__real_UnsharpMaskImage(rgba, 0.0, 2.0, 1.5, 0.02, ex)));
    {
      /* alpha unselected: the second pass weights with the unfiltered alpha -- must decline, and still be right */
      const long fb0 = B200ShimFallbacks();
      (void) SetPixelChannelMask(rgba, (ChannelType) (RedChannel | BlueChannel));
      CHECK("BlurImage(0,2) -channel RB (declines)", 0, BlurImage(rgba, 0.0, 2.0, ex), CPU(__real_BlurImage(rgba, 0.0, 2.0, ex)));
      if (mb200_device_count() > 0 && B200ShimFallbacks() <= fb0) { printf("FAIL: alpha-less selection was not declined\n"); failures++; }
    }
    (void) SetPixelChannelMask(rgb, (ChannelType) (RedChannel | BlueChannel));
    CHECK("BlurImage(0,2) -channel RB RGB (no alpha)", 1, BlurImage(rgb, 0.0, 2.0, ex), CPU(__real_BlurImage(rgb, 0.0, 2.0, ex)));
    (void) SetPixelChannelMask(rgb, DefaultChannels);
    (void) SetPixelChannelMask(rgba, (ChannelType) (AlphaChannel | GreenChannel));
    CHECK("MorphologyImage Dilate -channel GA", 0, MorphologyImage(rgba, DilateMorphology, 1, uk, ex), CPU(__real_MorphologyImage(rgba, DilateMorphology, 1, uk, ex)));
    CHECK("ResizeImage up -channel GA RGBA", 1, ResizeImage(rgba, 700, 500, MitchellFilter, ex), CPU(__real_ResizeImage(rgba, 700, 500, MitchellFilter, ex)));
    (void) SetPixelChannelMask(rgba, DefaultChannels);
    if (mb200_device_count() > 0 && B200ShimHits() - hits0 < 5) { printf("FAIL: channel selections did not reach the GPU path\n"); failures++; }
    /* convolve:bias / convolve:scale (morphology.c:4156-4183) are restated by the wrapper */
    (void) SetImageArtifact(rgb, "convolve:bias", "10%");
    (void) SetImageArtifact(rgb, "convolve:scale", "0.05,20%");
    CHECK("ConvolveImage bias 10% scale 0.05,20% RGB", 1, ConvolveImage(rgb, uk, ex), CPU(__real_ConvolveImage(rgb, uk, ex)));
    (void) DeleteImageArtifact(rgb, "convolve:bias");
    (void) DeleteImageArtifact(rgb, "convolve:scale");
    if (mb200_device_count() > 0 && B200ShimHits() - hits0 < 6) { printf("FAIL: convolve artifacts did not reach the GPU path\n"); failures++; }
    uk = DestroyKernelInfo(uk);
  }
  {
    /* `-channel RGB -threshold` on an opaque RGB image leaves every trait at its default, but the reference
       then thresholds each channel on its own value -- the intensity-driven kernel must decline */
    const long fb = B200ShimFallbacks();
    (void) SetPixelChannelMask(rgb, (ChannelType) (RedChannel | GreenChannel | BlueChannel));
    a = CloneImage(rgb, 0, 0, MagickTrue, ex); b = CloneImage(rgb, 0, 0, MagickTrue, ex);
    (void) SetPixelChannelMask(a, (ChannelType) (RedChannel | GreenChannel | BlueChannel));
    (void) SetPixelChannelMask(b, (ChannelType) (RedChannel | GreenChannel | BlueChannel));
    if (BilevelImage(a, 30000.0, ex) == MagickFalse) failures++;
    B200ShimEnable(0); (void) __real_BilevelImage(b, 30000.0, ex); B200ShimEnable(1);
    CHECK("BilevelImage -channel RGB (declines)", 0, a, b);
    if (mb200_device_count() > 0 && B200ShimFallbacks() <= fb) { printf("FAIL: per-channel threshold was not declined\n"); failures++; }
    (void) SetPixelChannelMask(rgb, DefaultChannels);
  }
  {
    /* a declined case must silently take the CPU path: tiled virtual pixels are not eligible */
    long fb = B200ShimFallbacks();
    Image *t = CloneImage(rgb, 0, 0, MagickTrue, ex);
    (void) SetImageVirtualPixelMethod(t, TileVirtualPixelMethod, ex);
    a = BlurImage(t, 0.0, 1.0, ex); b = CPU(__real_BlurImage(t, 0.0, 1.0, ex));
    CHECK("fallback: tile virtual pixels", 0, a, b);
    if (B200ShimFallbacks() <= fb) { printf("expected a fallback\n"); failures++; }
    t = DestroyImage(t);
  }
  printf("gpu hits %ld, cpu fallbacks %ld\n", B200ShimHits(), B200ShimFallbacks());
  if (mb200_device_count() > 0 && B200ShimHits() < 32) { printf("FAIL: operators did not reach the GPU path\n"); failures++; }
  rgba = DestroyImage(rgba); rgb = DestroyImage(rgb);
  ex = DestroyExceptionInfo(ex);
  MagickCoreTerminus();
  return failures ? 1 : 0;
}
