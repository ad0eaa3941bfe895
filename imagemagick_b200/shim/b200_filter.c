/*
  b200_filter.c -- the third way in: a `-process` image filter module for a stock `magick`
  built --with-modules, no relinking of MagickCore:

      magick in.png -process "b200 blur 0x4 resize 50% colorspace Lab" out.png

  The loader (MagickCore/module.c:942 InvokeDynamicImageFilter, static.c:152-161) resolves the tag "b200" to
  `b200Image`, the signature of filters/analyze.c:112.  argv is a flat list of  <operator> <argument>  pairs;
  every image of the list is replaced by the operator's result.  Each operator first tries the B200Accelerate*
  function of magick_b200_shim.c and, when that declines (ineligible image, unsupported variant, no sm_100
  device), calls the stock MagickCore function -- the same contract as the accelerate hooks.

  Operators (arguments are the CLI's own geometry strings, parsed with MagickCore's public parsers):
      blur RxS | gaussian-blur RxS | sharpen RxS | unsharp RxS+gain+threshold | edge R
      resize <geometry> | sample <geometry> | morphology Method:Kernel | colorspace Name
      threshold V[%] | black-threshold T | white-threshold T | clamp -
*/
#include "MagickCore/studio.h"
#include "MagickCore/MagickCore.h"
#include <string.h>

extern Image *B200AccelerateBlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *B200AccelerateGaussianBlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *B200AccelerateUnsharpMaskImage(const Image *, const double, const double, const double, const double,
                                             ExceptionInfo *);
extern Image *B200AccelerateSharpenImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *B200AccelerateEdgeImage(const Image *, const double, ExceptionInfo *);
extern Image *B200AccelerateMorphologyImage(const Image *, const MorphologyMethod, const ssize_t, const KernelInfo *,
                                            ExceptionInfo *);
extern Image *B200AccelerateResizeImage(const Image *, const size_t, const size_t, const FilterType, ExceptionInfo *);
extern Image *B200AccelerateSampleImage(const Image *, const size_t, const size_t, ExceptionInfo *);
extern MagickBooleanType B200AccelerateTransformImageColorspace(Image *, const ColorspaceType, ExceptionInfo *);
extern MagickBooleanType B200AccelerateBilevelImage(Image *, const double, ExceptionInfo *);
extern MagickBooleanType B200AccelerateBlackThresholdImage(Image *, const char *, ExceptionInfo *);
extern MagickBooleanType B200AccelerateWhiteThresholdImage(Image *, const char *, ExceptionInfo *);
extern MagickBooleanType B200AccelerateClampImage(Image *, ExceptionInfo *);

/* One operator on one image.  Returns a NEW image (the caller swaps it in), `image` itself for in-place
   operators that succeeded, or NULL on error / unknown operator. */
static Image *apply_one(Image *image, const char *op, const char *arg, ExceptionInfo *exception)
{
  GeometryInfo g;
  MagickStatusType flags;
  Image *out = (Image *) NULL;

  (void) memset(&g, 0, sizeof(g));
  if (LocaleCompare(op, "blur") == 0 || LocaleCompare(op, "gaussian-blur") == 0 || LocaleCompare(op, "sharpen") == 0) {
    flags = ParseGeometry(arg, &g);
    if ((flags & SigmaValue) == 0) g.sigma = 1.0;                              /* mogrify.c's default */
    if (LocaleCompare(op, "blur") == 0) {
      out = B200AccelerateBlurImage(image, g.rho, g.sigma, exception);
      return out != (Image *) NULL ? out : BlurImage(image, g.rho, g.sigma, exception);
    }
    if (LocaleCompare(op, "gaussian-blur") == 0) {
      out = B200AccelerateGaussianBlurImage(image, g.rho, g.sigma, exception);
      return out != (Image *) NULL ? out : GaussianBlurImage(image, g.rho, g.sigma, exception);
    }
    out = B200AccelerateSharpenImage(image, g.rho, g.sigma, exception);
    return out != (Image *) NULL ? out : SharpenImage(image, g.rho, g.sigma, exception);
  }
  if (LocaleCompare(op, "unsharp") == 0) {
    flags = ParseGeometry(arg, &g);
    if ((flags & SigmaValue) == 0) g.sigma = 1.0;
    if ((flags & XiValue) == 0) g.xi = 1.0;
    if ((flags & PsiValue) == 0) g.psi = 0.05;
    out = B200AccelerateUnsharpMaskImage(image, g.rho, g.sigma, g.xi, g.psi, exception);
    return out != (Image *) NULL ? out : UnsharpMaskImage(image, g.rho, g.sigma, g.xi, g.psi, exception);
  }
  if (LocaleCompare(op, "edge") == 0) {
    (void) ParseGeometry(arg, &g);
    out = B200AccelerateEdgeImage(image, g.rho, exception);
    return out != (Image *) NULL ? out : EdgeImage(image, g.rho, exception);
  }
  if (LocaleCompare(op, "resize") == 0 || LocaleCompare(op, "sample") == 0) {
    ssize_t x = 0, y = 0;
    size_t width = image->columns, height = image->rows;
    (void) ParseMetaGeometry(arg, &x, &y, &width, &height);
    if (width == 0 || height == 0) return (Image *) NULL;
    if (LocaleCompare(op, "sample") == 0) {
      out = B200AccelerateSampleImage(image, width, height, exception);
      return out != (Image *) NULL ? out : SampleImage(image, width, height, exception);
    }
    out = B200AccelerateResizeImage(image, width, height, image->filter, exception);
    return out != (Image *) NULL ? out : ResizeImage(image, width, height, image->filter, exception);
  }
  if (LocaleCompare(op, "morphology") == 0) {                                   /* Method:Kernel, one iteration */
    char method_name[MagickPathExtent];
    const char *colon = strchr(arg, ':');
    ssize_t method;
    KernelInfo *kernel;
    if (colon == (const char *) NULL || (size_t) (colon - arg) >= sizeof(method_name)) return (Image *) NULL;
    (void) memcpy(method_name, arg, (size_t) (colon - arg));
    method_name[colon - arg] = '\0';
    method = ParseCommandOption(MagickMorphologyOptions, MagickFalse, method_name);
    if (method < 0) return (Image *) NULL;
    kernel = AcquireKernelInfo(colon + 1, exception);
    if (kernel == (KernelInfo *) NULL) return (Image *) NULL;
    out = B200AccelerateMorphologyImage(image, (MorphologyMethod) method, 1, kernel, exception);
    if (out == (Image *) NULL) out = MorphologyImage(image, (MorphologyMethod) method, 1, kernel, exception);
    kernel = DestroyKernelInfo(kernel);
    return out;
  }
  if (LocaleCompare(op, "colorspace") == 0) {
    const ssize_t cs = ParseCommandOption(MagickColorspaceOptions, MagickFalse, arg);
    if (cs < 0) return (Image *) NULL;
    if (B200AccelerateTransformImageColorspace(image, (ColorspaceType) cs, exception) != MagickFalse) return image;
    return TransformImageColorspace(image, (ColorspaceType) cs, exception) != MagickFalse ? image : (Image *) NULL;
  }
  if (LocaleCompare(op, "threshold") == 0) {
    double threshold;
    flags = ParseGeometry(arg, &g);
    threshold = g.rho;
    if ((flags & PercentValue) != 0) threshold *= ((double) QuantumRange / 100.0);
    if (B200AccelerateBilevelImage(image, threshold, exception) != MagickFalse) return image;
    return BilevelImage(image, threshold, exception) != MagickFalse ? image : (Image *) NULL;
  }
  if (LocaleCompare(op, "black-threshold") == 0) {
    if (B200AccelerateBlackThresholdImage(image, arg, exception) != MagickFalse) return image;
    return BlackThresholdImage(image, arg, exception) != MagickFalse ? image : (Image *) NULL;
  }
  if (LocaleCompare(op, "white-threshold") == 0) {
    if (B200AccelerateWhiteThresholdImage(image, arg, exception) != MagickFalse) return image;
    return WhiteThresholdImage(image, arg, exception) != MagickFalse ? image : (Image *) NULL;
  }
  if (LocaleCompare(op, "clamp") == 0) {
    if (B200AccelerateClampImage(image, exception) != MagickFalse) return image;
    return ClampImage(image, exception) != MagickFalse ? image : (Image *) NULL;
  }
  (void) ThrowMagickException(exception, GetMagickModule(), OptionError, "UnrecognizedOption", "`b200 %s'", op);
  return (Image *) NULL;
}

ModuleExport size_t b200Image(Image **images, const int argc, const char **argv, ExceptionInfo *exception)
{
  Image *image;
  int i;

  assert(images != (Image **) NULL);
  assert(*images != (Image *) NULL);
  assert((*images)->signature == MagickCoreSignature);
  if ((argc % 2) != 0) {
    (void) ThrowMagickException(exception, GetMagickModule(), OptionError, "MissingArgument", "`b200'");
    return (size_t) 0;
  }
  for (image = GetFirstImageInList(*images); image != (Image *) NULL; ) {
    Image *next = GetNextImageInList(image);
    for (i = 0; i + 1 < argc; i += 2) {
      Image *result = apply_one(image, argv[i], argv[i + 1], exception);
      if (result == (Image *) NULL) return (size_t) 0;
      if (result != image) {
        ReplaceImageInList(&image, result);          /* destroys the old image, `image` now is the result */
        *images = GetFirstImageInList(image);
      }
    }
    image = next;
  }
  return MagickImageFilterSignature;
}
