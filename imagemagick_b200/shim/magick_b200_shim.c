/*
  magick_b200_shim.c -- the drop-in boundary: ImageMagick's own MagickCore entry points of
  the hot path, backed by libmagickb200 (include/magick_b200.h).

  Link an application / the MagickCore library with
      -Wl,--wrap=BlurImage,--wrap=GaussianBlurImage,--wrap=ConvolveImage,--wrap=UnsharpMaskImage,\
          --wrap=MorphologyImage,--wrap=ResizeImage,--wrap=TransformImageColorspace,\
          --wrap=BilevelImage,--wrap=BlackThresholdImage,--wrap=WhiteThresholdImage,--wrap=ClampImage,\
          --wrap=SharpenImage,--wrap=EdgeImage,--wrap=SampleImage,--wrap=ThumbnailImage,--wrap=MinifyImage,--wrap=ResampleImage,--wrap=MotionBlurImage,\
          --wrap=EmbossImage,--wrap=EqualizeImage,--wrap=StatisticImage,--wrap=RotationalBlurImage,--wrap=BilateralBlurImage,--wrap=ScaleImage,--wrap=SelectiveBlurImage,--wrap=AdaptiveBlurImage,--wrap=AdaptiveSharpenImage
  and every caller of those exported functions (effect.c:765/1709/1170/4256, morphology.c:4129,
  resize.c:3761, colorspace.c:1751, threshold.c:805/927/2518/1087) reaches __wrap_X below.  Each wrapper follows the accelerate
  hook contract of effect.c:783-787 / resize.c:3818-3826: try the GPU; if the image is not
  eligible or the GPU path declines (returns NULL / MagickFalse without raising), run the stock
  CPU implementation (__real_X).  B200Accelerate*Image() are the same functions with the
  accelerate-private.h:35-62 signatures, for a build that patches the #if OPENCL call sites.

  Only public MagickCore API is used: GetVirtualPixels / GetAuthenticPixels return the pixel cache
  itself for full-frame requests on memory caches (cache.c:5126-5156), which is exactly the
  interleaved float Quantum layout the C-ABI takes.
*/
#include "MagickCore/studio.h"
#include "MagickCore/MagickCore.h"
#include "MagickCore/string-private.h"          /* StringToDoubleInterval (convolve:bias, morphology.c:4163) */
#include "magick_b200.h"
#include <string.h>

#if !defined(MAGICKCORE_HDRI_SUPPORT) || (MAGICKCORE_QUANTUM_DEPTH != 16)
# error "magick_b200_shim targets the reference's default Q16-HDRI build (Quantum == float)"
#endif

/* ---- eligibility: mirrors checkAccelerateCondition (accelerate.c:110-170) + SURVEY 8b --------- */
/* update_mask == NULL: every channel must carry its default traits (no `-channel` selection).  Otherwise unselected
   channels (Copy trait, pixel.c:6338-6393 SetPixelChannelMask) are accepted and reported: bit c set = channel c is updated. */
static int b200_channels_masked(const Image *image, unsigned *update_mask)
{
  const size_t n = GetPixelChannels(image);
  const MagickBooleanType gray = (image->colorspace == GRAYColorspace) ||
    (image->colorspace == LinearGRAYColorspace) ? MagickTrue : MagickFalse;
  if (image->storage_class != DirectClass) return 0;
  if ((image->channels & (ReadMaskChannel | WriteMaskChannel | CompositeMaskChannel)) != 0) return 0;
  if (image->number_meta_channels != 0) return 0;
  if ((GetImageVirtualPixelMethod(image) != UndefinedVirtualPixelMethod) &&
      (GetImageVirtualPixelMethod(image) != EdgeVirtualPixelMethod)) return 0;
  if (image->progress_monitor != (MagickProgressMonitor) NULL) return 0;
  if (n < 1 || n > 4) return 0;
  if (GetPixelChannelOffset(image, RedPixelChannel) != 0) return 0;
  {
    /* default traits (pixel.c:6356-6381), or Copy for the channels a -channel selection leaves out */
    ssize_t i;
    unsigned mask = 0;
    for (i = 0; i < (ssize_t) n; i++) {
      PixelChannel ch = GetPixelChannelChannel(image, i);
      PixelTrait want = (ch == AlphaPixelChannel || image->alpha_trait == UndefinedPixelTrait)
        ? UpdatePixelTrait : (PixelTrait) (UpdatePixelTrait | BlendPixelTrait);
      const PixelTrait have = GetPixelChannelTraits(image, ch);
      if (have == want) mask |= 1u << i;
      else if (have != CopyPixelTrait || update_mask == (unsigned *) NULL) return 0;
    }
    if (update_mask != (unsigned *) NULL) *update_mask = mask;
    if (mask == 0) return 0;                       /* nothing to compute: let the CPU path clone */
    /* An UNSELECTED alpha channel is copied by every stage of a multi-stage operator (row pass -> column pass, the two
       resize passes), so the later stages weight the colour channels with the ORIGINAL alpha instead of the filtered
       one: the selected channels then depend on the selection and a final restore pass cannot reproduce them.  Such
       selections stay on the CPU path; with alpha selected (or no alpha) the unselected channels feed nothing. */
    if (image->alpha_trait != UndefinedPixelTrait && mask != ((1u << n) - 1u) && (mask >> (n - 1) & 1u) == 0) return 0;
  }
  if (gray != MagickFalse) {
    if (n == 1 && image->alpha_trait == UndefinedPixelTrait) return 1;
    if (n == 2 && image->alpha_trait != UndefinedPixelTrait &&
        GetPixelChannelOffset(image, AlphaPixelChannel) == 1) return 2;
    return 0;
  }
  if (image->colorspace == CMYKColorspace) return 0;
  if (GetPixelChannelOffset(image, GreenPixelChannel) != 1 ||
      GetPixelChannelOffset(image, BluePixelChannel) != 2) return 0;
  if (n == 3 && image->alpha_trait == UndefinedPixelTrait) return 3;
  if (n == 4 && image->alpha_trait != UndefinedPixelTrait &&
      GetPixelChannelOffset(image, AlphaPixelChannel) == 3) return 4;
  return 0;
}

static int b200_channels(const Image *image) { return b200_channels_masked(image, (unsigned *) NULL); }

static MagickBooleanType has_artifact(const Image *image, const char *const *names)
{
  for (; *names != (const char *) NULL; names++)
    if (GetImageArtifact(image, *names) != (const char *) NULL) return MagickTrue;
  return MagickFalse;
}

static const char *const morphology_artifacts[] = { "convolve:bias", "convolve:scale",
  "morphology:compose", "morphology:showKernel", "debug", (const char *) NULL };
/* MorphologyImage itself restates convolve:bias / convolve:scale (B200AccelerateMorphologyImage); the others change the
   control flow (kernel-list merging, stderr output) */
static const char *const morphology_control_artifacts[] = { "morphology:compose", "morphology:showKernel", "debug",
  (const char *) NULL };
static const char *const compose_artifacts[] = { "compose:clamp", "compose:sync", "compose:args",
  "compose:outside-overlay", "compose:clip-to-self", (const char *) NULL };
static const char *const filter_artifacts[] = { "filter:filter", "filter:window", "filter:sigma",
  "filter:alpha", "filter:kaiser-beta", "filter:kaiser-alpha", "filter:lobes", "filter:blur",
  "filter:support", "filter:win-support", "filter:b", "filter:c", "filter:verbose",
  (const char *) NULL };

/* Output image with the reference's conventions (morphology.c:3926-3933, resize.c:3827, :3872). */
static Image *new_result(const Image *image, size_t columns, size_t rows, ExceptionInfo *exception)
{
  Image *out = CloneImage(image, columns, rows, MagickTrue, exception);
  if (out == (Image *) NULL) return out;
  if (SetImageStorageClass(out, DirectClass, exception) == MagickFalse) return DestroyImage(out);
  if (GetPixelChannels(out) != GetPixelChannels(image)) return DestroyImage(out);
  return out;
}

/* The pixel cache itself, for images whose cache is heap memory (cache.c:1273: the OpenCL path makes the same
   MemoryCache test; disk, map and distributed caches decline).  GetPixelCachePixels (cache.c:2310) hands out
   cache_info->pixels without passing one of the cache's sync sites -- what GetAuthenticOpenCLBuffer (cache.c:1259)
   is to the OpenCL path: in hook mode a result that still lives in HBM must not be pulled back just because the
   next operator asks where the pixels are. */
static float *b200_cache_pixels(const Image *image, int channels, ExceptionInfo *exception)
{
  MagickSizeType length = 0;
  void *pixels;
  if (GetImagePixelCacheType(image) != MemoryCache) return (float *) NULL;
  pixels = GetPixelCachePixels((Image *) image, &length, exception);
  if (pixels == (void *) NULL ||
      length != (MagickSizeType) image->columns * image->rows * (size_t) channels * sizeof(Quantum))
    return (float *) NULL;
  return (float *) pixels;
}

/* A GPU attempt reports into an ExceptionInfo of its own: when it declines, the CPU path must start with the
   caller's exception untouched (the "declined without raising" contract of effect.c:783-787). */
#define B200_ATTEMPT_BEGIN ExceptionInfo *attempt = AcquireExceptionInfo()
#define B200_ATTEMPT_END   attempt = DestroyExceptionInfo(attempt)

typedef int (*same_size_op)(const float *, float *, size_t, size_t, int, const void *);

/* src pixels -> new image through `op`; NULL == declined (caller falls back to the CPU).  allow_mask: the operator hands
   Copy-trait channels through from its source, so a -channel selection is served by one extra point pass. */
static Image *run_same_size_masked(const Image *image, same_size_op op, const void *args, int allow_mask,
                                   ExceptionInfo *exception)
{
  unsigned update_mask = 0xfu;
  const int ch = allow_mask ? b200_channels_masked(image, &update_mask) : b200_channels(image);
  const float *p;
  Quantum *q;
  Image *out;
  (void) exception;
  if (ch == 0 || mb200_device_count() <= 0) return (Image *) NULL;
  {
    B200_ATTEMPT_BEGIN;
    out = (Image *) NULL;
    p = b200_cache_pixels(image, ch, attempt);
    if (p != (const float *) NULL) out = new_result(image, image->columns, image->rows, attempt);
    if (out != (Image *) NULL) {
      q = GetAuthenticPixels(out, 0, 0, out->columns, out->rows, attempt);
      if (q == (Quantum *) NULL || b200_cache_pixels(out, ch, attempt) != (float *) q ||
          op(p, (float *) q, image->columns, image->rows, ch, args) != MB200_OK ||
          ((update_mask & ((1u << ch) - 1u)) != ((1u << ch) - 1u) &&
           mb200_restore_channels((float *) q, p, image->columns, image->rows, ch, update_mask) != MB200_OK) ||
          SyncAuthenticPixels(out, attempt) == MagickFalse)
        out = DestroyImage(out);
    }
    B200_ATTEMPT_END;
  }
  if (out != (Image *) NULL) out->type = image->type;
  return out;
}

static Image *run_same_size(const Image *image, same_size_op op, const void *args, ExceptionInfo *exception)
{ return run_same_size_masked(image, op, args, 0, exception); }

/* ---- BlurImage / GaussianBlurImage / UnsharpMaskImage ------------------------------------------ */
typedef struct { double radius, sigma, gain, threshold; } blur_args;
static int op_blur(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_blur_image(s, d, w, h, ch, b->radius, b->sigma); }
static int op_gaussian(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_gaussian_blur_image(s, d, w, h, ch, b->radius, b->sigma); }
static int op_unsharp(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a;
  return mb200_unsharp_mask_image(s, d, w, h, ch, b->radius, b->sigma, b->gain, b->threshold); }

Image *B200AccelerateBlurImage(const Image *image, const double radius, const double sigma,
                               ExceptionInfo *exception)
{
  blur_args a = { radius, sigma, 0.0, 0.0 };
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  return run_same_size_masked(image, op_blur, &a, 1, exception);
}

Image *B200AccelerateGaussianBlurImage(const Image *image, const double radius, const double sigma,
                                       ExceptionInfo *exception)
{
  blur_args a = { radius, sigma, 0.0, 0.0 };
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  return run_same_size_masked(image, op_gaussian, &a, 1, exception);
}

Image *B200AccelerateUnsharpMaskImage(const Image *image, const double radius, const double sigma,
                                      const double gain, const double threshold, ExceptionInfo *exception)
{
  blur_args a = { radius, sigma, gain, threshold };
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  return run_same_size_masked(image, op_unsharp, &a, 1, exception);
}

/* ---- MorphologyImage / ConvolveImage --------------------------------------------------------------- */
static int map_kernel_type(KernelInfoType t)
{
  switch (t) {                       /* only what RotateKernelInfo distinguishes (morphology.c:4281-4305) */
    case BlurKernel: return MB200_BlurKernel;
    case GaussianKernel: case DoGKernel: case LoGKernel: case DiskKernel: case PeaksKernel:
    case LaplacianKernel: case ChebyshevKernel: case ManhattanKernel: case EuclideanKernel:
      return MB200_GaussianKernel;
    case SquareKernel: case DiamondKernel: case PlusKernel: case CrossKernel: return MB200_SquareKernel;
    default: return MB200_UserDefinedKernel;
  }
}

typedef struct { int method; long iterations; const KernelInfo *kernel; double bias; } morph_args;
static int op_morphology(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{
  const morph_args *m = (const morph_args *) a;
  mb200_kernel_info nodes[64];
  const KernelInfo *k;
  int n = 0, rc;
  for (k = m->kernel; k != (const KernelInfo *) NULL; k = k->next) {
    if (n == 64) return MB200_EUNSUPPORTED;
    memset(&nodes[n], 0, sizeof(nodes[n]));
    nodes[n].type = map_kernel_type(k->type);
    nodes[n].width = k->width; nodes[n].height = k->height;
    nodes[n].x = (long) k->x; nodes[n].y = (long) k->y;
    nodes[n].values = (double *) k->values;      /* MagickRealType == double; read-only use */
    nodes[n].minimum = k->minimum; nodes[n].maximum = k->maximum;
    nodes[n].negative_range = k->negative_range; nodes[n].positive_range = k->positive_range;
    nodes[n].angle = k->angle;
    if (n > 0) nodes[n - 1].next = &nodes[n];
    n++;
  }
  if (n == 0) return MB200_EINVAL;
  rc = mb200_morphology_image(s, d, w, h, ch, m->method, m->iterations, &nodes[0], m->bias);
  return rc;
}

Image *B200AccelerateMorphologyImage(const Image *image, const MorphologyMethod method,
                                     const ssize_t iterations, const KernelInfo *kernel,
                                     ExceptionInfo *exception)
{
  morph_args a;
  KernelInfo *scaled = (KernelInfo *) NULL;
  Image *out;
  int allow_mask = 1;
  const char *artifact;
  if (kernel == (const KernelInfo *) NULL || iterations == 0) return (Image *) NULL;
  if (has_artifact(image, morphology_control_artifacts) != MagickFalse) return (Image *) NULL;
  a.bias = 0.0;
  switch (method) {
    case ConvolveMorphology: case CorrelateMorphology:
      /* convolve:bias / convolve:scale apply to these two methods only (morphology.c:4156-4183) */
      artifact = GetImageArtifact(image, "convolve:bias");
      if (artifact != (const char *) NULL) {
        if (IsGeometry(artifact) == MagickFalse) return (Image *) NULL;      /* the reference warns: let it */
        a.bias = StringToDoubleInterval(artifact, (double) QuantumRange + 1.0);
      }
      artifact = GetImageArtifact(image, "convolve:scale");
      if (artifact != (const char *) NULL) {
        if (IsGeometry(artifact) == MagickFalse) return (Image *) NULL;
        scaled = CloneKernelInfo(kernel);
        if (scaled == (KernelInfo *) NULL) return (Image *) NULL;
        ScaleGeometryKernelInfo(scaled, artifact);
      }
      break;
    case ErodeMorphology: case DilateMorphology: case OpenMorphology: case CloseMorphology: case SmoothMorphology: break;
    case EdgeInMorphology: case EdgeOutMorphology: case EdgeMorphology: case TopHatMorphology:
    case BottomHatMorphology:                /* end in CompositeImage(Difference), morphology.c:3995-4012 */
      if (kernel->next != (KernelInfo *) NULL) return (Image *) NULL;
      if (has_artifact(image, compose_artifacts) != MagickFalse) return (Image *) NULL;
      allow_mask = 0;                        /* the composite step treats a channel selection on its own terms */
      break;
    case ErodeIntensityMorphology: case DilateIntensityMorphology: case OpenIntensityMorphology:
    case CloseIntensityMorphology:           /* GetPixelIntensity (pixel.c:2356): the kernel implements the default Rec709 luma */
      if (image->intensity != UndefinedPixelIntensityMethod && image->intensity != Rec709LumaPixelIntensityMethod)
        return (Image *) NULL;
      if (image->colorspace == RGBColorspace || image->colorspace == LinearGRAYColorspace) return (Image *) NULL;
      allow_mask = 0;                        /* the intensity is taken from the masked-out channels as well */
      break;
    case IterativeDistanceMorphology: case ThinningMorphology: case ThickenMorphology: break;
    case HitAndMissMorphology:               /* a kernel list is united with CompositeImage(Lighten), morphology.c:4044 */
      if (kernel->next != (KernelInfo *) NULL) {
        if (has_artifact(image, compose_artifacts) != MagickFalse) return (Image *) NULL;
        allow_mask = 0;
      }
      break;
    default: return (Image *) NULL;          /* Distance / Voronoi: sequential two-pass primitives stay on the CPU */
  }
  a.method = (int) method; a.iterations = (long) iterations; a.kernel = scaled != (KernelInfo *) NULL ? scaled : kernel;
  out = run_same_size_masked(image, op_morphology, &a, allow_mask, exception);
  if (scaled != (KernelInfo *) NULL) scaled = DestroyKernelInfo(scaled);
  return out;
}

/* ---- ResizeImage -------------------------------------------------------------------------------------- */
/* The "filter:*" expert settings (-define filter:blur=0.8 ...), read exactly as AcquireResizeFilter reads them
   (resize.c:999-1226: GetImageArtifact + IsStringTrue / ParseCommandOption / StringToDouble / StringToLong) and handed to
   the library as values.  MagickFalse: a setting the GPU path does not serve (filter:verbose prints the filter). */
static MagickBooleanType b200_filter_options(const Image *image, mb200_filter_options *o)
{
  const char *a;
  (void) memset(o, 0, sizeof(*o));
  if (IsStringTrue(GetImageArtifact(image, "filter:verbose")) != MagickFalse) return MagickFalse;
  o->keep_filter = IsStringTrue(GetImageArtifact(image, "filter:filter")) != MagickFalse ? 1 : 0;
  if (o->keep_filter != 0) {
    /* :1000-1011: a truthy string is parsed as a filter name; a name that parses would replace the filter itself */
    ssize_t option = ParseCommandOption(MagickFilterOptions, MagickFalse, GetImageArtifact(image, "filter:filter"));
    if ((UndefinedFilter < option) && (option < SentinelFilter)) return MagickFalse;
  }
  a = GetImageArtifact(image, "filter:window");
  if (a != (const char *) NULL) {
    ssize_t option = ParseCommandOption(MagickFilterOptions, MagickFalse, a);
    if ((UndefinedFilter < option) && (option < SentinelFilter)) { o->window = (int) option; o->set |= MB200_FO_WINDOW; }
  }
  a = GetImageArtifact(image, "filter:sigma");
  if (a != (const char *) NULL) { o->sigma = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_SIGMA; }
  a = GetImageArtifact(image, "filter:alpha");
  if (a != (const char *) NULL) { o->kaiser_beta = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_KAISER_BETA; }
  a = GetImageArtifact(image, "filter:kaiser-beta");
  if (a != (const char *) NULL) { o->kaiser_beta = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_KAISER_BETA; }
  a = GetImageArtifact(image, "filter:kaiser-alpha");
  if (a != (const char *) NULL) { o->kaiser_beta = StringToDouble(a, (char **) NULL) * 3.14159265358979323846264338327950288419716939937510; o->set |= MB200_FO_KAISER_BETA; }
  a = GetImageArtifact(image, "filter:lobes");
  if (a != (const char *) NULL) { o->lobes = (long) StringToLong(a); o->set |= MB200_FO_LOBES; }
  a = GetImageArtifact(image, "filter:blur");
  if (a != (const char *) NULL) { o->blur = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_BLUR; }
  a = GetImageArtifact(image, "filter:support");
  if (a != (const char *) NULL) { o->support = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_SUPPORT; }
  a = GetImageArtifact(image, "filter:win-support");
  if (a != (const char *) NULL) { o->win_support = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_WIN_SUPPORT; }
  a = GetImageArtifact(image, "filter:b");
  if (a != (const char *) NULL) { o->b = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_B; }
  a = GetImageArtifact(image, "filter:c");
  if (a != (const char *) NULL) { o->c = StringToDouble(a, (char **) NULL); o->set |= MB200_FO_C; }
  if ((o->set & MB200_FO_WINDOW) == 0) o->keep_filter = 0;
  return MagickTrue;
}

Image *B200AccelerateResizeImage(const Image *image, const size_t columns, const size_t rows,
                                 const FilterType filter, ExceptionInfo *exception)
{
  mb200_filter_options fopt;
  unsigned update_mask = 0xfu;
  const int ch = b200_channels_masked(image, &update_mask);
  const float *p;
  Quantum *q;
  Image *out;
  (void) exception;
  if (ch == 0 || columns == 0 || rows == 0 || mb200_device_count() <= 0) return (Image *) NULL;
  if (b200_filter_options(image, &fopt) == MagickFalse) return (Image *) NULL;       /* filter:verbose: stdout belongs to the CPU path */
  if (fopt.set != 0 && (update_mask & ((1u << ch) - 1u)) != ((1u << ch) - 1u)) return (Image *) NULL;
  if (image->storage_class == PseudoClass) return (Image *) NULL;
  {
    B200_ATTEMPT_BEGIN;
    out = (Image *) NULL;
    p = b200_cache_pixels(image, ch, attempt);
    if (p != (const float *) NULL) out = new_result(image, columns, rows, attempt);
    if (out != (Image *) NULL) {
      q = GetAuthenticPixels(out, 0, 0, columns, rows, attempt);
      if (q == (Quantum *) NULL || b200_cache_pixels(out, ch, attempt) != (float *) q ||
          mb200_resize_image_ex(p, image->columns, image->rows, ch, (float *) q, columns, rows, (int) filter,
                                fopt.set != 0 ? &fopt : (const mb200_filter_options *) NULL) != MB200_OK ||
          ((update_mask & ((1u << ch) - 1u)) != ((1u << ch) - 1u) &&
           mb200_resize_copy_channels(p, image->columns, image->rows, ch, (float *) q, columns, rows, (int) filter,
                                      update_mask) != MB200_OK) ||
          SyncAuthenticPixels(out, attempt) == MagickFalse)
        out = DestroyImage(out);
    }
    B200_ATTEMPT_END;
  }
  if (out != (Image *) NULL) out->type = image->type;
  return out;
}

/* ---- SampleImage (resize.c:3907) ------------------------------------------------------------------------------ */
Image *B200AccelerateSampleImage(const Image *image, const size_t columns, const size_t rows, ExceptionInfo *exception)
{
  const int ch = b200_channels(image);
  const float *p;
  Quantum *q;
  Image *out;
  (void) exception;
  if (ch == 0 || columns == 0 || rows == 0 || mb200_device_count() <= 0) return (Image *) NULL;
  if ((columns == image->columns) && (rows == image->rows)) return (Image *) NULL;      /* plain clone: CPU */
  if (GetImageArtifact(image, "sample:offset") != (const char *) NULL) return (Image *) NULL;
  {
    B200_ATTEMPT_BEGIN;
    out = (Image *) NULL;
    p = b200_cache_pixels(image, ch, attempt);
    if (p != (const float *) NULL) out = new_result(image, columns, rows, attempt);
    if (out != (Image *) NULL) {
      q = GetAuthenticPixels(out, 0, 0, columns, rows, attempt);
      if (q == (Quantum *) NULL || b200_cache_pixels(out, ch, attempt) != (float *) q ||
          mb200_sample_image(p, image->columns, image->rows, ch, (float *) q, columns, rows) != MB200_OK ||
          SyncAuthenticPixels(out, attempt) == MagickFalse)
        out = DestroyImage(out);
    }
    B200_ATTEMPT_END;
  }
  if (out != (Image *) NULL) out->type = image->type;
  return out;
}

/* ---- ScaleImage (resize.c:4106) --------------------------------------------------------------------------------------------- */
Image *B200AccelerateScaleImage(const Image *image, const size_t columns, const size_t rows, ExceptionInfo *exception)
{
  const int ch = b200_channels(image);
  const float *p;
  Quantum *q;
  Image *out;
  (void) exception;
  if (ch == 0 || columns == 0 || rows == 0 || mb200_device_count() <= 0) return (Image *) NULL;
  if ((columns == image->columns) && (rows == image->rows)) return (Image *) NULL;      /* plain clone: CPU */
  {
    B200_ATTEMPT_BEGIN;
    out = (Image *) NULL;
    p = b200_cache_pixels(image, ch, attempt);
    if (p != (const float *) NULL) out = new_result(image, columns, rows, attempt);
    if (out != (Image *) NULL) {
      q = GetAuthenticPixels(out, 0, 0, columns, rows, attempt);
      if (q == (Quantum *) NULL || b200_cache_pixels(out, ch, attempt) != (float *) q ||
          mb200_scale_image(p, image->columns, image->rows, ch, (float *) q, columns, rows) != MB200_OK ||
          SyncAuthenticPixels(out, attempt) == MagickFalse)
        out = DestroyImage(out);
    }
    B200_ATTEMPT_END;
  }
  if (out != (Image *) NULL) out->type = image->type;
  return out;
}

/* ---- TransformImageColorspace (in place) ------------------------------------------------------------ */
static int map_colorspace(ColorspaceType c)
{
  switch (c) {
    case sRGBColorspace: return MB200_sRGBColorspace;
    case RGBColorspace: return MB200_RGBColorspace;
    case LabColorspace: return MB200_LabColorspace;
    case XYZColorspace: return MB200_XYZColorspace;
    case CMYColorspace: return MB200_CMYColorspace;
    case OHTAColorspace: return MB200_OHTAColorspace;
    case Rec601YCbCrColorspace: return MB200_Rec601YCbCrColorspace;
    case Rec709YCbCrColorspace: return MB200_Rec709YCbCrColorspace;
    case YCbCrColorspace: return MB200_YCbCrColorspace;
    case YDbDrColorspace: return MB200_YDbDrColorspace;
    case YIQColorspace: return MB200_YIQColorspace;
    case YPbPrColorspace: return MB200_YPbPrColorspace;
    case YUVColorspace: return MB200_YUVColorspace;
    case LCHColorspace: return MB200_LCHColorspace;
    case LCHabColorspace: return MB200_LCHabColorspace;
    case LCHuvColorspace: return MB200_LCHuvColorspace;
    case LogColorspace: return MB200_LogColorspace;
    case YCCColorspace: return MB200_YCCColorspace;
    case JzazbzColorspace: return MB200_JzazbzColorspace;
    case OklabColorspace: return MB200_OklabColorspace;
    case OklchColorspace: return MB200_OklchColorspace;
    case LMSColorspace: return MB200_LMSColorspace;
    case LuvColorspace: return MB200_LuvColorspace;
    case xyYColorspace: return MB200_xyYColorspace;
    case DisplayP3Colorspace: return MB200_DisplayP3Colorspace;
    case Adobe98Colorspace: return MB200_Adobe98Colorspace;
    case ProPhotoColorspace: return MB200_ProPhotoColorspace;
    case CAT02LMSColorspace: return MB200_CAT02LMSColorspace;
    case HCLColorspace: return MB200_HCLColorspace;
    case HCLpColorspace: return MB200_HCLpColorspace;
    case HSBColorspace: return MB200_HSBColorspace;
    case HSIColorspace: return MB200_HSIColorspace;
    case HSLColorspace: return MB200_HSLColorspace;
    case HSVColorspace: return MB200_HSVColorspace;
    case HWBColorspace: return MB200_HWBColorspace;
    default: return -1;
  }
}

/* The image settings sRGBTransformImage / TransformsRGBImage read (colorspace.c:761-773, :996, :1081-1095), parsed with
   the reference's own functions.  False: leave the image to the CPU path. */
static MagickBooleanType b200_colorspace_options(const Image *image, mb200_colorspace_options *o, ExceptionInfo *exception)
{
  const char *value;
  (void) memset(o, 0, sizeof(*o));
  value = GetImageArtifact(image, "color:illuminant");
  if (value != (const char *) NULL) {
    const ssize_t type = ParseCommandOption(MagickIlluminantOptions, MagickFalse, value);
    o->illuminant = type < 0 ? (int) UndefinedIlluminant : (int) type;      /* :769-772 */
    o->set |= MB200_CO_ILLUMINANT;
  }
  value = GetImageProperty(image, "white-luminance", exception);
  if (value != (const char *) NULL) { o->white_luminance = StringToDouble(value, (char **) NULL); o->set |= MB200_CO_WHITE_LUMINANCE; }
  if (GetImageProperty(image, "gamma", exception) != (const char *) NULL) return MagickFalse;   /* unreachable through SetImageProperty */
  value = GetImageProperty(image, "film-gamma", exception);
  if (value != (const char *) NULL) { o->film_gamma = StringToDouble(value, (char **) NULL); o->set |= MB200_CO_FILM_GAMMA; }
  value = GetImageProperty(image, "reference-black", exception);
  if (value != (const char *) NULL) { o->reference_black = StringToDouble(value, (char **) NULL); o->set |= MB200_CO_REFERENCE_BLACK; }
  value = GetImageProperty(image, "reference-white", exception);
  if (value != (const char *) NULL) { o->reference_white = StringToDouble(value, (char **) NULL); o->set |= MB200_CO_REFERENCE_WHITE; }
  /* the reference's table loops index past MaxMap for values outside the 10-bit scale */
  if ((o->set & MB200_CO_REFERENCE_BLACK) != 0 && !(o->reference_black >= 0.0 && o->reference_black <= 1024.0)) return MagickFalse;
  if ((o->set & MB200_CO_REFERENCE_WHITE) != 0 && !(o->reference_white >= 0.0 && o->reference_white <= 1024.0)) return MagickFalse;
  return MagickTrue;
}

MagickBooleanType B200AccelerateTransformImageColorspace(Image *image, const ColorspaceType colorspace,
                                                         ExceptionInfo *exception)
{
  const int from = map_colorspace(image->colorspace), to = map_colorspace(colorspace);
  ColorspaceType saved = image->colorspace;
  mb200_colorspace_options copt;
  Quantum *q;
  int ch;
  if (from < 0 || to < 0 || from == to || mb200_device_count() <= 0) return MagickFalse;
  if (b200_colorspace_options(image, &copt, exception) == MagickFalse) return MagickFalse;
  /* the channel layout test is colourspace-agnostic for 3/4-channel images */
  image->colorspace = sRGBColorspace;
  ch = b200_channels(image);
  image->colorspace = saved;
  if (ch != 3 && ch != 4) return MagickFalse;
  {
    MagickBooleanType ok = MagickFalse;
    B200_ATTEMPT_BEGIN;
    /* GetAuthenticPixels un-shares a copy-on-write cache (GetImagePixelCache, cache.c:1715) before it is written */
    q = GetAuthenticPixels(image, 0, 0, image->columns, image->rows, attempt);
    if (q != (Quantum *) NULL && b200_cache_pixels(image, ch, attempt) == (float *) q &&
        mb200_transform_colorspace_ex((float *) q, image->columns, image->rows, ch, from, to,
                                      copt.set != 0 ? &copt : (const mb200_colorspace_options *) NULL) == MB200_OK &&
        SyncAuthenticPixels(image, attempt) != MagickFalse)
      ok = MagickTrue;
    B200_ATTEMPT_END;
    if (ok == MagickFalse) return MagickFalse;    /* nothing was written back: the CPU path starts from the same pixels */
  }
  (void) DeleteImageProfile(image, "icc");                 /* colorspace.c:1763-1764 */
  (void) DeleteImageProfile(image, "icm");
  return SetImageColorspace(image, colorspace, exception);  /* colorspace.c:1051 */
}

/* ---- threshold.c point operators (in place, bit exact) ---------------------------------------------- */
static MagickBooleanType default_intensity(const Image *image)
{
  return ((image->intensity == UndefinedPixelIntensityMethod) ||
          (image->intensity == Rec709LumaPixelIntensityMethod)) ? MagickTrue : MagickFalse;
}

/* op: 0 BilevelImage, 1 BlackThresholdImage, 2 WhiteThresholdImage, 3 ClampImage */
static MagickBooleanType run_threshold(Image *image, int op, double threshold, const char *thresholds,
                                       ExceptionInfo *exception)
{
  Quantum *q;
  int ch, rc, cs;
  if (mb200_device_count() <= 0) return MagickFalse;
  /* With a channel mask the reference thresholds every selected channel on ITS OWN value (threshold.c:871, :1032,
     :2623); the kernel implements the default, intensity-driven form only.  `-channel RGB` on an opaque RGB image
     leaves every trait at its default, so the trait test of b200_channels() cannot see the mask. */
  if (op != 3 && image->channel_mask != AllChannels) return MagickFalse;
  if (op != 3 && default_intensity(image) == MagickFalse) return MagickFalse;
  if (image->colorspace == LinearGRAYColorspace) return MagickFalse;       /* intensity would need EncodePixelGamma */
  if ((op == 1 || op == 2) && thresholds == (const char *) NULL) return MagickFalse;
  ch = b200_channels(image);
  if (ch == 0) return MagickFalse;
  if ((op == 1 || op == 2) && (ch < 3 || image->colorspace == RGBColorspace)) return MagickFalse;  /* :949, pixel.c:2421 */
  if (op == 0 && image->colorspace != GRAYColorspace && image->colorspace != LinearGRAYColorspace)
    (void) SetImageColorspace(image, sRGBColorspace, exception);            /* threshold.c:827 */
  if (GetPixelChannels(image) != (size_t) ch) return MagickFalse;
  cs = image->colorspace == RGBColorspace ? MB200_RGBColorspace : MB200_sRGBColorspace;
  {
    MagickBooleanType ok = MagickFalse;
    B200_ATTEMPT_BEGIN;
    q = GetAuthenticPixels(image, 0, 0, image->columns, image->rows, attempt);
    rc = MB200_EINVAL;
    if (q != (Quantum *) NULL && b200_cache_pixels(image, ch, attempt) == (float *) q)
      switch (op) {
        case 0: rc = mb200_bilevel_image((float *) q, image->columns, image->rows, ch, threshold); break;
        case 1: rc = mb200_black_threshold_image((float *) q, image->columns, image->rows, ch, cs, thresholds); break;
        case 2: rc = mb200_white_threshold_image((float *) q, image->columns, image->rows, ch, cs, thresholds); break;
        default: rc = mb200_clamp_image((float *) q, image->columns, image->rows, ch); break;
      }
    /* on failure nothing was written back: the CPU path starts from the same pixels */
    if (rc == MB200_OK && SyncAuthenticPixels(image, attempt) != MagickFalse) ok = MagickTrue;
    B200_ATTEMPT_END;
    return ok;
  }
}

MagickBooleanType B200AccelerateBilevelImage(Image *image, const double threshold, ExceptionInfo *exception)
{ return run_threshold(image, 0, threshold, (const char *) NULL, exception); }
MagickBooleanType B200AccelerateBlackThresholdImage(Image *image, const char *thresholds, ExceptionInfo *exception)
{ return run_threshold(image, 1, 0.0, thresholds, exception); }
MagickBooleanType B200AccelerateWhiteThresholdImage(Image *image, const char *thresholds, ExceptionInfo *exception)
{ return run_threshold(image, 2, 0.0, thresholds, exception); }
MagickBooleanType B200AccelerateClampImage(Image *image, ExceptionInfo *exception)
{ return run_threshold(image, 3, 0.0, (const char *) NULL, exception); }

/* ---- SharpenImage / EdgeImage (effect.c:3991, :1520): inline kernel + ConvolveImage -------------------------- */
static int op_sharpen(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_sharpen_image(s, d, w, h, ch, b->radius, b->sigma); }
static int op_edge(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_edge_image(s, d, w, h, ch, b->radius); }

Image *B200AccelerateSharpenImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  blur_args a;
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  a.radius = radius; a.sigma = sigma; a.gain = 0.0; a.threshold = 0.0;
  return run_same_size_masked(image, op_sharpen, &a, 1, exception);
}

Image *B200AccelerateEdgeImage(const Image *image, const double radius, ExceptionInfo *exception)
{
  blur_args a;
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  a.radius = radius; a.sigma = 0.0; a.gain = 0.0; a.threshold = 0.0;
  return run_same_size_masked(image, op_edge, &a, 1, exception);
}

/* ---- EqualizeImage (enhance.c:2040; the reference's hook is AccelerateEqualizeImage) and EmbossImage (effect.c:1600) ----- */
static MagickBooleanType equalize_eligible(const Image *image)
{
  /* the histogram is indexed by GetPixelIntensity (pixel.c:2356): only its default method on non-linear images is
     restated; linear RGB / gray would go through EncodePixelGamma */
  if (default_intensity(image) == MagickFalse) return MagickFalse;
  if (image->colorspace == RGBColorspace || image->colorspace == LinearGRAYColorspace) return MagickFalse;
  return MagickTrue;
}

MagickBooleanType B200AccelerateEqualizeImage(Image *image, ExceptionInfo *exception)
{
  const int ch = b200_channels(image);
  const int sync = (image->channel_mask & SyncChannels) != 0 ? 1 : 0;
  MagickBooleanType ok = MagickFalse;
  Quantum *q;
  (void) exception;
  if (ch == 0 || mb200_device_count() <= 0 || equalize_eligible(image) == MagickFalse) return MagickFalse;
  {
    B200_ATTEMPT_BEGIN;
    q = GetAuthenticPixels(image, 0, 0, image->columns, image->rows, attempt);
    if (q != (Quantum *) NULL && b200_cache_pixels(image, ch, attempt) == (float *) q &&
        mb200_equalize_image((float *) q, image->columns, image->rows, ch, sync) == MB200_OK &&
        SyncAuthenticPixels(image, attempt) != MagickFalse)
      ok = MagickTrue;
    B200_ATTEMPT_END;
  }
  return ok;
}

static int op_emboss(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_emboss_image(s, d, w, h, ch, b->radius, b->sigma); }

Image *B200AccelerateEmbossImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  blur_args a;
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  if (equalize_eligible(image) == MagickFalse || (image->channel_mask & SyncChannels) == 0) return (Image *) NULL;
  a.radius = radius; a.sigma = sigma; a.gain = 0.0; a.threshold = 0.0;
  return run_same_size(image, op_emboss, &a, exception);
}

/* ---- StatisticImage (statistic.c:2918), RotationalBlurImage (effect.c:3129), BilateralBlurImage (effect.c:821) -------------- */
typedef struct { int type; size_t width, height; double a, b, c; } stencil_args;
static int op_statistic(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const stencil_args *t = (const stencil_args *) a; return mb200_statistic_image(s, d, w, h, ch, t->type, t->width, t->height); }
static int op_rotational(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const stencil_args *t = (const stencil_args *) a; return mb200_rotational_blur_image(s, d, w, h, ch, t->a); }
static int op_bilateral(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const stencil_args *t = (const stencil_args *) a; return mb200_bilateral_blur_image(s, d, w, h, ch, t->width, t->height, t->a, t->b); }

Image *B200AccelerateStatisticImage(const Image *image, const StatisticType type, const size_t width, const size_t height,
                                    ExceptionInfo *exception)
{
  stencil_args a;
  switch (type) {
    case GradientStatistic: case MaximumStatistic: case MeanStatistic: case MedianStatistic: case MinimumStatistic:
    case ModeStatistic: case NonpeakStatistic:
    case RootMeanSquareStatistic: case StandardDeviationStatistic: case ContrastStatistic: break;
    default: return (Image *) NULL;
  }
  a.type = (int) type; a.width = width; a.height = height; a.a = a.b = a.c = 0.0;
  return run_same_size(image, op_statistic, &a, exception);
}

Image *B200AccelerateRotationalBlurImage(const Image *image, const double angle, ExceptionInfo *exception)
{
  stencil_args a;
  a.type = 0; a.width = a.height = 0; a.a = angle; a.b = a.c = 0.0;
  return run_same_size(image, op_rotational, &a, exception);
}

Image *B200AccelerateBilateralBlurImage(const Image *image, const size_t width, const size_t height,
                                        const double intensity_sigma, const double spatial_sigma, ExceptionInfo *exception)
{
  stencil_args a;
  if (equalize_eligible(image) == MagickFalse) return (Image *) NULL;        /* tonal weights use GetPixelIntensity */
  a.type = 0; a.width = width; a.height = height; a.a = intensity_sigma; a.b = spatial_sigma; a.c = 0.0;
  return run_same_size(image, op_bilateral, &a, exception);
}

/* AdaptiveBlurImage (effect.c:128) / AdaptiveSharpenImage (:447).  AutoLevelImage takes its all-channels branch only for the
   default mask (histogram.c:942) and EdgeImage / BlurImage read the convolve artifacts: anything else declines. */
static int op_adaptive_blur(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_adaptive_blur_image(s, d, w, h, ch, b->radius, b->sigma); }
static int op_adaptive_sharpen(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const blur_args *b = (const blur_args *) a; return mb200_adaptive_sharpen_image(s, d, w, h, ch, b->radius, b->sigma); }

static Image *adaptive(const Image *image, same_size_op op, const double radius, const double sigma, ExceptionInfo *exception)
{
  blur_args a;
  if (has_artifact(image, morphology_artifacts) != MagickFalse) return (Image *) NULL;
  if (equalize_eligible(image) == MagickFalse || image->channel_mask != AllChannels) return (Image *) NULL;
  a.radius = radius; a.sigma = sigma; a.gain = 0.0; a.threshold = 0.0;
  return run_same_size(image, op, &a, exception);
}

Image *B200AccelerateAdaptiveBlurImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{ return adaptive(image, op_adaptive_blur, radius, sigma, exception); }
Image *B200AccelerateAdaptiveSharpenImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{ return adaptive(image, op_adaptive_sharpen, radius, sigma, exception); }

/* SelectiveBlurImage (effect.c:3406): one stage, so channel selections are exact (unselected channels copy the centre) */
static int op_selective(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const stencil_args *t = (const stencil_args *) a; return mb200_selective_blur_image(s, d, w, h, ch, t->a, t->b, t->c); }

Image *B200AccelerateSelectiveBlurImage(const Image *image, const double radius, const double sigma, const double threshold,
                                        ExceptionInfo *exception)
{
  stencil_args a;
  if (equalize_eligible(image) == MagickFalse) return (Image *) NULL;        /* contrast uses GetPixelIntensity */
  a.type = 0; a.width = a.height = 0; a.a = radius; a.b = sigma; a.c = threshold;
  return run_same_size_masked(image, op_selective, &a, 1, exception);
}

/* ---- ld --wrap entry points ------------------------------------------------------------------------------ */
extern Image *__real_BlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_GaussianBlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_ConvolveImage(const Image *, const KernelInfo *, ExceptionInfo *);
extern Image *__real_UnsharpMaskImage(const Image *, const double, const double, const double, const double,
                                      ExceptionInfo *);
extern Image *__real_MorphologyImage(const Image *, const MorphologyMethod, const ssize_t, const KernelInfo *,
                                     ExceptionInfo *);
extern Image *__real_ResizeImage(const Image *, const size_t, const size_t, const FilterType, ExceptionInfo *);
extern MagickBooleanType __real_TransformImageColorspace(Image *, const ColorspaceType, ExceptionInfo *);
extern Image *__real_SampleImage(const Image *, const size_t, const size_t, ExceptionInfo *);
extern Image *__real_SharpenImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_EdgeImage(const Image *, const double, ExceptionInfo *);
extern MagickBooleanType __real_BilevelImage(Image *, const double, ExceptionInfo *);
extern MagickBooleanType __real_BlackThresholdImage(Image *, const char *, ExceptionInfo *);
extern MagickBooleanType __real_WhiteThresholdImage(Image *, const char *, ExceptionInfo *);
extern MagickBooleanType __real_ClampImage(Image *, ExceptionInfo *);

static long b200_hits = 0, b200_fallbacks = 0;          /* updated with atomic adds: the entry points are re-entrant */
static int b200_enabled = -1;       /* -1: not yet read from the environment */
#define B200_COUNT(var) ((void) __atomic_fetch_add(&(var), 1L, __ATOMIC_RELAXED))
long B200ShimHits(void) { return __atomic_load_n(&b200_hits, __ATOMIC_RELAXED); }
long B200ShimFallbacks(void) { return __atomic_load_n(&b200_fallbacks, __ATOMIC_RELAXED); }
/* Runtime switch (also: environment MAGICK_B200_DISABLE=1), e.g. to A/B against the CPU path. */
void B200ShimEnable(int on) { b200_enabled = on ? 1 : 0; }
static int b200_on(void)
{
  if (b200_enabled < 0) {
    const char *e = getenv("MAGICK_B200_DISABLE");
    b200_enabled = (e != (const char *) NULL && *e != '\0' && *e != '0') ? 0 : 1;
  }
  return b200_enabled;
}
#define TRY(expr) do { if (b200_on()) { Image *r_ = (expr); if (r_ != (Image *) NULL) { B200_COUNT(b200_hits); return r_; } B200_COUNT(b200_fallbacks); } } while (0)

Image *__wrap_BlurImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  TRY(B200AccelerateBlurImage(image, radius, sigma, exception));
  return __real_BlurImage(image, radius, sigma, exception);
}

Image *__wrap_GaussianBlurImage(const Image *image, const double radius, const double sigma,
                                ExceptionInfo *exception)
{
  TRY(B200AccelerateGaussianBlurImage(image, radius, sigma, exception));
  return __real_GaussianBlurImage(image, radius, sigma, exception);
}

Image *__wrap_ConvolveImage(const Image *image, const KernelInfo *kernel, ExceptionInfo *exception)
{
  TRY(B200AccelerateMorphologyImage(image, ConvolveMorphology, 1, kernel, exception));
  return __real_ConvolveImage(image, kernel, exception);
}

Image *__wrap_UnsharpMaskImage(const Image *image, const double radius, const double sigma, const double gain,
                               const double threshold, ExceptionInfo *exception)
{
  TRY(B200AccelerateUnsharpMaskImage(image, radius, sigma, gain, threshold, exception));
  return __real_UnsharpMaskImage(image, radius, sigma, gain, threshold, exception);
}

Image *__wrap_MorphologyImage(const Image *image, const MorphologyMethod method, const ssize_t iterations,
                              const KernelInfo *kernel, ExceptionInfo *exception)
{
  TRY(B200AccelerateMorphologyImage(image, method, iterations, kernel, exception));
  return __real_MorphologyImage(image, method, iterations, kernel, exception);
}

Image *__wrap_ResizeImage(const Image *image, const size_t columns, const size_t rows, const FilterType filter,
                          ExceptionInfo *exception)
{
  TRY(B200AccelerateResizeImage(image, columns, rows, filter, exception));
  return __real_ResizeImage(image, columns, rows, filter, exception);
}

MagickBooleanType __wrap_TransformImageColorspace(Image *image, const ColorspaceType colorspace,
                                                  ExceptionInfo *exception)
{
  if (b200_on()) {
    if (B200AccelerateTransformImageColorspace(image, colorspace, exception) != MagickFalse) {
      B200_COUNT(b200_hits);
      return MagickTrue;
    }
    B200_COUNT(b200_fallbacks);
  }
  return __real_TransformImageColorspace(image, colorspace, exception);
}

#define TRY_BOOL(expr) do { if (b200_on()) { if ((expr) != MagickFalse) { B200_COUNT(b200_hits); return MagickTrue; } B200_COUNT(b200_fallbacks); } } while (0)

MagickBooleanType __wrap_BilevelImage(Image *image, const double threshold, ExceptionInfo *exception)
{
  TRY_BOOL(B200AccelerateBilevelImage(image, threshold, exception));
  return __real_BilevelImage(image, threshold, exception);
}

MagickBooleanType __wrap_BlackThresholdImage(Image *image, const char *thresholds, ExceptionInfo *exception)
{
  TRY_BOOL(B200AccelerateBlackThresholdImage(image, thresholds, exception));
  return __real_BlackThresholdImage(image, thresholds, exception);
}

MagickBooleanType __wrap_WhiteThresholdImage(Image *image, const char *thresholds, ExceptionInfo *exception)
{
  TRY_BOOL(B200AccelerateWhiteThresholdImage(image, thresholds, exception));
  return __real_WhiteThresholdImage(image, thresholds, exception);
}

MagickBooleanType __wrap_ClampImage(Image *image, ExceptionInfo *exception)
{
  TRY_BOOL(B200AccelerateClampImage(image, exception));
  return __real_ClampImage(image, exception);
}

Image *__wrap_SharpenImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  TRY(B200AccelerateSharpenImage(image, radius, sigma, exception));
  return __real_SharpenImage(image, radius, sigma, exception);
}

Image *__wrap_EdgeImage(const Image *image, const double radius, ExceptionInfo *exception)
{
  TRY(B200AccelerateEdgeImage(image, radius, exception));
  return __real_EdgeImage(image, radius, exception);
}

extern Image *__real_StatisticImage(const Image *, const StatisticType, const size_t, const size_t, ExceptionInfo *);
extern Image *__real_RotationalBlurImage(const Image *, const double, ExceptionInfo *);
extern Image *__real_BilateralBlurImage(const Image *, const size_t, const size_t, const double, const double, ExceptionInfo *);

Image *__wrap_StatisticImage(const Image *image, const StatisticType type, const size_t width, const size_t height,
                             ExceptionInfo *exception)
{
  TRY(B200AccelerateStatisticImage(image, type, width, height, exception));
  return __real_StatisticImage(image, type, width, height, exception);
}

Image *__wrap_RotationalBlurImage(const Image *image, const double angle, ExceptionInfo *exception)
{
  TRY(B200AccelerateRotationalBlurImage(image, angle, exception));
  return __real_RotationalBlurImage(image, angle, exception);
}

Image *__wrap_BilateralBlurImage(const Image *image, const size_t width, const size_t height, const double intensity_sigma,
                                 const double spatial_sigma, ExceptionInfo *exception)
{
  TRY(B200AccelerateBilateralBlurImage(image, width, height, intensity_sigma, spatial_sigma, exception));
  return __real_BilateralBlurImage(image, width, height, intensity_sigma, spatial_sigma, exception);
}

extern Image *__real_AdaptiveBlurImage(const Image *, const double, const double, ExceptionInfo *);
extern Image *__real_AdaptiveSharpenImage(const Image *, const double, const double, ExceptionInfo *);
Image *__wrap_AdaptiveBlurImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  TRY(B200AccelerateAdaptiveBlurImage(image, radius, sigma, exception));
  return __real_AdaptiveBlurImage(image, radius, sigma, exception);
}
Image *__wrap_AdaptiveSharpenImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  TRY(B200AccelerateAdaptiveSharpenImage(image, radius, sigma, exception));
  return __real_AdaptiveSharpenImage(image, radius, sigma, exception);
}

extern Image *__real_SelectiveBlurImage(const Image *, const double, const double, const double, ExceptionInfo *);
Image *__wrap_SelectiveBlurImage(const Image *image, const double radius, const double sigma, const double threshold,
                                 ExceptionInfo *exception)
{
  TRY(B200AccelerateSelectiveBlurImage(image, radius, sigma, threshold, exception));
  return __real_SelectiveBlurImage(image, radius, sigma, threshold, exception);
}

extern Image *__real_EmbossImage(const Image *, const double, const double, ExceptionInfo *);
extern MagickBooleanType __real_EqualizeImage(Image *, ExceptionInfo *);

Image *__wrap_EmbossImage(const Image *image, const double radius, const double sigma, ExceptionInfo *exception)
{
  TRY(B200AccelerateEmbossImage(image, radius, sigma, exception));
  return __real_EmbossImage(image, radius, sigma, exception);
}

MagickBooleanType __wrap_EqualizeImage(Image *image, ExceptionInfo *exception)
{
  TRY_BOOL(B200AccelerateEqualizeImage(image, exception));
  return __real_EqualizeImage(image, exception);
}

extern Image *__real_ScaleImage(const Image *, const size_t, const size_t, ExceptionInfo *);
Image *__wrap_ScaleImage(const Image *image, const size_t columns, const size_t rows, ExceptionInfo *exception)
{
  TRY(B200AccelerateScaleImage(image, columns, rows, exception));
  return __real_ScaleImage(image, columns, rows, exception);
}

Image *__wrap_SampleImage(const Image *image, const size_t columns, const size_t rows, ExceptionInfo *exception)
{
  TRY(B200AccelerateSampleImage(image, columns, rows, exception));
  return __real_SampleImage(image, columns, rows, exception);
}

/* ---- ThumbnailImage (resize.c:4591) ------------------------------------------------------------------------------
   Its SampleImage / ResizeImage calls are made from inside resize.o, which --wrap does not redirect, so the pixel
   cascade (:4617-4645) is re-issued here through the wrapped entry points (each stage takes the GPU or declines
   to the CPU on its own).  The metadata half of the reference (page geometry, depth, profile stripping, the
   Thumb::* properties) is NOT restated: the real ThumbnailImage is called on the finished thumbnail with its
   own size -- a same-size call skips the resize block and only applies the metadata, which it derives from
   fields the cascade's CloneImage-based results inherit from the source (filenames, magick_columns/rows, blob,
   properties).  Only the page count refers to the source's image list and is set afterwards. */
extern Image *__real_ThumbnailImage(const Image *, const size_t, const size_t, ExceptionInfo *);

Image *__wrap_ThumbnailImage(const Image *image, const size_t columns, const size_t rows, ExceptionInfo *exception)
{
  Image *clone_image, *stage, *result;
  ssize_t x_factor, y_factor;
  if (!b200_on() || columns == 0 || rows == 0 || ((columns == image->columns) && (rows == image->rows)))
    return __real_ThumbnailImage(image, columns, rows, exception);
  x_factor = (ssize_t) image->columns / (ssize_t) columns;
  y_factor = (ssize_t) image->rows / (ssize_t) rows;
  clone_image = (Image *) NULL;                       /* result of the previous stage (NULL: still the source) */
  if ((x_factor > 4) && (y_factor > 4)) {
    stage = SampleImage(image, 4 * columns, 4 * rows, exception);
    if (stage != (Image *) NULL) clone_image = stage;
  }
  if ((x_factor > 2) && (y_factor > 2)) {
    stage = ResizeImage(clone_image != (Image *) NULL ? clone_image : image, 2 * columns, 2 * rows, BoxFilter, exception);
    if (stage != (Image *) NULL) {
      if (clone_image != (Image *) NULL) clone_image = DestroyImage(clone_image);
      clone_image = stage;
    }
  }
  stage = ResizeImage(clone_image != (Image *) NULL ? clone_image : image, columns, rows,
                      image->filter == UndefinedFilter ? LanczosSharpFilter : image->filter, exception);
  if (clone_image != (Image *) NULL) clone_image = DestroyImage(clone_image);
  if (stage == (Image *) NULL) return (Image *) NULL;
  result = __real_ThumbnailImage(stage, columns, rows, exception);     /* same size: metadata only */
  stage = DestroyImage(stage);
  if (result != (Image *) NULL)
    (void) FormatImageProperty(result, "Thumb::Document::Pages", "%.20g", (double) GetImageListLength(image));
  return result;
}

/* ---- MinifyImage (resize.c:3158) and ResampleImage (:3209): one-line callers of ResizeImage from inside resize.o,
   which --wrap does not redirect; re-issued here through the wrapped ResizeImage. */
extern Image *__real_MinifyImage(const Image *, ExceptionInfo *);
extern Image *__real_ResampleImage(const Image *, const double, const double, const FilterType, ExceptionInfo *);

Image *__wrap_MinifyImage(const Image *image, ExceptionInfo *exception)
{
  if (!b200_on() || image->columns < 2 || image->rows < 2) return __real_MinifyImage(image, exception);
  return ResizeImage(image, image->columns / 2, image->rows / 2, SplineFilter, exception);          /* :3170 */
}

Image *__wrap_ResampleImage(const Image *image, const double x_resolution, const double y_resolution,
                            const FilterType filter, ExceptionInfo *exception)
{
  Image *out;
  size_t width, height;
  if (!b200_on()) return __real_ResampleImage(image, x_resolution, y_resolution, filter, exception);
  width = (size_t) (x_resolution * image->columns / (image->resolution.x == 0.0 ? 72.0 : image->resolution.x) + 0.5);   /* :3230 */
  height = (size_t) (y_resolution * image->rows / (image->resolution.y == 0.0 ? 72.0 : image->resolution.y) + 0.5);
  out = ResizeImage(image, width, height, filter, exception);
  if (out != (Image *) NULL) { out->resolution.x = x_resolution; out->resolution.y = y_resolution; }
  return out;
}

/* ---- MotionBlurImage (effect.c:2347; the reference's own hook is AccelerateMotionBlurImage, :2401) ------------------ */
typedef struct { double radius, sigma, angle; } motion_args;
static int op_motion(const float *s, float *d, size_t w, size_t h, int ch, const void *a)
{ const motion_args *m = (const motion_args *) a; return mb200_motion_blur_image(s, d, w, h, ch, m->radius, m->sigma, m->angle); }

Image *B200AccelerateMotionBlurImage(const Image *image, const double radius, const double sigma, const double angle,
                                     ExceptionInfo *exception)
{
  motion_args a;
  a.radius = radius; a.sigma = sigma; a.angle = angle;
  return run_same_size(image, op_motion, &a, exception);
}

extern Image *__real_MotionBlurImage(const Image *, const double, const double, const double, ExceptionInfo *);
Image *__wrap_MotionBlurImage(const Image *image, const double radius, const double sigma, const double angle,
                              ExceptionInfo *exception)
{
  TRY(B200AccelerateMotionBlurImage(image, radius, sigma, angle, exception));
  return __real_MotionBlurImage(image, radius, sigma, angle, exception);
}

/* ---- pixel caches in pinned host memory --------------------------------------------------------------------------------
   OpenPixelCache takes a memory cache's pixels from AcquireAlignedMemory (cache.c:3757), and that function honours the
   public SetMagickAlignedMemoryMethods hook (memory.c:376, :1541).  B200ShimInstallPixelCachePool() installs an allocator
   that serves large blocks (pixel caches) from a recycling pool of CUDA-pinned host memory and attaches each block to
   libmagickb200's residency registry: the operators then move pixels at PCIe speed (55 GB/s instead of the 9 / 19 GB/s of
   pageable cudaMemcpy, tools/micro/staging.cu) straight from / to the cache, and every pixel cache keeps one HBM copy for
   its lifetime.  Pinning costs ~300 ms per GiB, which is why freed blocks are recycled instead of being unpinned.  Small
   blocks and everything allocated before the installation stay with posix_memalign / free.
   Enabled by calling the function, or by MAGICK_B200_PINNED_CACHE=1 in the environment (checked when the shim is loaded). */
#include <pthread.h>

#define B200_POOL_MIN_BYTES ((size_t) 1 << 20)
#define B200_POOL_MAX_BLOCKS 256
static struct { void *ptr; size_t capacity; int in_use; } b200_pool[B200_POOL_MAX_BLOCKS];
static size_t b200_pool_idle_bytes = 0, b200_pool_idle_limit = (size_t) 8 << 30;
static pthread_mutex_t b200_pool_mutex = PTHREAD_MUTEX_INITIALIZER;
static long b200_pool_reused = 0, b200_pool_pinned = 0;

static void *b200_acquire_aligned(const size_t size, const size_t alignment)
{
  void *memory = (void *) NULL;
  if (size >= B200_POOL_MIN_BYTES && mb200_device_count() > 0) {
    int i, best = -1, slot = -1;
    pthread_mutex_lock(&b200_pool_mutex);
    for (i = 0; i < B200_POOL_MAX_BLOCKS; i++) {
      if (b200_pool[i].ptr == (void *) NULL) { if (slot < 0) slot = i; continue; }
      if (b200_pool[i].in_use == 0 && b200_pool[i].capacity >= size && b200_pool[i].capacity <= size + size / 4 &&
          (best < 0 || b200_pool[i].capacity < b200_pool[best].capacity)) best = i;
    }
    if (best >= 0) {
      b200_pool[best].in_use = 1;
      b200_pool_idle_bytes -= b200_pool[best].capacity;
      b200_pool_reused++;
      memory = b200_pool[best].ptr;
    } else if (slot >= 0) {
      const size_t capacity = (size + ((size_t) 1 << 21) - 1) & ~(((size_t) 1 << 21) - 1);
      b200_pool[slot].in_use = 1;                 /* reserve the slot while the (slow) pinning runs unlocked */
      b200_pool[slot].ptr = (void *) &b200_pool;  /* placeholder: not a block anybody can hold */
      pthread_mutex_unlock(&b200_pool_mutex);
      if (mb200_malloc_host(&memory, capacity) != MB200_OK) memory = (void *) NULL;
      pthread_mutex_lock(&b200_pool_mutex);
      if (memory != (void *) NULL) { b200_pool[slot].ptr = memory; b200_pool[slot].capacity = capacity; b200_pool_pinned++; }
      else { b200_pool[slot].ptr = (void *) NULL; b200_pool[slot].in_use = 0; }
    }
    pthread_mutex_unlock(&b200_pool_mutex);
    if (memory != (void *) NULL) {
      (void) mb200_cache_attach(memory, size, 0);
      return memory;
    }
  }
  if (posix_memalign(&memory, alignment < sizeof(void *) ? sizeof(void *) : alignment, size) != 0) return (void *) NULL;
  return memory;
}

static void b200_relinquish_aligned(void *memory)
{
  int i, ours = 0;
  void *release = (void *) NULL;
  if (memory == (void *) NULL) return;
  pthread_mutex_lock(&b200_pool_mutex);
  for (i = 0; i < B200_POOL_MAX_BLOCKS; i++)
    if (b200_pool[i].ptr == memory && b200_pool[i].in_use != 0) {
      ours = 1;
      b200_pool[i].in_use = 0;
      if (b200_pool_idle_bytes + b200_pool[i].capacity > b200_pool_idle_limit) { release = memory; b200_pool[i].ptr = (void *) NULL; }
      else b200_pool_idle_bytes += b200_pool[i].capacity;
      break;
    }
  pthread_mutex_unlock(&b200_pool_mutex);
  if (ours == 0) { free(memory); return; }
  (void) mb200_cache_detach(memory);              /* the image is gone: drop its HBM copy */
  if (release != (void *) NULL) (void) mb200_free_host(release);
}

void B200ShimInstallPixelCachePool(void)
{
  SetMagickAlignedMemoryMethods(b200_acquire_aligned, b200_relinquish_aligned);
}
void B200ShimPixelCachePoolStats(long *pinned_blocks, long *reused_blocks)
{
  if (pinned_blocks) *pinned_blocks = b200_pool_pinned;
  if (reused_blocks) *reused_blocks = b200_pool_reused;
}

/* ---- lazy synchronisation (hook mode) -----------------------------------------------------------------------------------
   The reference keeps an OpenCL result in its cl_mem until somebody looks at the pixels: CopyOpenCLBuffer() is called at
   three places of cache.c -- GetImagePixelCache (:1710, before the host writes), GetVirtualPixelCacheNexus (:2771, before the
   host reads) and PersistPixelCache (:4079).  A build that adds B200PixelCacheHook(cache_info->pixels, for_write) at the
   same three places (INTEGRATION.md has the patch; oracle/Makefile generates such a cache.c for the hooked harness) gets the
   same behaviour here: call B200ShimSetLazySync(1) once and chained operators (-blur ... -resize ...) upload once and
   download once.  Without the hooks lazy mode must stay off: nothing would bring a result back to the host. */
void B200PixelCacheHook(void *pixels, int for_write)
{
  if (pixels == (void *) NULL) return;
  (void) mb200_cache_sync(pixels);
  if (for_write != 0) (void) mb200_cache_host_written(pixels);
}
void B200ShimSetLazySync(int on) { (void) mb200_cache_set_lazy(on); }

__attribute__((constructor)) static void b200_shim_init(void)
{
  const char *e = getenv("MAGICK_B200_PINNED_CACHE");
  if (e != (const char *) NULL && *e != '\0' && *e != '0') B200ShimInstallPixelCachePool();
  e = getenv("MAGICK_B200_LAZY_SYNC");          /* only for builds that carry the cache.c hooks */
  if (e != (const char *) NULL && *e != '\0' && *e != '0') B200ShimSetLazySync(1);
}
