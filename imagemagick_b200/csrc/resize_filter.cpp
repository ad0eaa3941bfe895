// resize_filter.cpp -- host side of the resize path: the reference's filter table
// and the per-output contribution (start / count / normalised weights) lists.
//
// Behavioural mirror of MagickCore/resize.c: AcquireResizeFilter :803 (mapping table
// :835-876, function table :888-942, sharpening :1064, cubic coefficients :1181-1226),
// GetResizeFilterWeight :1690, GetResizeFilterSupport :1656, and the contribution
// set-up that opens every iteration of HorizontalFilter (:3398-3443) and
// VerticalFilter (:3614-3657).  No "filter:*" artifacts (the shim declines when any
// is set), never cylindrical (ResizeImage passes MagickFalse, :3817).
//
// Weights are evaluated on the host in double with the same operation order as the
// reference, so the table uploaded to the GPU is bit-identical to what the CPU path
// uses; the device only does the weighted sums.
#include "mb200_internal.h"

#include <cmath>
#include <utility>
#include <vector>
#include <vector>

namespace {

constexpr double kEps = 1.0e-12;
constexpr double kPi = 3.1415926535897932384626433832795028841971693993751058209749445923078164062;
constexpr double kPi2 = 1.57079632679489661923132169163975144209858469968755;
constexpr double k2Pi = 6.28318530717958647692528676655900576839433879875020;

inline double perceptible_reciprocal(double x) {
  const double sign = x < 0.0 ? -1.0 : 1.0;
  return (sign * x) >= kEps ? 1.0 / x : sign / kEps;
}

enum class Fn { Box, Triangle, CubicBC, Hann, Hamming, Blackman, Gaussian, Quadratic, Sinc, SincFast,
                Welch, Bohman, Lagrange, Cosine, CubicSpline, Mks2013, Mks2021, Unsupported };

struct FnEntry { Fn fn; double support, scale, B, C; };

// resize.c:888-942, indexed by FilterType
const FnEntry kFunctions[MB200_SentinelFilter] = {
  {Fn::Box, 0.5, 0.5, 0, 0},            // Undefined
  {Fn::Box, 0.0, 0.5, 0, 0},            // Point
  {Fn::Box, 0.5, 0.5, 0, 0},            // Box
  {Fn::Triangle, 1.0, 1.0, 0, 0},       // Triangle
  {Fn::CubicBC, 1.0, 1.0, 0, 0},        // Hermite
  {Fn::Hann, 1.0, 1.0, 0, 0},
  {Fn::Hamming, 1.0, 1.0, 0, 0},
  {Fn::Blackman, 1.0, 1.0, 0, 0},
  {Fn::Gaussian, 2.0, 1.5, 0, 0},
  {Fn::Quadratic, 1.5, 1.5, 0, 0},
  {Fn::CubicBC, 2.0, 2.0, 1.0, 0.0},    // Cubic
  {Fn::CubicBC, 2.0, 1.0, 0.0, 0.5},    // Catrom
  {Fn::CubicBC, 2.0, 8.0 / 7.0, 1. / 3., 1. / 3.},   // Mitchell
  {Fn::Unsupported, 3.0, 1.2196698912665045, 0, 0},  // Jinc (Bessel): not on the 1-D path
  {Fn::Sinc, 4.0, 1.0, 0, 0},
  {Fn::SincFast, 4.0, 1.0, 0, 0},
  {Fn::Unsupported, 1.0, 1.0, 0, 0},    // Kaiser (needs I0)
  {Fn::Welch, 1.0, 1.0, 0, 0},
  {Fn::CubicBC, 2.0, 2.0, 1.0, 0.0},    // Parzen
  {Fn::Bohman, 1.0, 1.0, 0, 0},
  {Fn::Triangle, 1.0, 1.0, 0, 0},       // Bartlett
  {Fn::Lagrange, 2.0, 1.0, 0, 0},
  {Fn::SincFast, 3.0, 1.0, 0, 0},       // Lanczos
  {Fn::SincFast, 3.0, 1.0, 0, 0},       // LanczosSharp
  {Fn::SincFast, 2.0, 1.0, 0, 0},       // Lanczos2
  {Fn::SincFast, 2.0, 1.0, 0, 0},       // Lanczos2Sharp
  {Fn::CubicBC, 2.0, 1.1685777620836932, 0.37821575509399867, 0.31089212245300067},  // Robidoux
  {Fn::CubicBC, 2.0, 1.105822933719019, 0.2620145123990142, 0.3689927438004929},     // RobidouxSharp
  {Fn::Cosine, 1.0, 1.0, 0, 0},
  {Fn::CubicBC, 2.0, 2.0, 1.0, 0.0},    // Spline
  {Fn::SincFast, 3.0, 1.0, 0, 0},       // LanczosRadius
  {Fn::CubicSpline, 2.0, 0.5, 0, 0},
  {Fn::Mks2013, 2.5, 1.0, 0, 0},
  {Fn::Mks2021, 4.5, 1.0, 0, 0},
};

// resize.c:835-876: requested filter -> {weighting function's filter, window's filter}
struct Mapping { int filter, window; };
const Mapping kMapping[MB200_SentinelFilter] = {
  {MB200_UndefinedFilter, MB200_BoxFilter}, {MB200_PointFilter, MB200_BoxFilter},
  {MB200_BoxFilter, MB200_BoxFilter}, {MB200_TriangleFilter, MB200_BoxFilter},
  {MB200_HermiteFilter, MB200_BoxFilter}, {MB200_SincFastFilter, MB200_HannFilter},
  {MB200_SincFastFilter, MB200_HammingFilter}, {MB200_SincFastFilter, MB200_BlackmanFilter},
  {MB200_GaussianFilter, MB200_BoxFilter}, {MB200_QuadraticFilter, MB200_BoxFilter},
  {MB200_CubicFilter, MB200_BoxFilter}, {MB200_CatromFilter, MB200_BoxFilter},
  {MB200_MitchellFilter, MB200_BoxFilter}, {MB200_JincFilter, MB200_BoxFilter},
  {MB200_SincFilter, MB200_BoxFilter}, {MB200_SincFastFilter, MB200_BoxFilter},
  {MB200_SincFastFilter, MB200_KaiserFilter}, {MB200_LanczosFilter, MB200_WelchFilter},
  {MB200_SincFastFilter, MB200_CubicFilter}, {MB200_SincFastFilter, MB200_BohmanFilter},
  {MB200_SincFastFilter, MB200_TriangleFilter}, {MB200_LagrangeFilter, MB200_BoxFilter},
  {MB200_LanczosFilter, MB200_LanczosFilter}, {MB200_LanczosSharpFilter, MB200_LanczosSharpFilter},
  {MB200_Lanczos2Filter, MB200_Lanczos2Filter}, {MB200_Lanczos2SharpFilter, MB200_Lanczos2SharpFilter},
  {MB200_RobidouxFilter, MB200_BoxFilter}, {MB200_RobidouxSharpFilter, MB200_BoxFilter},
  {MB200_LanczosFilter, MB200_CosineFilter}, {MB200_SplineFilter, MB200_BoxFilter},
  {MB200_LanczosRadiusFilter, MB200_LanczosFilter}, {MB200_CubicSplineFilter, MB200_BoxFilter},
  {MB200_MagicKernelSharp2013Filter, MB200_BoxFilter}, {MB200_MagicKernelSharp2021Filter, MB200_BoxFilter},
};

struct ResizeFilter {
  Fn filter = Fn::Box, window = Fn::Box;
  double support = 0, window_support = 0, scale = 1, blur = 1, coefficient[7] = {0};
  bool valid = false;

  explicit ResizeFilter(int requested) {
    if (requested <= MB200_UndefinedFilter || requested >= MB200_SentinelFilter) return;
    const int ft = kMapping[requested].filter, wt = kMapping[requested].window;
    filter = kFunctions[ft].fn;
    window = kFunctions[wt].fn;
    if (filter == Fn::Unsupported || window == Fn::Unsupported) return;
    support = kFunctions[ft].support;
    scale = kFunctions[wt].scale;
    if (ft == MB200_LanczosSharpFilter) blur *= 0.9812505644269356;
    if (ft == MB200_Lanczos2SharpFilter) blur *= 0.9549963639785485;
    if (filter == Fn::Gaussian || window == Fn::Gaussian) {
      const double sigma = 0.5;
      coefficient[0] = sigma;
      coefficient[1] = perceptible_reciprocal(2.0 * sigma * sigma);
      coefficient[2] = perceptible_reciprocal(k2Pi * sigma * sigma);
    }
    if (blur < kEps) blur = kEps;
    window_support = support;
    scale *= perceptible_reciprocal(window_support);
    if (filter == Fn::CubicBC || window == Fn::CubicBC) {
      double B = kFunctions[ft].B, C = kFunctions[ft].C;
      if (kFunctions[wt].fn == Fn::CubicBC) { B = kFunctions[wt].B; C = kFunctions[wt].C; }
      const double twoB = B + B;
      coefficient[0] = 1.0 - (1.0 / 3.0) * B;
      coefficient[1] = -3.0 + twoB + C;
      coefficient[2] = 2.0 - 1.5 * B - C;
      coefficient[3] = (4.0 / 3.0) * B + 4.0 * C;
      coefficient[4] = -8.0 * C - twoB;
      coefficient[5] = B + 5.0 * C;
      coefficient[6] = (-1.0 / 6.0) * B - C;
    }
    valid = true;
  }

  // resize.c:493-587, Q16 coefficient set (:547-563)
  static double sinc_fast(double x) {
    if (x > 4.0) {
      const double alpha = kPi * x;
      return std::sin(alpha) / alpha;
    }
    static const double c[10] = {
      0.173611107357320220183368594093166520811e-2L, -0.384240921114946632192116762889211361285e-3L,
      0.394201182359318128221229891724947048771e-4L, -0.250963301609117217660068889165550534856e-5L,
      0.111902032818095784414237782071368805120e-6L, -0.372895101408779549368465614321137048875e-8L,
      0.957694196677572570319816780188718518330e-10L, -0.187208577776590710853865174371617338991e-11L,
      0.253524321426864752676094495396308636823e-13L, -0.177084805010701112639035485248501049364e-15L};
    const double xx = x * x;
    double p = c[9];
    for (int i = 8; i >= 0; --i) p = c[i] + xx * p;      // Horner, same association as the reference
    return (xx - 1.0) * (xx - 4.0) * (xx - 9.0) * (xx - 16.0) * p;
  }

  double eval(Fn fn, double x) const {
    switch (fn) {
      case Fn::Box: return 1.0;
      case Fn::Triangle: return x < 1.0 ? 1.0 - x : 0.0;
      case Fn::CubicBC:
        if (x < 1.0) return coefficient[0] + x * (x * (coefficient[1] + x * coefficient[2]));
        if (x < 2.0) return coefficient[3] + x * (coefficient[4] + x * (coefficient[5] + x * coefficient[6]));
        return 0.0;
      case Fn::Hann: { const double c = std::cos(kPi * x); return 0.5 + 0.5 * c; }
      case Fn::Hamming: { const double c = std::cos(kPi * x); return 0.54 + 0.46 * c; }
      case Fn::Blackman: { const double c = std::cos(kPi * x); return 0.34 + c * (0.5 + c * 0.16); }
      case Fn::Gaussian: return std::exp(-coefficient[1] * x * x);
      case Fn::Quadratic:
        if (x < 0.5) return 0.75 - x * x;
        if (x < 1.5) return 0.5 * (x - 1.5) * (x - 1.5);
        return 0.0;
      case Fn::Sinc:
        if (x != 0.0) { const double a = kPi * x; return std::sin(a) / a; }
        return 1.0;
      case Fn::SincFast: return sinc_fast(x);
      case Fn::Welch: return x < 1.0 ? 1.0 - x * x : 0.0;
      case Fn::Bohman: {
        const double c = std::cos(kPi * x);
        const double s = std::sqrt(1.0 - c * c);
        return (1.0 - x) * c + (1.0 / kPi) * s;
      }
      case Fn::Cosine: return std::cos(kPi2 * x);
      case Fn::Lagrange: {
        if (x > support) return 0.0;
        const long order = static_cast<long>(2.0 * window_support);
        const long n = static_cast<long>(window_support + x);
        double value = 1.0f;
        for (long i = 0; i < order; ++i)
          if (i != n) value *= (n - i - x) / (n - i);
        return value;
      }
      case Fn::CubicSpline:   // 2-lobe form; support is never overridden here
        if (x < 1.0) return ((x - 9.0 / 5.0) * x - 1.0 / 5.0) * x + 1.0;
        if (x < 2.0) return ((-1.0 / 3.0 * (x - 1.0) + 4.0 / 5.0) * (x - 1.0) - 7.0 / 15.0) * (x - 1.0);
        return 0.0;
      case Fn::Mks2013:
        if (x < 0.5) return 0.625 + 1.75 * (0.5 - x) * (0.5 + x);
        if (x < 1.5) return (1.0 - x) * (1.75 - x);
        if (x < 2.5) return -0.125 * (2.5 - x) * (2.5 - x);
        return 0.0;
      case Fn::Mks2021:
        if (x < 0.5) return 577.0 / 576.0 - 239.0 / 144.0 * x * x;
        if (x < 1.5) return 35.0 / 36.0 * (x - 1.0) * (x - 239.0 / 140.0);
        if (x < 2.5) return 1.0 / 6.0 * (x - 2.0) * (65.0 / 24.0 - x);
        if (x < 3.5) return 1.0 / 36.0 * (x - 3.0) * (x - 3.75);
        if (x < 4.5) return -1.0 / 288.0 * (x - 4.5) * (x - 4.5);
        return 0.0;
      default: return 0.0;
    }
  }

  double weight(double x) const {           // resize.c:1690
    const double x_blur = std::fabs(x) * perceptible_reciprocal(blur);
    double s;
    if (window_support < kEps || window == Fn::Box) s = 1.0;
    else s = eval(window, x_blur * scale);
    return s * eval(filter, x_blur);
  }
  double practical_support() const { return support * blur; }   // resize.c:1656
};

}  // namespace

extern "C" {

double mb200_resize_filter_weight(int filter, double x) {
  ResizeFilter rf(filter);
  if (!rf.valid) return std::nan("");
  return rf.weight(x);
}

double mb200_resize_filter_support(int filter) {
  ResizeFilter rf(filter);
  if (!rf.valid) return std::nan("");
  return rf.practical_support();
}

long mb200_resize_contributions(int filter, size_t in_n, size_t out_n, double factor, long *start,
                                int *count, double *weights, size_t max_taps) {
  ResizeFilter rf(filter);
  if (!rf.valid) return mb200::fail(MB200_EUNSUPPORTED, "resize filter %d is not supported on the 1-D GPU path", filter);
  if (in_n == 0 || out_n == 0 || !(factor > 0.0)) return mb200::fail(MB200_EINVAL, "bad resize geometry");
  // resize.c:3363-3386 / :3578-3601
  double scale = std::fmax(1.0 / factor + kEps, 1.0);
  double support = scale * rf.practical_support();
  if (support < 0.5) { support = 0.5; scale = 1.0; }
  const long need = static_cast<long>(2.0 * support + 3.0);
  if (!start || !count || !weights) return need;
  if (static_cast<long>(max_taps) < need) return mb200::fail(MB200_EINVAL, "max_taps %zu < %ld", max_taps, need);
  scale = perceptible_reciprocal(scale);
  for (size_t o = 0; o < out_n; ++o) {
    const double bisect = static_cast<double>(o + 0.5) / factor + kEps;
    const long first = static_cast<long>(std::fmax(bisect - support + 0.5, 0.0));
    const long last = static_cast<long>(std::fmin(bisect + support + 0.5, static_cast<double>(in_n)));
    const long n = last - first > 0 ? last - first : 0;
    double *w = weights + o * max_taps;
    double density = 0.0;
    for (long j = 0; j < n; ++j) {
      w[j] = rf.weight(scale * (static_cast<double>(first + j) - bisect + 0.5));
      density += w[j];
    }
    if (n > 0 && density != 0.0 && density != 1.0) {
      density = perceptible_reciprocal(density);
      for (long j = 0; j < n; ++j) w[j] *= density;
    }
    for (size_t j = static_cast<size_t>(n); j < max_taps; ++j) w[j] = 0.0;
    start[o] = first;
    count[o] = static_cast<int>(n);
  }
  return need;
}

// ScaleImage (resize.c:4106-4530) is a sequential state machine whose weights do not depend on the pixels: the running
// (span, scale) pairs are simulated here in the reference's own order and arithmetic, and recorded per output as a list of
// (source index, weight) terms in accumulation order.  axis 1 = rows (the y machine pulls source rows per output row,
// :4229-4330), axis 0 = columns (the x machine pushes every source column into the outputs it overlaps, :4378-4440).
// offsets has out_n + 1 entries; index / weight hold offsets[out_n] terms (query the count with index == NULL).
long mb200_scale_contributions(int axis, size_t in_n, size_t out_n, long *offsets, int *index, double *weight, size_t max_terms) {
  if (in_n == 0 || out_n == 0 || !offsets) return mb200::fail(MB200_EINVAL, "bad scale geometry");
  std::vector<std::vector<std::pair<int, double>>> lists(out_n);
  const double factor = static_cast<double>(out_n) / static_cast<double>(in_n);
  if (in_n == out_n) {
    for (size_t o = 0; o < out_n; ++o) lists[o].emplace_back(static_cast<int>(o), 1.0);      // taken as is (:4190, :4336)
  } else if (axis == 1) {
    long number_rows = 0, next = 0;
    int cur = 0;
    bool next_row = true;
    double span = 1.0, scale = factor;
    for (size_t y = 0; y < out_n; ++y) {
      while (scale < span) {
        if (next_row && number_rows < static_cast<long>(in_n)) { cur = static_cast<int>(next++); ++number_rows; }
        lists[y].emplace_back(cur, scale);                       // y_vector += scale.y * x_vector
        span -= scale;
        scale = factor;
        next_row = true;
      }
      if (next_row && number_rows < static_cast<long>(in_n)) { cur = static_cast<int>(next++); ++number_rows; next_row = false; }
      lists[y].emplace_back(cur, span);                          // pixel = y_vector + span.y * x_vector
      scale -= span;
      if (scale <= 0) { scale = factor; next_row = true; }
      span = 1.0;
    }
  } else {
    long t = 0;
    bool next_column = false;
    double span = 1.0;
    auto add = [&](long out, size_t x, double w) { if (out >= 0 && out < static_cast<long>(out_n)) lists[static_cast<size_t>(out)].emplace_back(static_cast<int>(x), w); };
    for (size_t x = 0; x < in_n; ++x) {
      double scale = factor;
      while (scale >= span) {
        if (next_column) ++t;                                    // pixel = 0 for a new output
        add(t, x, span);
        scale -= span;
        span = 1.0;
        next_column = true;
      }
      if (scale > 0) {
        if (next_column) { next_column = false; ++t; }
        add(t, x, scale);
        span -= scale;
      }
    }
    if (span > 0 && !next_column) add(t, in_n - 1, span);         // :4433-4440 (only a still-open output is stored again)
  }
  long total = 0;
  for (size_t o = 0; o < out_n; ++o) { offsets[o] = total; total += static_cast<long>(lists[o].size()); }
  offsets[out_n] = total;
  if (!index || !weight) return total;
  if (static_cast<long>(max_terms) < total) return mb200::fail(MB200_EINVAL, "max_terms %zu < %ld", max_terms, total);
  long k = 0;
  for (size_t o = 0; o < out_n; ++o)
    for (const auto &term : lists[o]) { index[k] = term.first; weight[k] = term.second; ++k; }
  return total;
}

// The source sample a Copy-trait channel takes for every output of one axis (resize.c:3697-3707):
// j = (ssize_t) (min(max(bisect, start), stop - 1) + 0.5), with bisect / start / stop as in the contribution list.
int mb200_resize_nearest(int filter, size_t in_n, size_t out_n, double factor, long *nearest) {
  ResizeFilter rf(filter);
  if (!rf.valid) return mb200::fail(MB200_EUNSUPPORTED, "resize filter %d is not supported on the 1-D GPU path", filter);
  if (in_n == 0 || out_n == 0 || !(factor > 0.0) || !nearest) return mb200::fail(MB200_EINVAL, "bad resize geometry");
  double scale = std::fmax(1.0 / factor + kEps, 1.0);
  double support = scale * rf.practical_support();
  if (support < 0.5) support = 0.5;
  for (size_t o = 0; o < out_n; ++o) {
    const double bisect = static_cast<double>(o + 0.5) / factor + kEps;
    const long first = static_cast<long>(std::fmax(bisect - support + 0.5, 0.0));
    const long last = static_cast<long>(std::fmin(bisect + support + 0.5, static_cast<double>(in_n)));
    nearest[o] = static_cast<long>(std::fmin(std::fmax(bisect, static_cast<double>(first)), static_cast<double>(last) - 1.0) + 0.5);
  }
  return MB200_OK;
}

}  // extern "C"
