// conv1d.cu -- one pass of a separable (1-D) convolution: the kernel behind
// BlurImage's two passes (MagickCore/effect.c:788 "blur:RxS;blur:RxS+90").
//
// Semantics follow MorphologyPrimitive's ConvolveMorphology for height-1 kernels (row
// path, MagickCore/morphology.c:2897-2979) and width-1 kernels (column fast path,
// :2654-2807): reflected taps, edge-clamped source (cache.c:2663), double accumulation,
// alpha-weighted blending of the colour channels when the image has alpha, result
// cast to float (quantum.h:88).
//
// B200 mapping.  The op is a memory-streaming stencil whose co-limit is the FP64 pipe
// (33 taps x 4 channels of double FMA per pixel for sigma=4).  Design:
//  * one thread per float component line; it walks along the filter axis keeping NT
//    rotating FP64 accumulators in registers, so every source sample is loaded and
//    converted to double ONCE and feeds NT FMAs whose tap operands are immediates in
//    the kernel-parameter constant bank (no tap loads, no shared-memory traffic per FMA);
//  * alpha weighting is done by premultiplying the sample once (q = A*p); the weight
//    sum gamma is exactly the alpha component's own accumulator, fetched with a warp
//    shuffle from the pixel's alpha lane -- 4 FMAs per tap per pixel instead of the
//    reference's 7 flops;
//  * column pass: lanes span 32 consecutive components of a row => every load/store
//    is one fully coalesced 128-byte line; a register ring keeps NT row loads in
//    flight per thread;
//  * row pass: a (rows x segment) tile is staged in shared memory with coalesced
//    float4 loads (odd pixel pitch => conflict-free LDS for the 8-rows x 4-channels
//    lane layout).
// No tensor cores: this is not a dense contraction.
#include "mb200_internal.h"

#include <cuda_runtime.h>

namespace mb200 {
namespace {

constexpr double kQuantumScale = 1.0 / 65535.0;
constexpr double kEpsilon = 1.0e-12;

template <int NT>
struct Taps { double k[NT]; };

struct Conv1dArgs {
  const float *src;
  float *dst;
  int width, height, channels;
  int rc;        // width * channels (floats per row)
  int off;       // samples before the output position covered by the window (ox / oy)
  int strip;     // outputs per thread along the filter axis
  int seg_w;     // row pass: source pixels staged per tile row (strip + NT - 1)
  int pitch;     // row pass: shared-memory tile pitch in pixels (odd)
  double bias;
  unsigned long long *changed;
};

__device__ __forceinline__ double fast_reciprocal(double g) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(g));
  double e = fma(-g, r, 1.0);
  r = fma(r, e, r);
  e = fma(-g, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ double shfl_double(double v, int lane) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, lane);
  hi = __shfl_sync(0xffffffffu, hi, lane);
  return __hiloint2double(hi, lo);
}

// Turns the finished accumulator(s) into the output Quantum.
//  plain : out = bias + sum
//  blend : pixel = bias + QS*sum', gamma = QS*gsum'  (sum' = sum K*A*p, gsum' = sum K*A)
//          out = PerceptibleReciprocal(gamma) * pixel          (morphology.c:3197)
template <int MODE>
__device__ __forceinline__ float finish(double sum, double gsum, bool is_alpha, double bias,
                                        double *unnormalised) {
  if (MODE == 0 || is_alpha) {
    const double pixel = bias + sum;
    *unnormalised = pixel;
    return static_cast<float>(pixel);
  }
  const double pixel = fma(kQuantumScale, sum, bias);
  const double gamma = kQuantumScale * gsum;
  *unnormalised = pixel;
  double r;
  if (fabs(gamma) >= kEpsilon) r = fast_reciprocal(gamma);
  else r = gamma < 0.0 ? -1.0 / kEpsilon : 1.0 / kEpsilon;
  return static_cast<float>(r * pixel);
}

__device__ __forceinline__ void count_changed(bool changed, unsigned long long *counter) {
  const unsigned mask = __ballot_sync(__activemask(), changed);
  if (mask != 0 && (threadIdx.x & 31) == (__ffs(mask) - 1)) atomicAdd(counter, (unsigned long long) __popc(mask));
}

// ---------------------------------------------------------------- column pass
// grid: (ceil(rc / THREADS), ceil(height / strip)); one thread per component column.
template <int NT, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) conv_col_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  const int col_raw = blockIdx.x * THREADS + threadIdx.x;
  const bool active = col_raw < a.rc;
  const int col = active ? col_raw : a.rc - 1;
  const int lane = threadIdx.x & 31;
  const int alane = MODE ? (lane | (MODE - 1)) : lane;
  const bool is_alpha = MODE ? ((col % MODE) == MODE - 1) : false;
  const int y0 = blockIdx.y * a.strip;
  const int nout = min(a.strip, a.height - y0);
  const int total = nout + NT - 1;
  const int hmax = a.height - 1;
  const size_t pitch = static_cast<size_t>(a.rc);
  const float *base = a.src + col;

  double acc[NT];
  float pre[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) acc[q] = 0.0;
#pragma unroll
  for (int s = 0; s < NT; ++s) {
    const int yy = min(max(y0 + s - a.off, 0), hmax);
    pre[s] = __ldg(base + static_cast<size_t>(yy) * pitch);
  }

  for (int mb = 0; mb < total; mb += NT) {
#pragma unroll
    for (int s = 0; s < NT; ++s) {
      const int m = mb + s;
      if (m >= total) break;
      const float vf = pre[s];
      {  // keep NT row loads in flight: refill this ring slot with the sample NT steps ahead
        const int yy = min(max(y0 + m + NT - a.off, 0), hmax);
        pre[s] = __ldg(base + static_cast<size_t>(yy) * pitch);
      }
      double v = static_cast<double>(vf);
      if (MODE) {
        const float af = __shfl_sync(0xffffffffu, vf, alane);
        v *= is_alpha ? 1.0 : static_cast<double>(af);
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = fma(taps.k[(s - q + NT) % NT], v, acc[q]);
      constexpr int kSlots = NT;
      const int qf = (s + 1) % kSlots;
      const double sum = acc[qf];
      acc[qf] = 0.0;
      double gsum = 0.0;
      if (MODE) gsum = shfl_double(sum, alane);
      const int j = m - (NT - 1);
      if (j >= 0 && active) {
        double unnorm;
        const float out = finish<MODE>(sum, gsum, is_alpha, a.bias, &unnorm);
        const size_t o = static_cast<size_t>(y0 + j) * pitch + col;
        a.dst[o] = out;
        if (a.changed != nullptr)
          count_changed(fabs(unnorm - static_cast<double>(__ldg(a.src + o))) >= kEpsilon, a.changed);
      }
    }
  }
}

// ------------------------------------------------------------------- row pass
// block: 128 threads = 4 warps; a warp covers RPW = 32/channels rows x channels lanes.
// grid: (ceil(width / strip), ceil(height / rows_per_cta)).
template <int NT, int MODE>
__global__ void __launch_bounds__(128) conv_row_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  extern __shared__ __align__(16) float tile[];
  const int ch = a.channels;
  const int rpw = 32 / ch;
  const int rows_per_cta = 4 * rpw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * a.strip;
  const int ybase = blockIdx.y * rows_per_cta;
  const int nout = min(a.strip, a.width - x0);
  const int total = nout + NT - 1;
  const int wmax = a.width - 1, hmax = a.height - 1;

  // ---- stage the tile: rows_per_cta x total source pixels, x edge-clamped
  if (ch == 4) {
    const int n = rows_per_cta * total;
    for (int idx = threadIdx.x; idx < n; idx += 128) {
      const int r = idx / total, px = idx - r * total;
      const int yy = min(ybase + r, hmax);
      const int xx = min(max(x0 + px - a.off, 0), wmax);
      const float4 v = __ldg(reinterpret_cast<const float4 *>(a.src) + (static_cast<size_t>(yy) * a.width + xx));
      *reinterpret_cast<float4 *>(tile + (static_cast<size_t>(r) * a.pitch + px) * 4) = v;
    }
  } else {
    const int n = rows_per_cta * total * ch;
    const int rowf = total * ch;
    for (int idx = threadIdx.x; idx < n; idx += 128) {
      const int r = idx / rowf, e = idx - r * rowf;
      const int px = e / ch, c = e - px * ch;
      const int yy = min(ybase + r, hmax);
      const int xx = min(max(x0 + px - a.off, 0), wmax);
      tile[(static_cast<size_t>(r) * a.pitch + px) * ch + c] =
          __ldg(a.src + (static_cast<size_t>(yy) * a.width + xx) * ch + c);
    }
  }
  __syncthreads();

  const bool lane_ok = lane < rpw * ch;
  const int lr = lane_ok ? lane / ch : 0, c = lane_ok ? lane % ch : 0;
  const int r = warp * rpw + lr;
  const int y = ybase + r;
  const bool active = lane_ok && y < a.height;
  const int alane = MODE ? (lane | (MODE - 1)) : lane;
  const bool is_alpha = MODE ? (c == MODE - 1) : false;
  const float *trow = tile + static_cast<size_t>(r) * a.pitch * ch + c;
  float *drow = a.dst + (static_cast<size_t>(min(y, hmax)) * a.width + x0) * ch + c;
  const float *crow = a.src + (static_cast<size_t>(min(y, hmax)) * a.width + x0) * ch + c;

  double acc[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) acc[q] = 0.0;

  for (int mb = 0; mb < total; mb += NT) {
#pragma unroll
    for (int s = 0; s < NT; ++s) {
      const int m = mb + s;
      if (m >= total) break;
      const float vf = trow[m * ch];
      double v = static_cast<double>(vf);
      if (MODE) {
        const float af = __shfl_sync(0xffffffffu, vf, alane);
        v *= is_alpha ? 1.0 : static_cast<double>(af);
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = fma(taps.k[(s - q + NT) % NT], v, acc[q]);
      const int qf = (s + 1) % NT;
      const double sum = acc[qf];
      acc[qf] = 0.0;
      double gsum = 0.0;
      if (MODE) gsum = shfl_double(sum, alane);
      const int j = m - (NT - 1);
      if (j >= 0 && active) {
        double unnorm;
        const float out = finish<MODE>(sum, gsum, is_alpha, a.bias, &unnorm);
        drow[j * ch] = out;
        if (a.changed != nullptr)
          count_changed(fabs(unnorm - static_cast<double>(__ldg(crow + j * ch))) >= kEpsilon, a.changed);
      }
    }
  }
}

template <int NT, int MODE>
int launch_nt(const Conv1dArgs &base, int axis, const double *taps_host, int ntaps, cudaStream_t stream) {
  Taps<NT> taps;
  for (int i = 0; i < NT; ++i) taps.k[i] = i < ntaps ? taps_host[i] : 0.0;   // zero padding past the window
  Conv1dArgs a = base;
  if (axis == 1) {
    constexpr int kThreads = 128;
    a.strip = 8 * NT + 1;                       // strip + NT - 1 is a whole number of rotations
    dim3 grid((a.rc + kThreads - 1) / kThreads, (a.height + a.strip - 1) / a.strip);
    conv_col_kernel<NT, MODE, kThreads><<<grid, kThreads, 0, stream>>>(a, taps);
  } else {
    // strip + NT - 1 is a whole number of rotations and the strip is at least ~64 outputs
    constexpr int kRot = (63 + NT - 1) / NT < 2 ? 2 : (63 + NT - 1) / NT;
    a.strip = kRot * NT + 1;
    a.seg_w = a.strip + NT - 1;
    a.pitch = a.seg_w | 1;
    const int rows_per_cta = 4 * (32 / a.channels);
    const size_t smem = static_cast<size_t>(rows_per_cta) * a.pitch * a.channels * sizeof(float);
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
      cudaFuncSetAttribute(conv_row_kernel<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    dim3 grid((a.width + a.strip - 1) / a.strip, (a.height + rows_per_cta - 1) / rows_per_cta);
    conv_row_kernel<NT, MODE><<<grid, 128, smem, stream>>>(a, taps);
  }
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv1d launch");
  return MB200_OK;
}

template <int NT>
int launch_mode(const Conv1dArgs &a, int axis, const double *taps, int ntaps, cudaStream_t s) {
  if (a.channels == 4) return launch_nt<NT, 4>(a, axis, taps, ntaps, s);
  if (a.channels == 2) return launch_nt<NT, 2>(a, axis, taps, ntaps, s);
  return launch_nt<NT, 0>(a, axis, taps, ntaps, s);
}

}  // namespace

int launch_conv1d(const float *src, float *dst, size_t width, size_t height, int channels, int axis,
                  const double *taps, int ntaps, int origin_offset, double bias, double /*gamma_scale*/,
                  unsigned long long *d_changed, void *stream) {
  if (width == 0 || height == 0 || channels < 1 || channels > 4 || ntaps < 1)
    return fail(MB200_EINVAL, "conv1d: bad geometry");
  if (width * channels > 0x7fffffffull || height > 0x7fffffffull) return fail(MB200_EINVAL, "conv1d: image too large");
  Conv1dArgs a{};
  a.src = src; a.dst = dst;
  a.width = static_cast<int>(width); a.height = static_cast<int>(height); a.channels = channels;
  a.rc = a.width * channels;
  a.off = origin_offset;
  a.bias = bias;
  a.changed = d_changed;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ntaps <= 9) return launch_mode<9>(a, axis, taps, ntaps, s);
  if (ntaps <= 17) return launch_mode<17>(a, axis, taps, ntaps, s);
  if (ntaps <= 25) return launch_mode<25>(a, axis, taps, ntaps, s);
  if (ntaps <= 33) return launch_mode<33>(a, axis, taps, ntaps, s);
  if (ntaps <= 49) return launch_mode<49>(a, axis, taps, ntaps, s);
  if (ntaps <= 65) return launch_mode<65>(a, axis, taps, ntaps, s);
  return MB200_EUNSUPPORTED;   // caller falls back to the generic 2-D kernel
}

}  // namespace mb200
