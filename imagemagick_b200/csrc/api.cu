// api.cu -- the C-ABI operators (include/magick_b200.h): the host-side control flow of
// the hot path around the CUDA kernels.
//
// Mirrors the drivers in the reference (behaviour, not code):
//   MorphologyImage / MorphologyApply  MagickCore/morphology.c:4129, :3634  (kernel lists are
//     re-iterated, compound Open/Close/Smooth staging :3813-3893, `changed`-driven iteration :3919)
//   BlurImage :765, ConvolveImage :1170, GaussianBlurImage :1709, UnsharpMaskImage :4256
//     (MagickCore/effect.c)
//   ResizeImage MagickCore/resize.c:3761 (pass order :3846-3861, default filter :3806-3816)
//   TransformImageColorspace MagickCore/colorspace.c:1751
// Intermediates live in HBM as float Quantum, exactly like the reference's intermediate
// images (the rounding between passes is part of the semantics).
#include "mb200_internal.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

using namespace mb200;

namespace {

// Device-resident resize contribution tables, cached per (device, filter, in_n, out_n): batches of
// equally sized images (BASELINE configs[4]) rebuild neither the weights on the host nor upload them.
struct ResizeTables {
  int device = -1, filter = 0;
  mb200_filter_options options{};     // expert settings the table was built with (all zero: none)
  size_t in_n = 0, out_n = 0;
  long taps = 0;
  int max_span = 0, reg_stride = 0, reg_taps = 0;
  int *d_start = nullptr, *d_count = nullptr;
  double *d_weights = nullptr, *d_wreg = nullptr;
  // streaming kernels (resize_stream.cu): runs of outputs with bit-identical weights, and the
  // complement (borders, short runs) that stays with the gather kernels
  int nseg = 0, seg_o[MB200_RESIZE_MAX_SEGMENTS], seg_n[MB200_RESIZE_MAX_SEGMENTS], seg_src[MB200_RESIZE_MAX_SEGMENTS];
  double *d_wsets = nullptr;
  int nborder = 0;
  int *d_border = nullptr;
  // fused V+H kernel (resize_stream.cu): the runs cut into tiles of tile_w (this axis used as x) / tile_h (as y) outputs
  int *d_tiles_x = nullptr, *d_tiles_y = nullptr;
  int ntiles_x = 0, ntiles_y = 0;
  unsigned char *d_is_border = nullptr;       // [out_n]: output lies outside the runs
  ResizeTables() = default;
  ResizeTables(const ResizeTables &) = delete;
  ResizeTables &operator=(const ResizeTables &) = delete;
  ~ResizeTables() {            // cudaFree waits for launches that still read the buffers
    int cur = 0;
    cudaGetDevice(&cur);
    if (device >= 0 && cur != device) cudaSetDevice(device);
    cudaFree(d_start); cudaFree(d_count); cudaFree(d_weights); cudaFree(d_wreg); cudaFree(d_wsets); cudaFree(d_border);
    cudaFree(d_tiles_x); cudaFree(d_tiles_y); cudaFree(d_is_border);
    if (device >= 0 && cur != device) cudaSetDevice(cur);
  }
};
// Bounded LRU of shared entries: a call holds a reference while its launches are being queued, eviction drops the
// cache's reference and the last owner frees the device buffers.  Tables are built outside the lock.
std::mutex g_tables_mutex;
std::vector<std::shared_ptr<ResizeTables>> g_tables;
constexpr size_t kMaxCachedTables = 64;


struct StreamAlloc {           // stream-ordered temporary; freed (stream-ordered) on scope exit
  void *ptr = nullptr;
  cudaStream_t s;
  explicit StreamAlloc(cudaStream_t stream) : s(stream) {}
  int alloc(size_t bytes) {
    cudaError_t e = cudaMallocAsync(&ptr, bytes ? bytes : 1, temp_pool(), s);      // the library's private pool
    if (e != cudaSuccess) { ptr = nullptr; return cuda_fail(e, "cudaMallocAsync"); }
    return MB200_OK;
  }
  ~StreamAlloc() { if (ptr) cudaFreeAsync(ptr, s); }
  StreamAlloc(const StreamAlloc &) = delete;
  StreamAlloc &operator=(const StreamAlloc &) = delete;
};

int prepare(void *stream, cudaStream_t *out) {
  int rc = ensure_device();
  if (rc) return rc;
  *out = stream ? static_cast<cudaStream_t>(stream) : static_cast<cudaStream_t>(default_stream());
  return MB200_OK;
}

bool valid_image(size_t w, size_t h, int ch) { return w > 0 && h > 0 && ch >= 1 && ch <= 4; }

// Developer switches (environment, read once per process): force the generic kernels so that tests can compare them
// with the specialised ones.
struct Knobs {
  bool no_rank1, no_morph_stream, no_resize_stream, resize_regular_h, no_fused_unsharp, resize_fused;
  Knobs() {
    auto on = [](const char *name) { const char *v = std::getenv(name); return v != nullptr && *v != '\0' && *v != '0'; };
    no_rank1 = on("MB200_NO_RANK1");
    no_morph_stream = on("MB200_NO_MORPH_STREAM");
    no_resize_stream = on("MB200_NO_RESIZE_STREAM");
    resize_regular_h = on("MB200_RESIZE_REGULAR_H");
    no_fused_unsharp = on("MB200_NO_FUSED_UNSHARP");
    resize_fused = on("MB200_RESIZE_FUSED");          // opt-in: measured slower than the two passes (profiles/r02_resize_fused.md)
  }
};
Knobs &knobs() {
  static Knobs k;
  return k;
}

// One MorphologyPrimitive launch.  d_counter may be null.
int primitive(const float *src, float *dst, size_t w, size_t h, int ch, int method,
              const mb200_kernel_info *k, double bias, unsigned long long *d_counter, cudaStream_t s,
              const UnsharpEpilogue *epilogue = nullptr, bool *epilogue_fused = nullptr) {
  const int kw = static_cast<int>(k->width), kh = static_cast<int>(k->height);
  const size_t n = k->width * k->height;
  std::vector<double> win(n);
  int ox, oy;
  bool has_nan = false;
  if (method == MB200_ConvolveMorphology || method == MB200_DilateMorphology || method == MB200_DilateIntensityMorphology ||
      method == MB200_IterativeDistanceMorphology) {                              // reflected, :2612-2626
    for (size_t i = 0; i < n; ++i) win[i] = k->values[n - 1 - i];
    ox = kw - static_cast<int>(k->x) - 1;
    oy = kh - static_cast<int>(k->y) - 1;
  } else if (method == MB200_ErodeMorphology || method == MB200_ErodeIntensityMorphology ||
             method == MB200_HitAndMissMorphology || method == MB200_ThinningMorphology ||
             method == MB200_ThickenMorphology) {
    for (size_t i = 0; i < n; ++i) win[i] = k->values[i];
    ox = static_cast<int>(k->x);
    oy = static_cast<int>(k->y);
  } else {
    return fail(MB200_EUNSUPPORTED, "morphology primitive %d is not implemented on the GPU path", method);
  }
  for (double v : win) if (std::isnan(v)) has_nan = true;
  if (ox < 0 || oy < 0 || ox >= kw || oy >= kh) return fail(MB200_EINVAL, "kernel origin outside the kernel");
  if (method == MB200_ConvolveMorphology && !has_nan && (kw == 1 || kh == 1)) {
    // width-1 kernels take the reference's column path (:2654); for all-finite taps its extra
    // gamma*(height/count) factor is exactly 1.
    const int axis = (kw == 1) ? 1 : 0;
    const int rc = launch_conv1d(src, dst, w, h, ch, axis, win.data(), axis == 1 ? kh : kw, axis == 1 ? oy : ox,
                                 bias, 1.0, d_counter, s, 0, epilogue, epilogue_fused);
    if (rc != MB200_EUNSUPPORTED) return rc;
  }
  if (method == MB200_ConvolveMorphology && !has_nan && kw > 1 && kh > 1 && kw <= 33 && kh <= 33 && ch == 4 &&
      bias == 0.0 && d_counter == nullptr && !knobs().no_rank1) {
    // Rank-1 kernels with non-negative taps ("gaussian:RxS", "binomial", "square" ...): K[v][u] = a[v] * b[u]
    // to 1e-14, so the kw*kh-tap sum is evaluated as a row pass that keeps RAW double sums and a column
    // pass that normalises -- same double accumulation as MorphologyPrimitive's inner loop
    // (morphology.c:2837-2871) without the float rounding a two-kernel list would add between passes.
    size_t pivot = 0;
    bool nonneg = true;
    for (size_t i = 0; i < n; ++i) {
      if (win[i] < 0.0) nonneg = false;
      if (win[i] > win[pivot]) pivot = i;
    }
    if (nonneg && win[pivot] > 0.0) {
      const int pv = static_cast<int>(pivot) / kw, pu = static_cast<int>(pivot) % kw;
      std::vector<double> colf(kh), rowf(kw);
      for (int v = 0; v < kh; ++v) colf[v] = win[static_cast<size_t>(v) * kw + pu];
      for (int u = 0; u < kw; ++u) rowf[u] = win[static_cast<size_t>(pv) * kw + u] / win[pivot];
      bool rank1 = true;
      for (int v = 0; v < kh && rank1; ++v)
        for (int u = 0; u < kw; ++u)
          if (std::fabs(win[static_cast<size_t>(v) * kw + u] - colf[v] * rowf[u]) > 1.0e-14 * win[pivot]) { rank1 = false; break; }
      if (rank1) {
        StreamAlloc sums(s);
        int rc = sums.alloc(w * h * 4 * sizeof(double));
        if (rc == MB200_OK) {
          rc = launch_conv1d(src, static_cast<float *>(sums.ptr), w, h, ch, 0, rowf.data(), kw, ox, 0.0, 1.0, nullptr, s, 1);
          if (rc == MB200_OK)
            rc = launch_conv1d(static_cast<const float *>(sums.ptr), dst, w, h, ch, 1, colf.data(), kh, oy, 0.0, 1.0,
                               nullptr, s, 2);
          if (rc != MB200_EUNSUPPORTED) return rc;
        }
        // allocation failure / unsupported geometry: fall through to the direct 2-D kernel
      }
    }
  }
  if ((method == MB200_ErodeMorphology || method == MB200_DilateMorphology) && d_counter == nullptr &&
      !knobs().no_morph_stream) {                            // register-streaming kernel for the built-in shapes
    const int rc = launch_morph_stream(src, dst, w, h, ch, method, win.data(), kw, kh, ox, oy, s);
    if (rc != MB200_EUNSUPPORTED) return rc;
  }
  double gamma_scale = 1.0;
  if (method == MB200_ConvolveMorphology && kw == 1) {
    size_t count = 0;
    for (double v : win) if (!std::isnan(v)) ++count;
    if (count != 0) gamma_scale = static_cast<double>(kh) / static_cast<double>(count);
  }
  return launch_morph2d(src, dst, w, h, ch, method, win.data(), kw, kh, ox, oy, bias, gamma_scale, d_counter, s);
}

// Reads back a device counter (synchronises the stream) and converts it to the reference's
// `changed` = per-channel changes / number of Update channels (image-private.h:147).
int read_changed(unsigned long long *d_counter, int channels, long long *out, cudaStream_t s) {
  unsigned long long hostv = 0;
  cudaError_t e = cudaMemcpyAsync(&hostv, d_counter, sizeof(hostv), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return cuda_fail(e, "changed readback");
  *out = static_cast<long long>(hostv / static_cast<unsigned long long>(channels));
  return MB200_OK;
}

// 180-degree reflection used by Correlate/Close/Smooth (:3779-3793, RotateKernelInfo :4258)
mb200_kernel_info *reflected_clone(const mb200_kernel_info *kernel) {
  mb200_kernel_info *r = mb200_clone_kernel_info(kernel);
  if (r) rotate_kernel_info(r, 180.0);
  return r;
}

struct Stage { int primitive; const mb200_kernel_info *kernel; };

// HitAndMiss / Thinning / Thicken (morphology.c:3722-3729): `iterations` repeats the WHOLE method (every kernel of the list
// once per round) while anything changes.  Thinning / Thicken re-iterate: each kernel works on the previous kernel's
// result.  HitAndMiss unites the kernels' results with LightenCompositeOp (:4016-4052): the first result is kept, every
// further one -- computed from the ORIGINAL image -- is composited onto it; a one-kernel list iterates on its own result.
int apply_hit_and_miss_family(const float *src, float *dst, size_t w, size_t h, int ch, int method, long iterations,
                              const mb200_kernel_info *kernel, cudaStream_t s) {
  const size_t method_limit = iterations < 0 ? (w > h ? w : h) : static_cast<size_t>(iterations);
  const bool unite = method == MB200_HitAndMissMorphology && kernel->next != nullptr;
  const size_t bytes = w * h * static_cast<size_t>(ch) * sizeof(float), npix = w * h;
  StreamAlloc a(s), b(s), united(s), counter(s);
  int rc = a.alloc(bytes);
  if (!rc) rc = b.alloc(bytes);
  if (!rc && unite) rc = united.alloc(bytes);
  if (!rc) rc = counter.alloc(sizeof(unsigned long long));
  if (rc) return rc;
  float *bufs[2] = {static_cast<float *>(a.ptr), static_cast<float *>(b.ptr)};
  unsigned long long *d_counter = static_cast<unsigned long long *>(counter.ptr);
  const float *cur = src;
  int next = 0;
  bool have_united = false;
  size_t round = 0;
  long long round_changed = 1;
  while (round < method_limit && round_changed > 0) {
    ++round;
    round_changed = 0;
    for (const mb200_kernel_info *k = kernel; k; k = k->next) {
      const bool count = method_limit > 1;                   // the count only decides whether another round runs
      if (count) cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), s);
      rc = primitive(cur, bufs[next], w, h, ch, method, k, 0.0, count ? d_counter : nullptr, s);
      if (rc) return rc;
      if (count) {
        long long changed = 0;
        rc = read_changed(d_counter, ch, &changed, s);
        if (rc) return rc;
        round_changed += changed;
      }
      cur = bufs[next];
      next ^= 1;
      if (unite) {
        if (!have_united) {
          const cudaError_t e = cudaMemcpyAsync(united.ptr, cur, bytes, cudaMemcpyDeviceToDevice, s);
          if (e != cudaSuccess) return cuda_fail(e, "hit-and-miss: first result");
          have_united = true;
        } else {
          rc = launch_composite_lighten(static_cast<float *>(united.ptr), cur, npix, ch, s);
          if (rc) return rc;
        }
        cur = src;
      }
    }
  }
  const cudaError_t e = cudaMemcpyAsync(dst, unite ? united.ptr : cur, bytes, cudaMemcpyDeviceToDevice, s);
  return e == cudaSuccess ? MB200_OK : cuda_fail(e, "hit-and-miss: result copy");
}

int morphology_apply(const float *src, float *dst, size_t w, size_t h, int ch, int method, long iterations,
                     const mb200_kernel_info *kernel, double bias, cudaStream_t s,
                     const UnsharpEpilogue *epilogue = nullptr, bool *epilogue_fused = nullptr) {
  // Methods that end in "difference with the original" (staging :3813-3893, CompositeImage :3995-4012):
  // the morphological part is one of the methods below, then one Difference composite.
  if (method == MB200_EdgeInMorphology || method == MB200_EdgeOutMorphology || method == MB200_EdgeMorphology ||
      method == MB200_TopHatMorphology || method == MB200_BottomHatMorphology) {
    if (kernel->next != nullptr)
      return fail(MB200_EUNSUPPORTED, "compound difference methods with a multi-kernel list stay on the CPU path");
    const size_t npix = w * h;
    if (method == MB200_EdgeMorphology) {          // dilate; erode the ORIGINAL; canvas = eroded, source = dilated
      StreamAlloc dil(s);
      int rc = dil.alloc(npix * static_cast<size_t>(ch) * sizeof(float));
      if (rc) return rc;
      rc = morphology_apply(src, static_cast<float *>(dil.ptr), w, h, ch, MB200_DilateMorphology, iterations, kernel, bias, s);
      if (!rc) rc = morphology_apply(src, dst, w, h, ch, MB200_ErodeMorphology, iterations, kernel, bias, s);
      if (!rc) rc = launch_composite_difference(dst, static_cast<const float *>(dil.ptr), npix, ch, s);
      return rc;
    }
    const int base = method == MB200_EdgeInMorphology ? MB200_ErodeMorphology
                     : method == MB200_EdgeOutMorphology ? MB200_DilateMorphology
                     : method == MB200_TopHatMorphology ? MB200_OpenMorphology : MB200_CloseMorphology;
    int rc = morphology_apply(src, dst, w, h, ch, base, iterations, kernel, bias, s);
    if (!rc) rc = launch_composite_difference(dst, src, npix, ch, s);
    return rc;
  }
  if (iterations == 0) return fail(MB200_EINVAL, "iterations == 0 is a null operation (reference returns NULL)");
  if (method == MB200_HitAndMissMorphology || method == MB200_ThinningMorphology || method == MB200_ThickenMorphology)
    return apply_hit_and_miss_family(src, dst, w, h, ch, method, iterations, kernel, s);
  size_t kernel_limit = iterations < 0 ? (w > h ? w : h) : static_cast<size_t>(iterations);
  int stage_limit = 1;
  switch (method) {
    case MB200_SmoothMorphology: stage_limit = 4; break;
    case MB200_OpenMorphology: case MB200_CloseMorphology:
    case MB200_OpenIntensityMorphology: case MB200_CloseIntensityMorphology: stage_limit = 2; break;
    case MB200_ConvolveMorphology: case MB200_CorrelateMorphology:
    case MB200_ErodeMorphology: case MB200_DilateMorphology:
    case MB200_ErodeIntensityMorphology: case MB200_DilateIntensityMorphology:
    case MB200_IterativeDistanceMorphology: break;
    default:
      return fail(MB200_EUNSUPPORTED, "morphology method %d (Distance / Voronoi: sequential two-pass primitives) is not on "
                  "the GPU path", method);
  }
  mb200_kernel_info *reflected = nullptr;
  if (method == MB200_CorrelateMorphology || method == MB200_CloseMorphology || method == MB200_SmoothMorphology ||
      method == MB200_CloseIntensityMorphology) {
    reflected = reflected_clone(kernel);
    if (!reflected) return fail(MB200_ENOMEM, "kernel clone failed");
  }
  std::vector<Stage> stages;
  const mb200_kernel_info *rk = reflected;
  for (const mb200_kernel_info *nk = kernel; nk; nk = nk->next, rk = rk ? rk->next : nullptr) {
    for (int stage = 1; stage <= stage_limit; ++stage) {
      Stage st{method, nk};
      switch (method) {
        case MB200_OpenMorphology: st.primitive = stage == 2 ? MB200_DilateMorphology : MB200_ErodeMorphology; break;
        case MB200_CloseMorphology: st.kernel = rk; st.primitive = stage == 2 ? MB200_ErodeMorphology : MB200_DilateMorphology; break;
        case MB200_OpenIntensityMorphology:
          st.primitive = stage == 2 ? MB200_DilateIntensityMorphology : MB200_ErodeIntensityMorphology;
          break;
        case MB200_CloseIntensityMorphology:
          st.kernel = rk;
          st.primitive = stage == 2 ? MB200_ErodeIntensityMorphology : MB200_DilateIntensityMorphology;
          break;
        case MB200_SmoothMorphology:
          if (stage == 1) st.primitive = MB200_ErodeMorphology;
          else if (stage == 2) st.primitive = MB200_DilateMorphology;
          else if (stage == 3) { st.kernel = rk; st.primitive = MB200_DilateMorphology; }
          else { st.kernel = rk; st.primitive = MB200_ErodeMorphology; }
          break;
        case MB200_CorrelateMorphology: st.kernel = rk; st.primitive = MB200_ConvolveMorphology; break;
        default: break;
      }
      stages.push_back(st);
    }
  }
  const size_t bytes = w * h * static_cast<size_t>(ch) * sizeof(float);
  int rc = MB200_OK;
  if (kernel_limit == 1) {
    // Every stage runs exactly once: ping-pong between one temporary and dst so that the last
    // primitive writes dst (no trailing copy).
    const size_t total = stages.size();
    StreamAlloc tmp(s);
    if (total > 1) { rc = tmp.alloc(bytes); }
    const float *cur = src;
    for (size_t i = 0; i < total && rc == MB200_OK; ++i) {
      float *out = ((total - 1 - i) % 2 == 0) ? dst : static_cast<float *>(tmp.ptr);
      const bool last = i + 1 == total;
      rc = primitive(cur, out, w, h, ch, stages[i].primitive, stages[i].kernel, bias, nullptr, s,
                     last ? epilogue : nullptr, last ? epilogue_fused : nullptr);
      cur = out;
    }
  } else {
    // `changed`-driven iteration (:3919-3962): needs the count after every primitive.
    StreamAlloc a(s), b(s), counter(s);
    rc = a.alloc(bytes);
    if (!rc) rc = b.alloc(bytes);
    if (!rc) rc = counter.alloc(sizeof(unsigned long long));
    const float *cur = src;
    float *bufs[2] = {static_cast<float *>(a.ptr), static_cast<float *>(b.ptr)};
    int next = 0;
    for (size_t i = 0; i < stages.size() && rc == MB200_OK; ++i) {
      size_t loop = 0;
      long long changed = 1;
      while (loop < kernel_limit && changed > 0 && rc == MB200_OK) {
        ++loop;
        cudaMemsetAsync(counter.ptr, 0, sizeof(unsigned long long), s);
        rc = primitive(cur, bufs[next], w, h, ch, stages[i].primitive, stages[i].kernel, bias,
                       static_cast<unsigned long long *>(counter.ptr), s);
        if (rc) break;
        rc = read_changed(static_cast<unsigned long long *>(counter.ptr), ch, &changed, s);
        cur = bufs[next];
        next ^= 1;
      }
    }
    if (rc == MB200_OK) {
      cudaError_t e = cudaMemcpyAsync(dst, cur, bytes, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) rc = cuda_fail(e, "result copy");
    }
  }
  mb200_destroy_kernel_info(reflected);
  return rc;
}

// Runs `op(d_src, d_dst, stream)` on the HBM copies of host buffers (cache.cu): an attached pixel cache keeps its HBM
// copy between calls (and, in lazy mode, its result stays there until mb200_cache_sync); anything else is staged
// through stream-ordered temporaries.  Pageable memory travels through the threaded pinned bounce ring.
template <typename Op>
int with_staging(const float *src, size_t src_bytes, float *dst, size_t dst_bytes, Op op) {
  cudaStream_t s;
  int rc = prepare(nullptr, &s);
  if (rc) return rc;
  StageRef in, out;
  rc = stage_input(src, src_bytes, s, &in);
  if (!rc) rc = stage_output(dst, dst_bytes, s, &out);
  if (!rc) rc = op(static_cast<const float *>(in.dev), static_cast<float *>(out.dev), s);
  if (rc) cudaStreamSynchronize(s);
  else rc = finish_output(&out, s);
  release_stage(&in, s);
  release_stage(&out, s);
  return rc;
}

// In-place operators on a host buffer.
template <typename Op>
int in_place_host(float *buf, size_t bytes, Op op) {
  cudaStream_t s;
  int rc = prepare(nullptr, &s);
  if (rc) return rc;
  StageRef io;
  rc = stage_input(buf, bytes, s, &io);
  if (!rc) rc = op(static_cast<float *>(io.dev), s);
  if (rc) cudaStreamSynchronize(s);
  else rc = finish_output(&io, s);
  release_stage(&io, s);
  return rc;
}

}  // namespace

extern "C" {

// Test / developer hook: the switches above (initialised from the MB200_* environment variables of the same upper-case
// names) can be flipped at run time, e.g. to compare a specialised kernel with the generic one in one process.
int mb200_set_option(const char *name, int value) {
  if (!name) return fail(MB200_EINVAL, "set_option: null name");
  Knobs &k = knobs();
  const bool v = value != 0;
  const std::string n(name);
  if (n == "no_rank1") k.no_rank1 = v;
  else if (n == "no_morph_stream") k.no_morph_stream = v;
  else if (n == "no_resize_stream") k.no_resize_stream = v;
  else if (n == "resize_regular_h") k.resize_regular_h = v;
  else if (n == "no_fused_unsharp") k.no_fused_unsharp = v;
  else if (n == "resize_fused") k.resize_fused = v;
  else if (n == "conv_mma") set_conv_mma(value);
  else return fail(MB200_EINVAL, "set_option: unknown option '%s'", name);
  return MB200_OK;
}

int mb200_get_option(const char *name, int *value) {
  if (!name || !value) return fail(MB200_EINVAL, "get_option: null argument");
  const Knobs &k = knobs();
  const std::string n(name);
  if (n == "no_rank1") *value = k.no_rank1;
  else if (n == "no_morph_stream") *value = k.no_morph_stream;
  else if (n == "no_resize_stream") *value = k.no_resize_stream;
  else if (n == "resize_regular_h") *value = k.resize_regular_h;
  else if (n == "no_fused_unsharp") *value = k.no_fused_unsharp;
  else if (n == "resize_fused") *value = k.resize_fused;
  else if (n == "conv_mma") *value = conv_mma_enabled();
  else if (n == "conv_mma_launches") *value = static_cast<int>(conv_mma_launches() & 0x7fffffff);
  else return fail(MB200_EINVAL, "get_option: unknown option '%s'", name);
  return MB200_OK;
}

int mb200_morphology_primitive_dev(const float *src, float *dst, size_t width, size_t height, int channels,
                                   int method, const mb200_kernel_info *kernel, double bias, long long *changed,
                                   void *stream) {
  if (!src || !dst || !kernel || !kernel->values || !valid_image(width, height, channels))
    return fail(MB200_EINVAL, "morphology_primitive: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  if (!changed) return primitive(src, dst, width, height, channels, method, kernel, bias, nullptr, s);
  StreamAlloc counter(s);
  rc = counter.alloc(sizeof(unsigned long long));
  if (rc) return rc;
  cudaMemsetAsync(counter.ptr, 0, sizeof(unsigned long long), s);
  rc = primitive(src, dst, width, height, channels, method, kernel, bias,
                 static_cast<unsigned long long *>(counter.ptr), s);
  if (rc) return rc;
  return read_changed(static_cast<unsigned long long *>(counter.ptr), channels, changed, s);
}

int mb200_morphology_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
                               int method, long iterations, const mb200_kernel_info *kernel, double bias,
                               void *stream) {
  if (!src || !dst || src == dst || !kernel || !valid_image(width, height, channels))
    return fail(MB200_EINVAL, "morphology_image: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return morphology_apply(src, dst, width, height, channels, method, iterations, kernel, bias, s);
}

int mb200_convolve_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
                             const mb200_kernel_info *kernel, void *stream) {
  return mb200_morphology_image_dev(src, dst, width, height, channels, MB200_ConvolveMorphology, 1, kernel, 0.0,
                                    stream);
}

int mb200_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                         double sigma, void *stream) {
  // effect.c:788: "blur:RxS;blur:RxS+90"
  mb200_kernel_info *k = mb200_acquire_kernel_builtin(MB200_BlurKernel, radius, sigma, 0.0, 0.0);
  if (!k) return fail(MB200_ENOMEM, "blur kernel");
  k->next = mb200_acquire_kernel_builtin(MB200_BlurKernel, radius, sigma, 90.0, 0.0);
  if (!k->next) { mb200_destroy_kernel_info(k); return fail(MB200_ENOMEM, "blur kernel"); }
  const int rc = mb200_convolve_image_dev(src, dst, width, height, channels, k, stream);
  mb200_destroy_kernel_info(k);
  return rc;
}

int mb200_gaussian_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
                                  double radius, double sigma, void *stream) {
  mb200_kernel_info *k = mb200_acquire_kernel_builtin(MB200_GaussianKernel, radius, sigma, 0.0, 0.0);
  if (!k) return fail(MB200_ENOMEM, "gaussian kernel");
  const int rc = mb200_convolve_image_dev(src, dst, width, height, channels, k, stream);
  mb200_destroy_kernel_info(k);
  return rc;
}

int mb200_unsharp_mask_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
                                 double radius, double sigma, double gain, double threshold, void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels))
    return fail(MB200_EINVAL, "unsharp: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  // BlurImage (effect.c:4295) + the point pass (:4310-4384).  With the RGBA pair kernels the point pass is the
  // epilogue of the blur's column pass: it is applied to the float-ROUNDED blur value, exactly what the reference
  // reads back from its blurred image, so the fused and the two-launch forms produce the same bits while the fused one
  // saves the 48 B/pixel of the separate pass.
  mb200_kernel_info *k = mb200_acquire_kernel_builtin(MB200_BlurKernel, radius, sigma, 0.0, 0.0);
  if (!k) return fail(MB200_ENOMEM, "blur kernel");
  k->next = mb200_acquire_kernel_builtin(MB200_BlurKernel, radius, sigma, 90.0, 0.0);
  if (!k->next) { mb200_destroy_kernel_info(k); return fail(MB200_ENOMEM, "blur kernel"); }
  const UnsharpEpilogue epi{src, gain, 65535.0 * threshold};
  bool fused = false;
  rc = morphology_apply(src, dst, width, height, channels, MB200_ConvolveMorphology, 1, k, 0.0, s,
                        knobs().no_fused_unsharp ? nullptr : &epi, &fused);
  mb200_destroy_kernel_info(k);
  if (rc || fused) return rc;
  return launch_unsharp_combine(src, dst, width * height * static_cast<size_t>(channels), gain,
                                65535.0 * threshold, s);
}

int mb200_resize_image_dev(const float *src, size_t width, size_t height, int channels, float *dst,
                           size_t out_width, size_t out_height, int filter, void *stream) {
  return mb200_resize_image_ex_dev(src, width, height, channels, dst, out_width, out_height, filter, nullptr, stream);
}

int mb200_resize_image_ex_dev(const float *src, size_t width, size_t height, int channels, float *dst,
                              size_t out_width, size_t out_height, int filter, const mb200_filter_options *options,
                              void *stream) {
  mb200_filter_options opt{};           // normalised copy: the cache compares the bytes
  if (options) {
    opt.set = options->set;
    if (opt.set & MB200_FO_WINDOW) { opt.window = options->window; opt.keep_filter = options->keep_filter ? 1 : 0; }
    if (opt.set & MB200_FO_LOBES) opt.lobes = options->lobes;
    if (opt.set & MB200_FO_SIGMA) opt.sigma = options->sigma;
    if (opt.set & MB200_FO_KAISER_BETA) opt.kaiser_beta = options->kaiser_beta;
    if (opt.set & MB200_FO_BLUR) opt.blur = options->blur;
    if (opt.set & MB200_FO_SUPPORT) opt.support = options->support;
    if (opt.set & MB200_FO_WIN_SUPPORT) opt.win_support = options->win_support;
    if (opt.set & MB200_FO_B) opt.b = options->b;
    if (opt.set & MB200_FO_C) opt.c = options->c;
  }

  if (!src || !dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "resize: bad arguments");
  if (out_width == 0 || out_height == 0) return fail(MB200_EINVAL, "NegativeOrZeroImageSize");   // :3791
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  const size_t px = static_cast<size_t>(channels) * sizeof(float);
  if (out_width == width && out_height == height && filter == MB200_UndefinedFilter) {            // :3793-3795
    cudaError_t e = cudaMemcpyAsync(dst, src, width * height * px, cudaMemcpyDeviceToDevice, s);
    return e == cudaSuccess ? MB200_OK : cuda_fail(e, "resize: clone");
  }
  auto reciprocal = [](double x) { return std::fabs(x) >= 1.0e-12 ? 1.0 / x : (x < 0 ? -1.0e12 : 1.0e12); };
  const double x_factor = static_cast<double>(out_width) * reciprocal(static_cast<double>(width));
  const double y_factor = static_cast<double>(out_height) * reciprocal(static_cast<double>(height));
  int filter_type = MB200_LanczosFilter;                                                           // :3806-3816
  if (filter != MB200_UndefinedFilter) filter_type = filter;
  else if (x_factor == 1.0 && y_factor == 1.0) filter_type = MB200_PointFilter;
  else if (has_alpha(channels) || (x_factor * y_factor) > 1.0) filter_type = MB200_MitchellFilter;

  auto get_tables = [&](size_t in_n, size_t out_n, double factor, std::shared_ptr<ResizeTables> *out) -> int {
    int dev = 0;
    cudaGetDevice(&dev);
    auto find = [&]() -> std::shared_ptr<ResizeTables> {
      for (size_t i = 0; i < g_tables.size(); ++i) {
        const std::shared_ptr<ResizeTables> &t = g_tables[i];
        if (t->device == dev && t->filter == filter_type && t->in_n == in_n && t->out_n == out_n &&
            std::memcmp(&t->options, &opt, sizeof(opt)) == 0) {
          std::shared_ptr<ResizeTables> hit = t;
          if (i + 1 != g_tables.size()) { g_tables.erase(g_tables.begin() + i); g_tables.push_back(hit); }   // most recent last
          return hit;
        }
      }
      return nullptr;
    };
    {
      std::lock_guard<std::mutex> lock(g_tables_mutex);
      if ((*out = find())) return MB200_OK;
    }
    const long taps = mb200_resize_contributions_ex(filter_type, &opt, in_n, out_n, factor, nullptr, nullptr, nullptr, 0);
    if (taps < 0) return static_cast<int>(taps);
    std::vector<long> start(out_n);
    std::vector<int> istart(out_n), count(out_n);
    std::vector<double> w(out_n * static_cast<size_t>(taps)), wt(out_n * static_cast<size_t>(taps)), wreg, wsets;
    std::vector<int> border;
    const long r = mb200_resize_contributions_ex(filter_type, &opt, in_n, out_n, factor, start.data(), count.data(), w.data(),
                                                 static_cast<size_t>(taps));
    if (r < 0) return static_cast<int>(r);
    std::shared_ptr<ResizeTables> t = std::make_shared<ResizeTables>();
    t->device = dev; t->filter = filter_type; t->options = opt; t->in_n = in_n; t->out_n = out_n; t->taps = taps;
    // widest source span of any aligned block of 32 outputs (tile width of the tiled horizontal kernel)
    for (size_t o = 0; o < out_n; o += 32) {
      const size_t last = std::min(o + 32, out_n) - 1;
      long hi = 0;
      for (size_t k = o; k <= last; ++k) hi = std::max(hi, start[k] + count[k]);
      t->max_span = std::max(t->max_span, static_cast<int>(hi - start[o]));
    }
    // regular pattern (integer-ratio reduction): constant tap count and window stride in the interior
    if (out_n >= 64) {
      const size_t mid = out_n / 2;
      const int n = count[mid];
      const long st = start[mid + 1] - start[mid];
      size_t regular = 0;
      for (size_t o = 0; o + 1 < out_n; ++o)
        if (count[o] == n && count[o + 1] == n && start[o + 1] - start[o] == st) ++regular;
      if (st >= 2 && n > 0 && regular * 10 >= out_n * 9) {
        t->reg_stride = static_cast<int>(st);
        t->reg_taps = n;
        wreg.assign(out_n * static_cast<size_t>(n), 0.0);
        for (size_t o = 0; o < out_n; ++o)
          for (int j = 0; j < n && j < count[o]; ++j) wreg[o * n + j] = w[o * taps + j];
        // longest run [lo, lo+len) with count == n and start advancing by st
        // runs of outputs whose weights are bit-identical (same binade of bisect, resize.c:3398-3443)
        struct Run { size_t lo, len; };
        std::vector<Run> runs;
        size_t lo = 0;
        for (size_t o = 1; o <= out_n; ++o) {
          const bool same = o < out_n && count[o] == n && count[lo] == n && start[o] - start[o - 1] == st &&
                            std::memcmp(&w[o * taps], &w[lo * taps], static_cast<size_t>(n) * sizeof(double)) == 0;
          if (!same) {
            if (count[lo] == n && o - lo >= 8) runs.push_back({lo, o - lo});
            lo = o;
          }
        }
        std::sort(runs.begin(), runs.end(), [](const Run &a, const Run &b) { return a.len > b.len; });
        if (runs.size() > MB200_RESIZE_MAX_SEGMENTS) runs.resize(MB200_RESIZE_MAX_SEGMENTS);
        std::sort(runs.begin(), runs.end(), [](const Run &a, const Run &b) { return a.lo < b.lo; });
        size_t covered = 0;
        for (const Run &r : runs) covered += r.len;
        if (!runs.empty() && covered * 10 >= out_n * 6) {
          long next = 0;
          for (const Run &r : runs) {
            t->seg_o[t->nseg] = static_cast<int>(r.lo);
            t->seg_n[t->nseg] = static_cast<int>(r.len);
            t->seg_src[t->nseg] = static_cast<int>(start[r.lo]);
            ++t->nseg;
            wsets.insert(wsets.end(), &w[r.lo * taps], &w[r.lo * taps] + n);
            for (long o = next; o < static_cast<long>(r.lo); ++o) border.push_back(static_cast<int>(o));
            next = static_cast<long>(r.lo + r.len);
          }
          for (long o = next; o < static_cast<long>(out_n); ++o) border.push_back(static_cast<int>(o));
          t->nborder = static_cast<int>(border.size());
          if (border.empty()) border.push_back(0);      // keep the buffer non-null
        }
      }
    }
    std::vector<int> tiles_x, tiles_y;
    std::vector<unsigned char> is_border;
    if (t->nseg > 0) {
      int tw = 0, th = 0;
      resize_fused_tile(t->reg_stride, t->reg_taps, &tw, &th);
      if (tw > 0) {
        auto cut = [&](int tile, std::vector<int> &out) {
          for (int k = 0; k < t->nseg; ++k)
            for (int rel = 0; rel < t->seg_n[k]; rel += tile) {
              out.push_back(t->seg_o[k] + rel);
              out.push_back(std::min(tile, t->seg_n[k] - rel));
              out.push_back(t->seg_src[k] + t->reg_stride * rel);
              out.push_back(k);
            }
        };
        cut(tw, tiles_x);
        cut(th, tiles_y);
        t->ntiles_x = static_cast<int>(tiles_x.size() / 4);
        t->ntiles_y = static_cast<int>(tiles_y.size() / 4);
        is_border.assign(out_n, 0);
        for (int k = 0; k < t->nborder; ++k) is_border[static_cast<size_t>(border[k])] = 1;
      }
    }
    for (size_t o = 0; o < out_n; ++o) {          // tap-major transpose for coalesced weight loads
      istart[o] = static_cast<int>(start[o]);
      for (long j = 0; j < taps; ++j) wt[static_cast<size_t>(j) * out_n + o] = w[o * taps + j];
    }
    cudaError_t e = cudaMalloc(&t->d_start, out_n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&t->d_count, out_n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&t->d_weights, wt.size() * sizeof(double));
    if (e == cudaSuccess && !wreg.empty()) e = cudaMalloc(&t->d_wreg, wreg.size() * sizeof(double));
    if (e == cudaSuccess && !wsets.empty()) e = cudaMalloc(&t->d_wsets, wsets.size() * sizeof(double));
    if (e == cudaSuccess && !wsets.empty())
      e = cudaMemcpy(t->d_wsets, wsets.data(), wsets.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !border.empty()) e = cudaMalloc(&t->d_border, border.size() * sizeof(int));
    if (e == cudaSuccess && !border.empty())
      e = cudaMemcpy(t->d_border, border.data(), border.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_start, istart.data(), out_n * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_count, count.data(), out_n * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_weights, wt.data(), wt.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !wreg.empty())
      e = cudaMemcpy(t->d_wreg, wreg.data(), wreg.size() * sizeof(double), cudaMemcpyHostToDevice);
    auto upload = [&](const void *host, size_t bytes, void **dev) {
      if (e == cudaSuccess && bytes) e = cudaMalloc(dev, bytes);
      if (e == cudaSuccess && bytes) e = cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice);
    };
    upload(tiles_x.data(), tiles_x.size() * sizeof(int), reinterpret_cast<void **>(&t->d_tiles_x));
    upload(tiles_y.data(), tiles_y.size() * sizeof(int), reinterpret_cast<void **>(&t->d_tiles_y));
    upload(is_border.data(), is_border.size(), reinterpret_cast<void **>(&t->d_is_border));
    if (e != cudaSuccess) return cuda_fail(e, "resize: table upload");     // ~ResizeTables frees what was allocated
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    if ((*out = find())) return MB200_OK;          // another thread built the same table meanwhile: keep theirs
    if (g_tables.size() >= kMaxCachedTables) g_tables.erase(g_tables.begin());     // least recently used
    g_tables.push_back(t);
    *out = t;
    return MB200_OK;
  };
  std::shared_ptr<ResizeTables> tx_owner, ty_owner;
  rc = get_tables(width, out_width, x_factor, &tx_owner);
  if (!rc) rc = get_tables(height, out_height, y_factor, &ty_owner);
  if (rc) return rc;
  const ResizeTables *tx = tx_owner.get(), *ty = ty_owner.get();
  StreamAlloc tmp(s);
  const bool reg_h = knobs().resize_regular_h;         // tiled regular H kernel: opt-in (r01: slower)
  const bool no_stream = knobs().no_resize_stream;
  auto run_axis = [&](const float *in, size_t w, size_t h, float *out, int axis, const ResizeTables *t) -> int {
    if (channels == 4 && t->d_wsets != nullptr && !no_stream) {        // streaming kernels (+ border gather CTAs)
      const int rs = launch_resize_stream(in, w, h, out, t->out_n, axis, t->reg_stride, t->reg_taps, t->nseg, t->seg_o,
                                          t->seg_n, t->seg_src, t->d_wsets, t->nborder, t->d_border, t->d_start,
                                          t->d_count, t->d_weights, s);
      if (rs != MB200_EUNSUPPORTED) return rs;
    }
    const bool use_reg = t->d_wreg != nullptr && (axis == 1 || reg_h);
    return launch_resize_axis(in, w, h, channels, out, t->out_n, axis, t->d_start, t->d_count, t->d_weights,
                              static_cast<int>(t->taps), t->max_span, use_reg ? t->reg_stride : 0, t->reg_taps,
                              use_reg ? t->d_wreg : nullptr, s);
  };
  // Equal integer reduction on both axes (the reference filters vertically first when x_factor <= y_factor, :3854-3861):
  // one fused launch keeps the vertically filtered intermediate of every output tile in shared memory.  Parity-green and
  // bit-identical to the two passes, DRAM traffic 1.32 GB instead of 2.45 GB for 8192^2 -> 4096^2 -- but 0.565 ms against
  // 0.489 ms: the passes are bound by the FP64 and conversion (XU) pipes, not by HBM, and the fused kernel adds 17 % of
  // halo work.  Opt-in (MB200_RESIZE_FUSED=1 / mb200_set_option("resize_fused", 1)); A/B in profiles/r02_resize_fused.md.
  if (channels == 4 && x_factor == y_factor && !no_stream && knobs().resize_fused && tx->ntiles_x > 0 && ty->ntiles_y > 0 &&
      tx->reg_stride == ty->reg_stride && tx->reg_taps == ty->reg_taps) {
    rc = launch_resize_fused(src, width, height, dst, out_width, out_height, tx->reg_stride, tx->reg_taps, tx->d_tiles_x,
                             tx->ntiles_x, ty->d_tiles_y, ty->ntiles_y, tx->d_wsets, ty->d_wsets, tx->d_start, tx->d_count,
                             tx->d_weights, ty->d_start, ty->d_count, ty->d_weights, tx->d_border, tx->nborder, ty->d_border,
                             ty->nborder, ty->d_is_border, s);
    if (rc != MB200_EUNSUPPORTED) return rc;
  }
  if (x_factor > y_factor) {                                                                       // :3846-3853
    rc = tmp.alloc(out_width * height * px);
    if (rc) return rc;
    rc = run_axis(src, width, height, static_cast<float *>(tmp.ptr), 0, tx);
    if (rc) return rc;
    rc = run_axis(static_cast<float *>(tmp.ptr), out_width, height, dst, 1, ty);
  } else {                                                                                         // :3854-3861
    rc = tmp.alloc(width * out_height * px);
    if (rc) return rc;
    rc = run_axis(src, width, height, static_cast<float *>(tmp.ptr), 1, ty);
    if (rc) return rc;
    rc = run_axis(static_cast<float *>(tmp.ptr), width, out_height, dst, 0, tx);
  }
  return rc;
}

int mb200_transform_colorspace_ex_dev(float *buf, size_t width, size_t height, int channels, int from, int to,
                                      const mb200_colorspace_options *options, void *stream) {
  if (!buf || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "colorspace: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_colorspace(buf, width * height, channels, from, to, options, s);
}

int mb200_transform_colorspace_dev(float *buf, size_t width, size_t height, int channels, int from, int to,
                                   void *stream) {
  return mb200_transform_colorspace_ex_dev(buf, width, height, channels, from, to, nullptr, stream);
}

// ------------------------------------------------------------ host buffers

int mb200_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_blur_image_dev(s, d, w, h, ch, radius, sigma, st);
  });
}

int mb200_gaussian_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius,
                              double sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "gaussian blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_gaussian_blur_image_dev(s, d, w, h, ch, radius, sigma, st);
  });
}

int mb200_convolve_image(const float *src, float *dst, size_t w, size_t h, int ch, const mb200_kernel_info *kernel) {
  if (!src || !dst || !kernel || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "convolve: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_convolve_image_dev(s, d, w, h, ch, kernel, st);
  });
}

int mb200_morphology_image(const float *src, float *dst, size_t w, size_t h, int ch, int method, long iterations,
                           const mb200_kernel_info *kernel, double bias) {
  if (!src || !dst || !kernel || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "morphology: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_morphology_image_dev(s, d, w, h, ch, method, iterations, kernel, bias, st);
  });
}

int mb200_unsharp_mask_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma,
                             double gain, double threshold) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "unsharp: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_unsharp_mask_image_dev(s, d, w, h, ch, radius, sigma, gain, threshold, st);
  });
}

int mb200_resize_image_ex(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh, int filter,
                          const mb200_filter_options *options) {
  if (!src || !dst || !valid_image(w, h, ch) || ow == 0 || oh == 0) return fail(MB200_EINVAL, "resize: bad arguments");
  return with_staging(src, w * h * ch * sizeof(float), dst, ow * oh * ch * sizeof(float),
                      [&](const float *s, float *d, cudaStream_t st) {
                        return mb200_resize_image_ex_dev(s, w, h, ch, d, ow, oh, filter, options, st);
                      });
}
int mb200_resize_image(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh, int filter) {
  return mb200_resize_image_ex(src, w, h, ch, dst, ow, oh, filter, nullptr);
}

int mb200_transform_colorspace_ex(float *buf, size_t w, size_t h, int ch, int from, int to,
                                  const mb200_colorspace_options *options) {
  if (!buf || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "colorspace: bad arguments");
  return in_place_host(buf, w * h * ch * sizeof(float),
                       [&](float *d, cudaStream_t st) { return launch_colorspace(d, w * h, ch, from, to, options, st); });
}

int mb200_transform_colorspace(float *buf, size_t w, size_t h, int ch, int from, int to) {
  return mb200_transform_colorspace_ex(buf, w, h, ch, from, to, nullptr);
}

}  // extern "C"

// ---- threshold.c point operators ---------------------------------------------------------
namespace {
// The ParseGeometry step of Black/WhiteThresholdImage (threshold.c:955-985) for "v[,v[,v[,v]]][%]".
int parse_thresholds(const char *spec, double (&t)[4]) {
  if (spec == nullptr) return fail(MB200_EINVAL, "thresholds == NULL (the reference returns MagickTrue untouched)");
  double v[4] = {0, 0, 0, 0};
  int n = 0;
  bool percent = false;
  const char *p = spec;
  while (*p) {
    while (*p == ' ') ++p;
    if (*p == '\0') break;
    if (n == 4) return fail(MB200_EUNSUPPORTED, "threshold geometry '%s' has more than four values", spec);
    char *end = nullptr;
    v[n] = std::strtod(p, &end);
    if (end == p) return fail(MB200_EUNSUPPORTED, "threshold geometry '%s' is not of the form v[,v[,v[,v]]][%%]", spec);
    ++n;
    p = end;
    while (*p == ' ') ++p;
    if (*p == '%') { percent = true; ++p; }
    while (*p == ' ') ++p;
    if (*p == ',') ++p;
    else if (*p != '\0') return fail(MB200_EUNSUPPORTED, "threshold geometry '%s' is not of the form v[,v[,v[,v]]][%%]", spec);
  }
  if (n == 0) return fail(MB200_EUNSUPPORTED, "empty threshold geometry");
  t[0] = v[0];
  t[1] = n > 1 ? v[1] : v[0];
  t[2] = n > 2 ? v[2] : v[0];
  t[3] = n > 3 ? v[3] : 100.0;
  if (percent)
    for (double &x : t) x *= (65535.0 / 100.0);
  return MB200_OK;
}

int threshold_dev(float *buf, size_t width, size_t height, int channels, int op, const double *t, void *stream) {
  if (!buf || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "threshold: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_threshold(buf, width * height, channels, op, t, s);
}

int black_white_dev(float *buf, size_t width, size_t height, int channels, int colorspace, const char *thresholds, int op,
                    void *stream) {
  if (!buf || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "threshold: bad arguments");
  if (channels < 3)
    return fail(MB200_EUNSUPPORTED, "Black/WhiteThresholdImage promote gray images to sRGB (threshold.c:949)");
  if (colorspace == MB200_RGBColorspace)
    return fail(MB200_EUNSUPPORTED, "linear RGB: the intensity needs EncodePixelGamma (pixel.c:2421)");
  double t[4];
  const int rc = parse_thresholds(thresholds, t);
  if (rc) return rc;
  return threshold_dev(buf, width, height, channels, op, t, stream);
}

}  // namespace

extern "C" {

int mb200_bilevel_image_dev(float *buf, size_t width, size_t height, int channels, double threshold, void *stream) {
  const double t[4] = {threshold, 0, 0, 0};
  return threshold_dev(buf, width, height, channels, 0, t, stream);
}
int mb200_black_threshold_image_dev(float *buf, size_t width, size_t height, int channels, int colorspace,
                                    const char *thresholds, void *stream) {
  return black_white_dev(buf, width, height, channels, colorspace, thresholds, 1, stream);
}
int mb200_white_threshold_image_dev(float *buf, size_t width, size_t height, int channels, int colorspace,
                                    const char *thresholds, void *stream) {
  return black_white_dev(buf, width, height, channels, colorspace, thresholds, 2, stream);
}
int mb200_clamp_image_dev(float *buf, size_t width, size_t height, int channels, void *stream) {
  return threshold_dev(buf, width, height, channels, 3, nullptr, stream);
}

int mb200_bilevel_image(float *buf, size_t w, size_t h, int ch, double threshold) {
  if (!buf || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "bilevel: bad arguments");
  return in_place_host(buf, w * h * ch * sizeof(float),
                       [&](float *d, cudaStream_t st) { return mb200_bilevel_image_dev(d, w, h, ch, threshold, st); });
}
int mb200_black_threshold_image(float *buf, size_t w, size_t h, int ch, int colorspace, const char *thresholds) {
  if (!buf || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "black threshold: bad arguments");
  double t[4];
  if (ch < 3 || colorspace == MB200_RGBColorspace || parse_thresholds(thresholds, t) != MB200_OK)
    return black_white_dev(buf, w, h, ch, colorspace, thresholds, 1, nullptr);   // reports the decline before staging
  return in_place_host(buf, w * h * ch * sizeof(float), [&](float *d, cudaStream_t st) {
    return mb200_black_threshold_image_dev(d, w, h, ch, colorspace, thresholds, st);
  });
}
int mb200_white_threshold_image(float *buf, size_t w, size_t h, int ch, int colorspace, const char *thresholds) {
  if (!buf || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "white threshold: bad arguments");
  double t[4];
  if (ch < 3 || colorspace == MB200_RGBColorspace || parse_thresholds(thresholds, t) != MB200_OK)
    return black_white_dev(buf, w, h, ch, colorspace, thresholds, 2, nullptr);
  return in_place_host(buf, w * h * ch * sizeof(float), [&](float *d, cudaStream_t st) {
    return mb200_white_threshold_image_dev(d, w, h, ch, colorspace, thresholds, st);
  });
}
int mb200_clamp_image(float *buf, size_t w, size_t h, int ch) {
  if (!buf || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "clamp: bad arguments");
  return in_place_host(buf, w * h * ch * sizeof(float),
                       [&](float *d, cudaStream_t st) { return mb200_clamp_image_dev(d, w, h, ch, st); });
}

}  // extern "C"

// ---- SharpenImage / EdgeImage: effect.c builds a kernel inline and calls ConvolveImage ------------------------
extern "C" {

int mb200_sharpen_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                            double sigma, void *stream) {
  mb200_kernel_info *k = mb200_sharpen_kernel(radius, sigma);
  if (!k) return fail(MB200_ENOMEM, "sharpen kernel");
  const int rc = mb200_convolve_image_dev(src, dst, width, height, channels, k, stream);
  mb200_destroy_kernel_info(k);
  return rc;
}

int mb200_edge_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                         void *stream) {
  mb200_kernel_info *k = mb200_edge_kernel(radius);
  if (!k) return fail(MB200_ENOMEM, "edge kernel");
  const int rc = mb200_convolve_image_dev(src, dst, width, height, channels, k, stream);
  mb200_destroy_kernel_info(k);
  return rc;
}

int mb200_sharpen_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "sharpen: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_sharpen_image_dev(s, d, w, h, ch, radius, sigma, st);
  });
}

int mb200_edge_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "edge: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_edge_image_dev(s, d, w, h, ch, radius, st);
  });
}

}  // extern "C"

// ---- Copy-trait channels of a `-channel` selection ------------------------------------------------------------------------
extern "C" int mb200_resize_nearest(int filter, size_t in_n, size_t out_n, double factor, long *nearest);

extern "C" {

int mb200_restore_channels_dev(float *dst, const float *src, size_t width, size_t height, int channels, unsigned update_mask,
                               void *stream) {
  if (!dst || !src || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "restore channels: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  if ((update_mask & ((1u << channels) - 1u)) == ((1u << channels) - 1u)) return MB200_OK;      // nothing to restore
  return launch_restore_channels(dst, src, width * height, channels, update_mask, s);
}

int mb200_resize_copy_channels_dev(const float *src, size_t width, size_t height, int channels, float *dst, size_t out_width,
                                   size_t out_height, int filter, unsigned update_mask, void *stream) {
  if (!dst || !src || !valid_image(width, height, channels) || out_width == 0 || out_height == 0)
    return fail(MB200_EINVAL, "resize copy channels: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  if ((update_mask & ((1u << channels) - 1u)) == ((1u << channels) - 1u)) return MB200_OK;
  auto reciprocal = [](double x) { return std::fabs(x) >= 1.0e-12 ? 1.0 / x : (x < 0 ? -1.0e12 : 1.0e12); };
  const double x_factor = static_cast<double>(out_width) * reciprocal(static_cast<double>(width));
  const double y_factor = static_cast<double>(out_height) * reciprocal(static_cast<double>(height));
  int filter_type = MB200_LanczosFilter;                                                           // resize.c:3806-3816
  if (filter != MB200_UndefinedFilter) filter_type = filter;
  else if (x_factor == 1.0 && y_factor == 1.0) filter_type = MB200_PointFilter;
  else if (has_alpha(channels) || (x_factor * y_factor) > 1.0) filter_type = MB200_MitchellFilter;
  std::vector<long> nx(out_width), ny(out_height);
  rc = mb200_resize_nearest(filter_type, width, out_width, x_factor, nx.data());
  if (!rc) rc = mb200_resize_nearest(filter_type, height, out_height, y_factor, ny.data());
  if (rc) return rc;
  std::vector<int> table(out_width + out_height);
  for (size_t i = 0; i < out_width; ++i) table[i] = static_cast<int>(nx[i]);
  for (size_t i = 0; i < out_height; ++i) table[out_width + i] = static_cast<int>(ny[i]);
  StreamAlloc d_table(s);
  rc = d_table.alloc(table.size() * sizeof(int));
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(d_table.ptr, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cuda_fail(e, "resize copy channels: table upload");
  const int *d = static_cast<const int *>(d_table.ptr);
  return launch_resize_copy_channels(dst, src, width, out_width, out_height, channels, d, d + out_width, update_mask, s);
}

// host buffers: dst already holds the operator's result (resident when attached), src is the operator's source
int mb200_restore_channels(float *dst, const float *src, size_t w, size_t h, int ch, unsigned update_mask) {
  if (!dst || !src || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "restore channels: bad arguments");
  cudaStream_t s;
  int rc = prepare(nullptr, &s);
  if (rc) return rc;
  const size_t bytes = w * h * ch * sizeof(float);
  StageRef io, in;
  rc = stage_input(dst, bytes, s, &io);
  if (!rc) rc = stage_input(src, bytes, s, &in);
  if (!rc) rc = mb200_restore_channels_dev(static_cast<float *>(io.dev), static_cast<const float *>(in.dev), w, h, ch, update_mask, s);
  if (rc) cudaStreamSynchronize(s);
  else rc = finish_output(&io, s);
  release_stage(&in, s);
  release_stage(&io, s);
  return rc;
}

int mb200_resize_copy_channels(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh, int filter,
                               unsigned update_mask) {
  if (!dst || !src || !valid_image(w, h, ch) || ow == 0 || oh == 0) return fail(MB200_EINVAL, "resize copy channels: bad arguments");
  cudaStream_t s;
  int rc = prepare(nullptr, &s);
  if (rc) return rc;
  StageRef io, in;
  rc = stage_input(dst, ow * oh * ch * sizeof(float), s, &io);
  if (!rc) rc = stage_input(src, w * h * ch * sizeof(float), s, &in);
  if (!rc) rc = mb200_resize_copy_channels_dev(static_cast<const float *>(in.dev), w, h, ch, static_cast<float *>(io.dev), ow, oh, filter,
                                               update_mask, s);
  if (rc) cudaStreamSynchronize(s);
  else rc = finish_output(&io, s);
  release_stage(&in, s);
  release_stage(&io, s);
  return rc;
}

}  // extern "C"

// ---- EqualizeImage (enhance.c:2040) and EmbossImage (effect.c:1600: inline kernel + ConvolveImage + EqualizeImage) -----------
extern "C" {

int mb200_equalize_image_dev(float *buf, size_t width, size_t height, int channels, int sync_channels, void *stream) {
  if (!buf || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "equalize: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_equalize(buf, width * height, channels, sync_channels, s);
}

int mb200_emboss_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                           double sigma, void *stream) {
  mb200_kernel_info *k = mb200_emboss_kernel(radius, sigma);
  if (!k) return fail(MB200_ENOMEM, "emboss kernel");
  int rc = mb200_convolve_image_dev(src, dst, width, height, channels, k, stream);
  mb200_destroy_kernel_info(k);
  if (rc) return rc;
  return mb200_equalize_image_dev(dst, width, height, channels, 1, stream);      // effect.c:1679, default channel mask
}

int mb200_equalize_image(float *buf, size_t w, size_t h, int ch, int sync_channels) {
  if (!buf || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "equalize: bad arguments");
  return in_place_host(buf, w * h * ch * sizeof(float),
                       [&](float *d, cudaStream_t st) { return mb200_equalize_image_dev(d, w, h, ch, sync_channels, st); });
}

int mb200_emboss_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "emboss: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_emboss_image_dev(s, d, w, h, ch, radius, sigma, st);
  });
}

}  // extern "C"

// ---- StatisticImage (statistic.c:2918), RotationalBlurImage (effect.c:3129), BilateralBlurImage (effect.c:821) ----------
extern "C" {

int mb200_statistic_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, int type,
                              size_t window_width, size_t window_height, void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "statistic: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_statistic(src, dst, width, height, channels, type, window_width, window_height, s);
}

int mb200_rotational_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double angle,
                                    void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "rotational blur: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_rotational_blur(src, dst, width, height, channels, angle, s);
}

int mb200_bilateral_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels,
                                   size_t window_width, size_t window_height, double intensity_sigma, double spatial_sigma,
                                   void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "bilateral blur: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_bilateral_blur(src, dst, width, height, channels, window_width, window_height, intensity_sigma, spatial_sigma, s);
}

int mb200_adaptive_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                                  double sigma, void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "adaptive blur: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_adaptive(src, dst, width, height, channels, radius, sigma, 0, s);
}

int mb200_adaptive_sharpen_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                                     double sigma, void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "adaptive sharpen: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_adaptive(src, dst, width, height, channels, radius, sigma, 1, s);
}

int mb200_adaptive_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "adaptive blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_adaptive_blur_image_dev(s, d, w, h, ch, radius, sigma, st);
  });
}

int mb200_adaptive_sharpen_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "adaptive sharpen: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_adaptive_sharpen_image_dev(s, d, w, h, ch, radius, sigma, st);
  });
}

int mb200_selective_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                                   double sigma, double threshold, void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "selective blur: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  return launch_selective_blur(src, dst, width, height, channels, radius, sigma, threshold, s);
}

int mb200_selective_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma,
                               double threshold) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "selective blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_selective_blur_image_dev(s, d, w, h, ch, radius, sigma, threshold, st);
  });
}

int mb200_statistic_image(const float *src, float *dst, size_t w, size_t h, int ch, int type, size_t ww, size_t wh) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "statistic: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_statistic_image_dev(s, d, w, h, ch, type, ww, wh, st);
  });
}

int mb200_rotational_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, double angle) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "rotational blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_rotational_blur_image_dev(s, d, w, h, ch, angle, st);
  });
}

int mb200_bilateral_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, size_t ww, size_t wh,
                               double intensity_sigma, double spatial_sigma) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "bilateral blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_bilateral_blur_image_dev(s, d, w, h, ch, ww, wh, intensity_sigma, spatial_sigma, st);
  });
}

}  // extern "C"

// ---- ScaleImage (resize.c:4106) ---------------------------------------------------------------------------------------------
extern "C" long mb200_scale_contributions(int axis, size_t in_n, size_t out_n, long *offsets, int *index, double *weight,
                                          size_t max_terms);
extern "C" {

int mb200_scale_image_dev(const float *src, size_t width, size_t height, int channels, float *dst, size_t out_width,
                          size_t out_height, void *stream) {
  if (!src || !dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "scale: bad arguments");
  if (out_width == 0 || out_height == 0) return fail(MB200_EINVAL, "NegativeOrZeroImageSize");      // :4153
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  if (out_width == width && out_height == height) {                                                  // :4155: clone
    cudaError_t e = cudaMemcpyAsync(dst, src, width * height * channels * sizeof(float), cudaMemcpyDeviceToDevice, s);
    return e == cudaSuccess ? MB200_OK : cuda_fail(e, "scale: clone");
  }
  // contribution lists of both axes (host, the reference's state machines), packed into one upload:
  // [xoff (ow+1) | yoff (oh+1) | xidx | yidx] ints, then [xwt | ywt] doubles
  std::vector<long> xoff(out_width + 1), yoff(out_height + 1);
  const long nx = mb200_scale_contributions(0, width, out_width, xoff.data(), nullptr, nullptr, 0);
  const long ny = mb200_scale_contributions(1, height, out_height, yoff.data(), nullptr, nullptr, 0);
  if (nx < 0 || ny < 0) return static_cast<int>(nx < 0 ? nx : ny);
  std::vector<int> ints(out_width + 1 + out_height + 1 + static_cast<size_t>(nx + ny));
  std::vector<double> wts(static_cast<size_t>(nx + ny));
  int *xo = ints.data(), *yo = xo + out_width + 1, *xi = yo + out_height + 1, *yi = xi + nx;
  if (mb200_scale_contributions(0, width, out_width, xoff.data(), xi, wts.data(), static_cast<size_t>(nx)) < 0 ||
      mb200_scale_contributions(1, height, out_height, yoff.data(), yi, wts.data() + nx, static_cast<size_t>(ny)) < 0)
    return MB200_EINVAL;
  for (size_t i = 0; i <= out_width; ++i) xo[i] = static_cast<int>(xoff[i]);
  for (size_t i = 0; i <= out_height; ++i) yo[i] = static_cast<int>(yoff[i]);
  StreamAlloc d_ints(s), d_wts(s);
  rc = d_ints.alloc(ints.size() * sizeof(int));
  if (!rc) rc = d_wts.alloc(wts.size() * sizeof(double));
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(d_ints.ptr, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_wts.ptr, wts.data(), wts.size() * sizeof(double), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cuda_fail(e, "scale: table upload");
  const int *di = static_cast<const int *>(d_ints.ptr);
  const double *dw = static_cast<const double *>(d_wts.ptr);
  return launch_scale(src, width, height, channels, dst, out_width, out_height, di, di + out_width + 1 + out_height + 1, dw,
                      di + out_width + 1, di + out_width + 1 + out_height + 1 + nx, dw + nx, s);
}

int mb200_scale_image(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh) {
  if (!src || !dst || !valid_image(w, h, ch) || ow == 0 || oh == 0) return fail(MB200_EINVAL, "scale: bad arguments");
  return with_staging(src, w * h * ch * sizeof(float), dst, ow * oh * ch * sizeof(float),
                      [&](const float *s, float *d, cudaStream_t st) { return mb200_scale_image_dev(s, w, h, ch, d, ow, oh, st); });
}

}  // extern "C"

// ---- SampleImage (resize.c:3907) --------------------------------------------------------------------------------
extern "C" {

int mb200_sample_image_dev(const float *src, size_t width, size_t height, int channels, float *dst, size_t out_width,
                           size_t out_height, void *stream) {
  if (!src || !dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "sample: bad arguments");
  if (out_width == 0 || out_height == 0) return fail(MB200_EINVAL, "NegativeOrZeroImageSize");   // :3946
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  if (out_width == width && out_height == height) {                                              // :3948: clone
    cudaError_t e = cudaMemcpyAsync(dst, src, width * height * channels * sizeof(float), cudaMemcpyDeviceToDevice, s);
    return e == cudaSuccess ? MB200_OK : cuda_fail(e, "sample: clone");
  }
  return launch_sample(src, width, height, channels, dst, out_width, out_height, s);
}

int mb200_sample_image(const float *src, size_t w, size_t h, int ch, float *dst, size_t ow, size_t oh) {
  if (!src || !dst || !valid_image(w, h, ch) || ow == 0 || oh == 0) return fail(MB200_EINVAL, "sample: bad arguments");
  return with_staging(src, w * h * ch * sizeof(float), dst, ow * oh * ch * sizeof(float),
                      [&](const float *s, float *d, cudaStream_t st) {
                        return mb200_sample_image_dev(s, w, h, ch, d, ow, oh, st);
                      });
}

}  // extern "C"

// ---- ThumbnailImage (resize.c:4591-4650), pixel path: SampleImage to 4x the target when both integer reduction
// factors exceed 4, ResizeImage(Box) to 2x when they exceed 2, then ResizeImage(filter; the reference passes
// image->filter, LanczosSharp when undefined).  The metadata the reference attaches afterwards (profile stripping,
// Thumb::* properties) is control plane and stays with the caller.
extern "C" {

int mb200_thumbnail_image_dev(const float *src, size_t width, size_t height, int channels, float *dst, size_t columns,
                              size_t rows, int filter, void *stream) {
  if (!src || !dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "thumbnail: bad arguments");
  if (columns == 0 || rows == 0) return fail(MB200_EINVAL, "NegativeOrZeroImageSize");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  const size_t px = static_cast<size_t>(channels) * sizeof(float);
  if (columns == width && rows == height) {
    cudaError_t e = cudaMemcpyAsync(dst, src, width * height * px, cudaMemcpyDeviceToDevice, s);
    return e == cudaSuccess ? MB200_OK : cuda_fail(e, "thumbnail: clone");
  }
  const long x_factor = static_cast<long>(width) / static_cast<long>(columns);
  const long y_factor = static_cast<long>(height) / static_cast<long>(rows);
  const float *cur = src;
  size_t cw = width, chh = height;
  StreamAlloc a(s), b(s);
  if (x_factor > 4 && y_factor > 4) {
    rc = a.alloc(4 * columns * 4 * rows * px);
    if (rc) return rc;
    rc = mb200_sample_image_dev(cur, cw, chh, channels, static_cast<float *>(a.ptr), 4 * columns, 4 * rows, s);
    if (rc) return rc;
    cur = static_cast<const float *>(a.ptr); cw = 4 * columns; chh = 4 * rows;
  }
  if (x_factor > 2 && y_factor > 2) {
    rc = b.alloc(2 * columns * 2 * rows * px);
    if (rc) return rc;
    rc = mb200_resize_image_dev(cur, cw, chh, channels, static_cast<float *>(b.ptr), 2 * columns, 2 * rows, MB200_BoxFilter, s);
    if (rc) return rc;
    cur = static_cast<const float *>(b.ptr); cw = 2 * columns; chh = 2 * rows;
  }
  return mb200_resize_image_dev(cur, cw, chh, channels, dst, columns, rows,
                                filter == MB200_UndefinedFilter ? MB200_LanczosSharpFilter : filter, s);
}

int mb200_thumbnail_image(const float *src, size_t w, size_t h, int ch, float *dst, size_t columns, size_t rows, int filter) {
  if (!src || !dst || !valid_image(w, h, ch) || columns == 0 || rows == 0) return fail(MB200_EINVAL, "thumbnail: bad arguments");
  return with_staging(src, w * h * ch * sizeof(float), dst, columns * rows * ch * sizeof(float),
                      [&](const float *s, float *d, cudaStream_t st) {
                        return mb200_thumbnail_image_dev(s, w, h, ch, d, columns, rows, filter, st);
                      });
}

}  // extern "C"

// ---- MotionBlurImage (effect.c:2347) ------------------------------------------------------------------------------
extern "C" {

int mb200_motion_blur_image_dev(const float *src, float *dst, size_t width, size_t height, int channels, double radius,
                                double sigma, double angle, void *stream) {
  if (!src || !dst || src == dst || !valid_image(width, height, channels)) return fail(MB200_EINVAL, "motion blur: bad arguments");
  cudaStream_t s;
  int rc = prepare(stream, &s);
  if (rc) return rc;
  double taps[129];
  long ox[129], oy[129];
  const long n = mb200_motion_blur_kernel(radius, sigma, angle, taps, ox, oy, 129);
  if (n < 0) return fail(MB200_EUNSUPPORTED, "motion blur: more than 129 taps");
  return launch_motion_blur(src, dst, width, height, channels, taps, ox, oy, static_cast<int>(n), s);
}

int mb200_motion_blur_image(const float *src, float *dst, size_t w, size_t h, int ch, double radius, double sigma,
                            double angle) {
  if (!src || !dst || !valid_image(w, h, ch)) return fail(MB200_EINVAL, "motion blur: bad arguments");
  const size_t bytes = w * h * ch * sizeof(float);
  return with_staging(src, bytes, dst, bytes, [&](const float *s, float *d, cudaStream_t st) {
    return mb200_motion_blur_image_dev(s, d, w, h, ch, radius, sigma, angle, st);
  });
}

}  // extern "C"

