// mb200_internal.h -- shared declarations inside libmagickb200 (not installed).
#pragma once

#include "../../include/magick_b200.h"

#include <cstddef>
#include <cstdint>

struct CUmemPoolHandle_st;      // cudaMemPool_t == CUmemPoolHandle_st * (driver_types.h)

namespace mb200 {

// ---- error plumbing (runtime.cu) -------------------------------------------
int fail(int code, const char *fmt, ...);       // records thread-local message, returns code
int cuda_fail(int cuda_error, const char *what); // wraps a cudaError_t
void count_launch(unsigned n = 1);

// ---- per-device state (runtime.cu) -----------------------------------------
struct DeviceState;
int ensure_device();                     // lazily initialises the current device; 0 or error
void *default_stream();                  // library-owned stream of the current device
int scratch(void **ptr, size_t bytes, int slot);   // grow-only device scratch buffers
int sm_count();
// The library's PRIVATE stream-ordered memory pool of the current device (temporaries of the operators).  The host
// application's default pool is left alone; mb200_trim() gives the cached memory back.
::CUmemPoolHandle_st *temp_pool();

// ---- pixel-cache staging (cache.cu) ------------------------------------------
// copy_h2d returns once the host buffer has been consumed (the copy itself may still be in flight on `stream`);
// copy_d2h returns when the host buffer holds the data.  Pinned / registered / managed host memory is copied
// directly, pageable memory through the threaded pinned bounce ring.
int copy_h2d(void *dev, const void *host, size_t bytes, void *stream);
int copy_d2h(void *host, const void *dev, size_t bytes, void *stream);
struct StageRef {
  void *entry = nullptr;     // residency-registry entry when the host buffer is an attached pixel cache
  void *dev = nullptr;       // HBM copy to run the operator on
  void *host = nullptr;
  size_t bytes = 0;
  bool temporary = false;    // dev is a stream-ordered temporary (unattached host buffer)
};
int stage_input(const void *host, size_t bytes, void *stream, StageRef *out);
int stage_output(void *host, size_t bytes, void *stream, StageRef *out);
int finish_output(StageRef *ref, void *stream);
void release_stage(StageRef *ref, void *stream);

// ---- kernel helpers (kernel_info.cpp) --------------------------------------
void rotate_kernel_info(mb200_kernel_info *k, double angle);

// ---- channel traits (pixel.c:6356-6381) ------------------------------------
inline bool has_alpha(int channels) { return channels == 2 || channels == 4; }

// ---- device launchers -------------------------------------------------------
// All take device pointers and a cudaStream_t (as void*).  Return 0 / MB200_E*.

// conv1d.cu: one 1-D convolution pass (width-1 or height-1 kernel) with the
// reference's edge clamp, reflected taps and alpha blending.  taps[] is in window
// order (tap t multiplies the source sample at offset t - origin_offset).
// axis 0 = along x (row path, morphology.c:2815), axis 1 = along y (column path :2654).
// d_changed: device counter (may be null) incremented per changed channel value.
// epilogue: UnsharpMaskImage's point pass (effect.c:4358-4364) fused into the output stage when the RGBA pair kernels
// run the pass (*epilogue_fused tells; otherwise the caller runs launch_unsharp_combine afterwards).
struct UnsharpEpilogue { const float *source; double gain, quantum_threshold; };
int launch_conv1d(const float *src, float *dst, size_t width, size_t height, int channels,
                  int axis, const double *taps_window_order, int ntaps, int origin_offset,
                  double bias, double gamma_scale, unsigned long long *d_changed, void *stream,
                  int io = 0,    // io: 0 float->float, 1 float->raw double sums, 2 raw double sums->float (RGBA only)
                  const UnsharpEpilogue *epilogue = nullptr, bool *epilogue_fused = nullptr);

// conv_mma.cu: the same pass on the FP64 matrix path (mma.sync m8n8k4) for RGBA images, bias 0, <= 33 taps; src / dst are
// float4 pixels (io 0), double4 sums out (io 1) or in (io 2).  MB200_EUNSUPPORTED => use launch_conv1d's DFMA kernels.
int launch_conv_mma(const void *src, void *dst, size_t width, size_t height, int axis, const double *taps_window_order,
                    int ntaps, int origin_offset, void *stream, int io, const UnsharpEpilogue *epilogue,
                    bool *epilogue_fused);
void set_conv_mma(int enable);
int conv_mma_enabled();
unsigned long long conv_mma_launches();

// conv2d.cu: general 2-D convolution / erode / dilate (MorphologyPrimitive row path)
int launch_morph2d(const float *src, float *dst, size_t width, size_t height, int channels,
                   int method, const double *kernel_window_order, int kw, int kh, int ox, int oy,
                   double bias, double gamma_scale, unsigned long long *d_changed, void *stream);

// morph_stream.cu: erode / dilate for the built-in structuring elements, register streaming
// (MB200_EUNSUPPORTED => shape not in the table; use launch_morph2d)
int launch_morph_stream(const float *src, float *dst, size_t width, size_t height, int channels, int method,
                        const double *kernel_window_order, int kw, int kh, int ox, int oy, void *stream);

// resize.cu: one axis of ResizeImage.  Contribution table lives in device memory.
int launch_resize_axis(const float *src, size_t width, size_t height, int channels, float *dst,
                       size_t out_n, int axis, const int *d_start, const int *d_count,
                       const double *d_weights, int max_taps, int max_span, int reg_stride, int reg_taps,
                       const double *d_wreg, void *stream, long o_begin = -1, long o_end = -1);

// resize_stream.cu: streaming kernels for runs of outputs with bit-identical weights (integer-ratio
// reductions, RGBA).  seg_*: first output / number of outputs / start[] of each run; d_wsets[nseg][taps];
// d_border: the outputs outside the runs (gathered by extra CTAs of the same launch).
#define MB200_RESIZE_MAX_SEGMENTS 8
int launch_resize_stream(const float *src, size_t width, size_t height, float *dst, size_t out_n, int axis,
                         int stride, int taps, int nseg, const int *seg_o, const int *seg_n, const int *seg_src,
                         const double *d_wsets, int nborder, const int *d_border, const int *d_start,
                         const int *d_count, const double *d_weights, void *stream);

// Fused vertical + horizontal pass for equal integer reductions on both axes (RGBA): tile lists of both axes' runs,
// the two-pass path's contribution / border lists for the outputs outside the runs.
void resize_fused_tile(int stride, int taps, int *tile_w, int *tile_h);
int launch_resize_fused(const float *src, size_t width, size_t height, float *dst, size_t out_w, size_t out_h, int stride,
                        int taps, const int *d_xtiles, int nxt, const int *d_ytiles, int nyt, const double *d_wx,
                        const double *d_wy, const int *d_xstart, const int *d_xcount, const double *d_xweights,
                        const int *d_ystart, const int *d_ycount, const double *d_yweights, const int *d_xborder, int nxborder,
                        const int *d_yborder, int nyborder, const unsigned char *d_row_is_border, void *stream);

// colorspace.cu
int launch_colorspace(float *buf, size_t npixels, int channels, int from, int to, const mb200_colorspace_options *options,
                      void *stream);

// hexcone.cu: HCL, HCLp, HSB, HSI, HSL, HSV, HWB (one leg: sRGB -> space or space -> sRGB), in place
bool is_hexcone_colorspace(int cs);
int launch_hexcone_leg(float *buf, size_t npixels, int channels, int space, bool forward, void *stream);

// pointwise.cu
int launch_unsharp_combine(const float *src, float *blur_inout, size_t n, double gain,
                           double quantum_threshold, void *stream);

// Copy-trait channels (a `-channel` selection): dst[c] = src[c] for the channels NOT in update_mask; the resize form takes
// the nearest source sample of each axis (resize.c:3697-3707)
int launch_restore_channels(float *dst, const float *src, size_t npixels, int channels, unsigned update_mask, void *stream);
int launch_resize_copy_channels(float *dst, const float *src, size_t w, size_t ow, size_t oh, int channels, const int *d_nearest_x,
                                const int *d_nearest_y, unsigned update_mask, void *stream);

// CompositeImage(canvas, source, DifferenceCompositeOp) for same-size images (Edge/TopHat/BottomHat), in place on canvas
int launch_composite_difference(float *canvas, const float *source, size_t npixels, int channels, void *stream);
// ... LightenCompositeOp: the union of the results of a HitAndMiss kernel list (morphology.c:3722, :4044-4046)
int launch_composite_lighten(float *canvas, const float *source, size_t npixels, int channels, void *stream);

// MotionBlurImage (effect.c:2347): taps + integer offsets from the host, gather along the blur direction
int launch_motion_blur(const float *src, float *dst, size_t w, size_t h, int channels, const double *taps, const long *ox,
                       const long *oy, int width, void *stream);

// stencils.cu: StatisticImage (statistic.c:2918), RotationalBlurImage (effect.c:3129), BilateralBlurImage (effect.c:821)
int launch_statistic(const float *src, float *dst, size_t w, size_t h, int channels, int type, size_t width, size_t height,
                     void *stream);
int launch_rotational_blur(const float *src, float *dst, size_t w, size_t h, int channels, double angle, void *stream);
int launch_bilateral_blur(const float *src, float *dst, size_t w, size_t h, int channels, size_t width, size_t height,
                          double intensity_sigma, double spatial_sigma, void *stream);
// AdaptiveBlurImage (effect.c:128) / AdaptiveSharpenImage (:447): edge map + per-pixel kernel size, bit exact
int launch_adaptive(const float *src, float *dst, size_t w, size_t h, int channels, double radius, double sigma, int sharpen,
                    void *stream);
// SelectiveBlurImage (effect.c:3406)
int launch_selective_blur(const float *src, float *dst, size_t w, size_t h, int channels, double radius, double sigma,
                          double threshold, void *stream);

// ScaleImage (resize.c:4106): CSR contribution lists of both axes (mb200_scale_contributions), bit exact
int launch_scale(const float *src, size_t w, size_t h, int channels, float *dst, size_t ow, size_t oh, const int *d_xoff,
                 const int *d_xidx, const double *d_xwt, const int *d_yoff, const int *d_yidx, const double *d_ywt, void *stream);

// SampleImage (resize.c:3907): nearest-sample gather, bit exact
int launch_sample(const float *src, size_t w, size_t h, int channels, float *dst, size_t ow, size_t oh, void *stream);

// equalize.cu: EqualizeImage (enhance.c:2040) in place; sync_channels = the channel mask carries SyncChannels (the
// default): one intensity-driven histogram for all channels.  Synchronises the stream (host step between the kernels).
int launch_equalize(float *buf, size_t npixels, int channels, int sync_channels, void *stream);

// threshold.c point operators in place; op: 0 bilevel (t[0]), 1 black, 2 white (t = r,g,b,a), 3 clamp
int launch_threshold(float *buf, size_t npixels, int channels, int op, const double *thresholds, void *stream);

}  // namespace mb200
