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

#include <cstdint>
#include <cstdlib>
#include <type_traits>

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

// 1/g to ~1 ulp: MUFU.RCP64H seed (one XU op, ~20 bits) + two FP64 Newton steps.
__device__ __forceinline__ double fast_reciprocal(double g) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(g));
  const double e0 = fma(-g, r0, 1.0);
  const double r1 = fma(r0, e0, r0);              // ~2^-40
  const double e1 = fma(-g, r1, 1.0);
  return fma(r1, e1, r1);                         // ~1 ulp of double: keeps the first pass of a two-pass operator
}                                                 // bit-identical to the reference's quotient in all but ~1e-8 of the samples

__device__ __forceinline__ double shfl_double(double v, int lane) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, lane);
  hi = __shfl_sync(0xffffffffu, hi, lane);
  return __hiloint2double(hi, lo);
}

// Per-thread constants of the output stage.  With sum' = sum K*(A*p) and gsum' = sum K*A the
// reference's  PerceptibleReciprocal(QS*gsum') * (bias + QS*sum')  (morphology.c:3197) equals
// (bias/QS + sum') / gsum'; alpha / plain lanes use a denominator of 1.
struct Finish {
  double bias_eff;   // bias (plain, alpha lane) or bias/QS (blend lane)
  bool blend;
};

// branch-free: out = (bias_eff + sum) * 1/den, den = blend ? gsum : 1, with the reference's
// |gamma| < MagickEpsilon clamp (PerceptibleReciprocal == 1/clamp(gamma)) applied to QS*gsum by
// comparing the high word of |den| against the threshold (exact up to the low 32 mantissa bits).
__device__ __forceinline__ float finish(const Finish &f, double sum, double gsum) {
  const double pixel = f.bias_eff + sum;
  double den = f.blend ? gsum : 1.0;
  constexpr double kTiny = kEpsilon / kQuantumScale;           // 6.5535e-8
  const int hi = __double2hiint(den);
  if ((hi & 0x7fffffff) < __double2hiint(kTiny))
    den = __hiloint2double((hi & 0x80000000) | __double2hiint(kTiny), __double2loint(kTiny));
  return static_cast<float>(fast_reciprocal(den) * pixel);
}

template <int NT> struct Ring { static constexpr int value = NT; };
template <> struct Ring<25> { static constexpr int value = 5; };
template <> struct Ring<33> { static constexpr int value = 11; };
template <> struct Ring<49> { static constexpr int value = 7; };
template <> struct Ring<65> { static constexpr int value = 13; };

// ---------------------------------------------------------------- column pass
// grid: (ceil(rc / THREADS), ceil(height / strip)); one thread per component column.
// strip + NT - 1 is a whole number of NT-step rotations, so the unrolled body has no exits.
template <int NT, int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) conv_col_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  constexpr int PF = Ring<NT>::value;          // source rows kept in flight per thread
  static_assert(NT % PF == 0, "ring must divide the rotation");
  const int col_raw = blockIdx.x * THREADS + threadIdx.x;
  const bool active = col_raw < a.rc;
  const int col = active ? col_raw : a.rc - 1;
  const int lane = threadIdx.x & 31;
  const int alane = MODE ? (lane | (MODE - 1)) : lane;
  const bool is_alpha = MODE ? ((col % MODE) == MODE - 1) : false;
  const int y0 = blockIdx.y * a.strip;
  const int nout = active ? min(a.strip, a.height - y0) : 0;
  const int total = a.strip + NT - 1;
  const int hmax = a.height - 1;
  const unsigned pitch_bytes = static_cast<unsigned>(a.rc) * 4u;
  const char *base = reinterpret_cast<const char *>(a.src + col);
  char *outp = reinterpret_cast<char *>(a.dst + col) + static_cast<size_t>(y0) * pitch_bytes;
  Finish fin;
  fin.blend = MODE && !is_alpha;
  fin.bias_eff = fin.blend ? a.bias * 65535.0 : a.bias;

  double acc[NT];
  float pre[PF];
#pragma unroll
  for (int q = 0; q < NT; ++q) acc[q] = 0.0;
  int ysrc = y0 - a.off;                       // source row of step 0
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    const unsigned yy = static_cast<unsigned>(min(max(ysrc + s, 0), hmax));
    pre[s] = __ldg(reinterpret_cast<const float *>(base + static_cast<size_t>(yy) * pitch_bytes));
  }
  ysrc += PF;                                  // next row to fetch

  int j = -(NT - 1);                           // output row (relative to y0) finished at this step
  // The rotation is unrolled PF steps at a time (not NT): after each block the accumulators are
  // physically rotated by PF slots, so the static tap pattern repeats and the loop body stays
  // small enough for the instruction cache (NT*NT FMAs would not).
#pragma unroll 1
  for (int mb = 0; mb < total; mb += PF) {
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      float vf = pre[s];
      {  // refill this ring slot with the row PF steps ahead (edge-clamped)
        const unsigned yy = static_cast<unsigned>(min(max(ysrc, 0), hmax));
        pre[s] = __ldg(reinterpret_cast<const float *>(base + static_cast<size_t>(yy) * pitch_bytes));
        ++ysrc;
      }
      double v = static_cast<double>(vf);
      if (MODE) {
        float af = __shfl_sync(0xffffffffu, vf, alane);
        af = is_alpha ? 1.0f : af;
        v *= static_cast<double>(af);
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = fma(taps.k[(s - q + NT) % NT], v, acc[q]);
      const int qf = (s + 1) % NT;
      const double sum = acc[qf];
      acc[qf] = 0.0;
      double gsum = 1.0;
      if (MODE) gsum = shfl_double(sum, alane);
      const float out = finish(fin, sum, gsum);
      if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<float *>(outp) = out;
      if (j >= 0) outp += pitch_bytes;
      ++j;
    }
    if (PF != NT) {                              // rotate: new acc[q] = old acc[(q + PF) % NT]
      double tmp[PF];
#pragma unroll
      for (int q = 0; q < PF; ++q) tmp[q] = acc[q];
#pragma unroll
      for (int q = 0; q < NT - PF; ++q) acc[q] = acc[q + PF];
#pragma unroll
      for (int q = 0; q < PF; ++q) acc[NT - PF + q] = tmp[q];
    }
  }
}

// ------------------------------------------------------------------- row pass
// block: 128 threads = 4 warps; a warp covers RPW = 32/channels rows x channels lanes.
// grid: (ceil(width / strip), ceil(height / rows_per_cta)).
template <int NT, int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_row_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  extern __shared__ __align__(16) float tile[];
  const int ch = a.channels;
  const int rpw = 32 / ch;
  const int rows_per_cta = 4 * rpw;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * a.strip;
  const int ybase = blockIdx.y * rows_per_cta;
  const int total = a.strip + NT - 1;
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
  const int nout = active ? min(a.strip, a.width - x0) : 0;
  const int alane = MODE ? (lane | (MODE - 1)) : lane;
  const bool is_alpha = MODE ? (c == MODE - 1) : false;
  const float *tp = tile + static_cast<size_t>(r) * a.pitch * ch + c;
  float *outp = a.dst + (static_cast<size_t>(min(y, hmax)) * a.width + x0) * ch + c;
  Finish fin;
  fin.blend = MODE && !is_alpha;
  fin.bias_eff = fin.blend ? a.bias * 65535.0 : a.bias;

  double acc[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) acc[q] = 0.0;

  int j = -(NT - 1);
  constexpr int BLK = Ring<NT>::value;         // unrolled steps per block; accumulators rotate by BLK after it
#pragma unroll 1
  for (int mb = 0; mb < total; mb += BLK) {
#pragma unroll
    for (int s = 0; s < BLK; ++s) {
      const float vf = *tp;
      tp += ch;
      double v = static_cast<double>(vf);
      if (MODE) {
        float af = __shfl_sync(0xffffffffu, vf, alane);
        af = is_alpha ? 1.0f : af;
        v *= static_cast<double>(af);
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = fma(taps.k[(s - q + NT) % NT], v, acc[q]);
      const int qf = (s + 1) % NT;
      const double sum = acc[qf];
      acc[qf] = 0.0;
      double gsum = 1.0;
      if (MODE) gsum = shfl_double(sum, alane);
      const float out = finish(fin, sum, gsum);
      if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *outp = out;
      if (j >= 0) outp += ch;
      ++j;
    }
    if (BLK != NT) {                             // rotate: new acc[q] = old acc[(q + BLK) % NT]
      double tmp[BLK];
#pragma unroll
      for (int q = 0; q < BLK; ++q) tmp[q] = acc[q];
#pragma unroll
      for (int q = 0; q < NT - BLK; ++q) acc[q] = acc[q + BLK];
#pragma unroll
      for (int q = 0; q < BLK; ++q) acc[NT - BLK + q] = tmp[q];
    }
  }
}

// ======================================================================================
// TMA (cp.async.bulk + mbarrier) staged variants for RGBA.  Every warp owns a private
// shared-memory ring of source chunks filled by the bulk-copy engine: no thread spends
// registers or issue slots on global loads, the prefetch depth is set by the ring (not by the
// register file), and warps never synchronise with each other (no __syncthreads in the loop).
// Edge clamping is done by the producer lanes: out-of-image rows / pixels are bulk-copied from
// the clamped source address.  Other layouts use the register-ring kernels above.
// ======================================================================================

__device__ __forceinline__ unsigned smem_u32(const void *p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Spin on the barrier phase.  The loop lives inside the asm block so that the compiler sees one
// convergent instruction; a bounded poll count traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 cnt;\n\t"
      "mov.u32 cnt, 0;\n\t"
      "MB200_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MB200_DONE;\n\t"
      "add.u32 cnt, cnt, 1;\n\t"
      "setp.gt.u32 p, cnt, 0x4000000;\n\t"
      "@p trap;\n\t"
      "bra MB200_WAIT;\n\t"
      "MB200_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int kTmaSlots = 4;      // ring depth (chunks in flight per warp)

// blend-lane output stage without bias: out = sum / clamp(den), den = blend ? gsum : 1; the
// reference's PerceptibleReciprocal clamp (|QS*gsum| < MagickEpsilon) is decided on the float copy.
__device__ __forceinline__ float finish_rgba(bool blend, double sum, double gsum) {
  const double den = blend ? gsum : 1.0;
  const float denf = static_cast<float>(den);
  float seed;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"(denf));
  const double r0 = static_cast<double>(seed);
  const double e = fma(-den, r0, 1.0);
  double r = fma(r0, e, r0);
  if (fabsf(denf) < static_cast<float>(kEpsilon / kQuantumScale))
    r = denf < 0.0f ? -(kQuantumScale / kEpsilon) : (kQuantumScale / kEpsilon);
  return static_cast<float>(r * sum);
}

// Producer step of the column pass (lanes 0..PF-1 of the warp): one edge-clamped 128-byte row
// segment per lane into the warp's ring slot.
template <int PF>
__device__ __forceinline__ void col_issue(float *slot_base, unsigned long long *bar, const char *gbase, int y_first,
                                       int hmax, unsigned pitch_bytes, unsigned row_bytes, int lane) {
  __syncwarp();                                           // every lane has finished reading this slot
  if (lane == 0) mbar_expect_tx(bar, row_bytes * PF);
  __syncwarp();
  if (lane < PF) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned yy = static_cast<unsigned>(min(max(y_first + lane, 0), hmax));
    bulk_g2s(slot_base + lane * 32, gbase + static_cast<size_t>(yy) * pitch_bytes, row_bytes, bar);
  }
}

// Producer step of the row pass (lanes 0..7 = the warp's 8 tile rows): PF pixels per row, the
// out-of-image pixels replicated from the edge pixel with 16-byte copies.
template <int PF>
__device__ __forceinline__ void row_issue(float *slot_base, unsigned long long *bar, const float4 *src, int ybase,
                                       int hmax, int xs, int width, int lane) {
  __syncwarp();
  if (lane == 0) mbar_expect_tx(bar, 8u * PF * 16u);
  __syncwarp();
  if (lane < 8) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const float4 *grow = src + static_cast<size_t>(min(ybase + lane, hmax)) * width;
    float *dst = slot_base + lane * (PF * 4);
    const int lo = max(xs, 0), hi = min(xs + PF, width);   // in-image part [lo, hi)
    if (hi > lo) bulk_g2s(dst + (lo - xs) * 4, grow + lo, static_cast<unsigned>(hi - lo) * 16u, bar);
#pragma unroll 1
    for (int x = xs; x < min(xs + PF, 0); ++x) bulk_g2s(dst + (x - xs) * 4, grow, 16u, bar);
#pragma unroll 1
    for (int x = max(xs, width); x < xs + PF; ++x) bulk_g2s(dst + (x - xs) * 4, grow + (width - 1), 16u, bar);
  }
}

// ---- column pass: CTA = 4 independent warps; a warp covers 32 consecutive components (8 RGBA
//      pixels, 128 B per row); chunk = PF rows.  grid: (ceil(rc/128), ceil(height/strip)).
template <int NT, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_col_tma_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  constexpr int PF = Ring<NT>::value;
  __shared__ __align__(128) float ring[4][kTmaSlots][PF][32];
  __shared__ __align__(8) unsigned long long full[4][kTmaSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 128 + warp * 32;
  if (col0 >= a.rc) return;                              // whole warp out of the image (warps are independent)
  const bool blend = (lane & 3) != 3;
  const int alpha_lane = lane | 3;
  const int y0 = blockIdx.y * a.strip;
  const int nout = min(a.strip, a.height - y0);
  const int total = a.strip + NT - 1;                 // multiple of NT, hence of PF
  const int nchunks = total / PF;
  const int hmax = a.height - 1;
  const unsigned pitch_bytes = static_cast<unsigned>(a.rc) * 4u;
  const char *gbase = reinterpret_cast<const char *>(a.src + col0);
  char *outp = reinterpret_cast<char *>(a.dst + col0 + lane) + static_cast<size_t>(y0) * pitch_bytes;
  unsigned long long *bars = &full[warp][0];
  float *wring = &ring[warp][0][0][0];

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaSlots; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int ysrc0 = y0 - a.off;
  for (int c = 0; c < kTmaSlots - 1 && c < nchunks; ++c)
    col_issue<PF>(wring + c * (PF * 32), &bars[c], gbase, ysrc0 + c * PF, hmax, pitch_bytes, 128u, lane);

  double acc[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) acc[q] = 0.0;
  int j = -(NT - 1);
  int chunk = 0, slot = 0, nslot = kTmaSlots - 1;     // nslot = slot of chunk + kTmaSlots - 1
  unsigned parity = 0;
#pragma unroll 1
  for (int mb = 0; mb < total; mb += NT) {
#pragma unroll
    for (int s = 0; s < NT; ++s) {
      if (s % PF == 0) {
        // the slot of chunk-1 is free now: refill it with chunk + kTmaSlots - 1, then wait for ours
        const int nxt = chunk + kTmaSlots - 1;
        if (nxt < nchunks)
          col_issue<PF>(wring + nslot * (PF * 32), &bars[nslot], gbase, ysrc0 + nxt * PF, hmax, pitch_bytes, 128u, lane);
        mbar_wait(&bars[slot], parity);
      }
      const float *rowp = wring + (slot * PF + (s % PF)) * 32;
      const float vf = rowp[lane];
      float af = rowp[alpha_lane];
      af = blend ? af : 1.0f;
      const double v = static_cast<double>(vf) * static_cast<double>(af);
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = fma(taps.k[(s - q + NT) % NT], v, acc[q]);
      const int qf = (s + 1) % NT;
      const double sum = acc[qf];
      acc[qf] = 0.0;
      const double gsum = shfl_double(sum, alpha_lane);
      const float out = finish_rgba(blend, sum, gsum);
      if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<float *>(outp) = out;
      if (j >= 0) outp += pitch_bytes;
      ++j;
      if (s % PF == PF - 1) {
        ++chunk;
        nslot = slot;
        if (++slot == kTmaSlots) { slot = 0; parity ^= 1u; }
      }
    }
  }
}

// ---- row pass: CTA = 4 independent warps; a warp covers 8 rows x 4 channels; chunk = PF pixels of
//      each row (pitch PF pixels, odd => conflict-free LDS).  grid: (ceil(width/strip), ceil(height/32)).
template <int NT, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_row_tma_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  constexpr int PF = Ring<NT>::value;
  static_assert(PF % 2 == 1, "odd chunk pitch keeps the LDS conflict-free");
  __shared__ __align__(128) float ring[4][kTmaSlots][8][PF * 4];
  __shared__ __align__(8) unsigned long long full[4][kTmaSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ybase = blockIdx.y * 32 + warp * 8;
  if (ybase >= a.height) return;
  const int lr = lane >> 2, c = lane & 3;
  const int x0 = blockIdx.x * a.strip;
  const int y = ybase + lr;
  const int hmax = a.height - 1;
  const int nout = y < a.height ? min(a.strip, a.width - x0) : 0;
  const int total = a.strip + NT - 1;
  const int nchunks = total / PF;
  const bool blend = c != 3;
  const int alpha_lane = lane | 3;
  float *outp = a.dst + (static_cast<size_t>(min(y, hmax)) * a.width + x0) * 4 + c;
  unsigned long long *bars = &full[warp][0];
  float *wring = &ring[warp][0][0][0];
  const float4 *src4 = reinterpret_cast<const float4 *>(a.src);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kTmaSlots; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int xsrc0 = x0 - a.off;
  for (int ck = 0; ck < kTmaSlots - 1 && ck < nchunks; ++ck)
    row_issue<PF>(wring + ck * (8 * PF * 4), &bars[ck], src4, ybase, hmax, xsrc0 + ck * PF, a.width, lane);

  double acc[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) acc[q] = 0.0;
  int j = -(NT - 1);
  int chunk = 0, slot = 0, nslot = kTmaSlots - 1;
  unsigned parity = 0;
#pragma unroll 1
  for (int mb = 0; mb < total; mb += NT) {
#pragma unroll
    for (int s = 0; s < NT; ++s) {
      if (s % PF == 0) {
        const int nxt = chunk + kTmaSlots - 1;
        if (nxt < nchunks)
          row_issue<PF>(wring + nslot * (8 * PF * 4), &bars[nslot], src4, ybase, hmax, xsrc0 + nxt * PF, a.width, lane);
        mbar_wait(&bars[slot], parity);
      }
      const float *pp = wring + ((slot * 8 + lr) * PF + (s % PF)) * 4;
      const float vf = pp[c];
      float af = pp[3];
      af = blend ? af : 1.0f;
      const double v = static_cast<double>(vf) * static_cast<double>(af);
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = fma(taps.k[(s - q + NT) % NT], v, acc[q]);
      const int qf = (s + 1) % NT;
      const double sum = acc[qf];
      acc[qf] = 0.0;
      const double gsum = shfl_double(sum, alpha_lane);
      const float out = finish_rgba(blend, sum, gsum);
      if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *outp = out;
      if (j >= 0) outp += 4;
      ++j;
      if (s % PF == PF - 1) {
        ++chunk;
        nslot = slot;
        if (++slot == kTmaSlots) { slot = 0; parity ^= 1u; }
      }
    }
  }
}

// ======================================================================================
// Two-components-per-thread kernels for RGBA ("pair" kernels, the default for 4 channels).
// A thread owns two adjacent components of a pixel -- (R,G) on even lanes, (B,A) on odd lanes --
// and therefore two independent rotations of NT accumulators.  Per pair of outputs this halves the
// per-step overhead the single-component kernels pay (one 64-bit load/store, one alpha shuffle,
// one reciprocal, three instead of four conversions), which matters because every half-rate
// FP64 / quarter-rate conversion instruction costs issue time (tools/micro/mix.cu).
// ======================================================================================

// Output stage of a pair: r = 1/clamp(den) (PerceptibleReciprocal on QS*den, morphology.c:3197),
// colour components are scaled by r, the alpha component (odd lane, .y) is stored unscaled.
__device__ __forceinline__ float2 finish_pair(bool odd, double sum0, double sum1, double gsum) {
  constexpr double kTiny = kEpsilon / kQuantumScale;
  double den = gsum;
  const int hi = __double2hiint(den);
  if ((hi & 0x7fffffff) < __double2hiint(kTiny))
    den = __hiloint2double((hi & 0x80000000) | __double2hiint(kTiny), __double2loint(kTiny));
  const double r = fast_reciprocal(den);
  const double m1 = odd ? 1.0 : r;
  return make_float2(static_cast<float>(sum0 * r), static_cast<float>(sum1 * m1));
}

// ---- pair stream kernel, both axes.  A thread owns one component pair and walks along the
//      filter axis with a register ring of PF samples in flight.
//      AXIS 1 (column pass): CTA = 128 threads = 256 consecutive components of a row (1 KB, fully
//        coalesced); step = one row.          grid (ceil(rc/256), ceil(height/strip))
//      AXIS 0 (row pass):    CTA = 128 threads = 64 rows x 2 pairs; step = one pixel (16 B); a
//        lane pair reads the 16 B of its pixel, consecutive steps of a lane fall in the same 128-B
//        line (L1-resident), so no shared-memory staging or barrier is needed.
//                                              grid (ceil(width/strip), ceil(height/64))
// IO selects the element types: 0 = float Quantum in, float Quantum out (one full pass);
// 1 = float in, RAW double sums out (first half of a rank-1 2-D kernel: no normalisation, no rounding);
// 2 = raw double sums in, float Quantum out (second half).  With IO 1 + 2 a separable 2-D kernel
// (e.g. "gaussian:RxS") is evaluated with kw + kh instead of kw * kh taps per sample while keeping the
// intermediate in double, i.e. without the float rounding a two-kernel list would introduce.
template <int NT, int MINB, int AXIS, int IO, bool L2PF = false>
__global__ void __launch_bounds__(128, MINB) conv_pair_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  constexpr int PF = Ring<NT>::value;
  constexpr unsigned kInB = (IO == 2) ? 8u : 4u, kOutB = (IO == 1) ? 8u : 4u;   // bytes per component
  using InT = typename std::conditional<IO == 2, double2, float2>::type;
  const int lane = threadIdx.x & 31;
  const bool odd = (threadIdx.x & 1) != 0;
  const int alpha_lane = lane | 1;
  const unsigned in_pitch = static_cast<unsigned>(a.rc) * kInB, out_pitch = static_cast<unsigned>(a.rc) * kOutB;
  int first, nout, limit;                 // first output / number of outputs / clamp limit along the axis
  const char *base;                       // address of sample 0 of this thread's line
  char *outp;
  unsigned step, ostep;                   // bytes between consecutive samples (input / output)
  if (AXIS == 1) {
    const int npairs = a.rc >> 1;
    const int pair_raw = blockIdx.x * 128 + threadIdx.x;
    const bool active = pair_raw < npairs;
    const int pair = active ? pair_raw : npairs - 1;
    first = blockIdx.y * a.strip;
    nout = active ? min(a.strip, a.height - first) : 0;
    limit = a.height - 1;
    step = in_pitch;
    ostep = out_pitch;
    base = reinterpret_cast<const char *>(a.src) + static_cast<size_t>(pair) * (2 * kInB);
    outp = reinterpret_cast<char *>(a.dst) + static_cast<size_t>(pair) * (2 * kOutB) + static_cast<size_t>(first) * out_pitch;
  } else {
    const int row_raw = blockIdx.y * 64 + (threadIdx.x >> 1);
    const bool active = row_raw < a.height;
    const int row = active ? row_raw : a.height - 1;
    first = blockIdx.x * a.strip;
    nout = active ? min(a.strip, a.width - first) : 0;
    limit = a.width - 1;
    step = 4 * kInB;
    ostep = 4 * kOutB;
    base = reinterpret_cast<const char *>(a.src) + static_cast<size_t>(row) * in_pitch + (odd ? 2 * kInB : 0);
    outp = reinterpret_cast<char *>(a.dst) + static_cast<size_t>(row) * out_pitch + (odd ? 2 * kOutB : 0) +
           static_cast<size_t>(first) * ostep;
  }
  const int total = a.strip + NT - 1;

  double acc0[NT], acc1[NT];
  InT pre[PF];
#pragma unroll
  for (int q = 0; q < NT; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
  int isrc = first - a.off;                 // source index of step 0
#pragma unroll
  for (int s = 0; s < PF; ++s) {
    const unsigned ii = static_cast<unsigned>(min(max(isrc + s, 0), limit));
    pre[s] = __ldg(reinterpret_cast<const InT *>(base + static_cast<size_t>(ii) * step));
  }
  isrc += PF;

  int j = -(NT - 1);
#pragma unroll 1
  for (int mb = 0; mb < total; mb += PF) {       // PF unrolled steps, then rotate the accumulators by PF
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const InT vf = pre[s];
      {
        const unsigned ii = static_cast<unsigned>(min(max(isrc, 0), limit));
        pre[s] = __ldg(reinterpret_cast<const InT *>(base + static_cast<size_t>(ii) * step));
        if (L2PF) {             // L2 prefetch far ahead: the ring's LDGs then complete at L2 latency
          const unsigned ip = static_cast<unsigned>(min(isrc + a.seg_w, limit));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + static_cast<size_t>(ip) * step));
        }
        ++isrc;
      }
      // (premultiplying one step ahead, as the cp.async kernel does, costs this kernel 5 %: 220 registers)
      double v0, v1;
      if (IO == 2) {
        v0 = vf.x; v1 = vf.y;                       // already premultiplied sums
      } else {
        const float af = __shfl_sync(0xffffffffu, static_cast<float>(vf.y), alpha_lane);
        const double da = static_cast<double>(af);
        v0 = static_cast<double>(vf.x) * da;
        v1 = static_cast<double>(vf.y) * (odd ? 1.0 : da);
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const double k = taps.k[(s - q + NT) % NT];
        acc0[q] = fma(k, v0, acc0[q]);
        acc1[q] = fma(k, v1, acc1[q]);
      }
      const int qf = (s + 1) % NT;
      const double sum0 = acc0[qf], sum1 = acc1[qf];
      acc0[qf] = 0.0;
      acc1[qf] = 0.0;
      if (IO == 1) {
        if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<double2 *>(outp) = make_double2(sum0, sum1);
      } else {
        const double gsum = shfl_double(sum1, alpha_lane);
        const float2 out = finish_pair(odd, sum0, sum1, gsum);
        if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<float2 *>(outp) = out;
      }
      if (j >= 0) outp += ostep;
      ++j;
    }
    if (PF != NT) {
      double t0[PF], t1[PF];
#pragma unroll
      for (int q = 0; q < PF; ++q) { t0[q] = acc0[q]; t1[q] = acc1[q]; }
#pragma unroll
      for (int q = 0; q < NT - PF; ++q) { acc0[q] = acc0[q + PF]; acc1[q] = acc1[q + PF]; }
#pragma unroll
      for (int q = 0; q < PF; ++q) { acc0[NT - PF + q] = t0[q]; acc1[NT - PF + q] = t1[q]; }
    }
  }
}

// ---- row pass: CTA = 4 independent warps; a warp covers 16 rows x 2 component pairs; every warp
//      streams its rows through a private cp.async.bulk ring (chunk = PF pixels per row, odd
//      pixel pitch => conflict-free LDS.64), so strips can be long and no block barrier is needed.
//      grid: (ceil(width/strip), ceil(height/64)).
template <int PF>
__device__ __forceinline__ void row_pair_issue(float *slot_base, unsigned long long *bar, const float4 *src, int ybase,
                                               int hmax, int xs, int width, int lane) {
  __syncwarp();
  if (lane == 0) mbar_expect_tx(bar, 16u * PF * 16u);
  __syncwarp();
  if (lane < 16) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const float4 *grow = src + static_cast<size_t>(min(ybase + lane, hmax)) * width;
    float *dst = slot_base + lane * (PF * 4);
    const int lo = max(xs, 0), hi = min(xs + PF, width);
    if (hi > lo) bulk_g2s(dst + (lo - xs) * 4, grow + lo, static_cast<unsigned>(hi - lo) * 16u, bar);
#pragma unroll 1
    for (int x = xs; x < min(xs + PF, 0); ++x) bulk_g2s(dst + (x - xs) * 4, grow, 16u, bar);
#pragma unroll 1
    for (int x = max(xs, width); x < xs + PF; ++x) bulk_g2s(dst + (x - xs) * 4, grow + (width - 1), 16u, bar);
  }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_row_pair_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  constexpr int PF = Ring<NT>::value;
  constexpr int kSlots = (PF > 11) ? 2 : 4;       // static shared memory stays below 48 KB
  static_assert(PF % 2 == 1, "odd chunk pitch keeps the LDS conflict-free");
  __shared__ __align__(128) float ring[4][kSlots][16][PF * 4];
  __shared__ __align__(8) unsigned long long full[4][kSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ybase = blockIdx.y * 64 + warp * 16;
  if (ybase >= a.height) return;
  const int lr = lane >> 1;
  const bool odd = (lane & 1) != 0;
  const int alpha_lane = lane | 1;
  const int x0 = blockIdx.x * a.strip;
  const int y = ybase + lr;
  const int hmax = a.height - 1;
  const int nout = y < a.height ? min(a.strip, a.width - x0) : 0;
  const int total = a.strip + NT - 1;
  const int nchunks = total / PF;
  float *outp = a.dst + (static_cast<size_t>(min(y, hmax)) * a.width + x0) * 4 + (odd ? 2 : 0);
  unsigned long long *bars = &full[warp][0];
  float *wring = &ring[warp][0][0][0];
  const float4 *src4 = reinterpret_cast<const float4 *>(a.src);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kSlots; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int xsrc0 = x0 - a.off;
  for (int ck = 0; ck < kSlots - 1 && ck < nchunks; ++ck)
    row_pair_issue<PF>(wring + ck * (16 * PF * 4), &bars[ck], src4, ybase, hmax, xsrc0 + ck * PF, a.width, lane);

  double acc0[NT], acc1[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
  int j = -(NT - 1);
  int chunk = 0, slot = 0, nslot = kSlots - 1;
  unsigned parity = 0;
#pragma unroll 1
  for (int mb = 0; mb < total; mb += PF) {       // one ring chunk == one unrolled block
    {
      const int nxt = chunk + kSlots - 1;
      if (nxt < nchunks)
        row_pair_issue<PF>(wring + nslot * (16 * PF * 4), &bars[nslot], src4, ybase, hmax, xsrc0 + nxt * PF, a.width,
                           lane);
      mbar_wait(&bars[slot], parity);
    }
    const float *cp = wring + ((slot * 16 + lr) * PF) * 4;
#pragma unroll
    for (int s = 0; s < PF; ++s) {
      const float *pp = cp + s * 4;
      const float2 vf = *reinterpret_cast<const float2 *>(pp + (odd ? 2 : 0));
      const float af = pp[3];
      const double da = static_cast<double>(af);
      const double v0 = static_cast<double>(vf.x) * da;
      const double v1 = static_cast<double>(vf.y) * (odd ? 1.0 : da);
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const double k = taps.k[(s - q + NT) % NT];
        acc0[q] = fma(k, v0, acc0[q]);
        acc1[q] = fma(k, v1, acc1[q]);
      }
      const int qf = (s + 1) % NT;
      const double sum0 = acc0[qf], sum1 = acc1[qf];
      acc0[qf] = 0.0;
      acc1[qf] = 0.0;
      const double gsum = shfl_double(sum1, alpha_lane);
      const float2 out = finish_pair(odd, sum0, sum1, gsum);
      if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<float2 *>(outp) = out;
      if (j >= 0) outp += 4;
      ++j;
    }
    ++chunk;
    nslot = slot;
    if (++slot == kSlots) { slot = 0; parity ^= 1u; }
    if (PF != NT) {
      double t0[PF], t1[PF];
#pragma unroll
      for (int q = 0; q < PF; ++q) { t0[q] = acc0[q]; t1[q] = acc1[q]; }
#pragma unroll
      for (int q = 0; q < NT - PF; ++q) { acc0[q] = acc0[q + PF]; acc1[q] = acc1[q + PF]; }
#pragma unroll
      for (int q = 0; q < PF; ++q) { acc0[NT - PF + q] = t0[q]; acc1[NT - PF + q] = t1[q]; }
    }
  }
}

// ---- pair stream kernel with a shared-memory prefetch ring (cp.async / LDGSTS), both axes.
// Same arithmetic and thread mapping as conv_pair_kernel; only the way source samples reach the
// thread differs.  With the register ring every in-flight LDG needs one of the six hardware
// scoreboards, which it has to share with the F2F / SHFL / MUFU results of the output stage: ncu
// showed 13 % (column) / 30 % (row) of all stall samples on output STGs that were waiting, through
// such a shared scoreboard, for an unrelated prefetch issued a few hundred cycles earlier.  cp.async
// groups are counted separately (DEPBAR.LE), cost no registers, and the ring depth is no longer tied
// to the rotation length.
//   AXIS 1: per-warp ring of NS rows x 256 B; a lane copies and later reads its own 8 bytes (no
//           cross-lane hazard, no barrier); LDS issued one step ahead.
//   AXIS 0: per-warp ring of 8-pixel chunks of its 16 rows (full 128-B line requests, 144-B row pitch
//           => conflict-free LDS.64); edge replication is applied by the loader.
template <int NT, int MINB, int AXIS, int IO>
__global__ void __launch_bounds__(128, MINB) conv_pair_async_kernel(const Conv1dArgs a, const Taps<NT> taps) {
  static_assert(IO == 0 || IO == 1, "float input only");
  // The asm statements below are volatile (ordered among themselves: copy -> commit -> wait -> LDS) but
  // carry no "memory" clobber: the ring is touched by nothing else, and the output STGs must stay free
  // to sink below the next step's copies (otherwise every step exposes the F2F -> STG latency).
  constexpr int UN = Ring<NT>::value;
  constexpr unsigned kOutB = (IO == 1) ? 8u : 4u;
  constexpr int NS = 16;                              // AXIS 1: rows in the ring
  constexpr int NC = 4, kPitch = 144, kSlot = 16 * kPitch;   // AXIS 0: chunk slots
  constexpr int kWarpBytes = AXIS == 1 ? NS * 256 : NC * kSlot;
  __shared__ __align__(128) unsigned char ring_all[4 * kWarpBytes];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool odd = (lane & 1) != 0;
  const int alpha_lane = lane | 1;
  const unsigned ring_s = static_cast<unsigned>(__cvta_generic_to_shared(ring_all + warp * kWarpBytes));
  const unsigned in_pitch = static_cast<unsigned>(a.rc) * 4u, out_pitch = static_cast<unsigned>(a.rc) * kOutB;
  int first, nout;
  char *outp;
  unsigned ostep;
  const int total = a.strip + NT - 1;

  // ---- loader state
  const char *lbase;            // AXIS 1: this lane's column; AXIS 0: image base
  int limit, isrc0;
  unsigned rd;                  // shared-memory address of this lane's next sample
  int chunk = 0, px_in = 0, c0 = 0, nchunks = 0, lrow0 = 0;
  if (AXIS == 1) {
    const int npairs = a.rc >> 1;
    const int pair_raw = blockIdx.x * 128 + threadIdx.x;
    const bool active = pair_raw < npairs;
    const int pair = active ? pair_raw : npairs - 1;
    first = blockIdx.y * a.strip;
    nout = active ? min(a.strip, a.height - first) : 0;
    limit = a.height - 1;
    ostep = out_pitch;
    lbase = reinterpret_cast<const char *>(a.src) + static_cast<size_t>(pair) * 8;
    outp = reinterpret_cast<char *>(a.dst) + static_cast<size_t>(pair) * (2 * kOutB) + static_cast<size_t>(first) * out_pitch;
    isrc0 = first - a.off;
    rd = ring_s + lane * 8;
  } else {
    const int row_raw = blockIdx.y * 64 + (threadIdx.x >> 1);
    const bool active = row_raw < a.height;
    first = blockIdx.x * a.strip;
    nout = active ? min(a.strip, a.width - first) : 0;
    limit = a.width - 1;
    ostep = 4 * kOutB;
    lbase = reinterpret_cast<const char *>(a.src);
    outp = reinterpret_cast<char *>(a.dst) + static_cast<size_t>(active ? row_raw : 0) * out_pitch + (odd ? 2 * kOutB : 0) +
           static_cast<size_t>(first) * ostep;
    isrc0 = first - a.off;
    c0 = isrc0 >= 0 ? isrc0 / 8 : -((7 - isrc0) / 8);            // floor(isrc0 / 8)
    px_in = isrc0 - 8 * c0;
    nchunks = (px_in + ((total + UN - 1) / UN) * UN + 7) / 8;
    lrow0 = blockIdx.y * 64 + warp * 16;
    rd = ring_s + (lane >> 1) * kPitch + (odd ? 8 : 0) + px_in * 16;
  }
  // AXIS 1: one row (step k) per group; AXIS 0: one 8-pixel chunk of the warp's 16 rows per group
  auto issue = [&](int k) {
    if (AXIS == 1) {
      const unsigned ii = static_cast<unsigned>(min(max(isrc0 + k, 0), limit));
      const unsigned dst = ring_s + static_cast<unsigned>(k & (NS - 1)) * 256u + lane * 8;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(lbase + static_cast<size_t>(ii) * in_pitch));
    } else if (k < nchunks) {
      const int x = min(max((c0 + k) * 8 + (lane & 7), 0), limit);
      const unsigned slot = ring_s + static_cast<unsigned>(k & (NC - 1)) * kSlot;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * i + (lane >> 3);
        const int y = min(lrow0 + r, a.height - 1);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot + r * kPitch + (lane & 7) * 16),
                     "l"(lbase + static_cast<size_t>(y) * in_pitch + static_cast<size_t>(x) * 16));
      }
    }
    asm volatile("cp.async.commit_group;" );
  };
  constexpr int kAhead = AXIS == 1 ? NS - 1 : NC;    // groups issued before the loop
#pragma unroll
  for (int k = 0; k < kAhead; ++k) issue(k);

  double acc0[NT], acc1[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
  float2 vnext;
  if (AXIS == 1) asm volatile("cp.async.wait_group %0;" ::"n"(NS - 2));
  else { asm volatile("cp.async.wait_group %0;" ::"n"(NC - 1)); __syncwarp(); }
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(vnext.x), "=f"(vnext.y) : "r"(rd));
  double nv0, nv1;                                 // premultiplied sample of the next step
  {
    const float af = __shfl_sync(0xffffffffu, vnext.y, alpha_lane);
    const double da = static_cast<double>(af);
    nv0 = static_cast<double>(vnext.x) * da;
    nv1 = static_cast<double>(vnext.y) * (odd ? 1.0 : da);
  }

  int j = -(NT - 1);
  int step = 0;
#pragma unroll 1
  for (int mb = 0; mb < total; mb += UN) {       // UN unrolled steps, then rotate the accumulators by UN
#pragma unroll
    for (int s = 0; s < UN; ++s) {
      const double v0 = nv0, v1 = nv1;
      // fetch the next step's sample and keep the ring full
      if (AXIS == 1) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NS - 3));      // row step+1 has landed
        rd = ring_s + static_cast<unsigned>((step + 1) & (NS - 1)) * 256u + lane * 8;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(vnext.x), "=f"(vnext.y) : "r"(rd));
        issue(step + NS - 1);                                                  // into the slot read at step-1
      } else {
        rd += 16;
        if (++px_in == 8) {                        // warp-uniform: chunk exhausted
          px_in = 0;
          __syncwarp();
          issue(chunk + NC);
          ++chunk;
          asm volatile("cp.async.wait_group %0;" ::"n"(NC - 1));
          __syncwarp();
          rd = ring_s + static_cast<unsigned>(chunk & (NC - 1)) * kSlot + (lane >> 1) * kPitch + (odd ? 8 : 0);
        }
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(vnext.x), "=f"(vnext.y) : "r"(rd));
      }
      ++step;
      {                                            // next step's SHFL -> F2F -> DMUL chain, under this step's FMAs
        const float af = __shfl_sync(0xffffffffu, vnext.y, alpha_lane);
        const double da = static_cast<double>(af);
        nv0 = static_cast<double>(vnext.x) * da;
        nv1 = static_cast<double>(vnext.y) * (odd ? 1.0 : da);
      }
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        const double k = taps.k[(s - q + NT) % NT];
        acc0[q] = fma(k, v0, acc0[q]);
        acc1[q] = fma(k, v1, acc1[q]);
      }
      const int qf = (s + 1) % NT;
      const double sum0 = acc0[qf], sum1 = acc1[qf];
      acc0[qf] = 0.0;
      acc1[qf] = 0.0;
      if (IO == 1) {
        if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<double2 *>(outp) = make_double2(sum0, sum1);
      } else {
        const double gsum = shfl_double(sum1, alpha_lane);
        const float2 out = finish_pair(odd, sum0, sum1, gsum);
        if (static_cast<unsigned>(j) < static_cast<unsigned>(nout)) *reinterpret_cast<float2 *>(outp) = out;
      }
      if (j >= 0) outp += ostep;
      ++j;
    }
    if (UN != NT) {
      double t0[UN], t1[UN];
#pragma unroll
      for (int q = 0; q < UN; ++q) { t0[q] = acc0[q]; t1[q] = acc1[q]; }
#pragma unroll
      for (int q = 0; q < NT - UN; ++q) { acc0[q] = acc0[q + UN]; acc1[q] = acc1[q + UN]; }
#pragma unroll
      for (int q = 0; q < UN; ++q) { acc0[NT - UN + q] = t0[q]; acc1[NT - UN + q] = t1[q]; }
    }
  }
  asm volatile("cp.async.wait_group 0;" );
}

// developer tuning knobs (environment, read on every launch; defaults are the tuned values)
int tuning(const char *name, int fallback) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : fallback;
}

template <int NT, int MODE>
int launch_nt(const Conv1dArgs &base, int axis, const double *taps_host, int ntaps, int io, cudaStream_t stream) {
  Taps<NT> taps;
  for (int i = 0; i < NT; ++i) taps.k[i] = i < ntaps ? taps_host[i] : 0.0;   // zero padding past the window
  Conv1dArgs a = base;
  const bool tma_ok = MODE == 4 && a.bias == 0.0 && tuning("MB200_TMA", 0) != 0 && (a.rc % 32) == 0 &&
                      ((reinterpret_cast<uintptr_t>(a.src) | reinterpret_cast<uintptr_t>(a.dst)) & 15) == 0;
  if (io != 0 && !(MODE == 4 && NT <= 33)) return MB200_EUNSUPPORTED;
  const bool pair_ok = MODE == 4 && a.bias == 0.0 && NT <= 33 && (io != 0 || tuning("MB200_PAIR", 1) != 0) &&
                       ((reinterpret_cast<uintptr_t>(a.src) | reinterpret_cast<uintptr_t>(a.dst)) & 15) == 0;
  if (io != 0 && !pair_ok) return MB200_EUNSUPPORTED;
  if (pair_ok && axis == 1) {
    if constexpr (NT <= 33) {
      // strip = 16 rotations.  (A list-scheduling model of the launch -- equal-cost CTAs on 2 slots per SM -- preferred
      // 27 rotations for 8192 rows, 6 % fewer modelled steps; measured it is 5 % SLOWER, 0.822 vs 0.778 ms: CTAs that
      // run alone on an SM in the last wave get the whole FP64 pipe, so the tail balances itself.)
      a.strip = tuning("MB200_COL_ROT", 16) * NT + 1;
      a.seg_w = 16;                                  // rows of L2 prefetch ahead of the register ring (L2PF kernels)
      dim3 grid((a.rc / 2 + 127) / 128, (a.height + a.strip - 1) / a.strip);
      // NT = 33 is FP64-bound: the register-ring kernel (+ L2 prefetch) wins; shorter windows are closer to
      // the HBM roof and gain from the cp.async ring (sigma=2: 1.36 -> 1.22 ms for the whole blur).
      const bool async = tuning("MB200_PAIR_ASYNC_COL", NT < 33 ? 1 : 0) != 0 && tuning("MB200_PAIR_ASYNC", 1) != 0;
      if (io == 2) conv_pair_kernel<NT, 2, 1, 2><<<grid, 128, 0, stream>>>(a, taps);
      else if (io == 1 && async) conv_pair_async_kernel<NT, 2, 1, 1><<<grid, 128, 0, stream>>>(a, taps);
      else if (io == 1) conv_pair_kernel<NT, 2, 1, 1, NT == 33><<<grid, 128, 0, stream>>>(a, taps);
      else if (async) conv_pair_async_kernel<NT, 2, 1, 0><<<grid, 128, 0, stream>>>(a, taps);
      else conv_pair_kernel<NT, 2, 1, 0, NT == 33><<<grid, 128, 0, stream>>>(a, taps);
    }
  } else if (pair_ok && axis == 0) {
    if constexpr (NT <= 33) {
      a.strip = tuning("MB200_ROW_PAIR_ROT", 16) * NT + 1;
      dim3 grid((a.width + a.strip - 1) / a.strip, (a.height + 63) / 64);
      const bool async = tuning("MB200_PAIR_ASYNC", 1) != 0;
      if (io == 1 && async) conv_pair_async_kernel<NT, 2, 0, 1><<<grid, 128, 0, stream>>>(a, taps);
      else if (io == 1) conv_pair_kernel<NT, 2, 0, 1><<<grid, 128, 0, stream>>>(a, taps);
      else if (io == 2) conv_pair_kernel<NT, 2, 0, 2><<<grid, 128, 0, stream>>>(a, taps);
      else if (async && !tuning("MB200_ROW_PAIR_TMA", 0)) conv_pair_async_kernel<NT, 2, 0, 0><<<grid, 128, 0, stream>>>(a, taps);
      else if (tuning("MB200_ROW_PAIR_TMA", 0)) conv_row_pair_kernel<NT, 2><<<grid, 128, 0, stream>>>(a, taps);
      else conv_pair_kernel<NT, 2, 0, 0><<<grid, 128, 0, stream>>>(a, taps);
    }
  } else if (tma_ok && axis == 1) {
    constexpr int kMinBlocks = NT <= 33 ? 4 : 2;
    a.strip = tuning("MB200_COL_ROT", 8) * NT + 1;
    dim3 grid((a.rc + 127) / 128, (a.height + a.strip - 1) / a.strip);
    if (NT == 33 && tuning("MB200_MINB", 3) == 3) conv_col_tma_kernel<NT, 3><<<grid, 128, 0, stream>>>(a, taps);
    else conv_col_tma_kernel<NT, kMinBlocks><<<grid, 128, 0, stream>>>(a, taps);
  } else if (tma_ok && axis == 0) {
    constexpr int kMinBlocks = NT <= 33 ? 4 : 2;
    a.strip = tuning("MB200_ROW_TMA_ROT", 8) * NT + 1;
    dim3 grid((a.width + a.strip - 1) / a.strip, (a.height + 31) / 32);
    if (NT == 33 && tuning("MB200_MINB", 3) == 3) conv_row_tma_kernel<NT, 3><<<grid, 128, 0, stream>>>(a, taps);
    else conv_row_tma_kernel<NT, kMinBlocks><<<grid, 128, 0, stream>>>(a, taps);
  } else if (axis == 1) {
    constexpr int kThreads = 128;
    constexpr int kMinBlocks = NT <= 33 ? 4 : 2;
    a.strip = tuning("MB200_COL_ROT", 16) * NT + 1;   // strip + NT - 1 is a whole number of rotations
    dim3 grid((a.rc + kThreads - 1) / kThreads, (a.height + a.strip - 1) / a.strip);
    conv_col_kernel<NT, MODE, kThreads, kMinBlocks><<<grid, kThreads, 0, stream>>>(a, taps);
  } else {
    constexpr int kMinBlocks = NT <= 33 ? 4 : 2;
    // strip + NT - 1 is a whole number of rotations and the strip is at least ~64 outputs
    constexpr int kRot = (63 + NT - 1) / NT < 3 ? 3 : (63 + NT - 1) / NT;
    a.strip = tuning("MB200_ROW_ROT", kRot) * NT + 1;
    a.seg_w = a.strip + NT - 1;
    a.pitch = a.seg_w | 1;
    const int rows_per_cta = 4 * (32 / a.channels);
    const size_t smem = static_cast<size_t>(rows_per_cta) * a.pitch * a.channels * sizeof(float);
    if (smem > 48 * 1024)            // per device attribute: set on every launch that needs it (a few microseconds)
      cudaFuncSetAttribute(conv_row_kernel<NT, MODE, kMinBlocks>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid((a.width + a.strip - 1) / a.strip, (a.height + rows_per_cta - 1) / rows_per_cta);
    conv_row_kernel<NT, MODE, kMinBlocks><<<grid, 128, smem, stream>>>(a, taps);
  }
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv1d launch");
  return MB200_OK;
}

template <int NT>
int launch_mode(const Conv1dArgs &a, int axis, const double *taps, int ntaps, int io, cudaStream_t s) {
  if (a.channels == 4) return launch_nt<NT, 4>(a, axis, taps, ntaps, io, s);
  if (io != 0) return MB200_EUNSUPPORTED;
  if (a.channels == 2) return launch_nt<NT, 2>(a, axis, taps, ntaps, io, s);
  return launch_nt<NT, 0>(a, axis, taps, ntaps, io, s);
}

}  // namespace

int launch_conv1d(const float *src, float *dst, size_t width, size_t height, int channels, int axis,
                  const double *taps, int ntaps, int origin_offset, double bias, double /*gamma_scale*/,
                  unsigned long long *d_changed, void *stream, int io) {
  if (width == 0 || height == 0 || channels < 1 || channels > 4 || ntaps < 1)
    return fail(MB200_EINVAL, "conv1d: bad geometry");
  if (width * channels > 0x1fffffffull || height > 0x7fffffffull) return MB200_EUNSUPPORTED;   // 32-bit byte pitch
  if (d_changed != nullptr) return MB200_EUNSUPPORTED;   // `changed` counting lives in the generic kernel
  Conv1dArgs a{};
  a.src = src; a.dst = dst;
  a.width = static_cast<int>(width); a.height = static_cast<int>(height); a.channels = channels;
  a.rc = a.width * channels;
  a.off = origin_offset;
  a.bias = bias;
  a.changed = d_changed;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ntaps <= 9) return launch_mode<9>(a, axis, taps, ntaps, io, s);
  if (ntaps <= 17) return launch_mode<17>(a, axis, taps, ntaps, io, s);
  if (ntaps <= 25) return launch_mode<25>(a, axis, taps, ntaps, io, s);
  if (ntaps <= 33) return launch_mode<33>(a, axis, taps, ntaps, io, s);
  if (ntaps <= 49) return launch_mode<49>(a, axis, taps, ntaps, io, s);
  if (ntaps <= 65) return launch_mode<65>(a, axis, taps, ntaps, io, s);
  return MB200_EUNSUPPORTED;   // caller falls back to the generic 2-D kernel
}

}  // namespace mb200
