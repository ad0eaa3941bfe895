// conv_mma.cu -- one pass of a separable (1-D) convolution on RGBA images, evaluated on the FP64 matrix path
// (mma.sync.m8n8k4.f64, SASS DMMA).  Same semantics as conv1d.cu (MorphologyPrimitive's ConvolveMorphology for
// height-1 / width-1 kernels, MagickCore/morphology.c:2897-2979 and :2654-2807: reflected taps, edge-clamped source,
// double accumulation, alpha-weighted colour channels, one rounding to float).
//
// Why a matrix formulation for a stencil.  The DFMA streaming kernels of conv1d.cu are bound by ISSUE SLOTS next to
// the half-rate FP64 pipe (profiles/r02_conv_pair.md: 124 instructions per 148 pipe cycles, two warps per scheduler
// because 66 FP64 accumulators need 206-228 registers).  DMMA shares the FP64 unit at the same FMA rate
// (tools/micro/dmma.cu) but one instruction carries 256 FMAs: ~7 instructions per DMMA instead of 1.7 per DFMA, 8
// accumulator registers per tile, four CTAs per SM.  The price is the band structure: 8 consecutive outputs of a 33-tap
// window touch 40 source samples, i.e. ten 8x4 Toeplitz tiles of which 17.5 % of the cells are zero (17 taps: 29 %,
// 9 taps: 44 %).  Measured (profiles/r02_conv_mma.md): the FP64 unit is saturated either way (math_pipe_throttle is the
// top stall), so at burst clocks the zero cells cancel what the freed issue slots buy for 33 taps (0.78-0.80 ms per
// pass on both paths) while short windows gain 4-19 %; under sustained load this path needs less power per FMA, keeps
// the SM clock at its maximum where the DFMA kernels are power-capped, and is ~5 % faster for 33 taps as well.
//
//   D[m][n] += A[m][k] * B[k][n]     m = 8 consecutive outputs along the filter axis
//                                    k = 4 consecutive source samples (one of NKS k-steps)
//                                    n = 8 independent component lines
//   A = taps as a banded Toeplitz tile: A_s[m][k] = tap[4 s + k - m] (zero outside the window) -- NKS doubles per
//       lane, loaded once per thread and reused for the whole strip;
//   B = source samples, converted to double and alpha-premultiplied ONCE when they are staged into a per-warp
//       shared-memory ring (q = A*p; the weight sum of the blend is the alpha line's own result, as in conv1d.cu);
//   D = two doubles per lane and tile.  Tiles come in (R,G) / (B,A) plane pairs whose n index is mapped so that a
//       lane ends up with all four sums of ONE pixel: one reciprocal per pixel, no shuffles, one 16-byte store.
//
// A warp is self-contained (its own ring, __syncwarp only): it owns 8 pixels x all rows of a strip (AXIS 1) or 8 image
// lines x all pixels of a strip (AXIS 0) and advances 8 outputs per iteration: stage 8 new source positions (2 pixels
// per lane, loaded one iteration ahead), 4 tiles x NKS DMMAs, output stage.
//
// Non-finite samples.  Every tile multiplies samples outside an output's window by a ZERO tap; 0 * (inf | NaN) would
// poison outputs the reference computes from finite samples only.  The loader therefore tests every sample's exponent
// field and keeps one flag per 8-position block of the ring; an output block whose window holds a flagged block is
// evaluated by a scalar DFMA loop over the window's own taps instead (rare: HDRI images with inf / NaN pixels).
#include "mb200_internal.h"
#include "conv_common.cuh"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>

namespace mb200 {
namespace {

template <int NKS>
struct MmaTaps { double k[4 * NKS]; };      // window order, zero past the window

struct MmaArgs {
  const void *src;
  void *dst;
  int width, height;      // pixels
  int off;                // samples of the window before the output position
  int ntaps;
  int strip;              // outputs per strip along the filter axis (multiple of 8)
  const float *aux;       // EPI = 1: UnsharpMaskImage's source image
  double gain, qthreshold;
  int l2pf;               // prefetch.global.L2 ahead of the register prefetch
};

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// Raw loads of one block (two pixels per lane) kept in registers for one iteration.
template <int IO> struct Raw;
template <> struct Raw<0> { float4 a; };
template <> struct Raw<2> { double2 a, b; };

template <int NKS, int AXIS, int IO, int EPI, int MINB>
__global__ void __launch_bounds__(128, MINB) conv_mma_kernel(const MmaArgs a, const MmaTaps<NKS> taps) {
  static_assert(NKS % 2 == 0, "the prologue stages whole 8-position blocks");
  static_assert(EPI == 0 || (AXIS == 1 && IO == 0), "the fused epilogue belongs to the final column pass");
  constexpr int NPRE = (4 * NKS - 8) / 8;          // blocks staged before the first output block
  constexpr int NB = NPRE + 2;                     // ring blocks: window (NPRE + 1) + the one being refilled
  constexpr int RR = 8 * NB;                       // ring positions along the filter axis
  constexpr int PW = 36;                           // AXIS 1: doubles per ring row: [RG of 8 px | BA of 8 px | pad 4]
  constexpr int PL = 2 * RR + 8;                   // AXIS 0: doubles per image line of a plane (== 8 mod 16)
  constexpr int kWarpDoubles = AXIS == 1 ? RR * PW : 16 * PL;
  constexpr int kInB = IO == 2 ? 32 : 16, kOutB = IO == 1 ? 32 : 16;      // bytes per pixel
  constexpr int RIO = IO == 2 ? 2 : 0;
  extern __shared__ __align__(16) double ring_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *ring = ring_all + warp * kWarpDoubles;
  const unsigned ring_s = static_cast<unsigned>(__cvta_generic_to_shared(ring));
  const int k4 = lane & 3, n8 = lane >> 2;         // fragment coordinates: A[m = n8][k = k4], B[k4][n8], D[n8][2 k4 + e]

  int first, nout, limit, par0;
  if (AXIS == 1) {
    par0 = (blockIdx.x * 4 + warp) * 8;            // first pixel column of this warp
    if (par0 >= a.width) return;                   // (no CTA-wide barrier anywhere: a warp may leave)
    first = blockIdx.y * a.strip;
    nout = min(a.strip, a.height - first);
    limit = a.height - 1;
  } else {
    par0 = (blockIdx.y * 4 + warp) * 8;            // first image line of this warp
    if (par0 >= a.height) return;
    first = blockIdx.x * a.strip;
    nout = min(a.strip, a.width - first);
    limit = a.width - 1;
  }
  const int nblocks = (nout + 7) >> 3;
  const bool mma_l2pf = a.l2pf != 0;
  const size_t in_pitch = static_cast<size_t>(a.width) * kInB, out_pitch = static_cast<size_t>(a.width) * kOutB;

  // ---- loader: two pixels per lane and block
  const int lq = lane & 7, lh = lane >> 3;
  const char *lbase[2];
  int uoff[2];
  unsigned st_off[2];                               // ring offset (doubles) of the RG pair inside block slot 0
  if (AXIS == 1) {
    const int xs = min(par0 + lq, a.width - 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      lbase[i] = static_cast<const char *>(a.src) + static_cast<size_t>(xs) * kInB;
      uoff[i] = lh + 4 * i;
      st_off[i] = static_cast<unsigned>(uoff[i] * PW + 2 * lq);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int line = lh + 4 * i;
      lbase[i] = static_cast<const char *>(a.src) + static_cast<size_t>(min(par0 + line, a.height - 1)) * in_pitch;
      uoff[i] = lq;
      st_off[i] = static_cast<unsigned>(line * PL + 2 * lq);
    }
  }
  constexpr unsigned kBlockStride = AXIS == 1 ? 8 * PW : 16;     // ring doubles per block slot
  constexpr unsigned kBA = AXIS == 1 ? 16 : 8 * PL;               // RG -> BA plane
  const size_t lstep = AXIS == 1 ? in_pitch : static_cast<size_t>(kInB);
  const int base = first - a.off;                   // source position of ring block 0, element 0

  Raw<RIO> raw[2];
  auto fetch = [&](int j) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned pos = static_cast<unsigned>(min(max(base + 8 * j + uoff[i], 0), limit));
      const char *p = lbase[i] + static_cast<size_t>(pos) * lstep;
      if constexpr (IO == 2) {
        raw[i].a = __ldg(reinterpret_cast<const double2 *>(p));
        raw[i].b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
      } else {
        raw[i].a = __ldg(reinterpret_cast<const float4 *>(p));
      }
    }
  };
  // L2 prefetch kL2Ahead blocks beyond the register prefetch: the LDGs of `fetch` then complete at L2 latency.  Short
  // windows run so few DMMAs per iteration that two blocks of loads in flight per warp do not cover DRAM latency
  // (9 taps: 0.52 ms per pass against 0.33 ms of HBM time before this).
  constexpr int kL2Ahead = 8;
  auto prefetch_l2 = [&](int j) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const unsigned pos = static_cast<unsigned>(min(max(base + 8 * j + uoff[i], 0), limit));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(lbase[i] + static_cast<size_t>(pos) * lstep));
    }
  };
  unsigned badmask = 0;
  auto stage = [&](int slot) {                      // registers -> ring block `slot`; updates the block's flag
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double v0, v1, v2, v3;
      if constexpr (IO == 2) {
        v0 = raw[i].a.x; v1 = raw[i].a.y; v2 = raw[i].b.x; v3 = raw[i].b.y;
        bad = bad || nonfinite_bits(v0) || nonfinite_bits(v1) || nonfinite_bits(v2) || nonfinite_bits(v3);
      } else {
        const float4 f = raw[i].a;
        const unsigned m = max(max(__float_as_uint(f.x) & 0x7fffffffu, __float_as_uint(f.y) & 0x7fffffffu),
                               max(__float_as_uint(f.z) & 0x7fffffffu, __float_as_uint(f.w) & 0x7fffffffu));
        bad = bad || m >= 0x7f800000u;
        const double da = static_cast<double>(f.w);
        v0 = static_cast<double>(f.x) * da;
        v1 = static_cast<double>(f.y) * da;
        v2 = static_cast<double>(f.z) * da;
        v3 = da;
      }
      double *q = ring + st_off[i] + static_cast<unsigned>(slot) * kBlockStride;
      *reinterpret_cast<double2 *>(q) = make_double2(v0, v1);
      *reinterpret_cast<double2 *>(q + kBA) = make_double2(v2, v3);
    }
    const bool any = __any_sync(0xffffffffu, bad);
    badmask = any ? (badmask | (1u << slot)) : (badmask & ~(1u << slot));
  };

  // ---- tap tiles: A_s[m][k] = tap[4 s + k - m]
  double afrag[NKS];
#pragma unroll
  for (int s = 0; s < NKS; ++s) {
    const int t = 4 * s + k4 - n8;
    const double v = taps.k[min(max(t, 0), 4 * NKS - 1)];
    afrag[s] = (t >= 0 && t < a.ntaps) ? v : 0.0;
  }

  // ---- B fragment addressing
  const unsigned frag_off = AXIS == 1 ? static_cast<unsigned>(k4 * PW + n8)
                                      : static_cast<unsigned>(2 * k4 + (n8 >> 1) * PL + (n8 & 1));
  constexpr unsigned kUStride = AXIS == 1 ? PW : 2;            // ring doubles per position
  constexpr unsigned kT1 = kBA;                                 // (group 0, BA)
  constexpr unsigned kT2 = AXIS == 1 ? 8 : 4 * PL;              // (group 1, RG)
  constexpr unsigned kT3 = kT2 + kBA;                           // (group 1, BA)
  // scalar path / output coordinates of this lane: output m = n8 of the block, pixel (AXIS 1) or line (AXIS 0) 4 g + k4
  const unsigned lane_px_off = AXIS == 1 ? static_cast<unsigned>(2 * k4) : static_cast<unsigned>(k4 * PL);
  constexpr unsigned kGroup = AXIS == 1 ? 8 : 4 * PL;

  // ---- output stage of one block: the lane holds pixel (m = n8, 4 g + k4) with all four sums
  char *outp;                                       // output of (block, m = n8, group 0); lags the MMAs by one block
  const char *epip = nullptr;
  if (AXIS == 1) {
    outp = static_cast<char *>(a.dst) + static_cast<size_t>(first + n8) * out_pitch + static_cast<size_t>(par0 + k4) * kOutB;
    if (EPI) epip = reinterpret_cast<const char *>(a.aux) + static_cast<size_t>(first + n8) * out_pitch + static_cast<size_t>(par0 + k4) * kOutB;
  } else {
    outp = static_cast<char *>(a.dst) + static_cast<size_t>(par0 + k4) * out_pitch + static_cast<size_t>(first + n8) * kOutB;
  }
  const size_t ostep = AXIS == 1 ? 8 * out_pitch : static_cast<size_t>(8 * kOutB);        // per block
  const size_t gstep = AXIS == 1 ? static_cast<size_t>(4 * kOutB) : 4 * out_pitch;        // group 0 -> group 1
  bool gvalid[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) gvalid[g] = (par0 + 4 * g + k4) < (AXIS == 1 ? a.width : a.height);
  auto output = [&](const double (&acc)[4][2], const float4 (&epi)[2], bool mvalid) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const double sr = acc[2 * g][0], sg = acc[2 * g][1], sbv = acc[2 * g + 1][0], sa = acc[2 * g + 1][1];
      char *o = outp + g * gstep;
      if (IO == 1) {
        if (mvalid && gvalid[g]) {
          *reinterpret_cast<double2 *>(o) = make_double2(sr, sg);
          *reinterpret_cast<double2 *>(o + 16) = make_double2(sbv, sa);
        }
      } else {
        const double r = fast_reciprocal(clamp_denominator(sa));
        float4 out = make_float4(static_cast<float>(sr * r), static_cast<float>(sg * r), static_cast<float>(sbv * r),
                                 static_cast<float>(sa));
        if (EPI) {
          out.x = unsharp_point(epi[g].x, out.x, a.gain, a.qthreshold);
          out.y = unsharp_point(epi[g].y, out.y, a.gain, a.qthreshold);
          out.z = unsharp_point(epi[g].z, out.z, a.gain, a.qthreshold);
          out.w = unsharp_point(epi[g].w, out.w, a.gain, a.qthreshold);
        }
        if (mvalid && gvalid[g]) *reinterpret_cast<float4 *>(o) = out;
      }
    }
  };

  // ---- prologue: the NPRE + 1 blocks of the first window (all loads in flight together), then the block after it
  {
    Raw<RIO> pre[NPRE + 1][2];
#pragma unroll
    for (int j = 0; j <= NPRE; ++j) { fetch(j); pre[j][0] = raw[0]; pre[j][1] = raw[1]; }
#pragma unroll
    for (int j = 0; j <= NPRE; ++j) { raw[0] = pre[j][0]; raw[1] = pre[j][1]; stage(j); }
  }
  Raw<RIO> ahead[2];                                 // the block after the one in `raw` (loads two iterations in flight)
  fetch(NPRE + 2);
  ahead[0] = raw[0]; ahead[1] = raw[1];
  fetch(NPRE + 1);

  // ---- main loop, software pipelined by hand: iteration b issues the DMMAs of block b and, in the same basic block,
  // stages the next source block into the ring slot no window of this iteration reads and runs the output stage of
  // block b - 1.  The warp issues in order, so the conversions / reciprocals / stores ride in the gaps between its DMMAs
  // instead of forming phases of their own in which the FP64 pipe is left to the other warps (first version, separate
  // phases: pipe 78 % busy).
  int st_slot = NPRE + 1, sb = 0;                    // slot staged in this iteration / first slot of the window
  double prev[4][2];
  float4 epi_prev[2], epi_cur[2];
#pragma unroll
  for (int t = 0; t < 4; ++t) { prev[t][0] = 0.0; prev[t][1] = 1.0; }
#pragma unroll
  for (int g = 0; g < 2; ++g) epi_prev[g] = epi_cur[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  bool prev_valid = false;                           // block -1 does not exist

#pragma unroll 1
  for (int b = 0; b < nblocks; ++b) {
    __syncwarp();
    const bool mvalid = 8 * b + n8 < nout;
    if (EPI) {
#pragma unroll
      for (int g = 0; g < 2; ++g)
        if (mvalid && gvalid[g]) epi_cur[g] = __ldg(reinterpret_cast<const float4 *>(epip + g * gstep));
      epip += ostep;
    }
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    {
      // B fragments through a 3-deep register ring, loaded two k-steps ahead of the DMMAs that consume them
      double bq[3][4];
      unsigned ua[NKS];
      {
        int ub = 8 * sb;
#pragma unroll
        for (int s = 0; s < NKS; ++s) {
          ua[s] = ring_s + (frag_off + static_cast<unsigned>(ub) * kUStride) * 8u;
          ub += 4;
          if (ub == RR) ub = 0;
        }
      }
      auto lds4 = [&](double (&bv)[4], unsigned addr) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bv[0]) : "r"(addr));
        asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(bv[1]) : "r"(addr), "n"(kT1 * 8));
        asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(bv[2]) : "r"(addr), "n"(kT2 * 8));
        asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(bv[3]) : "r"(addr), "n"(kT3 * 8));
      };
      lds4(bq[0], ua[0]);
      lds4(bq[1], ua[1]);
#pragma unroll
      for (int s = 0; s < NKS; ++s) {
        if (s + 2 < NKS) lds4(bq[(s + 2) % 3], ua[s + 2]);
        dmma(acc[0], afrag[s], bq[s % 3][0]);
        dmma(acc[1], afrag[s], bq[s % 3][1]);
        dmma(acc[2], afrag[s], bq[s % 3][2]);
        dmma(acc[3], afrag[s], bq[s % 3][3]);
      }
    }
    const unsigned window_bad = badmask & ~(1u << st_slot);     // flags of the slots this block's windows read
    // next source block -> ring (the slot outside this iteration's windows), the one after it -> registers
    stage(st_slot);
    raw[0] = ahead[0]; raw[1] = ahead[1];
    {
      const Raw<RIO> keep0 = raw[0], keep1 = raw[1];
      fetch(b + NPRE + 3);
      if (mma_l2pf) prefetch_l2(b + NPRE + 3 + kL2Ahead);
      ahead[0] = raw[0]; ahead[1] = raw[1];
      raw[0] = keep0; raw[1] = keep1;
    }
    // output stage of the previous block
    output(prev, epi_prev, prev_valid);
    if (b > 0) outp += ostep;

    if (window_bad != 0) {
      // a window of this block holds a non-finite sample: the window's own taps only, in scalar FMAs
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int u = 8 * sb + n8;
        if (u >= RR) u -= RR;
        for (int t = 0; t < a.ntaps; ++t) {
          const double *p = ring + static_cast<unsigned>(u) * kUStride + lane_px_off + g * kGroup;
          const double2 rg = *reinterpret_cast<const double2 *>(p);
          const double2 ba = *reinterpret_cast<const double2 *>(p + kBA);
          const double k = taps.k[t];
          s0 = fma(k, rg.x, s0); s1 = fma(k, rg.y, s1); s2 = fma(k, ba.x, s2); s3 = fma(k, ba.y, s3);
          if (++u == RR) u = 0;
        }
        acc[2 * g][0] = s0; acc[2 * g][1] = s1; acc[2 * g + 1][0] = s2; acc[2 * g + 1][1] = s3;
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) { prev[t][0] = acc[t][0]; prev[t][1] = acc[t][1]; }
    if (EPI) { epi_prev[0] = epi_cur[0]; epi_prev[1] = epi_cur[1]; }
    prev_valid = mvalid;
    st_slot = st_slot + 1 == NB ? 0 : st_slot + 1;
    if (++sb == NB) sb = 0;
  }
  output(prev, epi_prev, prev_valid);
}

struct MmaTuning {
  int enable, strip, minb, l2pf;
  MmaTuning() {
    auto get = [](const char *name, int fallback) {
      const char *v = getenv(name);
      return (v && *v) ? atoi(v) : fallback;
    };
    enable = get("MB200_MMA", -1);      // -1: automatic (float in / float out passes), 0: never, 1: whenever possible
    strip = get("MB200_MMA_STRIP", 512);
    minb = get("MB200_MMA_MINB", 4);
    l2pf = get("MB200_MMA_L2PF", -1);    // -1: windows of <= 9 taps (measured: 9 taps 1.05 -> 0.98 ms, 25 taps 1.37 -> 1.40 ms)
  }
};
MmaTuning &mma_tuning() {
  static MmaTuning t;
  return t;
}
std::atomic<unsigned long long> g_mma_launches{0};

template <int NKS, int AXIS, int IO, int EPI>
int launch_one(const MmaArgs &a, const MmaTaps<NKS> &taps, cudaStream_t stream) {
  constexpr int NB = (4 * NKS - 8) / 8 + 2, RR = 8 * NB;
  constexpr int kWarpDoubles = AXIS == 1 ? RR * 36 : 16 * (2 * RR + 8);
  constexpr size_t smem = 4 * kWarpDoubles * sizeof(double);
  dim3 grid;
  if (AXIS == 1) grid = dim3((a.width + 31) / 32, (a.height + a.strip - 1) / a.strip);
  else grid = dim3((a.width + a.strip - 1) / a.strip, (a.height + 31) / 32);
  auto go = [&](auto kernel) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kernel<<<grid, 128, smem, stream>>>(a, taps);
  };
  if (mma_tuning().minb >= 4) go(conv_mma_kernel<NKS, AXIS, IO, EPI, 4>);
  else go(conv_mma_kernel<NKS, AXIS, IO, EPI, 3>);
  count_launch();
  g_mma_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv_mma launch");
  return MB200_OK;
}

template <int NKS>
int launch_nks(const MmaArgs &a, const double *taps_host, int axis, int io, bool epi, cudaStream_t stream) {
  MmaTaps<NKS> taps;
  for (int i = 0; i < 4 * NKS; ++i) taps.k[i] = i < a.ntaps ? taps_host[i] : 0.0;
  if (axis == 1) {
    if (io == 1) return launch_one<NKS, 1, 1, 0>(a, taps, stream);
    if (io == 2) return launch_one<NKS, 1, 2, 0>(a, taps, stream);
    if (epi) return launch_one<NKS, 1, 0, 1>(a, taps, stream);
    return launch_one<NKS, 1, 0, 0>(a, taps, stream);
  }
  if (io == 1) return launch_one<NKS, 0, 1, 0>(a, taps, stream);
  if (io == 2) return launch_one<NKS, 0, 2, 0>(a, taps, stream);
  return launch_one<NKS, 0, 0, 0>(a, taps, stream);
}

}  // namespace

void set_conv_mma(int enable) { mma_tuning().enable = enable; }
int conv_mma_enabled() { return mma_tuning().enable; }
unsigned long long conv_mma_launches() { return g_mma_launches.load(std::memory_order_relaxed); }

// RGBA, bias 0, 16-byte aligned images, <= 33 taps.  MB200_EUNSUPPORTED => the caller uses the DFMA kernels of conv1d.cu.
int launch_conv_mma(const void *src, void *dst, size_t width, size_t height, int axis, const double *taps, int ntaps,
                    int origin_offset, void *stream, int io, const UnsharpEpilogue *epilogue, bool *epilogue_fused) {
  if (epilogue_fused) *epilogue_fused = false;
  // Default (-1): every float-in / float-out pass of <= 33 taps.  Measured on 8192^2 (profiles/r02_conv_mma.md): at burst
  // clocks the matrix path wins for short windows (9 taps 1.04 vs 1.24 ms, 17 taps 1.18 vs 1.22 ms) and is 3 % behind for
  // 33 taps (1.61 vs 1.56 ms: both paths saturate the FP64 unit, this one spends 17.5 % of it on zero cells) -- but it
  // draws less power per FMA: in a >= 1.3 s region the DFMA kernels sit on the 1000 W cap at 1845 MHz while this path
  // stays at 1965 MHz, and bench.py's step is 4.7 % faster (64.4 vs 67.4 ms per 32 images).  The rank-1 passes with a
  // double intermediate (io 1 / 2) are equal under load and slower at burst clocks: they keep the DFMA kernels.
  const int mode = mma_tuning().enable;
  if (mode == 0 || (mode < 0 && io != 0)) return MB200_EUNSUPPORTED;
  if (ntaps < 1 || ntaps > 33 || width * 32 > 0x7fffffffull || height > 0x3fffffffull) return MB200_EUNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) != 0) return MB200_EUNSUPPORTED;
  MmaArgs a{};
  a.src = src; a.dst = dst;
  a.width = static_cast<int>(width); a.height = static_cast<int>(height);
  a.off = origin_offset;
  a.ntaps = ntaps;
  a.strip = (mma_tuning().strip + 7) & ~7;
  a.l2pf = mma_tuning().l2pf < 0 ? (ntaps <= 9 ? 1 : 0) : mma_tuning().l2pf;
  const bool epi = axis == 1 && io == 0 && epilogue && epilogue->source && (reinterpret_cast<uintptr_t>(epilogue->source) & 15) == 0;
  if (epi) { a.aux = epilogue->source; a.gain = epilogue->gain; a.qthreshold = epilogue->quantum_threshold; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  if (ntaps <= 9) rc = launch_nks<4>(a, taps, axis, io, epi, s);
  else if (ntaps <= 17) rc = launch_nks<6>(a, taps, axis, io, epi, s);
  else if (ntaps <= 25) rc = launch_nks<8>(a, taps, axis, io, epi, s);
  else rc = launch_nks<10>(a, taps, axis, io, epi, s);
  if (rc == MB200_OK && epilogue_fused) *epilogue_fused = epi;
  return rc;
}

}  // namespace mb200
