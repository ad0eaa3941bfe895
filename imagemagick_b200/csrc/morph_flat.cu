// morph_flat.cu -- ErodeMorphology / DilateMorphology by run decomposition.
//
// Semantics are MorphologyPrimitive's (MagickCore/morphology.c:2980-3036): erode = min over the
// kernel cells >= 0.5 (kernel as is) starting from the centre value, dilate = max over the
// cells > 0.5 of the reflected kernel starting from 0.0; edge-clamped source; result is one of
// the input floats (bit exact), `changed` counts |result - centre| >= MagickEpsilon per channel.
//
// min/max are associative and idempotent, so the neighbourhood is factored: every kernel row is a
// union of horizontal runs [u0,u1]; distinct runs are few (Disk:3 -> 3, Square -> 1).  Per CTA tile
//   A. stage the source tile + halo in shared memory (float4, coalesced, edge-clamped),
//   B. for each distinct run d and each tile row: M_d[row][x] = op over the run   (shared across
//      all kernel rows and all outputs that use that run),
//   C. out[y][x] = op over kernel rows v of M_{d(v)}[y+v][x].
// Disk:3 costs 13+7 shared-memory reads and ~16 min/max per pixel-channel group instead of 29 and
// 29; the kernel is then limited by HBM traffic (16 B read + 16 B written per pixel), not by the ALU.
#include "mb200_internal.h"

#include <cuda_runtime.h>

#include <cmath>
#include <vector>

namespace mb200 {
namespace {

constexpr double kEpsilon = 1.0e-12;
constexpr int kMaxRuns = 8, kMaxEntries = 48;
constexpr int kTileW = 64, kTileH = 16, kThreads = 256;

struct FlatArgs {
  const float *src;
  float *dst;
  int width, height;
  int ox, oy, kw, kh;
  int nruns, nentries;
  short run_u0[kMaxRuns], run_u1[kMaxRuns];          // window columns of each distinct run
  unsigned char ent_v[kMaxEntries], ent_d[kMaxEntries];   // (kernel row, run id)
  int dilate;
  unsigned long long *changed;
};

template <int CH>
struct Px { float v[CH]; };

template <int CH>
__device__ __forceinline__ Px<CH> ld_shared(const float *p) {
  Px<CH> r;
  if (CH == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[CH - 1] = t.w; }
  else if (CH == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); r.v[0] = t.x; r.v[CH - 1] = t.y; }
  else {
#pragma unroll
    for (int c = 0; c < CH; ++c) r.v[c] = p[c];
  }
  return r;
}

template <int CH>
__device__ __forceinline__ void st_any(float *p, const Px<CH> &r) {
  if (CH == 4) *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[CH - 1]);
  else if (CH == 2) *reinterpret_cast<float2 *>(p) = make_float2(r.v[0], r.v[CH - 1]);
  else {
#pragma unroll
    for (int c = 0; c < CH; ++c) p[c] = r.v[c];
  }
}

// `if (p > pixel) pixel = p` / `if (p < pixel) pixel = p` (a NaN sample never replaces the value)
template <int CH, bool DILATE>
__device__ __forceinline__ void combine(Px<CH> &acc, const Px<CH> &p) {
#pragma unroll
  for (int c = 0; c < CH; ++c) acc.v[c] = DILATE ? fmaxf(acc.v[c], p.v[c]) : fminf(acc.v[c], p.v[c]);
}

template <int CH, bool DILATE>
__global__ void __launch_bounds__(kThreads) morph_flat_kernel(const FlatArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_ent_off[kMaxEntries];                     // float offset of (run d, row v) inside `runs`
  __shared__ int s_run_u0[kMaxRuns], s_run_len[kMaxRuns];
  const int tw = kTileW + a.kw - 1, th = kTileH + a.kh - 1;
  float *tile = smem;                                        // [th][tw][CH]
  float *runs = smem + static_cast<size_t>(th) * tw * CH;    // [nruns][th][kTileW][CH]
  const int bx = blockIdx.x * kTileW, by = blockIdx.y * kTileH;
  const int tid = threadIdx.x;
  const int wmax = a.width - 1, hmax = a.height - 1;
  if (tid < a.nentries) s_ent_off[tid] = ((a.ent_d[tid] * th + a.ent_v[tid]) * kTileW) * CH;
  if (tid < a.nruns) { s_run_u0[tid] = a.run_u0[tid]; s_run_len[tid] = a.run_u1[tid] - a.run_u0[tid]; }

  // A. stage: each thread walks one tile column segment (consecutive threads -> consecutive pixels)
  for (int ty = tid / 128; ty < th; ty += kThreads / 128) {
    const int sy = min(max(by - a.oy + ty, 0), hmax);
    const float *grow = a.src + static_cast<size_t>(sy) * a.width * CH;
    for (int tx = tid % 128; tx < tw; tx += 128) {
      const int sx = min(max(bx - a.ox + tx, 0), wmax);
      Px<CH> p;
      if (CH == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(grow) + sx); p.v[0] = t.x; p.v[1] = t.y; p.v[2] = t.z; p.v[CH - 1] = t.w; }
      else {
#pragma unroll
        for (int c = 0; c < CH; ++c) p.v[c] = __ldg(grow + static_cast<size_t>(sx) * CH + c);
      }
      st_any<CH>(tile + (static_cast<size_t>(ty) * tw + tx) * CH, p);
    }
  }
  __syncthreads();

  // B. horizontal run reductions (thread = fixed x, strided rows), shared by all users of the run
  const int x = tid % kTileW, rg = tid / kTileW;               // 4 row groups
  for (int d = 0; d < a.nruns; ++d) {
    const int u0 = s_run_u0[d], len = s_run_len[d];
    for (int ty = rg; ty < th; ty += kThreads / kTileW) {
      const float *p = tile + (static_cast<size_t>(ty) * tw + x + u0) * CH;
      Px<CH> acc = ld_shared<CH>(p);
      for (int u = 0; u < len; ++u) {
        p += CH;
        combine<CH, DILATE>(acc, ld_shared<CH>(p));
      }
      st_any<CH>(runs + ((static_cast<size_t>(d) * th + ty) * kTileW + x) * CH, acc);
    }
  }
  __syncthreads();

  // C. vertical combination
  unsigned nchanged = 0;
  const int gx = bx + x;
  if (gx < a.width) {
    for (int ly = rg; ly < kTileH; ly += kThreads / kTileW) {
      const int y = by + ly;
      if (y >= a.height) break;
      const Px<CH> centre = ld_shared<CH>(tile + (static_cast<size_t>(ly + a.oy) * tw + x + a.ox) * CH);
      Px<CH> acc;
#pragma unroll
      for (int c = 0; c < CH; ++c) acc.v[c] = DILATE ? 0.0f : centre.v[c];
      const float *rbase = runs + (static_cast<size_t>(ly) * kTileW + x) * CH;
      for (int e = 0; e < a.nentries; ++e) combine<CH, DILATE>(acc, ld_shared<CH>(rbase + s_ent_off[e]));
      st_any<CH>(a.dst + (static_cast<size_t>(y) * a.width + gx) * CH, acc);
#pragma unroll
      for (int c = 0; c < CH; ++c)
        nchanged += fabs(static_cast<double>(acc.v[c]) - static_cast<double>(centre.v[c])) >= kEpsilon;
    }
  }
  if (a.changed != nullptr && nchanged != 0) atomicAdd(a.changed, static_cast<unsigned long long>(nchanged));
}

template <int CH>
cudaError_t launch(const FlatArgs &a, dim3 grid, size_t smem, cudaStream_t s) {
  if (a.dilate) {
    cudaFuncSetAttribute(morph_flat_kernel<CH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    morph_flat_kernel<CH, true><<<grid, kThreads, smem, s>>>(a);
  } else {
    cudaFuncSetAttribute(morph_flat_kernel<CH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    morph_flat_kernel<CH, false><<<grid, kThreads, smem, s>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace

// Returns MB200_EUNSUPPORTED when the kernel does not decompose into few runs (caller then uses
// the generic cell-list kernel).  kernel_window_order: NaN / below-threshold cells are inactive.
int launch_morph_flat(const float *src, float *dst, size_t width, size_t height, int channels, int method,
                      const double *kernel_window_order, int kw, int kh, int ox, int oy,
                      unsigned long long *d_changed, void *stream) {
  if (method != MB200_ErodeMorphology && method != MB200_DilateMorphology) return MB200_EUNSUPPORTED;
  if (width > 0x3fffffffull || height > 0x3fffffffull || kw > 127 || kh > 127) return MB200_EUNSUPPORTED;
  FlatArgs a{};
  a.dilate = method == MB200_DilateMorphology;
  int nruns = 0, nent = 0;
  for (int v = 0; v < kh; ++v) {
    int u = 0;
    while (u < kw) {
      auto active = [&](int uu) {
        const double k = kernel_window_order[v * kw + uu];
        if (k != k) return false;
        return a.dilate ? (k > 0.5) : (k >= 0.5);
      };
      if (!active(u)) { ++u; continue; }
      int u1 = u;
      while (u1 + 1 < kw && active(u1 + 1)) ++u1;
      int d = -1;
      for (int i = 0; i < nruns; ++i)
        if (a.run_u0[i] == u && a.run_u1[i] == u1) d = i;
      if (d < 0) {
        if (nruns == kMaxRuns) return MB200_EUNSUPPORTED;
        d = nruns++;
        a.run_u0[d] = static_cast<short>(u);
        a.run_u1[d] = static_cast<short>(u1);
      }
      if (nent == kMaxEntries) return MB200_EUNSUPPORTED;
      a.ent_v[nent] = static_cast<unsigned char>(v);
      a.ent_d[nent] = static_cast<unsigned char>(d);
      ++nent;
      u = u1 + 1;
    }
  }
  if (nent == 0) return MB200_EUNSUPPORTED;
  a.src = src; a.dst = dst;
  a.width = static_cast<int>(width); a.height = static_cast<int>(height);
  a.ox = ox; a.oy = oy; a.kw = kw; a.kh = kh;
  a.nruns = nruns; a.nentries = nent;
  a.changed = d_changed;
  const int tw = kTileW + kw - 1, th = kTileH + kh - 1;
  const size_t smem = (static_cast<size_t>(th) * tw + static_cast<size_t>(nruns) * th * kTileW) * channels * sizeof(float);
  if (smem > 200 * 1024) return MB200_EUNSUPPORTED;
  dim3 grid((a.width + kTileW - 1) / kTileW, (a.height + kTileH - 1) / kTileH);
  if (grid.y > 65535) return MB200_EUNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (channels) {
    case 1: e = launch<1>(a, grid, smem, s); break;
    case 2: e = launch<2>(a, grid, smem, s); break;
    case 3: e = launch<3>(a, grid, smem, s); break;
    default: e = launch<4>(a, grid, smem, s); break;
  }
  count_launch();
  if (e != cudaSuccess) return cuda_fail(e, "morph_flat launch");
  return MB200_OK;
}

}  // namespace mb200
