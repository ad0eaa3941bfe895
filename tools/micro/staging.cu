// staging.cu -- how fast can a MagickCore pixel cache (ordinary host memory) reach HBM and come back?
// Measures, for a 1 GiB buffer (one 8192^2 RGBA float image):
//   pageable cudaMemcpy, pinned cudaMemcpy, cudaHostRegister cost, a threaded pinned-bounce pipeline,
//   managed memory (first touch, prefetch, GPU first touch, CPU fault-back, prefetch back),
//   and the stream-ordered allocator's steady-state cost for the temporaries of one blur + resize.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp staging.cu -o staging
#include <cuda_runtime.h>
#include <omp.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void scale_kernel(const float4 *in, float4 *out, size_t n) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 v = in[i];
    v.x *= 1.5f; v.y *= 1.5f; v.z *= 1.5f; v.w *= 1.5f;
    out[i] = v;
  }
}

static void par_memcpy(void *d, const void *s, size_t n, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static)
  for (long c = 0; c < (long) ((n + (1 << 20) - 1) >> 20); ++c) {
    size_t off = (size_t) c << 20, len = n - off < (1u << 20) ? n - off : (1u << 20);
    memcpy((char *) d + off, (const char *) s + off, len);
  }
}

int main(int argc, char **argv) {
  const size_t N = (size_t) 1 << 30;
  const int T = argc > 1 ? atoi(argv[1]) : 16;
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  cudaStream_t s, s2;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  void *d0, *d1;
  CK(cudaMalloc(&d0, N));
  CK(cudaMalloc(&d1, N));
  auto gbs = [&](double t) { return N / t / 1e9; };
  double t;

  // ---- pageable
  char *pg = (char *) aligned_alloc(4096, N);
  par_memcpy(pg, pg, 0, 1);
#pragma omp parallel for num_threads(T)
  for (long i = 0; i < (long) (N >> 12); ++i) pg[(size_t) i << 12] = (char) i;
  for (int r = 0; r < 2; ++r) {
    t = now(); CK(cudaMemcpy(d0, pg, N, cudaMemcpyHostToDevice)); t = now() - t;
    printf("pageable H2D cudaMemcpy            %8.2f ms  %6.1f GB/s\n", t * 1e3, gbs(t));
    t = now(); CK(cudaMemcpy(pg, d0, N, cudaMemcpyDeviceToHost)); t = now() - t;
    printf("pageable D2H cudaMemcpy            %8.2f ms  %6.1f GB/s\n", t * 1e3, gbs(t));
  }
  // ---- pinned
  char *pin;
  t = now(); CK(cudaMallocHost(&pin, N)); t = now() - t;
  printf("cudaMallocHost 1 GiB               %8.2f ms\n", t * 1e3);
  for (int r = 0; r < 2; ++r) {
    t = now(); CK(cudaMemcpyAsync(d0, pin, N, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); t = now() - t;
    printf("pinned H2D                         %8.2f ms  %6.1f GB/s\n", t * 1e3, gbs(t));
    t = now(); CK(cudaMemcpyAsync(pin, d0, N, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); t = now() - t;
    printf("pinned D2H                         %8.2f ms  %6.1f GB/s\n", t * 1e3, gbs(t));
  }
  // full duplex: H2D of one buffer while D2H of another
  {
    char *pin2; CK(cudaMallocHost(&pin2, N / 4));
    t = now();
    CK(cudaMemcpyAsync(d0, pin, N, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pin2, d1, N / 4, cudaMemcpyDeviceToHost, s2));
    CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(s2)); t = now() - t;
    printf("pinned H2D 1 GiB || D2H 0.25 GiB   %8.2f ms\n", t * 1e3);
    CK(cudaFreeHost(pin2));
  }
  t = now(); CK(cudaFreeHost(pin)); t = now() - t;
  printf("cudaFreeHost 1 GiB                 %8.2f ms\n", t * 1e3);
  // ---- register
  for (int r = 0; r < 2; ++r) {
    t = now(); CK(cudaHostRegister(pg, N, cudaHostRegisterDefault)); t = now() - t;
    printf("cudaHostRegister 1 GiB             %8.2f ms\n", t * 1e3);
    t = now(); CK(cudaMemcpyAsync(d0, pg, N, cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); t = now() - t;
    printf("registered H2D                     %8.2f ms  %6.1f GB/s\n", t * 1e3, gbs(t));
    t = now(); CK(cudaHostUnregister(pg)); t = now() - t;
    printf("cudaHostUnregister                 %8.2f ms\n", t * 1e3);
  }
  // chunked register + copy pipeline (register chunk k+1 on a helper thread while chunk k copies)
  {
    const size_t CH = (size_t) 64 << 20;
    t = now();
    std::thread reg([&] { for (size_t o = 0; o < N; o += CH) cudaHostRegister(pg + o, CH, cudaHostRegisterDefault); });
    reg.join();
    double t_reg = now() - t;
    for (size_t o = 0; o < N; o += CH) cudaHostUnregister(pg + o);
    printf("register in 64 MiB chunks (serial) %8.2f ms\n", t_reg * 1e3);
  }
  // ---- threaded bounce pipeline: pageable -> pinned ring (memcpy by `th` threads) -> cudaMemcpyAsync
  {
    const size_t CH = (size_t) 16 << 20;
    const int SLOTS = 4;
    char *ring; CK(cudaMallocHost(&ring, CH * SLOTS));
    cudaEvent_t ev[SLOTS];
    for (int i = 0; i < SLOTS; ++i) CK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    for (int th : {1, 2, 4, 8, 16, 32}) {
      t = now();
      size_t k = 0;
      for (size_t o = 0; o < N; o += CH, ++k) {
        const int sl = (int) (k % SLOTS);
        if (k >= SLOTS) CK(cudaEventSynchronize(ev[sl]));
        par_memcpy(ring + sl * CH, pg + o, CH, th);
        CK(cudaMemcpyAsync((char *) d0 + o, ring + sl * CH, CH, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(ev[sl], s));
      }
      CK(cudaStreamSynchronize(s)); t = now() - t;
      printf("bounce H2D, %2d copy threads        %8.2f ms  %6.1f GB/s\n", th, t * 1e3, gbs(t));
      // D2H
      t = now();
      k = 0;
      size_t done = 0;
      for (size_t o = 0; o < N + CH * SLOTS; o += CH, ++k) {
        const int sl = (int) (k % SLOTS);
        if (k >= SLOTS) {          // drain the chunk that used this slot
          CK(cudaEventSynchronize(ev[sl]));
          par_memcpy(pg + done, ring + sl * CH, CH, th);
          done += CH;
        }
        if (o < N) {
          CK(cudaMemcpyAsync(ring + sl * CH, (char *) d0 + o, CH, cudaMemcpyDeviceToHost, s));
          CK(cudaEventRecord(ev[sl], s));
        }
      }
      t = now() - t;
      printf("bounce D2H, %2d copy threads        %8.2f ms  %6.1f GB/s\n", th, t * 1e3, gbs(t));
    }
    t = now(); par_memcpy(pin = (char *) ring, pg, CH * SLOTS, 1); t = now() - t;
    printf("host memcpy 1 thread               %6.1f GB/s\n", CH * SLOTS / t / 1e9);
    CK(cudaFreeHost(ring));
  }
  // ---- managed memory
  {
    int conc = 0; CK(cudaDeviceGetAttribute(&conc, cudaDevAttrConcurrentManagedAccess, 0));
    int pageable = 0; CK(cudaDeviceGetAttribute(&pageable, cudaDevAttrPageableMemoryAccess, 0));
    printf("concurrentManagedAccess=%d pageableMemoryAccess=%d\n", conc, pageable);
    for (int r = 0; r < 2; ++r) {
      char *m, *mo;
      t = now(); CK(cudaMallocManaged(&m, N)); CK(cudaMallocManaged(&mo, N)); t = now() - t;
      printf("cudaMallocManaged 2 x 1 GiB        %8.2f ms\n", t * 1e3);
      t = now();
#pragma omp parallel for num_threads(T)
      for (long i = 0; i < (long) (N >> 20); ++i) memset(m + ((size_t) i << 20), 1, 1 << 20);
      t = now() - t;
      printf("managed CPU first touch (%d thr)   %8.2f ms  %6.1f GB/s\n", T, t * 1e3, gbs(t));
      cudaMemLocation loc; loc.type = cudaMemLocationTypeDevice; loc.id = 0;
      t = now(); CK(cudaMemPrefetchAsync(m, N, loc, 0, s)); CK(cudaStreamSynchronize(s)); t = now() - t;
      printf("managed prefetch -> device         %8.2f ms  %6.1f GB/s\n", t * 1e3, gbs(t));
      t = now(); CK(cudaMemPrefetchAsync(mo, N, loc, 0, s)); CK(cudaStreamSynchronize(s)); t = now() - t;
      printf("managed prefetch (unpopulated)     %8.2f ms\n", t * 1e3);
      for (int k = 0; k < 2; ++k) {
        t = now(); scale_kernel<<<148 * 8, 256, 0, s>>>((const float4 *) m, (float4 *) mo, N / 16); CK(cudaStreamSynchronize(s)); t = now() - t;
        printf("kernel managed->managed            %8.2f ms  %6.1f GB/s\n", t * 1e3, 2 * gbs(t));
      }
      for (int k = 0; k < 2; ++k) {
        t = now(); scale_kernel<<<148 * 8, 256, 0, s>>>((const float4 *) d0, (float4 *) d1, N / 16); CK(cudaStreamSynchronize(s)); t = now() - t;
        printf("kernel cudaMalloc->cudaMalloc      %8.2f ms  %6.1f GB/s\n", t * 1e3, 2 * gbs(t));
      }
      // CPU reads a quarter of the output through page faults
      t = now();
      double sum = 0;
#pragma omp parallel for num_threads(T) reduction(+ : sum)
      for (long i = 0; i < (long) (N / 4 >> 12); ++i) sum += mo[(size_t) i << 12];
      t = now() - t;
      printf("managed CPU fault-back 0.25 GiB    %8.2f ms  %6.1f GB/s (sum %g)\n", t * 1e3, N / 4 / t / 1e9, sum);
      cudaMemLocation host; host.type = cudaMemLocationTypeHost; host.id = 0;
      t = now(); CK(cudaMemPrefetchAsync(mo + N / 4, N / 4 * 3, host, 0, s)); CK(cudaStreamSynchronize(s)); t = now() - t;
      printf("managed prefetch -> host 0.75 GiB  %8.2f ms  %6.1f GB/s\n", t * 1e3, N / 4 * 3 / t / 1e9);
      // unpopulated output written by the GPU without a prefetch (GPU page faults)
      char *mf; CK(cudaMallocManaged(&mf, N / 4));
      t = now(); scale_kernel<<<148 * 8, 256, 0, s>>>((const float4 *) d0, (float4 *) mf, N / 64); CK(cudaStreamSynchronize(s)); t = now() - t;
      printf("kernel -> unpopulated managed 0.25 %8.2f ms\n", t * 1e3);
      t = now(); CK(cudaFree(mf)); CK(cudaFree(m)); CK(cudaFree(mo)); t = now() - t;
      printf("cudaFree managed                   %8.2f ms\n", t * 1e3);
    }
  }
  // ---- stream-ordered allocator, the temporaries of one blur (1 GiB) + resize (0.5 GiB) per step
  {
    cudaMemPool_t pool; CK(cudaDeviceGetDefaultMemPool(&pool, 0));
    unsigned long long thr = ~0ull; CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    for (int r = 0; r < 6; ++r) {
      void *a, *b;
      t = now();
      CK(cudaMallocAsync(&a, N, s)); CK(cudaFreeAsync(a, s));
      CK(cudaMallocAsync(&b, N / 2, s)); CK(cudaFreeAsync(b, s));
      double th = now() - t;
      CK(cudaStreamSynchronize(s)); t = now() - t;
      printf("mallocAsync 1G/free + 0.5G/free    host %8.3f ms, with sync %8.3f ms\n", th * 1e3, t * 1e3);
    }
  }
  return 0;
}
