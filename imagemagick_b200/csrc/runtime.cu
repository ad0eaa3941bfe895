// runtime.cu -- device selection, error reporting, staging buffers.
//
// The CUDA analogue of the reference's OpenCL environment plumbing
// (MagickCore/opencl.c: GetCurrentOpenCLEnv, RequestOpenCLDevice :2578, AcquireMagickCLCacheInfo
// :528) reduced to what the hot path needs: one stream and a few grow-only scratch
// buffers per device, guarded by a mutex (cf. openCL_lock, opencl.c:538).
#include "mb200_internal.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace mb200 {

namespace {
thread_local char t_error[512] = "";
std::atomic<unsigned long long> g_launches{0};

constexpr int kMaxDevices = 16;
constexpr int kScratchSlots = 8;

struct DeviceStateImpl {
  bool ready = false;
  cudaStream_t stream = nullptr;
  int sms = 0;
  void *scratch[kScratchSlots] = {nullptr};
  size_t scratch_bytes[kScratchSlots] = {0};
};
DeviceStateImpl g_dev[kMaxDevices];
std::mutex g_mutex;
}  // namespace

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(int e, const char *what) {
  const cudaError_t err = static_cast<cudaError_t>(e);
  if (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver)
    return fail(MB200_ENODEVICE, "%s: %s", what, cudaGetErrorString(err));
  if (err == cudaErrorMemoryAllocation) return fail(MB200_ENOMEM, "%s: %s", what, cudaGetErrorString(err));
  return fail(MB200_ECUDA, "%s: %s", what, cudaGetErrorString(err));
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ensure_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, "cudaGetDevice"); }
  if (dev < 0 || dev >= kMaxDevices) return fail(MB200_ENODEVICE, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_mutex);
  DeviceStateImpl &d = g_dev[dev];
  if (d.ready) return MB200_OK;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(MB200_ENODEVICE, "device %d is sm_%d%d; libmagickb200 carries sm_100a code only", dev,
                prop.major, prop.minor);
  d.sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
  d.ready = true;
  return MB200_OK;
}

static DeviceStateImpl *current() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  return &g_dev[dev];
}

void *default_stream() {
  DeviceStateImpl *d = current();
  return d ? d->stream : nullptr;
}

int sm_count() {
  DeviceStateImpl *d = current();
  return d && d->sms ? d->sms : 148;
}

int scratch(void **ptr, size_t bytes, int slot) {
  DeviceStateImpl *d = current();
  if (!d || slot < 0 || slot >= kScratchSlots) return fail(MB200_EINVAL, "bad scratch slot");
  std::lock_guard<std::mutex> lock(g_mutex);
  if (d->scratch_bytes[slot] < bytes) {
    if (d->scratch[slot]) {
      cudaStreamSynchronize(d->stream);
      cudaFree(d->scratch[slot]);
      d->scratch[slot] = nullptr;
      d->scratch_bytes[slot] = 0;
    }
    cudaError_t e = cudaMalloc(&d->scratch[slot], bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
    d->scratch_bytes[slot] = bytes;
  }
  *ptr = d->scratch[slot];
  return MB200_OK;
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int usable = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++usable;
  }
  return usable;
}

int mb200_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, "cudaSetDevice"); }
  return ensure_device();
}

const char *mb200_last_error(void) { return t_error; }
const char *mb200_version(void) { return "magick-b200 0.1 (sm_100a; ImageMagick 7.1.1-45 Q16-HDRI semantics)"; }
unsigned long long mb200_launch_count(void) { return g_launches.load(); }

int mb200_synchronize(void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : static_cast<cudaStream_t>(default_stream());
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
  return MB200_OK;
}

int mb200_malloc(void **dev_ptr, size_t bytes) {
  int rc = ensure_device();
  if (rc) return rc;
  cudaError_t e = cudaMalloc(dev_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  return MB200_OK;
}

int mb200_free(void *dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFree");
  return MB200_OK;
}

int mb200_malloc_host(void **host_ptr, size_t bytes) {
  int rc = ensure_device();
  if (rc) return rc;
  cudaError_t e = cudaMallocHost(host_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
  return MB200_OK;
}

int mb200_free_host(void *host_ptr) {
  cudaError_t e = cudaFreeHost(host_ptr);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFreeHost");
  return MB200_OK;
}

int mb200_upload(void *dev_dst, const void *host_src, size_t bytes, void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : static_cast<cudaStream_t>(default_stream());
  cudaError_t e = cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(H2D)");
  return MB200_OK;
}

int mb200_download(void *host_dst, const void *dev_src, size_t bytes, void *stream) {
  int rc = ensure_device();
  if (rc) return rc;
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : static_cast<cudaStream_t>(default_stream());
  cudaError_t e = cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(D2H)");
  return MB200_OK;
}

}  // extern "C"
