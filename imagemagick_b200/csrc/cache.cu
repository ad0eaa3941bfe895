// cache.cu -- the pixel cache "staged into HBM": host <-> HBM transfer engine and the residency
// registry behind the host-buffer operators.
//
// What the reference has for its OpenCL path and what this file is the CUDA counterpart of:
//   AcquireMagickCLCacheInfo (MagickCore/opencl.c:528-553)   mb200_cache_attach
//   GetAuthenticOpenCLBuffer (MagickCore/cache.c:1259-1292)  stage_input / stage_output (HBM copy of a pixel cache)
//   CopyOpenCLBuffer (cache.c:5341-5364, called at :1710, :2771, :4079)   mb200_cache_sync (lazy device -> host)
//   RelinquishMagickCLCacheInfo (cache.c:978-984)           mb200_cache_detach
//
// Transfers.  A MagickCore pixel cache is ordinary (pageable) host memory unless the shim's allocator
// hook put it into the pinned pool.  Measured on the B200 host (tools/micro/staging.cu, profiles/r02_staging.md):
// cudaMemcpy from pageable memory 9.4 GB/s up / 19 GB/s down, pinned 55.5 / 57.2 GB/s, cudaHostRegister of 1 GiB
// 67-450 ms.  So: pinned / registered / managed pointers are copied directly; anything else goes through a
// 4 x 16 MiB pinned ring filled (drained) by a small pool of memcpy threads while the previous chunk is on the
// wire: 49 GB/s up, 44 GB/s down with 16 threads.
#include "mb200_internal.h"

#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace mb200 {

namespace {

// ------------------------------------------------------------------ memcpy worker pool
class CopyPool {
 public:
  static CopyPool &get() {
    static CopyPool *pool = new CopyPool();      // intentionally leaked: workers may outlive static destruction
    return *pool;
  }
  // parallel memcpy; returns when every byte is copied.  One job at a time (callers hold the ring lock).
  void copy(void *dst, const void *src, size_t bytes) {
    const int n = workers_ + 1;
    if (n == 1 || bytes < (1u << 20)) { std::memcpy(dst, src, bytes); return; }
    const size_t slice = ((bytes + n - 1) / n + 4095) & ~static_cast<size_t>(4095);
    {
      std::lock_guard<std::mutex> lock(m_);
      dst_ = static_cast<char *>(dst); src_ = static_cast<const char *>(src); bytes_ = bytes; slice_ = slice;
      pending_ = workers_;
      ++generation_;
    }
    cv_.notify_all();
    run_slice(0);
    std::unique_lock<std::mutex> lock(m_);
    done_.wait(lock, [&] { return pending_ == 0; });
  }
  int threads() const { return workers_ + 1; }

 private:
  CopyPool() {
    int want = 16;
    if (const char *e = std::getenv("MB200_COPY_THREADS")) want = std::atoi(e);
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (hw > 0 && want > hw) want = hw;
    if (want < 1) want = 1;
    workers_ = want - 1;
    for (int i = 0; i < workers_; ++i) std::thread([this, i] { loop(i + 1); }).detach();
  }
  void run_slice(int idx) {
    const size_t off = slice_ * static_cast<size_t>(idx);
    if (off < bytes_) std::memcpy(dst_ + off, src_ + off, std::min(slice_, bytes_ - off));
  }
  void loop(int idx) {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lock(m_);
        cv_.wait(lock, [&] { return generation_ != seen; });
        seen = generation_;
      }
      run_slice(idx);
      std::lock_guard<std::mutex> lock(m_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_;
  int workers_ = 0, pending_ = 0;
  unsigned long long generation_ = 0;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
};

// ------------------------------------------------------------------ pinned bounce ring (per device)
constexpr int kSlots = 4;
constexpr size_t kChunk = static_cast<size_t>(16) << 20;
constexpr int kMaxDevices = 16;

struct Ring {
  std::mutex m;
  char *slot[kSlots] = {nullptr};
  cudaEvent_t ev[kSlots] = {nullptr};
  bool busy[kSlots] = {false};
  bool ready = false;
};
Ring g_ring[kMaxDevices];

int ring_for_current_device(Ring **out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fail(MB200_ENODEVICE, "bounce ring: no device");
  Ring &r = g_ring[dev];
  if (!r.ready) {
    std::lock_guard<std::mutex> lock(r.m);
    if (!r.ready) {
      for (int i = 0; i < kSlots; ++i) {
        e = cudaMallocHost(&r.slot[i], kChunk);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return cuda_fail(e, "bounce ring allocation");
      }
      r.ready = true;
    }
  }
  *out = &r;
  return MB200_OK;
}

bool directly_copyable(const void *host) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged || attr.type == cudaMemoryTypeDevice;
}

std::atomic<unsigned long long> g_uploads{0}, g_downloads{0}, g_upload_bytes{0}, g_download_bytes{0}, g_hits{0},
    g_bounced{0};

// ------------------------------------------------------------------ residency registry
struct Entry {
  void *host = nullptr;
  size_t bytes = 0;
  int device = 0;
  void *dev = nullptr;          // HBM copy (allocated on first use)
  bool host_valid = true;       // the host buffer holds the current pixels
  bool dev_valid = false;       // the HBM copy holds the current pixels
  bool registered = false;      // pinned by us with cudaHostRegister
  cudaEvent_t ready = nullptr;  // last use of the HBM copy (any stream): the next user's stream waits for it
};
std::mutex g_cache_mutex;
std::unordered_map<const void *, Entry *> g_entries;
std::atomic<int> g_lazy{0};
// entries whose HBM copy is current / whose host copy is stale: lets the cache hooks return without a lookup
std::atomic<long> g_dev_current{0}, g_host_stale{0};

void set_state(Entry *e, bool host_valid, bool dev_valid) {
  if (e->dev_valid != dev_valid) g_dev_current.fetch_add(dev_valid ? 1 : -1, std::memory_order_relaxed);
  if (e->host_valid != host_valid) g_host_stale.fetch_add(host_valid ? -1 : 1, std::memory_order_relaxed);
  e->host_valid = host_valid;
  e->dev_valid = dev_valid;
}

Entry *lookup(const void *host, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  auto it = g_entries.find(host);
  if (it == g_entries.end() || it->second->bytes != bytes) return nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  return it->second->device == dev ? it->second : nullptr;
}

int ensure_device_copy(Entry *e) {
  if (e->dev) return MB200_OK;
  cudaError_t err = cudaMalloc(&e->dev, e->bytes ? e->bytes : 1);
  if (err != cudaSuccess) { e->dev = nullptr; return cuda_fail(err, "pixel cache: cudaMalloc"); }
  if (!e->ready) cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming);
  return MB200_OK;
}

// Calls come in on per-thread streams: order this stream behind the last use of the entry's HBM copy.
void order_after_last_use(Entry *e, void *stream) {
  if (e->ready) cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), e->ready, 0);
}

}  // namespace

// ------------------------------------------------------------------ transfer engine (internal API)
int copy_h2d(void *dev, const void *host, size_t bytes, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  g_uploads.fetch_add(1, std::memory_order_relaxed);
  g_upload_bytes.fetch_add(bytes, std::memory_order_relaxed);
  if (bytes < (4u << 20) || directly_copyable(host)) {
    cudaError_t e = cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s);
    return e == cudaSuccess ? MB200_OK : cuda_fail(e, "H2D");
  }
  Ring *r;
  int rc = ring_for_current_device(&r);
  if (rc) return rc;
  g_bounced.fetch_add(bytes, std::memory_order_relaxed);
  CopyPool &pool = CopyPool::get();
  std::lock_guard<std::mutex> lock(r->m);
  size_t k = 0;
  for (size_t o = 0; o < bytes; o += kChunk, ++k) {
    const int sl = static_cast<int>(k % kSlots);
    const size_t len = std::min(kChunk, bytes - o);
    if (r->busy[sl]) { cudaError_t e = cudaEventSynchronize(r->ev[sl]); if (e != cudaSuccess) return cuda_fail(e, "H2D ring"); }
    pool.copy(r->slot[sl], static_cast<const char *>(host) + o, len);
    cudaError_t e = cudaMemcpyAsync(static_cast<char *>(dev) + o, r->slot[sl], len, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaEventRecord(r->ev[sl], s);
    if (e != cudaSuccess) return cuda_fail(e, "H2D ring");
    r->busy[sl] = true;
  }
  return MB200_OK;      // the host buffer has been read completely; the last chunks are still on the wire
}

// Device -> host.  Always complete on return (the host buffer holds the data).
int copy_d2h(void *host, const void *dev, size_t bytes, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  g_downloads.fetch_add(1, std::memory_order_relaxed);
  g_download_bytes.fetch_add(bytes, std::memory_order_relaxed);
  if (bytes < (4u << 20) || directly_copyable(host)) {
    cudaError_t e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    return e == cudaSuccess ? MB200_OK : cuda_fail(e, "D2H");
  }
  Ring *r;
  int rc = ring_for_current_device(&r);
  if (rc) return rc;
  g_bounced.fetch_add(bytes, std::memory_order_relaxed);
  CopyPool &pool = CopyPool::get();
  std::lock_guard<std::mutex> lock(r->m);
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  for (size_t k = 0; k < nchunks + kSlots; ++k) {
    const int sl = static_cast<int>(k % kSlots);
    if (k >= kSlots) {                                   // drain the chunk that used this slot
      const size_t o = (k - kSlots) * kChunk;
      cudaError_t e = cudaEventSynchronize(r->ev[sl]);
      if (e != cudaSuccess) return cuda_fail(e, "D2H ring");
      pool.copy(static_cast<char *>(host) + o, r->slot[sl], std::min(kChunk, bytes - o));
      r->busy[sl] = false;
    } else if (r->busy[sl]) {                            // slot still carries an earlier upload
      cudaError_t e = cudaEventSynchronize(r->ev[sl]);
      if (e != cudaSuccess) return cuda_fail(e, "D2H ring");
      r->busy[sl] = false;
    }
    if (k < nchunks) {
      const size_t o = k * kChunk;
      cudaError_t e = cudaMemcpyAsync(r->slot[sl], static_cast<const char *>(dev) + o, std::min(kChunk, bytes - o),
                                      cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaEventRecord(r->ev[sl], s);
      if (e != cudaSuccess) return cuda_fail(e, "D2H ring");
    }
  }
  return MB200_OK;
}

// ------------------------------------------------------------------ staging of operator arguments (internal API)
int stage_input(const void *host, size_t bytes, void *stream, StageRef *out) {
  out->entry = nullptr; out->dev = nullptr; out->temporary = false; out->host = const_cast<void *>(host); out->bytes = bytes;
  Entry *e = lookup(host, bytes);
  if (e) {
    int rc = ensure_device_copy(e);
    if (rc) return rc;
    out->entry = e;
    out->dev = e->dev;
    order_after_last_use(e, stream);
    // The HBM copy may be used without a fresh upload when it is the only current copy, or when the caller vouches
    // for host writes (lazy / hook mode: every host access goes through mb200_cache_sync / _host_written).
    if (e->dev_valid && (!e->host_valid || g_lazy.load(std::memory_order_relaxed))) {
      g_hits.fetch_add(1, std::memory_order_relaxed);
      return MB200_OK;
    }
    rc = copy_h2d(e->dev, host, bytes, stream);
    if (rc) return rc;
    set_state(e, e->host_valid, true);
    return MB200_OK;
  }
  cudaError_t err = cudaMallocAsync(&out->dev, bytes ? bytes : 1, temp_pool(), static_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) { out->dev = nullptr; return cuda_fail(err, "staging allocation"); }
  out->temporary = true;
  return copy_h2d(out->dev, host, bytes, stream);
}

int stage_output(void *host, size_t bytes, void *stream, StageRef *out) {
  out->entry = nullptr; out->dev = nullptr; out->temporary = false; out->host = host; out->bytes = bytes;
  Entry *e = lookup(host, bytes);
  if (e) {
    int rc = ensure_device_copy(e);
    if (rc) return rc;
    out->entry = e;
    out->dev = e->dev;
    order_after_last_use(e, stream);
    return MB200_OK;
  }
  cudaError_t err = cudaMallocAsync(&out->dev, bytes ? bytes : 1, temp_pool(), static_cast<cudaStream_t>(stream));
  if (err != cudaSuccess) { out->dev = nullptr; return cuda_fail(err, "staging allocation"); }
  out->temporary = true;
  return MB200_OK;
}

// The operator has written ref->dev.  Eager mode: copy to the host buffer now.  Lazy mode (attached buffers only): the
// HBM copy becomes the current one and the host copy is refreshed by mb200_cache_sync.
int finish_output(StageRef *ref, void *stream) {
  Entry *e = static_cast<Entry *>(ref->entry);
  if (e) {
    set_state(e, false, true);
    if (g_lazy.load(std::memory_order_relaxed)) return MB200_OK;
    const int rc = copy_d2h(ref->host, ref->dev, ref->bytes, stream);
    if (rc == MB200_OK) set_state(e, true, true);
    return rc;
  }
  return copy_d2h(ref->host, ref->dev, ref->bytes, stream);
}

void release_stage(StageRef *ref, void *stream) {
  Entry *e = static_cast<Entry *>(ref->entry);
  if (e && e->ready) cudaEventRecord(e->ready, static_cast<cudaStream_t>(stream));
  if (ref->temporary && ref->dev) cudaFreeAsync(ref->dev, static_cast<cudaStream_t>(stream));
  ref->dev = nullptr;
  ref->temporary = false;
}

}  // namespace mb200

using namespace mb200;

extern "C" {

int mb200_cache_attach(void *host_pixels, size_t bytes, int flags) {
  if (!host_pixels || bytes == 0) return fail(MB200_EINVAL, "cache_attach: bad arguments");
  int rc = ensure_device();
  if (rc) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  Entry *e = new Entry();
  e->host = host_pixels; e->bytes = bytes; e->device = dev;
  if ((flags & MB200_CACHE_REGISTER) && !directly_copyable(host_pixels)) {
    if (cudaHostRegister(host_pixels, bytes, cudaHostRegisterDefault) == cudaSuccess) e->registered = true;
    else cudaGetLastError();          // stays pageable: transfers take the bounce ring
  }
  Entry *old = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    auto it = g_entries.find(host_pixels);
    if (it != g_entries.end()) { old = it->second; it->second = e; }
    else g_entries.emplace(host_pixels, e);
  }
  if (old) {                       // a recycled address: the previous image is gone
    set_state(old, true, false);
    if (old->registered && !e->registered) cudaHostUnregister(old->host);
    if (old->dev) cudaFree(old->dev);
    if (old->ready) cudaEventDestroy(old->ready);
    delete old;
  }
  return MB200_OK;
}

int mb200_cache_detach(void *host_pixels) {
  Entry *e = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    auto it = g_entries.find(host_pixels);
    if (it == g_entries.end()) return MB200_OK;
    e = it->second;
    g_entries.erase(it);
  }
  set_state(e, true, false);
  if (e->dev) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != e->device) cudaSetDevice(e->device);
    cudaFree(e->dev);                 // synchronises with work that still reads / writes it
    if (cur != e->device) cudaSetDevice(cur);
  }
  if (e->registered) cudaHostUnregister(e->host);
  if (e->ready) cudaEventDestroy(e->ready);
  delete e;
  return MB200_OK;
}

int mb200_cache_sync(void *host_pixels) {
  if (g_host_stale.load(std::memory_order_relaxed) == 0) return MB200_OK;      // nothing lives in HBM only
  Entry *e = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    auto it = g_entries.find(host_pixels);
    if (it == g_entries.end()) return MB200_OK;
    e = it->second;
  }
  if (e->host_valid || !e->dev_valid) return MB200_OK;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != e->device) cudaSetDevice(e->device);
  int rc = ensure_device();
  if (rc == MB200_OK) {
    order_after_last_use(e, default_stream());
    rc = copy_d2h(e->host, e->dev, e->bytes, default_stream());
  }
  if (cur != e->device) cudaSetDevice(cur);
  if (rc == MB200_OK) set_state(e, true, e->dev_valid);
  return rc;
}

int mb200_cache_host_written(void *host_pixels) {
  if (g_dev_current.load(std::memory_order_relaxed) == 0) return MB200_OK;    // no HBM copy to invalidate
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  auto it = g_entries.find(host_pixels);
  if (it == g_entries.end()) return MB200_OK;
  set_state(it->second, true, false);
  return MB200_OK;
}

int mb200_cache_resident(const void *host_pixels) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  auto it = g_entries.find(host_pixels);
  if (it == g_entries.end()) return -1;
  return (it->second->dev_valid ? 1 : 0) | (it->second->host_valid ? 2 : 0);
}

int mb200_cache_set_lazy(int on) {
  return g_lazy.exchange(on ? 1 : 0);
}

void mb200_cache_stats(unsigned long long out[6]) {
  out[0] = g_uploads.load(); out[1] = g_upload_bytes.load(); out[2] = g_downloads.load();
  out[3] = g_download_bytes.load(); out[4] = g_hits.load(); out[5] = g_bounced.load();
}

int mb200_copy_threads(void) { return CopyPool::get().threads(); }

}  // extern "C"
