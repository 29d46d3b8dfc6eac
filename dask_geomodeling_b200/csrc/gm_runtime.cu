// Runtime part of the C ABI: device selection, stream, stream-ordered memory
// pool, pinned host memory, copies.  No compute here.
#include "gm_common.cuh"
#include <mutex>

namespace gm {

static thread_local std::string g_error;
static std::mutex g_mutex;
static bool g_ready = false;
static int g_device = 0;
static int g_sms = 148;
static cudaStream_t g_stream = nullptr;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_error = msg; }
int fail(const std::string& msg) { g_error = msg; return 1; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int sm_count() { return g_sms; }

cudaStream_t resolve_stream(void* stream) {
  return stream ? reinterpret_cast<cudaStream_t>(stream) : g_stream;
}

static int do_init(int device) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g_ready && device == g_device) {
    GM_CUDA(cudaSetDevice(g_device));
    return 0;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(std::string("no CUDA device available: ") + cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail("gm_init: device index out of range");
  GM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail("libgeokernels is built for sm_100a (Blackwell); found compute capability " +
                std::to_string(prop.major) + "." + std::to_string(prop.minor));
  g_sms = prop.multiProcessorCount;
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  GM_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  GM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = UINT64_MAX;  // cache freed blocks instead of returning them to the driver
  GM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  g_device = device;
  g_ready = true;
  return 0;
}

int ensure_init() {
  if (g_ready) return cudaSetDevice(g_device) == cudaSuccess ? 0 : fail("cudaSetDevice failed");
  return do_init(0);
}

int upload(void** dev, const void* host, int64_t bytes, cudaStream_t s) {
  GM_CUDA(cudaMallocAsync(dev, bytes > 0 ? bytes : 16, s));
  if (bytes > 0) GM_CUDA(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, s));
  return 0;
}

int Staged::open_input(const GmArray& a, cudaStream_t s) {
  stream = s;
  bytes = array_bytes(a);
  if (a.space == GM_DEVICE) { dev = a.data; owned = false; return 0; }
  host = a.data;
  owned = true;
  GM_CUDA(cudaMallocAsync(&dev, bytes > 0 ? bytes : 16, s));
  if (bytes > 0) GM_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s));
  return 0;
}

int Staged::open_output(const GmArray& a, cudaStream_t s) {
  stream = s;
  bytes = array_bytes(a);
  if (a.space == GM_DEVICE) { dev = a.data; owned = false; return 0; }
  host = a.data;
  owned = true;
  GM_CUDA(cudaMallocAsync(&dev, bytes > 0 ? bytes : 16, s));
  return 0;
}

int Staged::finish_output() {
  if (owned && host && bytes > 0)
    GM_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream));
  return 0;
}

void Staged::release() {
  if (owned && dev) cudaFreeAsync(dev, stream);
  dev = nullptr;
  owned = false;
}

}  // namespace gm

using namespace gm;

extern "C" {

int gm_abi_version(void) { return GM_ABI_VERSION; }

int gm_init(int device) { return do_init(device); }

int gm_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g_ready) {
    cudaStreamSynchronize(g_stream);
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    g_ready = false;
  }
  return 0;
}

const char* gm_last_error(void) { return g_error.c_str(); }

int gm_device_info(int* sms, int64_t* total_mem, int* cc_major, int* cc_minor) {
  if (ensure_init()) return 1;
  cudaDeviceProp prop;
  GM_CUDA(cudaGetDeviceProperties(&prop, g_device));
  if (sms) *sms = prop.multiProcessorCount;
  if (total_mem) *total_mem = (int64_t)prop.totalGlobalMem;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

int64_t gm_launch_count(void) { return g_launches.load(); }

void* gm_default_stream(void) { return ensure_init() ? nullptr : (void*)g_stream; }

int gm_stream_sync(void* stream) {
  if (ensure_init()) return 1;
  GM_CUDA(cudaStreamSynchronize(resolve_stream(stream)));
  return 0;
}

int gm_malloc(void** ptr, int64_t bytes, void* stream) {
  if (ensure_init()) return 1;
  if (bytes <= 0) bytes = 16;
  GM_CUDA(cudaMallocAsync(ptr, (size_t)bytes, resolve_stream(stream)));
  return 0;
}

int gm_free(void* ptr, void* stream) {
  if (!ptr) return 0;
  if (ensure_init()) return 1;
  GM_CUDA(cudaFreeAsync(ptr, resolve_stream(stream)));
  return 0;
}

int gm_host_alloc(void** ptr, int64_t bytes) {
  if (ensure_init()) return 1;
  GM_CUDA(cudaHostAlloc(ptr, (size_t)(bytes > 0 ? bytes : 16), cudaHostAllocDefault));
  return 0;
}

int gm_host_free(void* ptr) {
  if (!ptr) return 0;
  GM_CUDA(cudaFreeHost(ptr));
  return 0;
}

int gm_host_register(void* ptr, int64_t bytes) {
  if (ensure_init()) return 1;
  GM_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return 0;
}

int gm_host_unregister(void* ptr) {
  GM_CUDA(cudaHostUnregister(ptr));
  return 0;
}

int gm_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream) {
  if (ensure_init()) return 1;
  if (bytes > 0)
    GM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, resolve_stream(stream)));
  return 0;
}

int gm_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream) {
  if (ensure_init()) return 1;
  cudaStream_t s = resolve_stream(stream);
  if (bytes > 0) GM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, s));
  GM_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int gm_memcpy_d2h_async(void* dst, const void* src, int64_t bytes, void* stream) {
  if (ensure_init()) return 1;
  if (bytes > 0)
    GM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, resolve_stream(stream)));
  return 0;
}

int gm_stream_create(void** stream) {
  if (ensure_init()) return 1;
  if (!stream) return fail("gm_stream_create: null argument");
  cudaStream_t s;
  GM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void*)s;
  return 0;
}

int gm_stream_destroy(void* stream) {
  if (!stream) return 0;
  GM_CUDA(cudaStreamDestroy((cudaStream_t)stream));
  return 0;
}

int gm_memcpy_d2d(void* dst, const void* src, int64_t bytes, void* stream) {
  if (ensure_init()) return 1;
  if (bytes > 0)
    GM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, resolve_stream(stream)));
  return 0;
}

int gm_memcpy2d_h2d(void* dst, int64_t dpitch, const void* src, int64_t spitch,
                    int64_t row_bytes, int64_t rows, void* stream) {
  if (ensure_init()) return 1;
  if (row_bytes > 0 && rows > 0)
    GM_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)row_bytes,
                              (size_t)rows, cudaMemcpyHostToDevice, resolve_stream(stream)));
  return 0;
}

}  // extern "C"
