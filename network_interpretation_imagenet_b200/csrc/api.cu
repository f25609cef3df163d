// api.cu — error plumbing and device probing for the C ABI (include/nib.h).
#include "common.cuh"
#include <string.h>
#include <map>
#include <mutex>
#include <utility>

namespace nib {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_dev_state = 1;  // 1 = unknown, 0 = ok, <0 = error
static int g_num_sms = 0;

int check_device() {
  if (g_dev_state <= 0) {
    if (g_dev_state < 0) set_error("libnib: no usable sm_100 CUDA device (no CPU fallback exists)");
    return g_dev_state;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("libnib: no CUDA device visible (%s); there is no CPU fallback",
              e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    g_dev_state = NIB_ENODEVICE;
    return g_dev_state;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess || p.major != 10) {
    set_error("libnib: device %d is sm_%d%d; this library is built for sm_100a only", dev, p.major,
              p.minor);
    g_dev_state = NIB_ENODEVICE;
    return g_dev_state;
  }
  g_num_sms = p.multiProcessorCount;
  g_dev_state = 0;
  return 0;
}

int num_sms() {
  if (g_num_sms == 0) check_device();
  return g_num_sms > 0 ? g_num_sms : 148;
}

// Library scratch is keyed by (use, device, stream): work submitted to one stream is ordered, so the next call on that
// stream may reuse the buffer, while calls on other streams (classifier.py feeds stream copies concurrently) get their
// own.  Growing a buffer goes through cudaFree, which waits for the kernels still reading the old one.
struct ScratchBuf { void* ptr; size_t cap; };
static std::map<std::pair<std::pair<int, int>, cudaStream_t>, ScratchBuf> g_scratch_bufs;
static std::mutex g_scratch_mu;

int stream_scratch(int slot, cudaStream_t st, size_t bytes, size_t min_bytes, void** out) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  ScratchBuf& b = g_scratch_bufs[{{slot, dev}, st}];
  if (b.cap < bytes) {
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
    const size_t want = bytes < min_bytes ? min_bytes : bytes;
    NIB_CUDA(cudaMalloc(&b.ptr, want));
    b.cap = want;
  }
  *out = b.ptr;
  return NIB_OK;
}

}  // namespace nib

extern "C" {

int nib_abi_version(void) { return NIB_ABI_VERSION; }

const char* nib_last_error(void) { return nib::g_err; }

int nib_device_info(int dev, int* sm_major, int* sm_minor, int* num_sms) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    nib::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return NIB_ENODEVICE;
  }
  if (dev >= 0 && dev < n) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
      if (sm_major) *sm_major = p.major;
      if (sm_minor) *sm_minor = p.minor;
      if (num_sms) *num_sms = p.multiProcessorCount;
    }
  }
  return n;
}

}  // extern "C"
