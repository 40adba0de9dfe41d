// C-ABI plumbing shared by every entry point: error string, version, device gate.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void vsn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int vsn_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// Programmatic dependent launch for the kernels that support it (VSN_PDL=0 turns it off: measurements / debugging)
bool vsn_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VSN_PDL"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

// Number of kernel launches issued by this library since load (all threads); bench.py reports the
// difference over its timed region as "gpu_launches".
static unsigned long long g_launches = 0;
void vsn_count_launch() { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }
extern "C" long long vsn_launch_count() { return static_cast<long long>(__atomic_load_n(&g_launches, __ATOMIC_RELAXED)); }

extern "C" const char* vsn_last_error() { return g_err; }

extern "C" int vsn_version() { return 100; }

// 16-bit operand encoding this build computes in: 0 = bfloat16 (libvsn_b200.so), 1 = IEEE half (libvsn_b200_f16.so,
// the same sources compiled with -DVSN_F16).  The binding checks it against the dtype it allocates.
extern "C" int vsn_precision() {
#ifdef VSN_F16
  return 1;
#else
  return 0;
#endif
}

// 0 when the current device can run this library (compute capability 10.x); the Python side
// refuses to continue otherwise -- there is no fallback path.
extern "C" int vsn_check_device() {
  int dev = 0;
  VSN_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  VSN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  VSN_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  VSN_CHECK(major == 10, "vsn_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
  return 0;
}
