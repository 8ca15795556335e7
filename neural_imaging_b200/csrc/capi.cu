// C-ABI plumbing shared by all kernels: error string, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

#include "ni_common.cuh"

static thread_local char g_ni_err[512] = "";
unsigned long long g_ni_launches = 0;

void ni_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ni_err, sizeof(g_ni_err), fmt, ap);
    va_end(ap);
}

int ni_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

extern "C" const char* ni_last_error(void) { return g_ni_err; }
extern "C" unsigned long long ni_launch_count(void) { return g_ni_launches; }
extern "C" void ni_reset_launch_count(void) { g_ni_launches = 0; }
extern "C" int ni_version(void) { return 100; }

// Returns the compute capability (major*10+minor) of the current device, or a negative error.
extern "C" int ni_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    NI_CUDA(cudaGetDevice(&dev));
    NI_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    NI_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}
