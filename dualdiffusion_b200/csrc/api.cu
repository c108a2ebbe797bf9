// Library-level entry points of the C ABI: error reporting and device queries.
#include "common.cuh"
#include "conv_dx.cuh"
#include "dualdiffusion_b200.h"

#include <stdarg.h>

namespace {
thread_local char g_error[1024] = "";
}

void dd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int dd_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

void* dd_tensormap_encode_fn() {
    static void* fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = sym;
    }
    return fn;
}

extern "C" const char* dd_last_error(void) { return g_error; }

extern "C" int dd_abi_version(void) { return DD_ABI_VERSION; }

extern "C" int dd_device_info(int* num_sms, int* cc_major, int* cc_minor) {
    int dev = 0;
    DD_CHECK_CUDA(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    DD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DD_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DD_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (num_sms) *num_sms = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return 0;
}
