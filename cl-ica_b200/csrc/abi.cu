// abi.cu -- library-wide pieces of the C ABI: version, thread-local error string, device query.
#include "common.cuh"

#include <string.h>

namespace clica {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int get_device_info(DeviceInfo* out) {
    static DeviceInfo cache[64];
    int dev = 0;
    CLICA_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(CLICA_E_BADARG, "device ordinal %d out of range", dev);
    if (!cache[dev].ok) {
        DeviceInfo di;
        CLICA_CUDA_OK(cudaDeviceGetAttribute(&di.sm_count, cudaDevAttrMultiProcessorCount, dev));
        CLICA_CUDA_OK(cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
        CLICA_CUDA_OK(cudaDeviceGetAttribute(&di.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
        if (di.cc_major != 10)
            return fail(CLICA_E_ARCH, "libclica_sm100 is built for sm_100a only; device %d is sm_%d%d", dev,
                        di.cc_major, di.cc_minor);
        di.ok = 1;
        cache[dev] = di;
    }
    *out = cache[dev];
    return 0;
}

}  // namespace clica

extern "C" int clica_abi_version(void) { return CLICA_ABI_VERSION; }

extern "C" const char* clica_last_error(void) { return clica::last_error_buffer(); }

extern "C" int clica_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    clica::DeviceInfo di;
    int rc = clica::get_device_info(&di);
    if (rc) return rc;
    if (sm_count) *sm_count = di.sm_count;
    if (cc_major) *cc_major = di.cc_major;
    if (cc_minor) *cc_minor = di.cc_minor;
    return 0;
}
