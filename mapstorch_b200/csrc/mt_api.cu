// ABI bookkeeping: version, last-error text, the sm_100 device gate.
#include "mt_common.cuh"

namespace mt {

char *error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int require_sm100() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0) {
        cudaGetLastError();
        return fail(MT_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    }
    static int cached_major[64];
    static bool have[64];
    int major = 0;
    if (dev < 64 && have[dev]) {
        major = cached_major[dev];
    } else {
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(MT_ERR_NO_DEVICE, "cannot query device %d: %s", dev, cudaGetErrorString(e));
        }
        if (dev < 64) {
            cached_major[dev] = major;
            have[dev] = true;
        }
    }
    if (major != 10)
        return fail(MT_ERR_NO_DEVICE,
                    "device %d has compute capability %d.x; the kernels are built for sm_100a only", dev,
                    major);
    return MT_OK;
}

}  // namespace mt

extern "C" {

int mt_abi_version(void) { return MT_ABI_VERSION; }

const char *mt_last_error(void) { return mt::error_buffer(); }

int mt_device_check(void) { return mt::require_sm100(); }

}  // extern "C"
