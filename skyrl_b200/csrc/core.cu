// Library-wide state of libskyb200: thread-local error text, device capability check, SM count.
#include <mutex>

#include "common.cuh"

namespace skyb200 {

char* last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace skyb200

extern "C" int skyb200_abi_version(void) { return 2; }

extern "C" const char* skyb200_last_error(void) { return skyb200::last_error_buffer(); }

extern "C" int skyb200_check_device(void) {
    static int cached[64] = {0};  // 0 = unknown, 1 = ok, 2 = wrong arch
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return skyb200::fail(SKYB200_ECUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
    if (dev >= 0 && dev < 64 && cached[dev] == 1) return SKYB200_OK;
    int major = 0, minor = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess)
        return skyb200::fail(SKYB200_ECUDA, "cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    if (major != 10)
        return skyb200::fail(SKYB200_EARCH, "libskyb200 is built for sm_100a only; device %d is sm_%d%d", dev, major,
                             minor);
    if (dev >= 0 && dev < 64) cached[dev] = 1;
    return SKYB200_OK;
}
