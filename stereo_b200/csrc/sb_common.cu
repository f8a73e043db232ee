// sb_common.cu -- library-level entry points (include/stereo_b200.h, "library").
#include "sb_common.h"
#include <vector>

namespace sb {

static thread_local std::string t_last_error;
std::atomic<int64_t> g_launches{0};

void set_last_error(const std::string &m) { t_last_error = m; }

std::string format(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    char buf[1024];
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}

void require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        throw Error(SB_ENODEV, format("stereo_b200: no CUDA device available (%s); this library has no CPU fallback",
                                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    }
    // keep freed blocks in the default memory pool of the current device (see DevBuf)
    static thread_local int configured_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != configured_dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured_dev = dev;
    }
}

} // namespace sb

extern "C" {

const char *sb_version(void) { return "stereo_b200 0.1 (sm_100a)"; }

const char *sb_last_error(void) { return sb::t_last_error.c_str(); }

int sb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sb_set_device(int device)
{
    return sb::guarded([&] {
        sb::require_device();
        SB_CUDA(cudaSetDevice(device));
    });
}

int64_t sb_kernel_launches(void) { return sb::g_launches.load(); }

} // extern "C"
