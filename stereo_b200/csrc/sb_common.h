// sb_common.h -- shared host-side plumbing of libstereo_b200: error reporting
// through the C ABI (the reference reports through mexErrMsgTxt / ASSERT,
// cpp/utils/cppmatrix.h:20-24), CUDA error checks, RAII device buffers and the
// launch counter behind sb_kernel_launches().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string>
#include <stdexcept>
#include <atomic>
#include "../../include/stereo_b200.h"

namespace sb {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);
std::string format(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

inline void count_launch(int64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SB_REQUIRE(cond, code, ...)                                              \
    do {                                                                         \
        if (!(cond)) throw ::sb::Error((code), ::sb::format(__VA_ARGS__));       \
    } while (0)

#define SB_CUDA(expr)                                                            \
    do {                                                                         \
        cudaError_t _e = (expr);                                                 \
        if (_e != cudaSuccess) {                                                 \
            int _c = (_e == cudaErrorMemoryAllocation) ? SB_ENOMEM               \
                   : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? SB_ENODEV \
                   : SB_ECUDA;                                                   \
            throw ::sb::Error(_c, ::sb::format("%s failed: %s (%s:%d)", #expr,   \
                              cudaGetErrorString(_e), __FILE__, __LINE__));      \
        }                                                                        \
    } while (0)

// Fails loudly when no device is present: there is no CPU fallback.
void require_device();

// SB_PLAIN_MALLOC=1: cudaMalloc / cudaFree instead of the stream-ordered pool (GPU core dumps and
// some debuggers do not support cudaMallocAsync).
inline bool plain_malloc()
{
    static const bool on = [] { const char *e = getenv("SB_PLAIN_MALLOC"); return e && atoi(e) != 0; }();
    return on;
}

// Device buffer (cudaMalloc / cudaFree).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    // Stream-ordered allocation from the device's default pool (configured in require_device()
    // to keep freed blocks): repeated solves reuse their buffers instead of paying cudaMalloc /
    // cudaFree (which synchronises) every call.  Everything the library does runs on stream 0.
    void alloc(size_t count)
    {
        release();
        n = count;
        if (!count) return;
        if (plain_malloc()) SB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        else SB_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), 0));
    }
    void release()
    {
        if (p) {
            if (plain_malloc()) cudaFree(p);
            else cudaFreeAsync(p, 0);
        }
        p = nullptr;
        n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Wrap a C-ABI entry point body: exceptions -> error code + sb_last_error().
template <typename F>
int guarded(F &&f)
{
    try {
        f();
        return SB_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return SB_ECUDA;
    }
}

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t stream) : s(stream)
    {
        SB_CUDA(cudaEventCreate(&a));
        SB_CUDA(cudaEventCreate(&b));
    }
    ~EventTimer()
    {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
    void start() { SB_CUDA(cudaEventRecord(a, s)); }
    double stop_ms()
    {
        SB_CUDA(cudaEventRecord(b, s));
        SB_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        SB_CUDA(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
};

} // namespace sb
