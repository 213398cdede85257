// common.cuh -- error handling and small helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fruits_b200.h"

namespace fb {

// thread-local last error message returned by fb_last_error()
char *err_buf();
int set_err(int code, const char *fmt, ...);

#define FB_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess)                                                \
            return fb::set_err((int)_e, "%s failed: %s (%s:%d)", #expr,       \
                               cudaGetErrorString(_e), __FILE__, __LINE__);   \
    } while (0)

#define FB_REQUIRE(cond, ...)                                                 \
    do {                                                                      \
        if (!(cond)) return fb::set_err(FB_EINVAL, __VA_ARGS__);              \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double d_ninf() { return __longlong_as_double(0xfff0000000000000LL); }
__device__ __forceinline__ double d_max() { return __longlong_as_double(0x7fefffffffffffffLL); }

// np.nan_to_num(nan=0.0): nan -> 0, +-inf -> +-DBL_MAX (fruits/fruit.py:172)
__device__ __forceinline__ double nan_to_num(double v)
{
    if (v != v) return 0.0;
    if (v == d_inf()) return d_max();
    if (v == d_ninf()) return -d_max();
    return v;
}

}  // namespace fb
