// Shared helpers for libepn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/epn_b200.h"

#define EPN_API extern "C" __attribute__((visibility("default")))

namespace epn {

void set_error(const char *fmt, ...);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define EPN_REQUIRE_PTR(p)                                \
    do {                                                  \
        if ((p) == nullptr) {                             \
            epn::set_error("%s: %s is NULL", __func__, #p); \
            return EPN_ERR_NULL;                          \
        }                                                 \
    } while (0)

#define EPN_REQUIRE_POS(v)                                          \
    do {                                                            \
        if ((v) <= 0) {                                             \
            epn::set_error("%s: %s = %lld must be > 0", __func__, #v, (long long)(v)); \
            return EPN_ERR_SHAPE;                                   \
        }                                                           \
    } while (0)

#define EPN_REQUIRE(cond, code, msg)                      \
    do {                                                  \
        if (!(cond)) {                                    \
            epn::set_error("%s: %s", __func__, msg);      \
            return (code);                                \
        }                                                 \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// fp32 squared distance pinned to the reference kernels' SASS order
// (FMUL, FFMA, FFMA):  (dx*dx + dy*dy) + dz*dz -> fma(dz,dz, fma(dx,dx, dy*dy)).
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

}  // namespace epn
