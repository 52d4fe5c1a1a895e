// Shared helpers for libepn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/epn_b200.h"

#define EPN_API extern "C" __attribute__((visibility("default")))

namespace epn {

void set_error(const char *fmt, ...);
void count_launch();

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return 0;
}

// Kernel classes for the optional per-kernel CUDA-event timing (epn_profile_*).
enum KernelClass {
    KC_INDEX = 0,        // ball query, FPS, gather
    KC_INTER_GROUP = 1,  // inter grouping fwd (gather + kernel weights + spatial contraction)
    KC_INTER_SCATTER = 2,// inter grouping bwd (transposed contraction + RED scatter)
    KC_INTRA_GROUP = 3,  // intra grouping fwd/bwd
    KC_GEMM = 4,         // channel GEMMs (fwd, dX, dW)
    KC_OTHER = 5,
    KC_SPLIT = 5,        // fp32 -> bf16 hi/lo split-tile conversion passes
    KC_NORM = 6,         // fused normalisation + leaky_relu
    KC_INTER_FUSED = 7,  // fused inter conv forward (gather + contraction + channel GEMM)
    KC_COUNT = 8
};

// RAII: brackets the launches issued in its scope with a cudaEvent pair when profiling is on.
struct ProfScope {
    cudaStream_t s;
    void *slot;
    ProfScope(cudaStream_t stream, int kclass);
    ~ProfScope();
};

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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: a process that drives several GPUs
// (model.to('cuda:1'), nn.DataParallel as in vgtk/app/trainer.py:153-160) must set it on each of them.  One
// DynSmemOnce per launch site remembers, lock-free, on which devices the attribute is already raised; a race
// between two host threads only repeats an idempotent call.
struct DynSmemOnce {
    unsigned long long done[2] = {0ull, 0ull};  // bit d <-> device d (128 devices)
};
template <class Kernel>
inline int ensure_dyn_smem(DynSmemOnce &once, Kernel kernel, int bytes, const char *what) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int w = (dev >> 6) & 1;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&once.done[w], __ATOMIC_ACQUIRE) & bit) return 0;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s: cannot raise dynamic shared memory to %d bytes: %s", what, bytes, cudaGetErrorString(e));
        return (int)e;
    }
    __atomic_fetch_or(&once.done[w], bit, __ATOMIC_RELEASE);
    return 0;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// fp32 squared distance pinned to the reference kernels' SASS order
// (FMUL, FFMA, FFMA):  (dx*dx + dy*dy) + dz*dz -> fma(dz,dz, fma(dx,dx, dy*dy)).
__device__ __forceinline__ float sqdist3(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

}  // namespace epn
