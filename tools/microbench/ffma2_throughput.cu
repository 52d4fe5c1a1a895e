#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2 &d, float2 a, float2 b) {
    uint64_t dd = *reinterpret_cast<uint64_t*>(&d), aa = *reinterpret_cast<uint64_t*>(&a), bb = *reinterpret_cast<uint64_t*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 x = make_float2(a, a), y = make_float2(b, b + 1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(acc[i].x, x.x, y.x); acc[i].y = fmaf(acc[i].y, x.y, y.y); }
            else { float2 t = y; uint64_t dd = *reinterpret_cast<uint64_t*>(&acc[i]), aa = *reinterpret_cast<uint64_t*>(&x), bb = *reinterpret_cast<uint64_t*>(&t);
                   asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(dd) : "l"(aa), "l"(bb)); acc[i] = *reinterpret_cast<float2*>(&dd); }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 100000;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 16 * iters * 148 * 8 * 256;
        printf("mode %d (%s): %.3f ms  %.1f TFLOP/s\n", mode, mode ? "FFMA2" : "FFMA", ms, fl / ms / 1e9);
    }
    return 0;
}
