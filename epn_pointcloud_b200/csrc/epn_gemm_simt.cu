// Generic fp32 SIMT GEMM used (a) for shapes the tcgen05 path does not cover and
// (b) as the in-library cross-check of the tensor-core kernels in the GPU tests.
//   C[z][m,n] (+)= sum_k A[z](m,k) * B[z](k,n)
// Operands are addressed with explicit (row, col) strides so W, W^T, x, x^T and
// dout can all be used in place.  split_k > 1 accumulates with fp32 RED into a
// pre-zeroed C (used for dW, whose reduction runs over batch * points * anchors).
#include "epn_internal.cuh"

namespace epn {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
sgemm_kernel(GemmOperand A, GemmOperand B, float *__restrict__ C, long long c_stride_z, long long ldc,
             int M, int N, int K, int split_k, int accumulate) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int t = threadIdx.x;
    const int tx = t % 16, ty = t / 16;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int z = blockIdx.z / split_k, ks = blockIdx.z % split_k;
    const float *Ap = A.ptr + (size_t)z * A.stride_z;
    const float *Bp = B.ptr + (size_t)z * B.stride_z;
    const int k_per = ((K + split_k - 1) / split_k + TK - 1) / TK * TK;
    const int k_begin = ks * k_per, k_end = min(K, k_begin + k_per);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += TK) {
        // A tile: TM x TK.  Map threads so the contiguous operand axis is fastest.
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int m, k;
            if (A.stride_col == 1) { m = t / 4; k = (t % 4) * 4 + j; }  // k contiguous
            else { m = t % 64; k = t / 64 + 4 * j; }                    // m contiguous
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < k_end) ? __ldg(Ap + (size_t)gm * A.stride_row + (size_t)gk * A.stride_col) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n, k;
            if (B.stride_col == 1) { n = t % 64; k = t / 64 + 4 * j; }  // n contiguous
            else { n = t / 4; k = (t % 4) * 4 + j; }                    // k contiguous
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < k_end) ? __ldg(Bp + (size_t)gk * B.stride_row + (size_t)gn * B.stride_col) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *Cp = C + (size_t)z * c_stride_z;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float *dst = Cp + (size_t)gm * ldc + gn;
            if (split_k > 1 || accumulate == 2) atomicAdd(dst, acc[i][j]);
            else if (accumulate) *dst += acc[i][j];
            else *dst = acc[i][j];
        }
    }
}

// accumulate: 0 = overwrite, 1 = C += (single writer), 2 = atomic C += (several
// launches / z-slices target the same C, e.g. dW over clouds).
int launch_sgemm(const GemmOperand &A, const GemmOperand &B, float *C, long long c_stride_z, long long ldc,
                 int M, int N, int K, int batch, int split_k, int accumulate, cudaStream_t s) {
    if (split_k < 1) split_k = 1;
    dim3 grid(cdiv(N, TN), cdiv(M, TM), batch * split_k);
    if (grid.z > 65535 || grid.y > 65535) {
        set_error("sgemm: grid too large (M=%d batch*split=%d)", M, batch * split_k);
        return EPN_ERR_SHAPE;
    }
    ProfScope prof(s, KC_GEMM);
    sgemm_kernel<<<grid, 256, 0, s>>>(A, B, C, c_stride_z, ldc, M, N, K, split_k, accumulate);
    return check_launch("sgemm_kernel");
}

}  // namespace epn
