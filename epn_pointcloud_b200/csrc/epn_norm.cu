// Fused normalisation + leaky_relu of the block wrappers:
//   InterSO3ConvBlock / SeparableSO3ConvBlock : BatchNorm2d (training statistics, affine) -> leaky_relu
//                                               SPConvNets/utils/base_so3conv.py:107,119-125,193,209
//   IntraSO3ConvBlock                         : InstanceNorm2d(affine=False) -> leaky_relu     base_so3conv.py:43,55-57
// on feature maps x [b, c, n] with n = points * anchors contiguous.
//
// Forward: one pass for the row statistics (two sweeps over a row that stays in L1/L2: mean, then
// sum (x-mean)^2 -- no E[x^2]-E[x]^2 cancellation), a tiny combine over the batch (Chan's formula) for
// BatchNorm, one vectorised apply pass.  Backward: one reduction pass (sum g, sum g*xhat with
// g = dy * leaky'(z)), the combine, one apply pass.  All HBM-bound, float4 accesses.
#include "epn_internal.cuh"

namespace epn {

constexpr int NT = 256;

__device__ __forceinline__ float block_sum(float v, float *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (lane < NT / 32) ? s_red[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return __shfl_sync(0xffffffffu, t, 0);
}

__device__ __forceinline__ double block_sum_d(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = (lane < NT / 32) ? s_red[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return __shfl_sync(0xffffffffu, t, 0);
}

// part[row] = (mean, M2) of row (b,c), accumulated in fp64 (the kernel is HBM-bound: ~24 bytes per SM and cycle need
// 6 fp64 adds per cycle of the 64 an SM has): a row that is CONSTANT (block 0 of every shipped model: the skip
// convolution of the all-ones occupancy features, base_so3conv.py:206-211) must come out with mean == the constant
// and M2 == 0 exactly -- with an fp32 mean one ulp off, (x - mean) / sqrt(0 + eps) turns that ulp into a 1e-5 error.
__global__ void __launch_bounds__(NT)
norm_row_stats_kernel(const float *__restrict__ x, double2 *__restrict__ part, int n) {
    __shared__ double s_red[NT / 32];
    const float *row = x + (size_t)blockIdx.x * n;
    const int n4 = (n % 4 == 0) ? n / 4 : 0;
    double s = 0.0;
    for (int i = threadIdx.x; i < n4; i += NT) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + i);
        s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += NT) s += (double)__ldg(row + i);
    const double mean_d = block_sum_d(s, s_red) / (double)n;
    double m2 = 0.0;
    for (int i = threadIdx.x; i < n4; i += NT) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + i);
        const double a = (double)v.x - mean_d, b = (double)v.y - mean_d, c = (double)v.z - mean_d, d = (double)v.w - mean_d;
        m2 += (a * a + b * b) + (c * c + d * d);
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += NT) {
        const double a = (double)__ldg(row + i) - mean_d;
        m2 += a * a;
    }
    m2 = block_sum_d(m2, s_red);
    if (threadIdx.x == 0) part[blockIdx.x] = make_double2(mean_d, m2 > 0.0 ? m2 : 0.0);
}

// stats[g] = mean, stats[G + g] = rstd.  mode 0: g = row.  mode 1: g = channel, combined over the batch (fp64).
__global__ void norm_finalize_kernel(const double2 *__restrict__ part, float *__restrict__ stats, int b, int c, int n,
                                     int mode, float eps) {
    const int G = mode == 0 ? b * c : c;
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= G) return;
    if (mode == 0) {
        const double2 p = part[gi];
        stats[gi] = (float)p.x;
        stats[G + gi] = (float)(1.0 / sqrt(p.y / (double)n + (double)eps));
    } else {
        double mean = 0.0;
        for (int bi = 0; bi < b; ++bi) mean += part[bi * c + gi].x;
        mean /= (double)b;
        double m2 = 0.0;
        for (int bi = 0; bi < b; ++bi) {
            const double2 p = part[bi * c + gi];
            const double d = p.x - mean;
            m2 += p.y + (double)n * d * d;
        }
        stats[gi] = (float)mean;
        stats[G + gi] = (float)(1.0 / sqrt(m2 / ((double)b * (double)n) + (double)eps));
    }
}

__device__ __forceinline__ float lrelu(float z, float slope) { return z > 0.f ? z : z * slope; }

__global__ void __launch_bounds__(NT)
norm_apply_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                  const float *__restrict__ stats, const float *__restrict__ residual, float *__restrict__ y, int c, int n,
                  int G, int mode, float slope) {
    const int rowi = blockIdx.y;
    const int ch = rowi % c;
    const int gi = mode == 0 ? rowi : ch;
    const float mean = stats[gi], rstd = stats[G + gi];
    const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
    // (x - mean) * sc + beta, NOT x * sc + (beta - mean * sc): the subtraction first is exact for values near the mean,
    // the folded form loses |mean| * sc * 2^-24 (a constant row would come out as beta + noise instead of beta)
    const float sc = rstd * ga;
    const float *row = x + (size_t)rowi * n;
    const float *res = residual ? residual + (size_t)rowi * n : nullptr;   // skip connection: y = act(norm(x)) + residual
    float *out = y + (size_t)rowi * n;
    if (n % 4 == 0) {
        for (int i = blockIdx.x * NT + threadIdx.x; i < n / 4; i += gridDim.x * NT) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + i);
            float4 r;
            r.x = lrelu(fmaf(v.x - mean, sc, be), slope);
            r.y = lrelu(fmaf(v.y - mean, sc, be), slope);
            r.z = lrelu(fmaf(v.z - mean, sc, be), slope);
            r.w = lrelu(fmaf(v.w - mean, sc, be), slope);
            if (res != nullptr) {
                const float4 q = __ldg(reinterpret_cast<const float4 *>(res) + i);
                r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
            }
            reinterpret_cast<float4 *>(out)[i] = r;
        }
    } else {
        for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += gridDim.x * NT)
            out[i] = lrelu(fmaf(row[i] - mean, sc, be), slope) + (res ? res[i] : 0.f);
    }
}

// part[row] = (sum g, sum g * xhat),  g = dy * leaky'(z)
__global__ void __launch_bounds__(NT)
norm_bwd_reduce_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ gamma,
                       const float *__restrict__ beta, const float *__restrict__ stats, float2 *__restrict__ part, int c,
                       int n, int G, int mode, float slope) {
    __shared__ float s_red[NT / 32];
    const int rowi = blockIdx.x;
    const int ch = rowi % c;
    const int gi = mode == 0 ? rowi : ch;
    const float mean = stats[gi], rstd = stats[G + gi];
    const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
    const float *xr = x + (size_t)rowi * n, *dr = dy + (size_t)rowi * n;
    float s1 = 0.f, s2 = 0.f;
    auto term = [&](float xv, float dv) {
        const float xh = (xv - mean) * rstd;
        const float z = fmaf(xh, ga, be);
        const float gd = dv * (z > 0.f ? 1.f : slope);
        s1 += gd;
        s2 = fmaf(gd, xh, s2);
    };
    const int n4 = (n % 4 == 0 && (((uintptr_t)xr | (uintptr_t)dr) & 15) == 0) ? n / 4 : 0;
    for (int i = threadIdx.x; i < n4; i += NT) {
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(xr) + i), dv = __ldg(reinterpret_cast<const float4 *>(dr) + i);
        term(xv.x, dv.x); term(xv.y, dv.y); term(xv.z, dv.z); term(xv.w, dv.w);
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += NT) term(__ldg(xr + i), __ldg(dr + i));
    s1 = block_sum(s1, s_red);
    s2 = block_sum(s2, s_red);
    if (threadIdx.x == 0) part[rowi] = make_float2(s1, s2);
}

// mode 1: sums[c] = (S1, S2) over the batch, dgamma = S2, dbeta = S1
__global__ void norm_bwd_finalize_kernel(const float2 *__restrict__ part, float2 *__restrict__ sums,
                                         float *__restrict__ dgamma, float *__restrict__ dbeta, int b, int c) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    float s1 = 0.f, s2 = 0.f;
    for (int bi = 0; bi < b; ++bi) {
        const float2 p = part[bi * c + ch];
        s1 += p.x;
        s2 += p.y;
    }
    sums[ch] = make_float2(s1, s2);
    if (dgamma) dgamma[ch] = s2;
    if (dbeta) dbeta[ch] = s1;
}

__global__ void __launch_bounds__(NT)
norm_bwd_apply_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ gamma,
                      const float *__restrict__ beta, const float *__restrict__ stats, const float2 *__restrict__ sums,
                      float *__restrict__ dx, int c, int n, int G, int mode, float slope, float inv_count) {
    const int rowi = blockIdx.y;
    const int ch = rowi % c;
    const int gi = mode == 0 ? rowi : ch;
    const float mean = stats[gi], rstd = stats[G + gi];
    const float ga = gamma ? gamma[ch] : 1.f, be = beta ? beta[ch] : 0.f;
    const float2 sm = sums[gi];
    const float m1 = sm.x * inv_count, m2 = sm.y * inv_count, k = ga * rstd;
    const float *xr = x + (size_t)rowi * n, *dr = dy + (size_t)rowi * n;
    float *out = dx + (size_t)rowi * n;
    auto elem = [&](float xv, float dv) {
        const float xh = (xv - mean) * rstd;
        const float z = fmaf(xh, ga, be);
        const float gd = dv * (z > 0.f ? 1.f : slope);
        return k * (gd - m1 - xh * m2);
    };
    const int n4 = (n % 4 == 0 && (((uintptr_t)xr | (uintptr_t)dr | (uintptr_t)out) & 15) == 0) ? n / 4 : 0;
    for (int i = blockIdx.x * NT + threadIdx.x; i < n4; i += gridDim.x * NT) {
        const float4 xv = __ldg(reinterpret_cast<const float4 *>(xr) + i), dv = __ldg(reinterpret_cast<const float4 *>(dr) + i);
        reinterpret_cast<float4 *>(out)[i] = make_float4(elem(xv.x, dv.x), elem(xv.y, dv.y), elem(xv.z, dv.z), elem(xv.w, dv.w));
    }
    for (int i = n4 * 4 + blockIdx.x * NT + threadIdx.x; i < n; i += gridDim.x * NT) out[i] = elem(__ldg(xr + i), __ldg(dr + i));
}

// BatchNorm's running-statistics bookkeeping (nn.BatchNorm2d.forward in training mode) from the batch (mean, rstd) of
// the fused kernels, as ONE launch: the same arithmetic as eight elementwise torch kernels per BatchNorm layer, which at
// small per-GPU batches (strong scaling) were a visible share of the ~600 launches of a step.  One block; thread <->
// channel.  momentum < 0: cumulative moving average (momentum=None), factor 1 / (num_batches_tracked + 1).
__global__ void __launch_bounds__(256)
bn_track_kernel(const float *__restrict__ stats, const float *__restrict__ bias, float *__restrict__ running_mean,
                float *__restrict__ running_var, long long *__restrict__ num_batches_tracked, int c, float count, float momentum,
                float eps) {
    const long long nbt = *num_batches_tracked + 1;
    const float m = momentum >= 0.f ? momentum : 1.0f / (float)nbt;
    const float unbias = count / fmaxf(count - 1.0f, 1.0f);
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const float mean = stats[ch] + (bias ? bias[ch] : 0.f);
        const float rstd = stats[c + ch];
        const float var = (1.0f / (rstd * rstd) - eps) * unbias;
        if (momentum >= 0.f) {
            running_mean[ch] = running_mean[ch] * (1.0f - m) + m * mean;
            running_var[ch] = running_var[ch] * (1.0f - m) + m * var;
        } else {
            running_mean[ch] += (mean - running_mean[ch]) * m;
            running_var[ch] += (var - running_var[ch]) * m;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *num_batches_tracked = nbt;
}

}  // namespace epn

using namespace epn;

EPN_API int epn_bn_track_f32(const float *stats, const float *bias, float *running_mean, float *running_var,
                             long long *num_batches_tracked, int c, long long count, float momentum, float eps, void *stream) {
    EPN_REQUIRE_PTR(stats); EPN_REQUIRE_PTR(running_mean); EPN_REQUIRE_PTR(running_var); EPN_REQUIRE_PTR(num_batches_tracked);
    EPN_REQUIRE_POS(c);
    EPN_REQUIRE(count > 0, EPN_ERR_SHAPE, "count must be > 0");
    cudaStream_t s = as_stream(stream);
    ProfScope prof(s, KC_NORM);
    bn_track_kernel<<<1, 256, 0, s>>>(stats, bias, running_mean, running_var, num_batches_tracked, c, (float)count, momentum, eps);
    return check_launch("bn_track_kernel");
}

EPN_API size_t epn_norm_act_workspace_bytes(int b, int c) {
    if (b <= 0 || c <= 0) return 0;
    return (size_t)2 * b * c * sizeof(float2) + 256;
}

EPN_API int epn_norm_act_fwd_f32(const float *x, const float *gamma, const float *beta, const float *residual, float *y,
                                 float *stats, void *workspace, size_t workspace_bytes, int b, int c, int n, int mode,
                                 float eps, float slope, void *stream) {
    EPN_REQUIRE_PTR(x); EPN_REQUIRE_PTR(y); EPN_REQUIRE_PTR(stats); EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(n);
    EPN_REQUIRE(mode >= 0 && mode <= 2, EPN_ERR_SHAPE, "mode must be 0 (instance), 1 (batch) or 2 (batch, given statistics)");
    EPN_REQUIRE((long long)b * c <= 2147483647LL / 2, EPN_ERR_SHAPE, "b*c too large");
    EPN_REQUIRE(workspace_bytes >= epn_norm_act_workspace_bytes(b, c), EPN_ERR_WORKSPACE, "workspace too small");
    EPN_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual) & 15) == 0, EPN_ERR_ALIGN, "x, y and residual must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    double2 *part = static_cast<double2 *>(workspace);   // rows x 16 bytes = the first half of the workspace
    const int rows = b * c, G = mode == 0 ? rows : c;
    ProfScope prof(s, KC_NORM);
    int rc = 0;
    if (mode == 2) {
        mode = 1;   // evaluation-mode BatchNorm: the caller's per-channel (mean, rstd) are applied as they are
    } else {
        norm_row_stats_kernel<<<rows, NT, 0, s>>>(x, part, n);
        rc = check_launch("norm_row_stats_kernel");
        if (rc) return rc;
        norm_finalize_kernel<<<cdiv(G, 128), 128, 0, s>>>(part, stats, b, c, n, mode, eps);
        rc = check_launch("norm_finalize_kernel");
        if (rc) return rc;
    }
    dim3 grid(cdiv(n, NT * 4 * 4) > 0 ? cdiv(n, NT * 4 * 4) : 1, rows);
    if (rows > 65535) {
        set_error("epn_norm_act_fwd_f32: b*c = %d rows exceed the grid limit", rows);
        return EPN_ERR_SHAPE;
    }
    norm_apply_kernel<<<grid, NT, 0, s>>>(x, gamma, beta, stats, residual, y, c, n, G, mode, slope);
    return check_launch("norm_apply_kernel");
}

// statistics only: what epn_norm_act_fwd_f32 computes before its apply pass (for consumers that apply the
// normalisation themselves while loading, epn_intra_so3conv_fwd_norm_f32)
EPN_API int epn_norm_stats_f32(const float *x, float *stats, void *workspace, size_t workspace_bytes, int b, int c, int n,
                               int mode, float eps, void *stream) {
    EPN_REQUIRE_PTR(x); EPN_REQUIRE_PTR(stats); EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(n);
    EPN_REQUIRE(mode == 0 || mode == 1, EPN_ERR_SHAPE, "mode must be 0 (instance) or 1 (batch)");
    EPN_REQUIRE((long long)b * c <= 2147483647LL / 2, EPN_ERR_SHAPE, "b*c too large");
    EPN_REQUIRE(workspace_bytes >= epn_norm_act_workspace_bytes(b, c), EPN_ERR_WORKSPACE, "workspace too small");
    EPN_REQUIRE(((uintptr_t)x & 15) == 0, EPN_ERR_ALIGN, "x must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    double2 *part = static_cast<double2 *>(workspace);
    const int rows = b * c, G = mode == 0 ? rows : c;
    ProfScope prof(s, KC_NORM);
    norm_row_stats_kernel<<<rows, NT, 0, s>>>(x, part, n);
    int rc = check_launch("norm_row_stats_kernel");
    if (rc) return rc;
    norm_finalize_kernel<<<cdiv(G, 128), 128, 0, s>>>(part, stats, b, c, n, mode, eps);
    return check_launch("norm_finalize_kernel");
}

EPN_API int epn_norm_act_bwd_f32(const float *dy, const float *x, const float *gamma, const float *beta,
                                 const float *stats, float *dx, float *dgamma, float *dbeta, void *workspace,
                                 size_t workspace_bytes, int b, int c, int n, int mode, float slope, void *stream) {
    EPN_REQUIRE_PTR(dy); EPN_REQUIRE_PTR(x); EPN_REQUIRE_PTR(stats); EPN_REQUIRE_PTR(dx); EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(n);
    EPN_REQUIRE(mode == 0 || mode == 1, EPN_ERR_SHAPE, "mode must be 0 (instance) or 1 (batch)");
    EPN_REQUIRE(workspace_bytes >= epn_norm_act_workspace_bytes(b, c), EPN_ERR_WORKSPACE, "workspace too small");
    const int rows = b * c, G = mode == 0 ? rows : c;
    if (rows > 65535) {
        set_error("epn_norm_act_bwd_f32: b*c = %d rows exceed the grid limit", rows);
        return EPN_ERR_SHAPE;
    }
    cudaStream_t s = as_stream(stream);
    float2 *part = static_cast<float2 *>(workspace);
    float2 *sums = part + rows;
    ProfScope prof(s, KC_NORM);
    norm_bwd_reduce_kernel<<<rows, NT, 0, s>>>(dy, x, gamma, beta, stats, part, c, n, G, mode, slope);
    int rc = check_launch("norm_bwd_reduce_kernel");
    if (rc) return rc;
    const float2 *group_sums = part;
    if (mode == 1) {
        norm_bwd_finalize_kernel<<<cdiv(c, 128), 128, 0, s>>>(part, sums, dgamma, dbeta, b, c);
        rc = check_launch("norm_bwd_finalize_kernel");
        if (rc) return rc;
        group_sums = sums;
    }
    dim3 grid(cdiv(n, NT * 4 * 4) > 0 ? cdiv(n, NT * 4 * 4) : 1, rows);
    const float inv_count = mode == 0 ? 1.0f / (float)n : 1.0f / ((float)b * (float)n);
    norm_bwd_apply_kernel<<<grid, NT, 0, s>>>(dy, x, gamma, beta, stats, group_sums, dx, c, n, G, mode, slope, inv_count);
    return check_launch("norm_bwd_apply_kernel");
}
