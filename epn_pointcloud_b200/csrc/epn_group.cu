// Grouping stages of the SPConv hot path (unfused op surface):
//   inter kernel weights, inter feature grouping fwd/bwd, intra grouping fwd/bwd,
//   and the generic zpconv (5-D index) surface.
// Replaces the PyTorch op chains of vgtk/vgtk/so3conv/functional.py:180-268 and
// vgtk/vgtk/spconv/functional.py:361-390, and the atomicAdd kernels of
// vgtk/vgtk/cuda/zpconv_cuda_kernel.cu:32-195.
//
// Thread mapping shared by the inter kernels: lane <-> anchor a (feature rows
// [.., q, 0..na) are contiguous, so a warp reads/writes one 4*na-byte segment),
// warp-pair <-> group of KG kernel points.  A thread keeps its w[KG][NN] slice of
// the kernel-weight tensor in REGISTERS for one point and streams all channels
// through it, so inter_w never has to exist in memory.
#include "epn_internal.cuh"

namespace epn {

constexpr int ALANES = 64;  // anchor lanes per k-group (na <= 64)

// rk[i] = anchors[a] @ kernels[k0+i]   (so3conv/functional.py:190)
template <int KG>
__device__ __forceinline__ void rotated_kernels(const InterGeom &g, int a, int k0, int ks,
                                                float (&rx)[KG], float (&ry)[KG], float (&rz)[KG]) {
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + a * 9 + i);
#pragma unroll
    for (int i = 0; i < KG; ++i) {
        const int k = min(k0 + i, ks - 1);
        const float kx = __ldg(g.kernels + k * 3), ky = __ldg(g.kernels + k * 3 + 1),
                    kz = __ldg(g.kernels + k * 3 + 2);
        rx[i] = R[0] * kx + R[1] * ky + R[2] * kz;
        ry[i] = R[3] * kx + R[4] * ky + R[5] * kz;
        rz[i] = R[6] * kx + R[7] * ky + R[8] * kz;
    }
}

// -------------------------------------------------------------- inter weights
// inter_w[b,p,a,k,n] (so3conv/functional.py:198-200).  One CTA per (p, b);
// threads sweep (a,k,n) with n fastest -> coalesced 4-byte stores of the
// contiguous [na*ks*nn] slab of the point.
__global__ void __launch_bounds__(256)
inter_weights_kernel(InterGeom g, const int32_t *__restrict__ idx, float *__restrict__ inter_w,
                     int p_in, int p, int nn, int na, int ks) {
    extern __shared__ float s_dyn[];
    float *s_rk = s_dyn;                 // [na*ks][3]
    float *s_g = s_dyn + na * ks * 3;    // [nn][3]
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < na * ks; i += blockDim.x) {
        const int a = i / ks, k = i - a * ks;
        const float *R = g.anchors + a * 9;
        const float kx = g.kernels[k * 3], ky = g.kernels[k * 3 + 1], kz = g.kernels[k * 3 + 2];
        s_rk[i * 3 + 0] = R[0] * kx + R[1] * ky + R[2] * kz;
        s_rk[i * 3 + 1] = R[3] * kx + R[4] * ky + R[5] * kz;
        s_rk[i * 3 + 2] = R[6] * kx + R[7] * ky + R[8] * kz;
    }
    const float *X = g.xyz + (size_t)b * 3 * p_in;
    const float *Cn = g.centers + (size_t)b * 3 * p;
    for (int pi = blockIdx.x; pi < p; pi += gridDim.x) {
        __syncthreads();
        for (int n = threadIdx.x; n < nn; n += blockDim.x) {
            const int q = idx[((size_t)b * p + pi) * nn + n];
            s_g[n * 3 + 0] = X[q] - Cn[pi];
            s_g[n * 3 + 1] = X[p_in + q] - Cn[p + pi];
            s_g[n * 3 + 2] = X[2 * p_in + q] - Cn[2 * p + pi];
        }
        __syncthreads();
        float *dst = inter_w + ((size_t)b * p + pi) * na * ks * nn;
        const int total = na * ks * nn;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int ak = i / nn, n = i - ak * nn;
            const float dx = s_g[n * 3] - s_rk[ak * 3], dy = s_g[n * 3 + 1] - s_rk[ak * 3 + 1],
                        dz = s_g[n * 3 + 2] - s_rk[ak * 3 + 2];
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            dst[i] = fmaxf(__fsub_rn(1.0f, __fdiv_rn(d, g.sigma)), 0.0f);
        }
    }
}

// ---------------------------------------------------- w slice into registers
// Fills w[KG][NN] for anchor a, kernel points k0.., neighbours n0.. of point
// (b,pi): from inter_w when given, else from geometry (s_g = neighbour offsets).
template <int KG, int NN>
__device__ __forceinline__ void load_w(float (&w)[KG][NN], const float *__restrict__ inter_w_pt,
                                       const float *s_g, const float (&rx)[KG], const float (&ry)[KG],
                                       const float (&rz)[KG], float sigma, int a, int k0, int n0,
                                       int nn, int ks, bool a_ok) {
#pragma unroll
    for (int i = 0; i < KG; ++i) {
        const bool k_ok = a_ok && (k0 + i) < ks;
#pragma unroll
        for (int n = 0; n < NN; ++n) {
            float v = 0.f;
            if (k_ok && n0 + n < nn) {
                if (inter_w_pt != nullptr) {
                    v = __ldg(inter_w_pt + ((size_t)a * ks + k0 + i) * nn + n0 + n);
                } else {
                    v = kernel_weight(s_g[(n0 + n) * 3], s_g[(n0 + n) * 3 + 1], s_g[(n0 + n) * 3 + 2], rx[i], ry[i],
                                      rz[i], sigma);
                }
            }
            w[i][n] = v;
        }
    }
}

// ------------------------------------------------------ inter grouping forward
// out[b,c,k,p,a] = sum_n feats[b,c,idx[b,p,n],a] * w[b,p,a,k,n]
// (spconv/functional.py:372-390).  feats == nullptr: occupancy features (c == 1,
// feats == 1, so3conv/functional.py:25-44).  Output addressed with explicit
// strides so the fused conv can write a compact [chunk][c*ks][p_chunk*na] slab.
struct GroupOut {
    float *ptr;
    long long stride_b;   // elements between clouds
    long long stride_ck;  // elements between (c,k) rows
    int p_off;            // first point of the slab
    int p_cnt;            // points in the slab
};

template <int KG, int NN>
__global__ void __launch_bounds__(KG >= 6 ? 256 : 512)
inter_group_fwd_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx,
                       const float *__restrict__ inter_w, InterGeom g, GroupOut out, int c, int p_in,
                       int p, int nn, int na, int ks) {
    extern __shared__ float s_dyn[];
    float *s_g = s_dyn;                                          // [nn][3]
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_dyn + nn * 3);  // [nn]
    const int a = threadIdx.x % ALANES;
    const int grp = threadIdx.x / ALANES, ngrp = blockDim.x / ALANES;
    const int b = blockIdx.y;
    const bool a_ok = a < na;
    const int aa = a_ok ? a : 0;
    const float *F = feats ? feats + (size_t)b * c * p_in * na : nullptr;
    const int kgroups = (ks + KG - 1) / KG;

    for (int pl = blockIdx.x; pl < out.p_cnt; pl += gridDim.x) {
        const int pi = out.p_off + pl;
        __syncthreads();
        for (int n = threadIdx.x; n < nn; n += blockDim.x) {
            const int q = idx[((size_t)b * p + pi) * nn + n];
            s_idx[n] = q;
            if (inter_w == nullptr) {
                const float *X = g.xyz + (size_t)b * 3 * p_in;
                const float *Cn = g.centers + (size_t)b * 3 * p;
                s_g[n * 3 + 0] = X[q] - Cn[pi];
                s_g[n * 3 + 1] = X[p_in + q] - Cn[p + pi];
                s_g[n * 3 + 2] = X[2 * p_in + q] - Cn[2 * p + pi];
            }
        }
        __syncthreads();
        const float *wpt = inter_w ? inter_w + ((size_t)b * p + pi) * na * ks * nn : nullptr;
        for (int kg = grp; kg < kgroups; kg += ngrp) {
            const int k0 = kg * KG;
            float rx[KG], ry[KG], rz[KG];
            if (inter_w == nullptr) rotated_kernels<KG>(g, aa, k0, ks, rx, ry, rz);
            for (int n0 = 0; n0 < nn; n0 += NN) {
                float w[KG][NN];
                load_w<KG, NN>(w, wpt, s_g, rx, ry, rz, g.sigma, aa, k0, n0, nn, ks, a_ok);
                for (int ci = 0; ci < c; ++ci) {
                    float f[NN];
#pragma unroll
                    for (int n = 0; n < NN; ++n) {
                        f[n] = 1.0f;
                        if (F != nullptr) {
                            const int q = s_idx[min(n0 + n, nn - 1)];
                            f[n] = a_ok ? __ldg(F + ((size_t)ci * p_in + q) * na + a) : 0.f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < KG; ++i) {
                        float acc = 0.f;
#pragma unroll
                        for (int n = 0; n < NN; ++n) acc = fmaf(w[i][n], f[n], acc);
                        if (a_ok && k0 + i < ks) {
                            float *o = out.ptr + (size_t)b * out.stride_b +
                                       (size_t)(ci * ks + k0 + i) * out.stride_ck + (size_t)pl * na + a;
                            if (n0 == 0) *o = acc; else *o += acc;
                        }
                    }
                }
            }
        }
    }
}

// ----------------------------------------------------- inter grouping backward
// dfeats[b,c,idx[b,p,n],a] += sum_k w[b,p,a,k,n] * dout[b,c,k,p,a].
// Thread = (anchor lane, group of NB neighbours) holding w[ks<=KSMAX][NB] in
// registers; all channels stream through; one fp32 RED per (c, n, a) into the
// 4*na-byte feature row of the neighbour (contention only among points that
// share a neighbour).
template <int KSMAX, int NB>
__global__ void __launch_bounds__(512)
inter_group_bwd_kernel(const float *__restrict__ dgrouped, long long d_stride_b, long long d_stride_ck,
                       int p_off, int p_cnt, const int32_t *__restrict__ idx,
                       const float *__restrict__ inter_w, InterGeom g, float *__restrict__ dfeats, int c,
                       int p_in, int p, int nn, int na, int ks) {
    extern __shared__ float s_dyn[];
    float *s_g = s_dyn;
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_dyn + nn * 3);
    const int a = threadIdx.x % ALANES;
    const int grp = threadIdx.x / ALANES, ngrp = blockDim.x / ALANES;
    const int b = blockIdx.y;
    const bool a_ok = a < na;
    const int aa = a_ok ? a : 0;
    float *DF = dfeats + (size_t)b * c * p_in * na;
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = inter_w ? 0.f : __ldg(g.anchors + aa * 9 + i);
    const int ngroups = (nn + NB - 1) / NB;

    for (int pl = blockIdx.x; pl < p_cnt; pl += gridDim.x) {
        const int pi = p_off + pl;
        __syncthreads();
        for (int n = threadIdx.x; n < nn; n += blockDim.x) {
            const int q = idx[((size_t)b * p + pi) * nn + n];
            s_idx[n] = q;
            if (inter_w == nullptr) {
                const float *X = g.xyz + (size_t)b * 3 * p_in;
                const float *Cn = g.centers + (size_t)b * 3 * p;
                s_g[n * 3 + 0] = X[q] - Cn[pi];
                s_g[n * 3 + 1] = X[p_in + q] - Cn[p + pi];
                s_g[n * 3 + 2] = X[2 * p_in + q] - Cn[2 * p + pi];
            }
        }
        __syncthreads();
        const float *wpt = inter_w ? inter_w + ((size_t)b * p + pi) * na * ks * nn : nullptr;
        for (int ng = grp; ng < ngroups; ng += ngrp) {
            const int n0 = ng * NB;
            for (int kb = 0; kb < ks; kb += KSMAX) {
                float w[KSMAX][NB];
#pragma unroll
                for (int k = 0; k < KSMAX; ++k) {
                    const bool k_ok = a_ok && kb + k < ks;
                    float rx = 0.f, ry = 0.f, rz = 0.f;
                    if (k_ok && inter_w == nullptr) {
                        const float kx = __ldg(g.kernels + (kb + k) * 3), ky = __ldg(g.kernels + (kb + k) * 3 + 1),
                                    kz = __ldg(g.kernels + (kb + k) * 3 + 2);
                        rx = R[0] * kx + R[1] * ky + R[2] * kz;
                        ry = R[3] * kx + R[4] * ky + R[5] * kz;
                        rz = R[6] * kx + R[7] * ky + R[8] * kz;
                    }
#pragma unroll
                    for (int n = 0; n < NB; ++n) {
                        float v = 0.f;
                        if (k_ok && n0 + n < nn) {
                            if (inter_w != nullptr) {
                                v = __ldg(wpt + ((size_t)a * ks + kb + k) * nn + n0 + n);
                            } else {
                                v = kernel_weight(s_g[(n0 + n) * 3], s_g[(n0 + n) * 3 + 1], s_g[(n0 + n) * 3 + 2], rx, ry,
                                                  rz, g.sigma);
                            }
                        }
                        w[k][n] = v;
                    }
                }
                for (int ci = 0; ci < c; ++ci) {
                    float t[NB];
#pragma unroll
                    for (int n = 0; n < NB; ++n) t[n] = 0.f;
#pragma unroll
                    for (int k = 0; k < KSMAX; ++k) {
                        float dv = 0.f;
                        if (a_ok && kb + k < ks)
                            dv = __ldg(dgrouped + (size_t)b * d_stride_b +
                                       (size_t)(ci * ks + kb + k) * d_stride_ck + (size_t)pl * na + a);
#pragma unroll
                        for (int n = 0; n < NB; ++n) t[n] = fmaf(w[k][n], dv, t[n]);
                    }
#pragma unroll
                    for (int n = 0; n < NB; ++n)
                        if (a_ok && n0 + n < nn)
                            atomicAdd(DF + ((size_t)ci * p_in + s_idx[n0 + n]) * na + a, t[n]);
                }
            }
        }
    }
}

// ------------------------------------------------------------- intra grouping
// out[b,c,k,p,a] = feats[b,c,p,intra_idx[a,k]]  (so3conv/functional.py:221-268)
// One CTA row = one (b,c,p) feature row of na floats staged in smem and expanded
// to kn permuted copies; all global accesses are contiguous 4*na-byte segments.
struct IntraOut {
    float *ptr;
    long long stride_b, stride_ck;
    int p_off, p_cnt;
};

__global__ void __launch_bounds__(256)
intra_group_fwd_kernel(const float *__restrict__ feats, const int32_t *__restrict__ intra_idx,
                       IntraOut out, int c, int p, int na, int kn) {
    extern __shared__ float s_dyn[];
    int32_t *s_ix = reinterpret_cast<int32_t *>(s_dyn);  // [na*kn]
    float *s_row = s_dyn + na * kn;                      // [rows][na]
    const int rows = blockDim.y;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < na * kn; i += blockDim.x * blockDim.y)
        s_ix[i] = intra_idx[i];
    const int b = blockIdx.y;
    const long long nrow = (long long)c * out.p_cnt;
    for (long long r0 = (long long)blockIdx.x * rows; r0 < nrow; r0 += (long long)gridDim.x * rows) {
        const long long r = r0 + threadIdx.y;
        __syncthreads();
        int ci = 0, pl = 0;
        if (r < nrow) {
            ci = (int)(r / out.p_cnt);
            pl = (int)(r - (long long)ci * out.p_cnt);
            const float *src = feats + (((size_t)b * c + ci) * p + out.p_off + pl) * na;
            for (int a = threadIdx.x; a < na; a += blockDim.x) s_row[threadIdx.y * na + a] = __ldg(src + a);
        }
        __syncthreads();
        if (r < nrow) {
            for (int i = threadIdx.x; i < kn * na; i += blockDim.x) {
                const int k = i / na, a = i - k * na;
                out.ptr[(size_t)b * out.stride_b + (size_t)(ci * kn + k) * out.stride_ck + (size_t)pl * na + a] =
                    s_row[threadIdx.y * na + s_ix[a * kn + k]];
            }
        }
    }
}

// dfeats[b,c,p,a'] = sum_k dout[b,c,k,p,inv_k[a']]  -- requires each column of
// intra_idx to be a permutation (true for the icosahedral index); deterministic,
// no atomics.  accumulate != 0 adds into dfeats instead of overwriting.
__global__ void __launch_bounds__(256)
intra_group_bwd_kernel(const float *__restrict__ dgrouped, long long d_stride_b, long long d_stride_ck,
                       int p_off, int p_cnt, const int32_t *__restrict__ intra_idx,
                       float *__restrict__ dfeats, int c, int p, int na, int kn) {
    extern __shared__ float s_dyn[];
    int32_t *s_inv = reinterpret_cast<int32_t *>(s_dyn);  // [kn][na]
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < na * kn; i += blockDim.x * blockDim.y) {
        const int a = i / kn, k = i - a * kn;
        s_inv[k * na + intra_idx[i]] = a;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const long long nrow = (long long)c * p_cnt;
    for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < nrow;
         r += (long long)gridDim.x * blockDim.y) {
        const int ci = (int)(r / p_cnt), pl = (int)(r - (long long)ci * p_cnt);
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < kn; ++k)
                s += __ldg(dgrouped + (size_t)b * d_stride_b + (size_t)(ci * kn + k) * d_stride_ck +
                           (size_t)pl * na + s_inv[k * na + a]);
            dfeats[(((size_t)b * c + ci) * p + p_off + pl) * na + a] = s;
        }
    }
}

// ------------------------------------------------------------- zpconv surface
// Generic per-(p,a,k) neighbour lists (zpconv_cuda_kernel.cu:32-195).  Forward
// kernels own their output element (no atomics); backward kernels scatter with
// fp32 RED like the reference.
__global__ void zp_inter_fwd_kernel(const int32_t *__restrict__ nbr, const float *__restrict__ w,
                                    const float *__restrict__ feats, float *__restrict__ out, int c,
                                    int nq, int np, int na, int ks, int ann) {
    const int b = blockIdx.y;
    const long long total = (long long)ks * np * na;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t % na);
        const int pn = (int)((t / na) % np);
        const int k = (int)(t / ((long long)na * np));
        const size_t qi = ((((size_t)b * np + pn) * na + a) * ks + k) * ann;
        for (int ci = 0; ci < c; ++ci) {
            float acc = 0.f;
            for (int n = 0; n < ann; ++n)
                acc = fmaf(__ldg(w + qi + n),
                           __ldg(feats + (((size_t)b * c + ci) * nq + __ldg(nbr + qi + n)) * na + a), acc);
            out[((((size_t)b * c + ci) * ks + k) * np + pn) * na + a] = acc;
        }
    }
}

__global__ void zp_inter_bwd_kernel(const int32_t *__restrict__ nbr, const float *__restrict__ w,
                                    const float *__restrict__ dout, float *__restrict__ dfeats, int c,
                                    int nq, int np, int na, int ks, int ann) {
    const int b = blockIdx.y;
    const long long total = (long long)ks * np * na;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t % na);
        const int pn = (int)((t / na) % np);
        const int k = (int)(t / ((long long)na * np));
        const size_t qi = ((((size_t)b * np + pn) * na + a) * ks + k) * ann;
        for (int ci = 0; ci < c; ++ci) {
            const float dv = __ldg(dout + ((((size_t)b * c + ci) * ks + k) * np + pn) * na + a);
            for (int n = 0; n < ann; ++n)
                atomicAdd(dfeats + (((size_t)b * c + ci) * nq + __ldg(nbr + qi + n)) * na + a,
                          __ldg(w + qi + n) * dv);
        }
    }
}

__global__ void zp_intra_fwd_kernel(const int32_t *__restrict__ nbr, const float *__restrict__ w,
                                    const float *__restrict__ feats, float *__restrict__ out, int c,
                                    int np, int na_in, int na_out, int ks, int ann) {
    const int b = blockIdx.y;
    const long long total = (long long)c * ks * np * na_out;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t % na_out);
        const int pn = (int)((t / na_out) % np);
        const int k = (int)((t / ((long long)na_out * np)) % ks);
        const int ci = (int)(t / ((long long)na_out * np * ks));
        const float *frow = feats + (((size_t)b * c + ci) * np + pn) * na_in;
        float acc = 0.f;
        for (int n = 0; n < ann; ++n)
            acc = fmaf(__ldg(w + ((size_t)a * ks + k) * ann + n), __ldg(frow + __ldg(nbr + a * ann + n)), acc);
        out[((((size_t)b * c + ci) * ks + k) * np + pn) * na_out + a] = acc;
    }
}

__global__ void zp_intra_bwd_kernel(const int32_t *__restrict__ nbr, const float *__restrict__ w,
                                    const float *__restrict__ dout, float *__restrict__ dfeats, int c,
                                    int np, int na_in, int na_out, int ks, int ann) {
    const int b = blockIdx.y;
    const long long total = (long long)c * ks * np * na_out;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int a = (int)(t % na_out);
        const int pn = (int)((t / na_out) % np);
        const int k = (int)((t / ((long long)na_out * np)) % ks);
        const int ci = (int)(t / ((long long)na_out * np * ks));
        const float dv = __ldg(dout + t + (size_t)b * total);
        float *drow = dfeats + (((size_t)b * c + ci) * np + pn) * na_in;
        for (int n = 0; n < ann; ++n)
            atomicAdd(drow + __ldg(nbr + a * ann + n), __ldg(w + ((size_t)a * ks + k) * ann + n) * dv);
    }
}

// ------------------------------------------------------------ internal launchers
// (also used by the fused convs in epn_conv.cu)
static int grid_x_for(long long units, int per_cta_units, int b) {
    long long g = (units + per_cta_units - 1) / per_cta_units;
    const long long cap = (long long)148 * 16 / (b > 0 ? 1 : 1);
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_inter_group_fwd(const float *feats, const int32_t *idx, const float *inter_w, const InterGeom &g,
                           float *out, long long stride_b, long long stride_ck, int p_off, int p_cnt, int b,
                           int c, int p_in, int p, int nn, int na, int ks, cudaStream_t s) {
    GroupOut o{out, stride_b, stride_ck, p_off, p_cnt};
    ProfScope prof(s, KC_INTER_GROUP);
    const size_t smem = (size_t)nn * 4 * sizeof(float);
    dim3 grid(p_cnt, b);
    if (nn <= 16) {
        const int groups = min(4, cdiv(ks, 6));
        inter_group_fwd_kernel<6, 16><<<grid, ALANES * groups, smem, s>>>(feats, idx, inter_w, g, o, c, p_in, p, nn, na, ks);
    } else {
        const int groups = min(8, cdiv(ks, 3));
        inter_group_fwd_kernel<3, 32><<<grid, ALANES * groups, smem, s>>>(feats, idx, inter_w, g, o, c, p_in, p, nn, na, ks);
    }
    return check_launch("inter_group_fwd_kernel");
}

int launch_inter_group_bwd(const float *dgrouped, long long stride_b, long long stride_ck, int p_off, int p_cnt,
                           const int32_t *idx, const float *inter_w, const InterGeom &g, float *dfeats, int b,
                           int c, int p_in, int p, int nn, int na, int ks, cudaStream_t s) {
    const size_t smem = (size_t)nn * 4 * sizeof(float);
    ProfScope prof(s, KC_INTER_SCATTER);
    dim3 grid(p_cnt, b);
    const int groups = min(8, cdiv(nn, 4));
    inter_group_bwd_kernel<24, 4><<<grid, ALANES * groups, smem, s>>>(dgrouped, stride_b, stride_ck, p_off, p_cnt, idx,
                                                                   inter_w, g, dfeats, c, p_in, p, nn, na, ks);
    return check_launch("inter_group_bwd_kernel");
}

int launch_intra_group_fwd(const float *feats, const int32_t *intra_idx, float *out, long long stride_b,
                           long long stride_ck, int p_off, int p_cnt, int b, int c, int p, int na, int kn,
                           cudaStream_t s) {
    IntraOut o{out, stride_b, stride_ck, p_off, p_cnt};
    ProfScope prof(s, KC_INTRA_GROUP);
    const int rows = 4;
    dim3 block(64, rows);
    const size_t smem = ((size_t)na * kn + (size_t)rows * na) * sizeof(float);
    dim3 grid(grid_x_for((long long)c * p_cnt, rows * 4, b), b);
    intra_group_fwd_kernel<<<grid, block, smem, s>>>(feats, intra_idx, o, c, p, na, kn);
    return check_launch("intra_group_fwd_kernel");
}

int launch_intra_group_bwd(const float *dgrouped, long long stride_b, long long stride_ck, int p_off, int p_cnt,
                           const int32_t *intra_idx, float *dfeats, int b, int c, int p, int na, int kn,
                           cudaStream_t s) {
    ProfScope prof(s, KC_INTRA_GROUP);
    const int rows = 4;
    dim3 block(64, rows);
    const size_t smem = (size_t)na * kn * sizeof(int32_t);
    dim3 grid(grid_x_for((long long)c * p_cnt, rows * 4, b), b);
    intra_group_bwd_kernel<<<grid, block, smem, s>>>(dgrouped, stride_b, stride_ck, p_off, p_cnt, intra_idx, dfeats, c,
                                                    p, na, kn);
    return check_launch("intra_group_bwd_kernel");
}

}  // namespace epn

using namespace epn;

#define EPN_CHECK_B(b) EPN_REQUIRE((b) <= 65535, EPN_ERR_SHAPE, "batch > 65535")

EPN_API int epn_inter_weights_f32(const float *xyz, const float *centers, const int32_t *idx,
                                  const float *anchors, const float *kernels, float sigma, float *inter_w,
                                  int b, int p_in, int p, int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(anchors);
    EPN_REQUIRE_PTR(kernels); EPN_REQUIRE_PTR(inter_w);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(nn); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    const size_t smem = ((size_t)na * ks * 3 + (size_t)nn * 3) * sizeof(float);
    EPN_REQUIRE(smem <= 200 * 1024, EPN_ERR_SHAPE, "na*ks*3 + nn*3 floats exceed shared memory");
    InterGeom g{xyz, centers, anchors, kernels, sigma};
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(inter_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(p < 148 * 8 ? p : 148 * 8, b);
    inter_weights_kernel<<<grid, 256, smem, as_stream(stream)>>>(g, idx, inter_w, p_in, p, nn, na, ks);
    return check_launch("inter_weights_kernel");
}

EPN_API int epn_inter_group_fwd_f32(const float *feats, const int32_t *idx, const float *inter_w,
                                    const float *xyz, const float *centers, const float *anchors,
                                    const float *kernels, float sigma, float *out, int b, int c, int p_in,
                                    int p, int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(nn);
    EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(na <= ALANES, EPN_ERR_SHAPE, "na > 64 anchors not supported");
    EPN_REQUIRE(feats != nullptr || c == 1, EPN_ERR_NULL, "feats NULL requires c == 1 (occupancy features)");
    if (inter_w == nullptr) {
        EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(anchors); EPN_REQUIRE_PTR(kernels);
        EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    }
    InterGeom g{xyz, centers, anchors, kernels, sigma};
    return launch_inter_group_fwd(feats, idx, inter_w, g, out, (long long)c * ks * p * na, (long long)p * na, 0, p, b,
                                  c, p_in, p, nn, na, ks, as_stream(stream));
}

EPN_API int epn_inter_group_bwd_f32(const float *dout, const int32_t *idx, const float *inter_w,
                                    const float *xyz, const float *centers, const float *anchors,
                                    const float *kernels, float sigma, float *dfeats, int b, int c, int p_in,
                                    int p, int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(dfeats);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(nn);
    EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(na <= ALANES, EPN_ERR_SHAPE, "na > 64 anchors not supported");
    if (inter_w == nullptr) {
        EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(anchors); EPN_REQUIRE_PTR(kernels);
        EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    }
    InterGeom g{xyz, centers, anchors, kernels, sigma};
    return launch_inter_group_bwd(dout, (long long)c * ks * p * na, (long long)p * na, 0, p, idx, inter_w, g, dfeats, b,
                                  c, p_in, p, nn, na, ks, as_stream(stream));
}

EPN_API int epn_intra_group_fwd_f32(const float *feats, const int32_t *intra_idx, float *out, int b, int c,
                                    int p, int na, int kn, void *stream) {
    EPN_REQUIRE_PTR(feats); EPN_REQUIRE_PTR(intra_idx); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(kn);
    EPN_CHECK_B(b);
    return launch_intra_group_fwd(feats, intra_idx, out, (long long)c * kn * p * na, (long long)p * na, 0, p, b, c, p,
                                  na, kn, as_stream(stream));
}

EPN_API int epn_intra_group_bwd_f32(const float *dout, const int32_t *intra_idx, float *dfeats, int b, int c,
                                    int p, int na, int kn, void *stream) {
    EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(intra_idx); EPN_REQUIRE_PTR(dfeats);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(kn);
    EPN_CHECK_B(b);
    return launch_intra_group_bwd(dout, (long long)c * kn * p * na, (long long)p * na, 0, p, intra_idx, dfeats, b, c, p,
                                  na, kn, as_stream(stream));
}

static int zp_grid(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    return (int)(g < 1 ? 1 : g);
}

EPN_API int epn_zp_inter_fwd_f32(const int32_t *nbr, const float *w, const float *feats, float *out, int b,
                                 int c, int nq, int np, int na, int ks, int ann, void *stream) {
    EPN_REQUIRE_PTR(nbr); EPN_REQUIRE_PTR(w); EPN_REQUIRE_PTR(feats); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(nq); EPN_REQUIRE_POS(np); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(ks); EPN_REQUIRE_POS(ann); EPN_CHECK_B(b);
    dim3 grid(zp_grid((long long)ks * np * na), b);
    zp_inter_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(nbr, w, feats, out, c, nq, np, na, ks, ann);
    return check_launch("zp_inter_fwd_kernel");
}

EPN_API int epn_zp_inter_bwd_f32(const int32_t *nbr, const float *w, const float *dout, float *dfeats, int b,
                                 int c, int nq, int np, int na, int ks, int ann, void *stream) {
    EPN_REQUIRE_PTR(nbr); EPN_REQUIRE_PTR(w); EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(dfeats);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(nq); EPN_REQUIRE_POS(np); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(ks); EPN_REQUIRE_POS(ann); EPN_CHECK_B(b);
    dim3 grid(zp_grid((long long)ks * np * na), b);
    zp_inter_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(nbr, w, dout, dfeats, c, nq, np, na, ks, ann);
    return check_launch("zp_inter_bwd_kernel");
}

EPN_API int epn_zp_intra_fwd_f32(const int32_t *nbr, const float *w, const float *feats, float *out, int b,
                                 int c, int np, int na_in, int na_out, int ks, int ann, void *stream) {
    EPN_REQUIRE_PTR(nbr); EPN_REQUIRE_PTR(w); EPN_REQUIRE_PTR(feats); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(np); EPN_REQUIRE_POS(na_in); EPN_REQUIRE_POS(na_out);
    EPN_REQUIRE_POS(ks); EPN_REQUIRE_POS(ann); EPN_CHECK_B(b);
    dim3 grid(zp_grid((long long)c * ks * np * na_out), b);
    zp_intra_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(nbr, w, feats, out, c, np, na_in, na_out, ks, ann);
    return check_launch("zp_intra_fwd_kernel");
}

EPN_API int epn_zp_intra_bwd_f32(const int32_t *nbr, const float *w, const float *dout, float *dfeats, int b,
                                 int c, int np, int na_in, int na_out, int ks, int ann, void *stream) {
    EPN_REQUIRE_PTR(nbr); EPN_REQUIRE_PTR(w); EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(dfeats);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(np); EPN_REQUIRE_POS(na_in); EPN_REQUIRE_POS(na_out);
    EPN_REQUIRE_POS(ks); EPN_REQUIRE_POS(ann); EPN_CHECK_B(b);
    dim3 grid(zp_grid((long long)c * ks * np * na_out), b);
    zp_intra_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(nbr, w, dout, dfeats, c, np, na_in, na_out, ks, ann);
    return check_launch("zp_intra_bwd_kernel");
}
