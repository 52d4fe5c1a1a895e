/*
 * epn_oracle.c -- CPU restatement of the EPN / SPConv hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * library (libepn_b200.so).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may build, load or call it.
 * The product path never routes through it.
 *
 * Every function restates one function of the reference
 * (nintendops/EPN_PointCloud @ b625483); the file:line it follows is cited.
 * Pinned against the reference's own Python (run through oracle/ref_harness.py
 * in the build container) by oracle/make_golden.py -> tests/golden/.
 *
 * Index ops are bit-exact restatements (fmaf in the contraction order nvcc
 * emits for the reference kernels: FMUL, FFMA, FFMA).  Floating-point ops
 * accumulate in double and round once to float: they are the "true" value the
 * fp32 reference and the CUDA path are both compared to.
 *
 * Layouts (same as the reference): xyz [B,3,P]; feats [B,C,P,A] (A innermost);
 * idx [B,P,K] int32; inter_w [B,P,A,KS,K]; grouped [B,C,KS,P,A];
 * W [C_out, C_in*KS] (c major, k minor).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPN_ORACLE_API __attribute__((visibility("default")))

/* fp32 squared distance exactly as the reference kernels compute it after nvcc
 * contraction: (a*a + b*b) + c*c  ->  fma(c,c, fma(a,a, b*b)).
 * vgtk/vgtk/cuda/grouping_cuda_kernel.cu:95 and :388-389 */
static inline float sqdist3(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ---------------------------------------------------------------- ball query
 * vgtk/vgtk/cuda/grouping_cuda_kernel.cu:67-113 (kernel),
 * vgtk/vgtk/cuda/grouping_cuda.cpp:71-86 (zero-initialised output). */
EPN_ORACLE_API void epn_oracle_ball_query(int b, int n, int m, float radius, int nsample,
                                          const float *new_xyz, const float *xyz, int32_t *idx) {
    const float r2 = radius * radius;
    memset(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample);
    for (int bi = 0; bi < b; ++bi) {
        const float *q = new_xyz + (size_t)bi * 3 * m;
        const float *s = xyz + (size_t)bi * 3 * n;
        int32_t *out = idx + (size_t)bi * m * nsample;
        for (int j = 0; j < m; ++j) {
            const float qx = q[j], qy = q[m + j], qz = q[2 * m + j];
            int32_t *row = out + (size_t)j * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float d2 = sqdist3(qx - s[k], qy - s[n + k], qz - s[2 * n + k]);
                if (d2 < r2) row[cnt++] = k;
            }
            if (cnt < nsample - 1) /* cyclic repeat fill; cnt == nsample-1 leaves a 0 */
                for (int k = 0; k + cnt < nsample; ++k) row[k + cnt] = row[k];
        }
    }
}

/* ------------------------------------------------- furthest point sampling
 * vgtk/vgtk/cuda/grouping_cuda_kernel.cu:340-466 (kernel + __update),
 * :29-33 (block size), vgtk/vgtk/cuda/grouping_cuda.cpp:160-174 (temp = 1e10).
 * The thread/tree structure is emulated because it fixes the tie-breaking. */
EPN_ORACLE_API void epn_oracle_fps(int b, int n, int m, const float *xyz, int32_t *idx) {
    if (m <= 0) return;
    int T = 1;
    while (T * 2 <= n && T < 1024) T *= 2;
    float *temp = (float *)malloc(sizeof(float) * (size_t)n);
    float *dv = (float *)malloc(sizeof(float) * (size_t)T);
    int *di = (int *)malloc(sizeof(int) * (size_t)T);
    memset(idx, 0, sizeof(int32_t) * (size_t)b * m);
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * 3 * n;
        int32_t *out = idx + (size_t)bi * m;
        for (int k = 0; k < n; ++k) temp[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old], y1 = p[n + old], z1 = p[2 * n + old];
            for (int t = 0; t < T; ++t) {
                int besti = 0;
                float best = -1.0f;
                for (int k = t; k < n; k += T) {
                    const float x2 = p[k], y2 = p[n + k], z2 = p[2 * n + k];
                    const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
                    if ((double)mag <= 1e-3) continue;
                    const float d = sqdist3(x2 - x1, y2 - y1, z2 - z1);
                    const float d2 = d < temp[k] ? d : temp[k];
                    temp[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dv[t] = best;
                di[t] = besti;
            }
            for (int s = T / 2; s >= 1; s /= 2)
                for (int t = 0; t < s; ++t) {
                    const float v1 = dv[t], v2 = dv[t + s];
                    if (v2 > v1) { dv[t] = v2; di[t] = di[t + s]; }
                }
            old = di[0];
            out[j] = old;
        }
    }
    free(temp); free(dv); free(di);
}

/* ------------------------------------------------------------------ gather
 * vgtk/vgtk/cuda/gathering_cuda_kernel.cu:42-68 (fwd), :72-98 (bwd). */
EPN_ORACLE_API void epn_oracle_gather_fwd(int b, int c, int n, int m, const float *points,
                                          const int32_t *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                out[((size_t)bi * c + ci) * m + j] =
                    points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]];
}

EPN_ORACLE_API void epn_oracle_gather_bwd(int b, int c, int n, int m, const float *grad_out,
                                          const int32_t *idx, float *grad_points) {
    double *acc = (double *)calloc((size_t)b * c * n, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int j = 0; j < m; ++j)
                acc[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + j]] +=
                    grad_out[((size_t)bi * c + ci) * m + j];
    for (size_t i = 0; i < (size_t)b * c * n; ++i) grad_points[i] = (float)acc[i];
    free(acc);
}

/* ------------------------------------------------- inter kernel weights
 * vgtk/vgtk/so3conv/functional.py:180-218 with the grouped_xyz of
 * vgtk/vgtk/spconv/functional.py:412-421 (neighbour minus centre):
 *   rk[a,k,:]   = anchors[a] @ kernels[k]
 *   w[b,p,a,k,n] = relu(1 - |xyz[idx[b,p,n]] - center[b,p] - rk[a,k]|^2 / sigma)
 * xyz [B,3,P_in], centers [B,3,P], idx [B,P,K], anchors [A,3,3], kernels [KS,3]
 * -> inter_w [B,P,A,KS,K]. */
EPN_ORACLE_API void epn_oracle_inter_weights(int b, int p_in, int p, int nn, int na, int ks,
                                             const float *xyz, const float *centers,
                                             const int32_t *idx, const float *anchors,
                                             const float *kernels, float sigma, float *inter_w) {
    float *rk = (float *)malloc(sizeof(float) * (size_t)na * ks * 3);
    for (int a = 0; a < na; ++a)
        for (int k = 0; k < ks; ++k)
            for (int d = 0; d < 3; ++d) {
                float s = 0.f; /* fp32 matmul like torch.matmul(anchors, kernels.T) */
                for (int j = 0; j < 3; ++j) s += anchors[(a * 3 + d) * 3 + j] * kernels[k * 3 + j];
                rk[((size_t)a * ks + k) * 3 + d] = s;
            }
    for (int bi = 0; bi < b; ++bi)
        for (int pi = 0; pi < p; ++pi) {
            const float cx = centers[((size_t)bi * 3 + 0) * p + pi];
            const float cy = centers[((size_t)bi * 3 + 1) * p + pi];
            const float cz = centers[((size_t)bi * 3 + 2) * p + pi];
            for (int n = 0; n < nn; ++n) {
                const int q = idx[((size_t)bi * p + pi) * nn + n];
                /* grouped_xyz - sample_xyz is an fp32 subtraction in the reference */
                const float gx = xyz[((size_t)bi * 3 + 0) * p_in + q] - cx;
                const float gy = xyz[((size_t)bi * 3 + 1) * p_in + q] - cy;
                const float gz = xyz[((size_t)bi * 3 + 2) * p_in + q] - cz;
                for (int a = 0; a < na; ++a)
                    for (int k = 0; k < ks; ++k) {
                        const float *r = rk + ((size_t)a * ks + k) * 3;
                        const double dx = (double)gx - r[0], dy = (double)gy - r[1],
                                     dz = (double)gz - r[2];
                        const double w = 1.0 - (dx * dx + dy * dy + dz * dz) / (double)sigma;
                        inter_w[((((size_t)bi * p + pi) * na + a) * ks + k) * nn + n] =
                            (float)(w > 0.0 ? w : 0.0);
                    }
            }
        }
    free(rk);
}

/* ------------------------------------------------- inter feature grouping
 * vgtk/vgtk/spconv/functional.py:372-390 (inter_zpconv_grouping_naive):
 *   out[b,c,k,p,a] = sum_n feats[b,c,idx[b,p,n],a] * w[b,p,a,k,n]
 * feats [B,C,P_in,A], idx [B,P,K], inter_w [B,P,A,KS,K] -> out [B,C,KS,P,A] */
EPN_ORACLE_API void epn_oracle_inter_group_fwd(int b, int c, int p_in, int p, int nn, int na,
                                               int ks, const float *feats, const int32_t *idx,
                                               const float *inter_w, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int k = 0; k < ks; ++k)
                for (int pi = 0; pi < p; ++pi)
                    for (int a = 0; a < na; ++a) {
                        double s = 0.0;
                        for (int n = 0; n < nn; ++n) {
                            const int q = idx[((size_t)bi * p + pi) * nn + n];
                            s += (double)feats[(((size_t)bi * c + ci) * p_in + q) * na + a] *
                                 (double)inter_w[((((size_t)bi * p + pi) * na + a) * ks + k) * nn + n];
                        }
                        out[((((size_t)bi * c + ci) * ks + k) * p + pi) * na + a] = (float)s;
                    }
}

/* adjoint of the above w.r.t. feats (what autograd derives for the reference):
 *   dfeats[b,c,idx[b,p,n],a] += sum_k w[b,p,a,k,n] * dout[b,c,k,p,a] */
EPN_ORACLE_API void epn_oracle_inter_group_bwd(int b, int c, int p_in, int p, int nn, int na,
                                               int ks, const float *dout, const int32_t *idx,
                                               const float *inter_w, float *dfeats) {
    double *acc = (double *)calloc((size_t)b * c * p_in * na, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int pi = 0; pi < p; ++pi)
                for (int a = 0; a < na; ++a)
                    for (int n = 0; n < nn; ++n) {
                        const int q = idx[((size_t)bi * p + pi) * nn + n];
                        double s = 0.0;
                        for (int k = 0; k < ks; ++k)
                            s += (double)inter_w[((((size_t)bi * p + pi) * na + a) * ks + k) * nn + n] *
                                 (double)dout[((((size_t)bi * c + ci) * ks + k) * p + pi) * na + a];
                        acc[(((size_t)bi * c + ci) * p_in + q) * na + a] += s;
                    }
    for (size_t i = 0; i < (size_t)b * c * p_in * na; ++i) dfeats[i] = (float)acc[i];
    free(acc);
}

/* ------------------------------------------------- intra feature grouping
 * vgtk/vgtk/so3conv/functional.py:221-268:
 *   out[b,c,k,p,a] = feats[b,c,p,intra_idx[a,k]]
 * feats [B,C,P,A], intra_idx [A,KN] -> out [B,C,KN,P,A] */
EPN_ORACLE_API void epn_oracle_intra_group_fwd(int b, int c, int p, int na, int kn,
                                               const float *feats, const int32_t *intra_idx,
                                               float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int k = 0; k < kn; ++k)
                for (int pi = 0; pi < p; ++pi)
                    for (int a = 0; a < na; ++a)
                        out[((((size_t)bi * c + ci) * kn + k) * p + pi) * na + a] =
                            feats[(((size_t)bi * c + ci) * p + pi) * na + intra_idx[a * kn + k]];
}

EPN_ORACLE_API void epn_oracle_intra_group_bwd(int b, int c, int p, int na, int kn,
                                               const float *dout, const int32_t *intra_idx,
                                               float *dfeats) {
    double *acc = (double *)calloc((size_t)b * c * p * na, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int k = 0; k < kn; ++k)
                for (int pi = 0; pi < p; ++pi)
                    for (int a = 0; a < na; ++a)
                        acc[(((size_t)bi * c + ci) * p + pi) * na + intra_idx[a * kn + k]] +=
                            dout[((((size_t)bi * c + ci) * kn + k) * p + pi) * na + a];
    for (size_t i = 0; i < (size_t)b * c * p * na; ++i) dfeats[i] = (float)acc[i];
    free(acc);
}

/* ------------------------------------------------------ BasicSO3Conv GEMM
 * vgtk/vgtk/so3conv/modules.py:48-55:
 *   out[b,o,p,a] = sum_{c,k} W[o, c*KS+k] * x[b,c,k,p,a]
 * x [B, C*KS, PA], W [CO, C*KS] -> out [B, CO, PA] */
EPN_ORACLE_API void epn_oracle_basic_conv(int b, int ck, int co, int pa, const float *x,
                                          const float *W, float *out) {
    double *row = (double *)malloc(sizeof(double) * (size_t)pa);
    for (int bi = 0; bi < b; ++bi)
        for (int o = 0; o < co; ++o) {
            for (int j = 0; j < pa; ++j) row[j] = 0.0;
            for (int q = 0; q < ck; ++q) {
                const double w = W[(size_t)o * ck + q];
                const float *xr = x + ((size_t)bi * ck + q) * pa;
                for (int j = 0; j < pa; ++j) row[j] += w * (double)xr[j];
            }
            float *orow = out + ((size_t)bi * co + o) * pa;
            for (int j = 0; j < pa; ++j) orow[j] = (float)row[j];
        }
    free(row);
}

/* ------------------------------------------------ zpconv op surface (dead on
 * the model path, named by the op surface).
 * vgtk/vgtk/cuda/zpconv_cuda_kernel.cu:32-73 (inter fwd), :76-116 (inter bwd),
 * :119-156 (intra fwd), :159-195 (intra bwd).  The reference accumulates with
 * fp32 atomicAdd in non-deterministic order; the oracle accumulates in double. */
EPN_ORACLE_API void epn_oracle_zp_inter_fwd(int b, int c, int nq, int np, int na, int ks, int ann,
                                            const int32_t *nbr, const float *w, const float *feats,
                                            float *out) {
    double *acc = (double *)calloc((size_t)b * c * ks * np * na, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int pn = 0; pn < np; ++pn)
            for (int an = 0; an < na; ++an)
                for (int k = 0; k < ks; ++k)
                    for (int ni = 0; ni < ann; ++ni) {
                        const size_t qi = ((((size_t)bi * np + pn) * na + an) * ks + k) * ann + ni;
                        const int qn = nbr[qi];
                        for (int ci = 0; ci < c; ++ci)
                            acc[((((size_t)bi * c + ci) * ks + k) * np + pn) * na + an] +=
                                (double)feats[(((size_t)bi * c + ci) * nq + qn) * na + an] * (double)w[qi];
                    }
    for (size_t i = 0; i < (size_t)b * c * ks * np * na; ++i) out[i] = (float)acc[i];
    free(acc);
}

EPN_ORACLE_API void epn_oracle_zp_inter_bwd(int b, int c, int nq, int np, int na, int ks, int ann,
                                            const int32_t *nbr, const float *w, const float *dout,
                                            float *dfeats) {
    double *acc = (double *)calloc((size_t)b * c * nq * na, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int pn = 0; pn < np; ++pn)
            for (int an = 0; an < na; ++an)
                for (int k = 0; k < ks; ++k)
                    for (int ni = 0; ni < ann; ++ni) {
                        const size_t qi = ((((size_t)bi * np + pn) * na + an) * ks + k) * ann + ni;
                        const int qn = nbr[qi];
                        for (int ci = 0; ci < c; ++ci)
                            acc[(((size_t)bi * c + ci) * nq + qn) * na + an] +=
                                (double)dout[((((size_t)bi * c + ci) * ks + k) * np + pn) * na + an] *
                                (double)w[qi];
                    }
    for (size_t i = 0; i < (size_t)b * c * nq * na; ++i) dfeats[i] = (float)acc[i];
    free(acc);
}

EPN_ORACLE_API void epn_oracle_zp_intra_fwd(int b, int c, int np, int na_in, int na_out, int ks,
                                            int ann, const int32_t *nbr, const float *w,
                                            const float *feats, float *out) {
    double *acc = (double *)calloc((size_t)b * c * ks * np * na_out, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int pn = 0; pn < np; ++pn)
            for (int an = 0; an < na_out; ++an)
                for (int k = 0; k < ks; ++k)
                    for (int ni = 0; ni < ann; ++ni) {
                        const int qan = nbr[an * ann + ni];
                        const double ww = w[((size_t)an * ks + k) * ann + ni];
                        for (int ci = 0; ci < c; ++ci)
                            acc[((((size_t)bi * c + ci) * ks + k) * np + pn) * na_out + an] +=
                                (double)feats[(((size_t)bi * c + ci) * np + pn) * na_in + qan] * ww;
                    }
    for (size_t i = 0; i < (size_t)b * c * ks * np * na_out; ++i) out[i] = (float)acc[i];
    free(acc);
}

EPN_ORACLE_API void epn_oracle_zp_intra_bwd(int b, int c, int np, int na_in, int na_out, int ks,
                                            int ann, const int32_t *nbr, const float *w,
                                            const float *dout, float *dfeats) {
    double *acc = (double *)calloc((size_t)b * c * np * na_in, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int pn = 0; pn < np; ++pn)
            for (int an = 0; an < na_out; ++an)
                for (int k = 0; k < ks; ++k)
                    for (int ni = 0; ni < ann; ++ni) {
                        const int qan = nbr[an * ann + ni];
                        const double ww = w[((size_t)an * ks + k) * ann + ni];
                        for (int ci = 0; ci < c; ++ci)
                            acc[(((size_t)bi * c + ci) * np + pn) * na_in + qan] +=
                                (double)dout[((((size_t)bi * c + ci) * ks + k) * np + pn) * na_out + an] * ww;
                    }
    for (size_t i = 0; i < (size_t)b * c * np * na_in; ++i) dfeats[i] = (float)acc[i];
    free(acc);
}
