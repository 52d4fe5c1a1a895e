// Fused convolutions of the SPConv hot path:
//   InterSO3Conv  = ball-neighbour gather + kernel weights + spatial contraction
//                   + channel GEMM        (vgtk/vgtk/so3conv/modules.py:157-174)
//   IntraSO3Conv  = anchor-permutation gather + channel GEMM   (modules.py:197-200)
//   BasicSO3Conv  = channel GEMM on an already grouped tensor  (modules.py:48-55)
//
// Schedule: the batch is cut into slabs of (clouds x point range) whose grouped
// tensor G[c*ks, points*na] fits the L2 (EPN_SLAB_BYTES); the grouping kernel
// writes the slab, the GEMM consumes it while it is still L2-resident, so
// neither inter_w nor the gathered (B,C,P,K,A) tensor nor the full grouped
// tensor ever reaches HBM.  The workspace is one slab (+ weight staging).
#include <stdlib.h>

#include "epn_internal.cuh"

namespace epn {

static size_t slab_budget_bytes() {
    static size_t v = 0;
    if (v == 0) {
        const char *e = getenv("EPN_SLAB_BYTES");
        v = e ? (size_t)strtoull(e, nullptr, 10) : (size_t)48 << 20;
        if (v < ((size_t)1 << 20)) v = (size_t)1 << 20;
    }
    return v;
}

struct SlabPlan {
    int bc;  // clouds per slab
    int pc;  // points per slab (pc == p when bc > 1)
    size_t bytes;
};

static SlabPlan plan_slabs(int b, int ck, int p, int na) {
    const size_t per_point = (size_t)ck * na * sizeof(float);
    const size_t per_cloud = per_point * p;
    const size_t budget = slab_budget_bytes();
    SlabPlan s;
    if (per_cloud <= budget) {
        s.bc = (int)(budget / per_cloud);
        if (s.bc > b) s.bc = b;
        s.pc = p;
    } else {
        s.bc = 1;
        s.pc = (int)(budget / per_point);
        if (s.pc < 1) s.pc = 1;
        if (s.pc > p) s.pc = p;
    }
    s.bytes = (size_t)s.bc * s.pc * per_point;
    return s;
}

static int pick_split_k(int M, int N, int K, int batch) {
    const long long tiles = (long long)cdiv(M, 64) * cdiv(N, 64) * batch;
    long long sk = (148LL * 6 + tiles - 1) / tiles;
    const long long maxk = K / 256 > 0 ? K / 256 : 1;
    if (sk > maxk) sk = maxk;
    if (sk < 1) sk = 1;
    if (sk * batch > 65535) sk = 65535 / batch;
    return (int)sk;
}

}  // namespace epn

using namespace epn;

#define EPN_CHECK_B(b) EPN_REQUIRE((b) <= 65535, EPN_ERR_SHAPE, "batch > 65535")
#define EPN_TRY(expr)            \
    do {                         \
        const int rc__ = (expr); \
        if (rc__ != 0) return rc__; \
    } while (0)

// ------------------------------------------------------------------ BasicSO3Conv
EPN_API int epn_basic_conv_fwd_f32(const float *x, const float *W, float *out, int b, int ck, int co, int pa,
                                   void *stream) {
    EPN_REQUIRE_PTR(x); EPN_REQUIRE_PTR(W); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(ck); EPN_REQUIRE_POS(co); EPN_REQUIRE_POS(pa); EPN_CHECK_B(b);
    GemmOperand A{W, 0, ck, 1};
    GemmOperand B{x, (long long)ck * pa, pa, 1};
    return launch_sgemm(A, B, out, (long long)co * pa, pa, co, pa, ck, b, 1, 0, as_stream(stream));
}

EPN_API int epn_basic_conv_bwd_f32(const float *dout, const float *x, const float *W, float *dx, float *dW,
                                   int b, int ck, int co, int pa, void *stream) {
    EPN_REQUIRE_PTR(dout);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(ck); EPN_REQUIRE_POS(co); EPN_REQUIRE_POS(pa); EPN_CHECK_B(b);
    cudaStream_t s = as_stream(stream);
    if (dx != nullptr) {
        EPN_REQUIRE_PTR(W);
        GemmOperand A{W, 0, 1, ck};  // W^T: (m=ck index, k=co index)
        GemmOperand B{dout, (long long)co * pa, pa, 1};
        EPN_TRY(launch_sgemm(A, B, dx, (long long)ck * pa, pa, ck, pa, co, b, 1, 0, s));
    }
    if (dW != nullptr) {
        EPN_REQUIRE_PTR(x);
        cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)co * ck, s);
        GemmOperand A{dout, (long long)co * pa, pa, 1};
        GemmOperand B{x, (long long)ck * pa, 1, pa};  // x^T: (k=pa index, n=ck index)
        EPN_TRY(launch_sgemm(A, B, dW, 0, ck, co, ck, pa, b, pick_split_k(co, ck, pa, b), 2, s));
    }
    return 0;
}

// ------------------------------------------------------------------ InterSO3Conv
EPN_API size_t epn_inter_so3conv_workspace_bytes(int b, int c_in, int c_out, int p_in, int p, int nn, int na,
                                                 int ks, int backward) {
    (void)c_out; (void)p_in; (void)nn; (void)backward;
    if (b <= 0 || c_in <= 0 || p <= 0 || na <= 0 || ks <= 0) return 0;
    return plan_slabs(b, c_in * ks, p, na).bytes;
}

EPN_API int epn_inter_so3conv_fwd_f32(const float *feats, const float *xyz, const float *centers,
                                      const int32_t *idx, const float *anchors, const float *kernels,
                                      float sigma, const float *W, float *out, void *workspace,
                                      size_t workspace_bytes, int b, int c_in, int c_out, int p_in, int p,
                                      int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(anchors);
    EPN_REQUIRE_PTR(kernels); EPN_REQUIRE_PTR(W); EPN_REQUIRE_PTR(out); EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p);
    EPN_REQUIRE_POS(nn); EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(na <= 64, EPN_ERR_SHAPE, "na > 64 anchors not supported");
    EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    EPN_REQUIRE(feats != nullptr || c_in == 1, EPN_ERR_NULL, "feats NULL requires c_in == 1");
    const int ck = c_in * ks;
    const SlabPlan sp = plan_slabs(b, ck, p, na);
    EPN_REQUIRE(workspace_bytes >= sp.bytes, EPN_ERR_WORKSPACE, "workspace smaller than epn_inter_so3conv_workspace_bytes()");
    EPN_REQUIRE(((uintptr_t)workspace & 255) == 0, EPN_ERR_ALIGN, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    float *G = static_cast<float *>(workspace);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na;
            InterGeom g{xyz + (size_t)b0 * 3 * p_in, centers + (size_t)b0 * 3 * p, anchors, kernels, sigma};
            EPN_TRY(launch_inter_group_fwd(feats ? feats + (size_t)b0 * c_in * p_in * na : nullptr,
                                           idx + (size_t)b0 * p * nn, nullptr, g, G, (long long)ck * cols, cols, p0, pc,
                                           bc, c_in, p_in, p, nn, na, ks, s));
            GemmOperand A{W, 0, ck, 1};
            GemmOperand B{G, (long long)ck * cols, cols, 1};
            EPN_TRY(launch_sgemm(A, B, out + ((size_t)b0 * c_out * p + p0) * na, (long long)c_out * p * na,
                                 (long long)p * na, c_out, (int)cols, ck, bc, 1, 0, s));
        }
    }
    return 0;
}

EPN_API int epn_inter_so3conv_bwd_f32(const float *dout, const float *feats, const float *xyz,
                                      const float *centers, const int32_t *idx, const float *anchors,
                                      const float *kernels, float sigma, const float *W, float *dfeats,
                                      float *dW, void *workspace, size_t workspace_bytes, int b, int c_in,
                                      int c_out, int p_in, int p, int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(idx);
    EPN_REQUIRE_PTR(anchors); EPN_REQUIRE_PTR(kernels); EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p);
    EPN_REQUIRE_POS(nn); EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(na <= 64, EPN_ERR_SHAPE, "na > 64 anchors not supported");
    EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    EPN_REQUIRE(feats != nullptr || c_in == 1, EPN_ERR_NULL, "feats NULL requires c_in == 1");
    if (dfeats != nullptr) EPN_REQUIRE_PTR(W);
    const int ck = c_in * ks;
    const SlabPlan sp = plan_slabs(b, ck, p, na);
    EPN_REQUIRE(workspace_bytes >= sp.bytes, EPN_ERR_WORKSPACE, "workspace smaller than epn_inter_so3conv_workspace_bytes()");
    EPN_REQUIRE(((uintptr_t)workspace & 255) == 0, EPN_ERR_ALIGN, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    float *G = static_cast<float *>(workspace);
    if (dfeats != nullptr) cudaMemsetAsync(dfeats, 0, sizeof(float) * (size_t)b * c_in * p_in * na, s);
    if (dW != nullptr) cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)c_out * ck, s);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na;
            InterGeom g{xyz + (size_t)b0 * 3 * p_in, centers + (size_t)b0 * 3 * p, anchors, kernels, sigma};
            const float *dout_slab = dout + ((size_t)b0 * c_out * p + p0) * na;
            const int32_t *idx_b = idx + (size_t)b0 * p * nn;
            if (dfeats != nullptr) {
                // dG = W^T . dout_slab, then scatter through the transposed spatial contraction
                GemmOperand A{W, 0, 1, ck};
                GemmOperand B{dout_slab, (long long)c_out * p * na, (long long)p * na, 1};
                EPN_TRY(launch_sgemm(A, B, G, (long long)ck * cols, cols, ck, (int)cols, c_out, bc, 1, 0, s));
                EPN_TRY(launch_inter_group_bwd(G, (long long)ck * cols, cols, p0, pc, idx_b, nullptr, g,
                                               dfeats + (size_t)b0 * c_in * p_in * na, bc, c_in, p_in, p, nn, na, ks, s));
            }
            if (dW != nullptr) {
                // dW += dout_slab . G^T with G recomputed (never saved by the forward)
                EPN_TRY(launch_inter_group_fwd(feats ? feats + (size_t)b0 * c_in * p_in * na : nullptr, idx_b, nullptr,
                                               g, G, (long long)ck * cols, cols, p0, pc, bc, c_in, p_in, p, nn, na, ks, s));
                GemmOperand A{dout_slab, (long long)c_out * p * na, (long long)p * na, 1};
                GemmOperand B{G, (long long)ck * cols, 1, cols};
                EPN_TRY(launch_sgemm(A, B, dW, 0, ck, c_out, ck, (int)cols, bc,
                                     pick_split_k(c_out, ck, (int)cols, bc), 2, s));
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------ IntraSO3Conv
EPN_API size_t epn_intra_so3conv_workspace_bytes(int b, int c_in, int c_out, int p, int na, int kn,
                                                 int backward) {
    (void)c_out; (void)backward;
    if (b <= 0 || c_in <= 0 || p <= 0 || na <= 0 || kn <= 0) return 0;
    return plan_slabs(b, c_in * kn, p, na).bytes;
}

EPN_API int epn_intra_so3conv_fwd_f32(const float *feats, const int32_t *intra_idx, const float *W, float *out,
                                      void *workspace, size_t workspace_bytes, int b, int c_in, int c_out,
                                      int p, int na, int kn, void *stream) {
    EPN_REQUIRE_PTR(feats); EPN_REQUIRE_PTR(intra_idx); EPN_REQUIRE_PTR(W); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(kn); EPN_CHECK_B(b);
    const int ck = c_in * kn;
    const SlabPlan sp = plan_slabs(b, ck, p, na);
    EPN_REQUIRE(workspace_bytes >= sp.bytes, EPN_ERR_WORKSPACE, "workspace smaller than epn_intra_so3conv_workspace_bytes()");
    EPN_REQUIRE(((uintptr_t)workspace & 255) == 0, EPN_ERR_ALIGN, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    float *G = static_cast<float *>(workspace);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na;
            EPN_TRY(launch_intra_group_fwd(feats + (size_t)b0 * c_in * p * na, intra_idx, G, (long long)ck * cols, cols,
                                           p0, pc, bc, c_in, p, na, kn, s));
            GemmOperand A{W, 0, ck, 1};
            GemmOperand B{G, (long long)ck * cols, cols, 1};
            EPN_TRY(launch_sgemm(A, B, out + ((size_t)b0 * c_out * p + p0) * na, (long long)c_out * p * na,
                                 (long long)p * na, c_out, (int)cols, ck, bc, 1, 0, s));
        }
    }
    return 0;
}

EPN_API int epn_intra_so3conv_bwd_f32(const float *dout, const float *feats, const int32_t *intra_idx,
                                      const float *W, float *dfeats, float *dW, void *workspace,
                                      size_t workspace_bytes, int b, int c_in, int c_out, int p, int na, int kn,
                                      void *stream) {
    EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(intra_idx); EPN_REQUIRE_PTR(workspace);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(kn); EPN_CHECK_B(b);
    if (dfeats != nullptr) EPN_REQUIRE_PTR(W);
    if (dW != nullptr) EPN_REQUIRE_PTR(feats);
    const int ck = c_in * kn;
    const SlabPlan sp = plan_slabs(b, ck, p, na);
    EPN_REQUIRE(workspace_bytes >= sp.bytes, EPN_ERR_WORKSPACE, "workspace smaller than epn_intra_so3conv_workspace_bytes()");
    EPN_REQUIRE(((uintptr_t)workspace & 255) == 0, EPN_ERR_ALIGN, "workspace must be 256-byte aligned");
    cudaStream_t s = as_stream(stream);
    float *G = static_cast<float *>(workspace);
    if (dW != nullptr) cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)c_out * ck, s);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na;
            const float *dout_slab = dout + ((size_t)b0 * c_out * p + p0) * na;
            if (dfeats != nullptr) {
                GemmOperand A{W, 0, 1, ck};
                GemmOperand B{dout_slab, (long long)c_out * p * na, (long long)p * na, 1};
                EPN_TRY(launch_sgemm(A, B, G, (long long)ck * cols, cols, ck, (int)cols, c_out, bc, 1, 0, s));
                EPN_TRY(launch_intra_group_bwd(G, (long long)ck * cols, cols, p0, pc, intra_idx,
                                               dfeats + (size_t)b0 * c_in * p * na, bc, c_in, p, na, kn, s));
            }
            if (dW != nullptr) {
                EPN_TRY(launch_intra_group_fwd(feats + (size_t)b0 * c_in * p * na, intra_idx, G, (long long)ck * cols,
                                               cols, p0, pc, bc, c_in, p, na, kn, s));
                GemmOperand A{dout_slab, (long long)c_out * p * na, (long long)p * na, 1};
                GemmOperand B{G, (long long)ck * cols, 1, cols};
                EPN_TRY(launch_sgemm(A, B, dW, 0, ck, c_out, ck, (int)cols, bc,
                                     pick_split_k(c_out, ck, (int)cols, bc), 2, s));
            }
        }
    }
    return 0;
}
