// Fused convolutions of the SPConv hot path:
//   InterSO3Conv  = ball-neighbour gather + kernel weights + spatial contraction
//                   + channel GEMM        (vgtk/vgtk/so3conv/modules.py:157-174)
//   IntraSO3Conv  = anchor-permutation gather + channel GEMM   (modules.py:197-200)
//   BasicSO3Conv  = channel GEMM on an already grouped tensor  (modules.py:48-55)
//
// Schedule: the batch is cut into slabs of (clouds x point range) whose grouped tensor
// G[c*ks, clouds*points*na] fits the L2 (EPN_SLAB_BYTES); the grouping kernel writes the slab, a
// conversion pass turns it into bf16 hi/lo split tiles, and the tcgen05 GEMM consumes them while
// they are L2-resident, so neither inter_w nor the gathered (B,C,P,K,A) tensor nor the full
// grouped tensor ever reaches HBM.  The workspace holds one slab + its operand tiles.
// EPN_GEMM=simt (or epn_set_gemm_backend(1)) routes the same schedule through the fp32 SIMT GEMM
// (cross-check path of the GPU tests).
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "epn_internal.cuh"
#include "epn_umma.cuh"

#ifndef EPN_FUSED_BWD_DEFAULT
#define EPN_FUSED_BWD_DEFAULT 2
#endif

namespace epn {

static std::atomic<int> g_backend{-1};  // 0 = tcgen05, 1 = SIMT

static int gemm_backend() {
    int v = g_backend.load();
    if (v < 0) {
        const char *e = getenv("EPN_GEMM");
        v = (e && strcmp(e, "simt") == 0) ? 1 : 0;
        g_backend.store(v);
    }
    return v;
}

static std::atomic<int> g_fused{-1};

static int fused_enabled() {  // default ON; EPN_FUSED=0 / epn_set_fused_inter(0): grouping kernel + GEMM kernel instead
    int v = g_fused.load();
    if (v < 0) {
        v = (getenv("EPN_FUSED") && strcmp(getenv("EPN_FUSED"), "0") == 0) ? 0 : 1;
        g_fused.store(v);
    }
    return v;
}

static std::atomic<int> g_fused_bwd{-1};

// fused data gradient of the inter conv (epn_inter_bwd_fused.cu); EPN_FUSED_BWD = 0 off, 1 rows of <= 16 slots, 2 also
// rows of 17..32 slots (two CTAs per point pair: neutral on the classification network, +6 % on the rotation network)
static int fused_bwd_enabled() {
    int v = g_fused_bwd.load();
    if (v < 0) {
        const char *e = getenv("EPN_FUSED_BWD");
        v = e ? (atoi(e) < 0 ? 0 : (atoi(e) > 2 ? 2 : atoi(e))) : EPN_FUSED_BWD_DEFAULT;
        g_fused_bwd.store(v);
    }
    return v;
}

// Inter grouping with the bf16 split in registers + permuted K order (epn_group_direct.cu): 0 or the K' mode the
// forward will use for this call (the backward is told through the grouped_layout word, never re-derives it).
static int inter_direct(const float *feats, int c_in, int nn, int na, int ks) {
    if (gemm_backend() != 0 || c_in == 1) return 0;
    return inter_group_direct_mode(feats, c_in, nn, na, ks);
}

// Shapes whose forward writes operand tiles directly (and can therefore keep them for the weight gradient).
static bool inter_tiles_available(int c_in, int nn, int na, int ks) {
    if (c_in == 1) return inter_group_occ_ok(c_in, nn, na, ks);
    return inter_group_tiles_ok(nn, na, ks) || inter_group_direct_mode(reinterpret_cast<const float *>(16), c_in, nn, na, ks) != 0;
}

static std::atomic<size_t> g_slab_bytes{0};

// Operand format of the forward GEMMs of the calling THREAD (epn_set_forward_operands): thread-local, so concurrent
// callers (the reference's nn.DataParallel threads) cannot disturb each other.  Only forwards that keep no operand
// tiles honour it (kept tiles feed the weight-gradient GEMM, whose gradient operand needs bf16's exponent range).
static thread_local int t_fwd_fmt = 0;
static int fwd_fmt() { return gemm_backend() == 0 ? t_fwd_fmt : 0; }

static size_t slab_budget_bytes() {
    size_t v = g_slab_bytes.load();
    if (v == 0) {
        const char *e = getenv("EPN_SLAB_BYTES");
        v = e ? (size_t)strtoull(e, nullptr, 10) : (size_t)1 << 30;
        if (v < ((size_t)64 << 10)) v = (size_t)64 << 10;
        g_slab_bytes.store(v);
    }
    return v;
}

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

struct SlabPlan {
    int bc;  // clouds per slab
    int pc;  // points per slab (pc == p when bc > 1)
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
    return s;
}

// Bytes of the forward operand tiles of ALL slabs of one conv call (what a training forward may keep for the
// weight gradient); 0 when the shape cannot take that route (SIMT backend, or a slab whose column count is
// not a multiple of the 128-row tile, where padded rows would hold garbage).
static size_t grouped_tiles_bytes(int b, int ck, int p, int na, const SlabPlan &sp) {
    if (gemm_backend() != 0) return 0;
    size_t total = 0;
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long n = (long long)bc * pc * na;
            if (n % 128 != 0) return 0;
            total += split_tiles_bytes(n, ck, 128);
        }
    }
    return total;
}
static size_t grouped_tiles_bytes(int b, int ck, int p, int na) {
    return grouped_tiles_bytes(b, ck, p, na, plan_slabs(b, ck, p, na));
}

// "Grouped layout" word: what a forward that kept its operand tiles tells its backward about them -- the order of
// the K dimension (0 = c*ks+k, 1/2 = the permuted orders of epn_group_direct.cu) and the slab plan -- so that the
// backward never re-derives either from the process-wide knobs (which may have changed in between).
//   bits 0-3 K' mode | bits 4-23 clouds per slab | bits 24-55 points per slab
static unsigned long long encode_layout(int kperm, const SlabPlan &sp) {
    return (unsigned long long)(kperm & 15) | ((unsigned long long)(sp.bc & 0xFFFFF) << 4) | ((unsigned long long)(unsigned)sp.pc << 24);
}
static bool decode_layout(unsigned long long w, int b, int p, int *kperm, SlabPlan *sp) {
    *kperm = (int)(w & 15);
    sp->bc = (int)((w >> 4) & 0xFFFFF);
    sp->pc = (int)((w >> 24) & 0xFFFFFFFFull);
    return *kperm <= 2 && sp->bc >= 1 && sp->bc <= b && sp->pc >= 1 && sp->pc <= p && (sp->bc == 1 || sp->pc == p);
}

// Workspace carve-up shared by the three convs.
struct Workspace {
    float *slab;      // fp32 grouped slab  [ck][n_slab]
    uint8_t *tilesA;  // 128-row operand tiles (activations)
    uint8_t *tilesB;  // dout tiles for dW
    uint8_t *tilesW;  // weights  (rows = c_out, K = ck)
    uint8_t *tilesWT; // weights^T (rows = ck, K = c_out)
    size_t total;
};

static Workspace carve(void *base, int ck, int c_out, long long n_slab, bool need_slab) {
    Workspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t *p = base ? static_cast<uint8_t *>(base) + off : nullptr;
        off += up256(bytes);
        return p;
    };
    w.slab = reinterpret_cast<float *>(take(need_slab ? (size_t)ck * n_slab * sizeof(float) : 0));
    size_t a = split_tiles_bytes(n_slab, ck, 128);
    const size_t a2 = split_tiles_bytes(n_slab, c_out, 256) + split_tiles_bytes(256, c_out, 256),
                 a3 = split_tiles_bytes(ck, n_slab, 128);
    if (a2 > a) a = a2;
    if (a3 > a) a = a3;
    w.tilesA = take(a);
    w.tilesB = take(split_tiles_bytes(c_out, n_slab, umma_trb_for(c_out)));
    w.tilesW = take(split_tiles_bytes(c_out, ck, umma_trb_for(c_out)));
    w.tilesWT = take(split_tiles_bytes(ck, c_out, 128));
    w.total = off;
    return w;
}

// A [rows_k x (bc x cols)] fp32 matrix view: element (k, z, j) = ptr[z*stride_z + k*stride_k + j]
struct ColsView {
    float *ptr;
    long long stride_z, stride_k;
};

constexpr long long HUGE_Z = 1LL << 60;

static int prep_weights(const float *W, int c_out, int ck, const Workspace &ws, bool fwd, bool transposed, cudaStream_t s,
                        int kperm = 0, int step_layout = 0, int fmt = 0) {   // fmt: format of the FORWARD weight tiles
    if (gemm_backend() != 0) return 0;
    if (fwd && kperm) {
        int rc = launch_inter_w_tiles_kperm(W, ws.tilesW, c_out, ck, umma_trb_for(c_out), kperm, step_layout, s, fmt);
        if (rc) return rc;
    } else if (fwd) {
        SplitSrc src{W, HUGE_Z, 0, ck, HUGE_Z, 0, 1};
        int rc = launch_split_tiles(src, ws.tilesW, c_out, ck, umma_trb_for(c_out), s, fmt,
                                    fmt == umma::FMT_F16 ? umma::F16_W_SCALE : 1.0f);
        if (rc) return rc;
    }
    if (transposed) {
        SplitSrc src{W, HUGE_Z, 0, 1, HUGE_Z, 0, ck};
        int rc = launch_split_tiles(src, ws.tilesWT, ck, c_out, 128, s);
        if (rc) return rc;
    }
    return 0;
}

static int pick_split_k(int M, int N, long long K, int batch, int tile_m, int tile_n, int k_unit) {
    const long long tiles = (long long)cdiv(M, tile_m) * cdiv(N, tile_n) * batch;
    long long sk = (148LL * 4 + tiles - 1) / tiles;
    const long long maxk = K / k_unit > 0 ? K / k_unit : 1;
    if (sk > maxk) sk = maxk;
    if (sk < 1) sk = 1;
    if (sk * batch > 65535) sk = 65535 / batch;
    return (int)sk;
}

// out(c_out) = W . G with the activation operand already in ws.tilesA (rows = (z,j) columns, K = ck)
static int gemm_fwd_tiles(const void *tilesA, int c_out, int ck, int bc, long long cols, ColsView out,
                          const Workspace &ws, cudaStream_t s, int fmt = 0) {
    GemmEpilogue ep{out.ptr, cols, out.stride_z, 1, out.stride_k, false};
    return launch_umma_gemm(tilesA, ws.tilesW, (int)(bc * cols), c_out, ck, umma_trb_for(c_out), ep, 1, s, fmt);
}

// dW(c_out x ck) += dout . G with G = the forward operand tiles of this slab (rows = (z,j) columns, K = ck)
static int gemm_dw_grouped(const void *grouped, ColsView dout, int c_out, int ck, int bc, long long cols, float *dW,
                           const Workspace &ws, cudaStream_t s, int kperm = 0) {
    const long long n = bc * cols;
    const int trb = umma_trb_for(c_out);
    SplitSrc sb{dout.ptr, HUGE_Z, 0, dout.stride_k, cols, dout.stride_z, 1};
    int rc = launch_split_tiles(sb, ws.tilesB, c_out, n, trb, s);
    if (rc) return rc;
    return launch_umma_dw(grouped, ws.tilesB, ck, c_out, n, trb, dW, kperm, s);
}

// dW(c_out x ck) += dout . G^T with G^T already in ws.tilesA (rows = ck, K = (z,j) columns)
static int gemm_dw_tiles(ColsView dout, int c_out, int ck, int bc, long long cols, float *dW, const Workspace &ws,
                         cudaStream_t s) {
    const long long n = bc * cols;
    const int trb = umma_trb_for(c_out);
    SplitSrc sb{dout.ptr, HUGE_Z, 0, dout.stride_k, cols, dout.stride_z, 1};
    int rc = launch_split_tiles(sb, ws.tilesB, c_out, n, trb, s);
    if (rc) return rc;
    GemmEpilogue ep{dW, HUGE_Z, 0, 1, ck, true};
    const long long tiles = (long long)cdiv(ck, 128) * cdiv(c_out, trb);
    long long sk = (148LL * 3 + tiles - 1) / tiles;
    const long long maxk = n / 512 > 0 ? n / 512 : 1;
    if (sk > maxk) sk = maxk;
    return launch_umma_gemm(ws.tilesA, ws.tilesB, ck, c_out, n, trb, ep, (int)sk, s);
}

// out(c_out) = W . in(ck)
static int gemm_fwd(const float *W, int c_out, int ck, ColsView in, int bc, long long cols, ColsView out,
                    const Workspace &ws, cudaStream_t s, int fmt = 0) {
    if (gemm_backend() != 0) {
        GemmOperand A{W, 0, ck, 1};
        GemmOperand B{in.ptr, in.stride_z, in.stride_k, 1};
        return launch_sgemm(A, B, out.ptr, out.stride_z, out.stride_k, c_out, (int)cols, ck, bc, 1, 0, s);
    }
    const long long n = bc * cols;
    SplitSrc src{in.ptr, cols, in.stride_z, 1, HUGE_Z, 0, in.stride_k};
    int rc = launch_split_tiles(src, ws.tilesA, n, ck, 128, s, fmt);
    if (rc) return rc;
    return gemm_fwd_tiles(ws.tilesA, c_out, ck, bc, cols, out, ws, s, fmt);
}

// din(ck) = W^T . dout(c_out)
static int gemm_dx(const float *W, int c_out, int ck, ColsView dout, int bc, long long cols, ColsView din,
                   const Workspace &ws, cudaStream_t s) {
    if (gemm_backend() != 0) {
        GemmOperand A{W, 0, 1, ck};
        GemmOperand B{dout.ptr, dout.stride_z, dout.stride_k, 1};
        return launch_sgemm(A, B, din.ptr, din.stride_z, din.stride_k, ck, (int)cols, c_out, bc, 1, 0, s);
    }
    // orientation: D[rows = (c,k), cols = (z,j)] so that an epilogue thread owns one (c,k) row and writes
    // runs of consecutive columns as 16-byte vectors (din is column-contiguous)
    const long long n = bc * cols;
    const int trn = umma_trb_for((int)(n < 256 ? n : 256));
    SplitSrc src{dout.ptr, cols, dout.stride_z, 1, HUGE_Z, 0, dout.stride_k};
    int rc = launch_split_tiles(src, ws.tilesA, n, c_out, trn, s);
    if (rc) return rc;
    GemmEpilogue ep{din.ptr, HUGE_Z, 0, din.stride_k, 1, false};
    ep.cols_per_z = cols;
    ep.stride_cz = din.stride_z;
    return launch_umma_gemm(ws.tilesWT, ws.tilesA, ck, (int)n, c_out, trn, ep, 1, s);
}

// dW(c_out x ck) += dout(c_out) . in(ck)^T      (dW pre-zeroed by the caller)
static int gemm_dw(ColsView dout, ColsView in, int c_out, int ck, int bc, long long cols, float *dW,
                   const Workspace &ws, cudaStream_t s) {
    if (gemm_backend() != 0) {
        GemmOperand A{dout.ptr, dout.stride_z, dout.stride_k, 1};
        GemmOperand B{in.ptr, in.stride_z, 1, in.stride_k};
        return launch_sgemm(A, B, dW, 0, ck, c_out, ck, (int)cols, bc, pick_split_k(c_out, ck, cols, bc, 64, 64, 256),
                            2, s);
    }
    const long long n = bc * cols;
    SplitSrc sa{in.ptr, HUGE_Z, 0, in.stride_k, cols, in.stride_z, 1};
    int rc = launch_split_tiles(sa, ws.tilesA, ck, n, 128, s);
    if (rc) return rc;
    return gemm_dw_tiles(dout, c_out, ck, bc, cols, dW, ws, s);
}

}  // namespace epn

using namespace epn;

#define EPN_CHECK_B(b) EPN_REQUIRE((b) <= 65535, EPN_ERR_SHAPE, "batch > 65535")
#define EPN_TRY(expr)            \
    do {                         \
        const int rc__ = (expr); \
        if (rc__ != 0) return rc__; \
    } while (0)
#define EPN_CHECK_WS(need)                                                                                      \
    do {                                                                                                        \
        EPN_REQUIRE_PTR(workspace);                                                                             \
        EPN_REQUIRE(workspace_bytes >= (need), EPN_ERR_WORKSPACE, "workspace smaller than *_workspace_bytes()"); \
        EPN_REQUIRE(((uintptr_t)workspace & 255) == 0, EPN_ERR_ALIGN, "workspace must be 256-byte aligned");      \
    } while (0)

#define EPN_CHECK_GROUPED(need)                                                                               \
    do {                                                                                                      \
        const size_t need__ = (need);                                                                         \
        EPN_REQUIRE(need__ != 0, EPN_ERR_SHAPE, "grouped tiles are not available for this shape / backend");  \
        EPN_REQUIRE(grouped_bytes == need__, EPN_ERR_WORKSPACE, "grouped_bytes != *_grouped_bytes()");        \
        EPN_REQUIRE(((uintptr_t)grouped & 255) == 0, EPN_ERR_ALIGN, "grouped must be 256-byte aligned");      \
    } while (0)

EPN_API void epn_set_slab_bytes(size_t bytes) { g_slab_bytes.store(bytes < ((size_t)64 << 10) ? ((size_t)64 << 10) : bytes); }
EPN_API size_t epn_get_slab_bytes(void) { return slab_budget_bytes(); }
EPN_API void epn_set_gemm_backend(int simt) { g_backend.store(simt ? 1 : 0); }
EPN_API int epn_get_gemm_backend(void) { return gemm_backend(); }
EPN_API void epn_set_fused_inter(int on) { g_fused.store(on ? 1 : 0); }
EPN_API void epn_set_fused_inter_bwd(int on) { g_fused_bwd.store(on < 0 ? 0 : (on > 2 ? 2 : on)); }
EPN_API int epn_get_fused_inter_bwd(void) { return fused_bwd_enabled(); }
EPN_API void epn_set_forward_operands(int fmt) { t_fwd_fmt = fmt == umma::FMT_F16 ? umma::FMT_F16 : umma::FMT_BF16; }
EPN_API int epn_get_forward_operands(void) { return t_fwd_fmt; }
EPN_API int epn_get_fused_inter(void) { return fused_enabled(); }

// ------------------------------------------------------------------ BasicSO3Conv
EPN_API size_t epn_basic_conv_workspace_bytes(int b, int ck, int co, int pa) {
    if (b <= 0 || ck <= 0 || co <= 0 || pa <= 0) return 0;
    const SlabPlan sp = plan_slabs(b, ck, pa, 1);
    return carve(nullptr, ck, co, (long long)sp.bc * sp.pc, false).total;
}

EPN_API int epn_basic_conv_fwd_f32(const float *x, const float *W, float *out, void *workspace,
                                   size_t workspace_bytes, int b, int ck, int co, int pa, void *stream) {
    EPN_REQUIRE_PTR(x); EPN_REQUIRE_PTR(W); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(ck); EPN_REQUIRE_POS(co); EPN_REQUIRE_POS(pa); EPN_CHECK_B(b);
    const SlabPlan sp = plan_slabs(b, ck, pa, 1);
    const Workspace ws = carve(workspace, ck, co, (long long)sp.bc * sp.pc, false);
    EPN_CHECK_WS(ws.total);
    cudaStream_t s = as_stream(stream);
    const int fmt = fwd_fmt();
    EPN_TRY(prep_weights(W, co, ck, ws, true, false, s, 0, 0, fmt));
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int j0 = 0; j0 < pa; j0 += sp.pc) {
            const long long cols = pa - j0 < sp.pc ? pa - j0 : sp.pc;
            ColsView in{const_cast<float *>(x) + (size_t)b0 * ck * pa + j0, (long long)ck * pa, pa};
            ColsView o{out + (size_t)b0 * co * pa + j0, (long long)co * pa, pa};
            EPN_TRY(gemm_fwd(W, co, ck, in, bc, cols, o, ws, s, fmt));
        }
    }
    return 0;
}

EPN_API int epn_basic_conv_bwd_f32(const float *dout, const float *x, const float *W, float *dx, float *dW,
                                   void *workspace, size_t workspace_bytes, int b, int ck, int co, int pa,
                                   void *stream) {
    EPN_REQUIRE_PTR(dout);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(ck); EPN_REQUIRE_POS(co); EPN_REQUIRE_POS(pa); EPN_CHECK_B(b);
    if (dx != nullptr) EPN_REQUIRE_PTR(W);
    if (dW != nullptr) EPN_REQUIRE_PTR(x);
    const SlabPlan sp = plan_slabs(b, ck, pa, 1);
    const Workspace ws = carve(workspace, ck, co, (long long)sp.bc * sp.pc, false);
    EPN_CHECK_WS(ws.total);
    cudaStream_t s = as_stream(stream);
    if (dx != nullptr) EPN_TRY(prep_weights(W, co, ck, ws, false, true, s));
    if (dW != nullptr) cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)co * ck, s);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int j0 = 0; j0 < pa; j0 += sp.pc) {
            const long long cols = pa - j0 < sp.pc ? pa - j0 : sp.pc;
            ColsView d{const_cast<float *>(dout) + (size_t)b0 * co * pa + j0, (long long)co * pa, pa};
            if (dx != nullptr) {
                ColsView di{dx + (size_t)b0 * ck * pa + j0, (long long)ck * pa, pa};
                EPN_TRY(gemm_dx(W, co, ck, d, bc, cols, di, ws, s));
            }
            if (dW != nullptr) {
                ColsView in{const_cast<float *>(x) + (size_t)b0 * ck * pa + j0, (long long)ck * pa, pa};
                EPN_TRY(gemm_dw(d, in, co, ck, bc, cols, dW, ws, s));
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------ InterSO3Conv
EPN_API size_t epn_inter_so3conv_workspace_bytes(int b, int c_in, int c_out, int p_in, int p, int nn, int na,
                                                 int ks, int backward) {
    (void)p_in; (void)nn; (void)backward;
    if (b <= 0 || c_in <= 0 || c_out <= 0 || p <= 0 || na <= 0 || ks <= 0) return 0;
    const SlabPlan sp = plan_slabs(b, c_in * ks, p, na);
    return carve(nullptr, c_in * ks, c_out, (long long)sp.bc * sp.pc * na, true).total;
}

EPN_API size_t epn_inter_so3conv_grouped_bytes(int b, int c_in, int p, int nn, int na, int ks) {
    if (b <= 0 || c_in <= 0 || p <= 0 || nn <= 0 || na <= 0 || ks <= 0 || !inter_tiles_available(c_in, nn, na, ks)) return 0;
    return grouped_tiles_bytes(b, c_in * ks, p, na);
}

EPN_API int epn_inter_so3conv_fwd_f32(const float *feats, const float *xyz, const float *centers,
                                      const int32_t *idx, const float *anchors, const float *kernels,
                                      float sigma, const float *W, float *out, void *workspace,
                                      size_t workspace_bytes, void *grouped, size_t grouped_bytes,
                                      unsigned long long *grouped_layout, int b, int c_in, int c_out, int p_in, int p,
                                      int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(anchors);
    EPN_REQUIRE_PTR(kernels); EPN_REQUIRE_PTR(W); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p);
    EPN_REQUIRE_POS(nn); EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(na <= 64, EPN_ERR_SHAPE, "na > 64 anchors not supported");
    EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    EPN_REQUIRE(feats != nullptr || c_in == 1, EPN_ERR_NULL, "feats NULL requires c_in == 1");
    const int ck = c_in * ks;
    const SlabPlan sp = plan_slabs(b, ck, p, na);
    const Workspace ws = carve(workspace, ck, c_out, (long long)sp.bc * sp.pc * na, true);
    EPN_CHECK_WS(ws.total);
    if (grouped != nullptr) EPN_CHECK_GROUPED(epn_inter_so3conv_grouped_bytes(b, c_in, p, nn, na, ks));
    uint8_t *keep = static_cast<uint8_t *>(grouped);
    cudaStream_t s = as_stream(stream);
    // fused route: ONE kernel over every cloud, G stays in shared memory (its K' mode numbers are the direct kernels')
    int fused = 0;
    if (gemm_backend() == 0 && fused_enabled() && feats != nullptr && c_in > 1 && (sp.pc == p || grouped == nullptr))
        fused = inter_fused_mode(c_in, c_out, p, nn, na, ks, grouped != nullptr);
    const int kperm = fused ? fused : inter_direct(feats, c_in, nn, na, ks);
    if (grouped != nullptr) {
        EPN_REQUIRE_PTR(grouped_layout);
        *grouped_layout = encode_layout(kperm, sp);
    }
    // fp16 operands (epn_set_forward_operands): inference forwards on the fused route and on the one-channel route
    const bool occ_route = c_in == 1 && gemm_backend() == 0 && inter_group_occ_ok(c_in, nn, na, ks);
    const int fmt = (grouped == nullptr && (fused || occ_route)) ? fwd_fmt() : 0;
    EPN_TRY(prep_weights(W, c_out, ck, ws, true, false, s, kperm, fused ? 1 : 0, fmt));
    if (fused) {
        // kept tiles (training) keep the slab layout the weight-gradient pass expects
        InterGeom g{xyz, centers, anchors, kernels, sigma};
        const long long cols = (long long)p * na;
        const int rc = launch_inter_fused(feats, idx, g, ws.tilesW, out, (long long)c_out * p * na, (long long)p * na, keep,
                                          cdiv(ck, 32), cols, sp.bc, split_tiles_bytes((long long)sp.bc * cols, ck, 128), 0, p,
                                          b, c_in, c_out, p_in, p, nn, na, ks, s, fmt);
        return rc == 1 ? EPN_ERR_SHAPE : rc;
    }
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na, n_slab = bc * cols;
            InterGeom g{xyz + (size_t)b0 * 3 * p_in, centers + (size_t)b0 * 3 * p, anchors, kernels, sigma};
            const float *feats_b = feats ? feats + (size_t)b0 * c_in * p_in * na : nullptr;
            ColsView o{out + ((size_t)b0 * c_out * p + p0) * na, (long long)c_out * p * na, (long long)p * na};
            void *tiles = keep ? keep : ws.tilesA;  // kept tiles: every slab has its own region
            if (keep) keep += split_tiles_bytes(n_slab, ck, 128);
            int direct = 1;  // 0: the grouping kernel wrote the operand tiles itself
            if (occ_route) {
                direct = launch_inter_group_occ(feats_b, idx + (size_t)b0 * p * nn, g, tiles, cols, p0, pc, bc, p_in, p, nn, na,
                                                ks, s, fmt);
                if (direct != 0) return direct == 1 ? EPN_ERR_SHAPE : direct;
            } else if (kperm) {
                direct = launch_inter_group_direct(feats_b, idx + (size_t)b0 * p * nn, g, tiles, cdiv(ck, 32), cols, p0, pc, bc,
                                                   c_in, p_in, p, nn, na, ks, s);
                if (direct != 0) return direct == 1 ? EPN_ERR_SHAPE : direct;
            } else if (gemm_backend() == 0) {
                direct = launch_inter_group_tiles(feats_b, idx + (size_t)b0 * p * nn, g, tiles, cdiv(ck, 32), 0, cols, 0,
                                                  p0, pc, bc, c_in, p_in, p, nn, na, ks, s);
                if (direct != 0 && direct != 1) return direct;
            }
            EPN_REQUIRE(direct == 0 || grouped == nullptr, EPN_ERR_SHAPE, "grouped tiles requested for an unsupported shape");
            if (direct == 0) {
                EPN_TRY(gemm_fwd_tiles(tiles, c_out, ck, bc, cols, o, ws, s, occ_route ? fmt : 0));
            } else {
                EPN_TRY(launch_inter_group_fwd(feats_b, idx + (size_t)b0 * p * nn, nullptr, g, ws.slab, cols, n_slab, p0, pc,
                                               bc, c_in, p_in, p, nn, na, ks, s));
                ColsView in{ws.slab, cols, n_slab};
                EPN_TRY(gemm_fwd(W, c_out, ck, in, bc, cols, o, ws, s));
            }
        }
    }
    return 0;
}

EPN_API int epn_inter_so3conv_bwd_f32(const float *dout, const float *feats, const float *xyz,
                                      const float *centers, const int32_t *idx, const float *anchors,
                                      const float *kernels, float sigma, const float *W, float *dfeats,
                                      float *dW, void *workspace, size_t workspace_bytes, const void *grouped,
                                      size_t grouped_bytes, unsigned long long grouped_layout, int b, int c_in,
                                      int c_out, int p_in, int p, int nn, int na, int ks, void *stream) {
    EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(centers); EPN_REQUIRE_PTR(idx);
    EPN_REQUIRE_PTR(anchors); EPN_REQUIRE_PTR(kernels);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p_in); EPN_REQUIRE_POS(p);
    EPN_REQUIRE_POS(nn); EPN_REQUIRE_POS(na); EPN_REQUIRE_POS(ks); EPN_CHECK_B(b);
    EPN_REQUIRE(na <= 64, EPN_ERR_SHAPE, "na > 64 anchors not supported");
    EPN_REQUIRE(sigma > 0.f, EPN_ERR_SHAPE, "sigma must be > 0");
    EPN_REQUIRE(feats != nullptr || c_in == 1, EPN_ERR_NULL, "feats NULL requires c_in == 1");
    if (dfeats != nullptr) EPN_REQUIRE_PTR(W);
    const int ck = c_in * ks;
    SlabPlan sp = plan_slabs(b, ck, p, na);
    int kperm_kept = 0;
    if (grouped != nullptr) {  // slab plan and K order are the forward's, not whatever the knobs say now
        EPN_REQUIRE(decode_layout(grouped_layout, b, p, &kperm_kept, &sp), EPN_ERR_SHAPE,
                    "grouped_layout is not a value returned by the forward");
        EPN_CHECK_GROUPED(inter_tiles_available(c_in, nn, na, ks) ? grouped_tiles_bytes(b, ck, p, na, sp) : 0);
    }
    const Workspace ws = carve(workspace, ck, c_out, (long long)sp.bc * sp.pc * na, true);
    EPN_CHECK_WS(ws.total);
    const uint8_t *keep = static_cast<const uint8_t *>(grouped);
    cudaStream_t s = as_stream(stream);
    bool fused_bwd = false;   // data gradient by the fused kernel (all clouds in one launch): no dG slab, no scatter kernel
    if (dfeats != nullptr) {
        cudaMemsetAsync(dfeats, 0, sizeof(float) * (size_t)b * c_in * p_in * na, s);
        if (gemm_backend() == 0 && fused_bwd_enabled() >= (nn > 16 ? 2 : 1) && feats != nullptr && c_in > 1 &&
            inter_bwd_fused_ok(c_in, c_out, p, nn, na, ks)) {
            InterGeom ga{xyz, centers, anchors, kernels, sigma};
            const int rc = launch_inter_bwd_fused(dout, (long long)c_out * p * na, (long long)p * na, idx, ga, W, ws.tilesWT, dfeats, 0,
                                                  p, b, c_in, c_out, p_in, p, nn, na, ks, s);
            if (rc != 0 && rc != 1) return rc;
            fused_bwd = rc == 0;
        }
        if (!fused_bwd) EPN_TRY(prep_weights(W, c_out, ck, ws, false, true, s));
    }
    if (dW != nullptr) cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)c_out * ck, s);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na, n_slab = bc * cols;
            InterGeom g{xyz + (size_t)b0 * 3 * p_in, centers + (size_t)b0 * 3 * p, anchors, kernels, sigma};
            ColsView d{const_cast<float *>(dout) + ((size_t)b0 * c_out * p + p0) * na, (long long)c_out * p * na,
                       (long long)p * na};
            ColsView slab{ws.slab, cols, n_slab};
            const int32_t *idx_b = idx + (size_t)b0 * p * nn;
            if (dfeats != nullptr && !fused_bwd) {
                // dG = W^T . dout, then scatter through the transposed spatial contraction
                EPN_TRY(gemm_dx(W, c_out, ck, d, bc, cols, slab, ws, s));
                int sc = launch_inter_scatter(ws.slab, cols, n_slab, idx_b, g, dfeats + (size_t)b0 * c_in * p_in * na, p0, pc,
                                              bc, c_in, p_in, p, nn, na, ks, s);
                if (sc == 1)
                    sc = launch_inter_group_bwd(ws.slab, cols, n_slab, p0, pc, idx_b, nullptr, g,
                                                dfeats + (size_t)b0 * c_in * p_in * na, bc, c_in, p_in, p, nn, na, ks, s);
                if (sc != 0) return sc;
            }
            const uint8_t *kept = keep;
            if (keep) keep += split_tiles_bytes(n_slab, ck, 128);
            if (dW != nullptr && kept != nullptr) {
                // dW += dout . G with G = the operand tiles the forward kept (K possibly in the permuted order)
                EPN_TRY(gemm_dw_grouped(kept, d, c_out, ck, bc, cols, dW, ws, s, kperm_kept));
            } else if (dW != nullptr) {
                // dW += dout . G^T with G recomputed
                const float *feats_b = feats ? feats + (size_t)b0 * c_in * p_in * na : nullptr;
                int direct = 1;
                if (gemm_backend() == 0 && n_slab % 32 == 0) {
                    direct = launch_inter_group_tiles(feats_b, idx_b, g, ws.tilesA, (int)(n_slab / 32), cdiv(ck, 128) * 128,
                                                      cols, 1, p0, pc, bc, c_in, p_in, p, nn, na, ks, s);
                    if (direct != 0 && direct != 1) return direct;
                }
                if (direct == 0) {
                    EPN_TRY(gemm_dw_tiles(d, c_out, ck, bc, cols, dW, ws, s));
                } else {
                    EPN_TRY(launch_inter_group_fwd(feats_b, idx_b, nullptr, g, ws.slab, cols, n_slab, p0, pc, bc, c_in, p_in, p,
                                                   nn, na, ks, s));
                    EPN_TRY(gemm_dw(d, slab, c_out, ck, bc, cols, dW, ws, s));
                }
            }
        }
    }
    return 0;
}

// ------------------------------------------------------------------ IntraSO3Conv
EPN_API size_t epn_intra_so3conv_workspace_bytes(int b, int c_in, int c_out, int p, int na, int kn,
                                                 int backward) {
    (void)backward;
    if (b <= 0 || c_in <= 0 || c_out <= 0 || p <= 0 || na <= 0 || kn <= 0) return 0;
    const SlabPlan sp = plan_slabs(b, c_in * kn, p, na);
    return carve(nullptr, c_in * kn, c_out, (long long)sp.bc * sp.pc * na, true).total;
}

EPN_API size_t epn_intra_so3conv_grouped_bytes(int b, int c_in, int p, int na, int kn) {
    if (b <= 0 || c_in <= 0 || p <= 0 || na <= 0 || kn <= 0 || !intra_group_tiles_ok(na, kn)) return 0;
    return grouped_tiles_bytes(b, c_in * kn, p, na);
}

// pro != NULL: feats is the RAW output of the preceding conv and the normalisation + leaky_relu that follows it is applied
// while the operand tiles are built (tile routes only: anything else returns EPN_ERR_SHAPE and the caller runs the
// norm kernel itself)
static int intra_fwd_impl(const float *feats, const int32_t *intra_idx, const float *W, float *out,
                          void *workspace, size_t workspace_bytes, void *grouped, size_t grouped_bytes,
                          unsigned long long *grouped_layout, int b, int c_in, int c_out, int p, int na,
                          int kn, void *stream, const NormPrologue *pro) {
    EPN_REQUIRE_PTR(feats); EPN_REQUIRE_PTR(intra_idx); EPN_REQUIRE_PTR(W); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(kn); EPN_CHECK_B(b);
    const int ck = c_in * kn;
    const SlabPlan sp = plan_slabs(b, ck, p, na);
    const Workspace ws = carve(workspace, ck, c_out, (long long)sp.bc * sp.pc * na, true);
    EPN_CHECK_WS(ws.total);
    if (grouped != nullptr) {
        EPN_CHECK_GROUPED(epn_intra_so3conv_grouped_bytes(b, c_in, p, na, kn));
        EPN_REQUIRE_PTR(grouped_layout);
        *grouped_layout = encode_layout(0, sp);
    }
    uint8_t *keep = static_cast<uint8_t *>(grouped);
    cudaStream_t s = as_stream(stream);
    EPN_TRY(prep_weights(W, c_out, ck, ws, true, false, s));
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na, n_slab = bc * cols;
            ColsView o{out + ((size_t)b0 * c_out * p + p0) * na, (long long)c_out * p * na, (long long)p * na};
            if (grouped == nullptr && gemm_backend() == 0 && pc == p && intra_dx_fused_ok(n_slab, p, na, kn) &&
                intra_dx_wt_bytes(c_out, c_in) <= (size_t)ck * n_slab * sizeof(float)) {
                // inference (no operand tiles to keep for the weight gradient): Y_k = W_k . feats on the tensor cores,
                // out = sum_k Y_k permuted, reduced in shared memory -- the 12x larger grouped tensor never exists
                const float *fz = feats + (size_t)b0 * c_in * p * na;
                NormPrologue pz;
                if (pro != nullptr) {
                    pz = *pro;
                    if (pz.mode == 0) pz.stats += (size_t)b0 * c_in;   // instance statistics of the slab's first cloud
                }
                const int rc = launch_umma_intra_dx(fz, (long long)c_in * p * na, (long long)p * na, W, intra_idx,
                                                    out + (size_t)b0 * c_out * p * na, ws.slab, ws.tilesA, bc, c_in, c_out, p, 1, s,
                                                    fwd_fmt(), pro ? &pz : nullptr);
                if (rc == 0) continue;
                if (rc != 1) return rc;
            }
            void *tiles = keep ? keep : ws.tilesA;
            if (keep) keep += split_tiles_bytes(n_slab, ck, 128);
            int direct = 1;
            if (gemm_backend() == 0) {
                NormPrologue pz;
                if (pro != nullptr) {
                    pz = *pro;
                    if (pz.mode == 0) pz.stats += (size_t)b0 * c_in;
                }
                direct = launch_intra_group_tiles(feats + (size_t)b0 * c_in * p * na, intra_idx, tiles, 0, p0, pc, bc, c_in,
                                                  p, na, kn, s, pro ? &pz : nullptr);
                if (direct != 0 && direct != 1) return direct;
            }
            EPN_REQUIRE(direct == 0 || pro == nullptr, EPN_ERR_SHAPE, "fused norm prologue: shape not covered by the tile routes");
            EPN_REQUIRE(direct == 0 || grouped == nullptr, EPN_ERR_SHAPE, "grouped tiles requested for an unsupported shape");
            if (direct == 0) {
                EPN_TRY(gemm_fwd_tiles(tiles, c_out, ck, bc, cols, o, ws, s));
            } else {
                EPN_TRY(launch_intra_group_fwd(feats + (size_t)b0 * c_in * p * na, intra_idx, ws.slab, cols, n_slab, p0, pc,
                                               bc, c_in, p, na, kn, s));
                ColsView in{ws.slab, cols, n_slab};
                EPN_TRY(gemm_fwd(W, c_out, ck, in, bc, cols, o, ws, s));
            }
        }
    }
    return 0;
}

EPN_API int epn_intra_so3conv_fwd_f32(const float *feats, const int32_t *intra_idx, const float *W, float *out,
                                      void *workspace, size_t workspace_bytes, void *grouped, size_t grouped_bytes,
                                      unsigned long long *grouped_layout, int b, int c_in, int c_out, int p, int na,
                                      int kn, void *stream) {
    return intra_fwd_impl(feats, intra_idx, W, out, workspace, workspace_bytes, grouped, grouped_bytes, grouped_layout, b,
                          c_in, c_out, p, na, kn, stream, nullptr);
}

EPN_API int epn_intra_so3conv_fwd_norm_f32(const float *x, const float *stats, const float *gamma, const float *beta,
                                           int norm_mode, float slope, const int32_t *intra_idx, const float *W,
                                           float *out, void *workspace, size_t workspace_bytes, void *grouped,
                                           size_t grouped_bytes, unsigned long long *grouped_layout, int b, int c_in,
                                           int c_out, int p, int na, int kn, void *stream) {
    EPN_REQUIRE_PTR(stats);
    EPN_REQUIRE(norm_mode == 0 || norm_mode == 1, EPN_ERR_SHAPE, "norm_mode must be 0 (instance) or 1 (batch)");
    EPN_REQUIRE(gemm_backend() == 0, EPN_ERR_SHAPE, "fused norm prologue needs the tensor-core engine");
    NormPrologue pro;
    pro.stats = stats;
    pro.gamma = gamma;
    pro.beta = beta;
    pro.mode = norm_mode;
    pro.G = norm_mode == 0 ? b * c_in : c_in;
    pro.c = c_in;
    pro.slope = slope;
    return intra_fwd_impl(x, intra_idx, W, out, workspace, workspace_bytes, grouped, grouped_bytes, grouped_layout, b, c_in,
                          c_out, p, na, kn, stream, &pro);
}

EPN_API int epn_intra_so3conv_bwd_f32(const float *dout, const float *feats, const int32_t *intra_idx,
                                      const float *W, float *dfeats, float *dW, void *workspace,
                                      size_t workspace_bytes, const void *grouped, size_t grouped_bytes,
                                      unsigned long long grouped_layout, int b, int c_in, int c_out, int p, int na,
                                      int kn, void *stream) {
    EPN_REQUIRE_PTR(dout); EPN_REQUIRE_PTR(intra_idx);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c_in); EPN_REQUIRE_POS(c_out); EPN_REQUIRE_POS(p); EPN_REQUIRE_POS(na);
    EPN_REQUIRE_POS(kn); EPN_CHECK_B(b);
    if (dfeats != nullptr) EPN_REQUIRE_PTR(W);
    if (dW != nullptr && grouped == nullptr) EPN_REQUIRE_PTR(feats);
    const int ck = c_in * kn;
    SlabPlan sp = plan_slabs(b, ck, p, na);
    if (grouped != nullptr) {
        int kperm_kept = 0;
        EPN_REQUIRE(decode_layout(grouped_layout, b, p, &kperm_kept, &sp) && kperm_kept == 0, EPN_ERR_SHAPE,
                    "grouped_layout is not a value returned by the forward");
        EPN_CHECK_GROUPED(intra_group_tiles_ok(na, kn) ? grouped_tiles_bytes(b, ck, p, na, sp) : 0);
    }
    const Workspace ws = carve(workspace, ck, c_out, (long long)sp.bc * sp.pc * na, true);
    EPN_CHECK_WS(ws.total);
    const uint8_t *keep = static_cast<const uint8_t *>(grouped);
    cudaStream_t s = as_stream(stream);
    if (dfeats != nullptr) EPN_TRY(prep_weights(W, c_out, ck, ws, false, true, s));
    if (dW != nullptr) cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)c_out * ck, s);
    for (int b0 = 0; b0 < b; b0 += sp.bc) {
        const int bc = b - b0 < sp.bc ? b - b0 : sp.bc;
        for (int p0 = 0; p0 < p; p0 += sp.pc) {
            const int pc = p - p0 < sp.pc ? p - p0 : sp.pc;
            const long long cols = (long long)pc * na, n_slab = bc * cols;
            ColsView d{const_cast<float *>(dout) + ((size_t)b0 * c_out * p + p0) * na, (long long)c_out * p * na,
                       (long long)p * na};
            ColsView slab{ws.slab, cols, n_slab};
            if (dfeats != nullptr) {
                // fused: dG = W^T . dout reduced through the inverse anchor permutations inside the GEMM epilogue
                // (dG, 12x the size of dfeats, never reaches HBM); otherwise GEMM into the slab + gather kernel
                int rc = 1;
                if (gemm_backend() == 0 && pc == p && intra_dx_fused_ok(n_slab, p, na, kn) &&
                    intra_dx_wt_bytes(c_in, c_out) <= (size_t)ck * n_slab * sizeof(float))
                    rc = launch_umma_intra_dx(d.ptr, d.stride_z, d.stride_k, W, intra_idx, dfeats + (size_t)b0 * c_in * p * na,
                                              ws.slab, ws.tilesA, bc, c_in, c_out, p, 0, s);
                if (rc != 0 && rc != 1) return rc;
                if (rc == 1) {
                    EPN_TRY(gemm_dx(W, c_out, ck, d, bc, cols, slab, ws, s));
                    EPN_TRY(launch_intra_group_bwd(ws.slab, cols, n_slab, p0, pc, intra_idx,
                                                   dfeats + (size_t)b0 * c_in * p * na, bc, c_in, p, na, kn, s));
                }
            }
            const uint8_t *kept = keep;
            if (keep) keep += split_tiles_bytes(n_slab, ck, 128);
            if (dW != nullptr && kept != nullptr) {
                EPN_TRY(gemm_dw_grouped(kept, d, c_out, ck, bc, cols, dW, ws, s));
            } else if (dW != nullptr) {
                int direct = 1;
                if (gemm_backend() == 0) {
                    direct = launch_intra_group_tiles(feats + (size_t)b0 * c_in * p * na, intra_idx, ws.tilesA, 1, p0, pc, bc,
                                                      c_in, p, na, kn, s);
                    if (direct != 0 && direct != 1) return direct;
                }
                if (direct == 0) {
                    EPN_TRY(gemm_dw_tiles(d, c_out, ck, bc, cols, dW, ws, s));
                } else {
                    EPN_TRY(launch_intra_group_fwd(feats + (size_t)b0 * c_in * p * na, intra_idx, ws.slab, cols, n_slab, p0,
                                                   pc, bc, c_in, p, na, kn, s));
                    EPN_TRY(gemm_dw(d, slab, c_out, ck, bc, cols, dW, ws, s));
                }
            }
        }
    }
    return 0;
}
