// Internal launch interfaces shared between the translation units of libepn_b200.so.
#pragma once
#include "epn_common.cuh"

namespace epn {

struct InterGeom {
    const float *xyz;      // [B,3,P_in]
    const float *centers;  // [B,3,P]
    const float *anchors;  // [na,3,3]
    const float *kernels;  // [ks,3]
    float sigma;
};

// Normalisation + leaky_relu applied to a feature tensor [z, c, n] WHILE a consumer loads it (the
// "next prologue" half of the norm fusion): y = lrelu((x - mean_g) * rstd_g * gamma_ch + beta_ch), stats[g] = mean,
// stats[G + g] = rstd with g = z * c + ch (mode 0, InstanceNorm2d) or g = ch (mode 1, BatchNorm2d).  stats == NULL: off.
struct NormPrologue {
    const float *stats = nullptr;
    const float *gamma = nullptr, *beta = nullptr;   // NULL = 1 / 0
    int mode = 0, G = 0, c = 0;
    float slope = 0.01f;
    __device__ __forceinline__ float apply(float x, int z, int ch) const {
        const int g = mode == 0 ? z * c + ch : ch;
        const float mean = __ldg(stats + g), sc = __ldg(stats + G + g) * (gamma ? __ldg(gamma + ch) : 1.f);
        const float v = fmaf(x - mean, sc, beta ? __ldg(beta + ch) : 0.f);
        return v > 0.f ? v : v * slope;
    }
};

// Kernel weight relu(1 - |g - r|^2 / sigma) in the reference's fp32 operation order
// (so3conv/functional.py:198-200: square, (x+y)+z, true division, subtract; no FMA contraction), so the
// weights agree with the reference to the last bit for identical rotated kernel points.
__device__ __forceinline__ float kernel_weight(float gx, float gy, float gz, float rx, float ry, float rz,
                                               float sigma) {
    const float dx = gx - rx, dy = gy - ry, dz = gz - rz;
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return fmaxf(__fsub_rn(1.0f, __fdiv_rn(d, sigma)), 0.0f);
}

// Same weight with the division replaced by a multiplication with 1/sigma (<= 1 ulp of d/sigma away, i.e.
// ~6e-8 absolute on a weight in [0,1]); used where 96 weights per thread are derived per point and the IEEE
// division's slow path (FCHK + call) would cost ~30 instructions each.
__device__ __forceinline__ float kernel_weight_fast(float gx, float gy, float gz, float rx, float ry, float rz,
                                                    float inv_sigma) {
    const float dx = gx - rx, dy = gy - ry, dz = gz - rz;
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return fmaxf(fmaf(-d, inv_sigma, 1.0f), 0.0f);
}

// Packed fp32x2 arithmetic (sm_100: FFMA2 retires two fp32 FMAs per issue slot at the same peak FLOP/s as
// two FFMAs -- measured 73.8 vs 72.5 TFLOP/s -- so issue-bound FMA loops need half the instructions).
// Each half is an ordinary IEEE fp32 fma.rn.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Rotated kernel point r = R_a kappa_k broadcast into both halves of three fp32x2 registers (hoisted per (anchor, k)).
struct KPoint2 {
    uint64_t x, y, z;
};
__device__ __forceinline__ KPoint2 kpoint2(float rx, float ry, float rz) {
    return KPoint2{pack_f32x2(rx, rx), pack_f32x2(ry, ry), pack_f32x2(rz, rz)};
}
// Kernel weights of TWO neighbours against one rotated kernel point, multiplicities folded in, as one fp32x2 pair:
//   ( m0 * relu(1 - |g0 - r|^2 / sigma),  m1 * relu(1 - |g1 - r|^2 / sigma) )
// The arithmetic of kernel_weight_fast per half, issued as packed fp32x2 instructions (8 packed + 2 max instead of
// ~24 scalar ones; ptxas contracts the packed squares and adds into FFMA2, so |g - r|^2 carries two roundings fewer
// than the scalar form: the weights move by <= 2e-7 absolute, far inside the 1e-4 bar -- the op-surface function
// epn_inter_weights_f32 keeps the reference's exact operation order, kernel_weight above).  The
// multiplicity (>= 0, 0 for absent neighbours) is multiplied in before the max, which gives the same value.
// g*: the two neighbours' offsets (lo half = neighbour 0); neg_inv_sigma2 = (-1/sigma, -1/sigma); m2 = (m0, m1).
__device__ __forceinline__ uint64_t kernel_weight_pair(uint64_t gx, uint64_t gy, uint64_t gz, const KPoint2 &r,
                                                       uint64_t neg_inv_sigma2, uint64_t m2) {
    const uint64_t minus1 = pack_f32x2(-1.0f, -1.0f), one = pack_f32x2(1.0f, 1.0f);
    const uint64_t dx = fma_f32x2(r.x, minus1, gx), dy = fma_f32x2(r.y, minus1, gy), dz = fma_f32x2(r.z, minus1, gz);  // g - r, exact
    const uint64_t d = add_f32x2(add_f32x2(mul_f32x2(dx, dx), mul_f32x2(dy, dy)), mul_f32x2(dz, dz));
    const uint64_t t = mul_f32x2(fma_f32x2(d, neg_inv_sigma2, one), m2);
    float t0, t1;
    unpack_f32x2(t, t0, t1);
    return pack_f32x2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f));
}

// Matrix operand of the generic GEMM: element (row, col) of slice z lives at
// ptr + z*stride_z + row*stride_row + col*stride_col.  For A rows are m and
// cols are k; for B rows are k and cols are n.
struct GemmOperand {
    const float *ptr;
    long long stride_z, stride_row, stride_col;
};

// epn_group.cu -- grouped slabs are addressed as
//   slab[b*stride_b + (c*ks+k)*stride_ck + (p - p_off)*na + a],  p in [p_off, p_off+p_cnt)
int launch_inter_group_fwd(const float *feats, const int32_t *idx, const float *inter_w, const InterGeom &g,
                           float *out, long long stride_b, long long stride_ck, int p_off, int p_cnt, int b,
                           int c, int p_in, int p, int nn, int na, int ks, cudaStream_t s);
int launch_inter_group_bwd(const float *dgrouped, long long stride_b, long long stride_ck, int p_off, int p_cnt,
                           const int32_t *idx, const float *inter_w, const InterGeom &g, float *dfeats, int b,
                           int c, int p_in, int p, int nn, int na, int ks, cudaStream_t s);
int launch_intra_group_fwd(const float *feats, const int32_t *intra_idx, float *out, long long stride_b,
                           long long stride_ck, int p_off, int p_cnt, int b, int c, int p, int na, int kn,
                           cudaStream_t s);
int launch_intra_group_bwd(const float *dgrouped, long long stride_b, long long stride_ck, int p_off, int p_cnt,
                           const int32_t *intra_idx, float *dfeats, int b, int c, int p, int na, int kn,
                           cudaStream_t s);

// epn_group_tiles.cu -- grouping straight into operand tiles; returns 1 if the shape is unsupported
int launch_inter_group_tiles(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, int k_blocks,
                             int row_limit, long long cols_per_z, int mode, int p_off, int p_cnt, int bc, int c,
                             int p_in, int p, int nn, int na, int ks, cudaStream_t s);

// epn_inter_fused.cu -- gather + spatial contraction + channel GEMM in one kernel (forward); out element
// (z, o, pl, a) at out + z*out_stride_z + o*out_stride_o + pl*na + a.  keep_tiles (optional): the operand tiles
// for the weight gradient, clouds stored in slabs of keep_slab_clouds (keep_slab_bytes apart), rows
// (z % keep_slab_clouds)*keep_cols_per_z + pl*na + a.  Returns 1 if the shape is unsupported.
int inter_fused_mode(int c, int c_out, int p_cnt, int nn, int na, int ks, bool keep);  // 0 = not covered, else the K' mode
int launch_inter_fused(const float *feats, const int32_t *idx, const InterGeom &g, const void *w_tiles, float *out,
                       long long out_stride_z, long long out_stride_o, void *keep_tiles, int keep_k_blocks,
                       long long keep_cols_per_z, int keep_slab_clouds, size_t keep_slab_bytes, int p_off, int p_cnt,
                       int bc, int c, int c_out, int p_in, int p, int nn, int na, int ks, cudaStream_t s, int fmt = 0);

// epn_inter_bwd_fused.cu -- data gradient of the inter conv in one kernel (dG = dout . W^T in TMEM, transposed spatial
// contraction + scatter straight from TMEM); dout element (z, o, pl, a) at dout + z*stride_z + o*stride_o + pl*na + a;
// wt_scratch >= 4 * c*ks * c_out bytes; dfeats pre-zeroed.  Returns 1 if the shape is unsupported.
bool inter_bwd_fused_ok(int c, int c_out, int p_cnt, int nn, int na, int ks);
int launch_inter_bwd_fused(const float *dout, long long dout_stride_z, long long dout_stride_o, const int32_t *idx,
                           const InterGeom &g, const float *W, void *wt_scratch, float *dfeats, int p_off, int p_cnt, int bc,
                           int c, int c_out, int p_in, int p, int nn, int na, int ks, cudaStream_t s);

// epn_group_direct.cu -- inter grouping with the bf16 split in registers; the operand tiles use a permuted K order
// (24 kernel points):  mode 1 (K <= 16 neighbours, c % 4 == 0)  K'(c,k) = (c/4)*96  + (k/6)*24 + (c%4)*6 + (k%6)
//                      mode 2 (K <= 64 neighbours, c % 8 == 0)  K'(c,k) = (c/8)*192 + (k/3)*24 + (c%8)*3 + (k%3)
__host__ __device__ __forceinline__ int inter_kperm_inv(int kp, int mode) {  // K' -> c*24 + k
    if (mode == 1) {
        const int blk = kp / 96, r = kp - blk * 96, grp = r / 24, q = r - grp * 24, cl4 = q / 6, i = q - cl4 * 6;
        return (blk * 4 + cl4) * 24 + grp * 6 + i;
    }
    const int blk = kp / 192, r = kp - blk * 192, grp = r / 24, q = r - grp * 24, c8 = q / 3, i = q - c8 * 3;
    return (blk * 8 + c8) * 24 + grp * 3 + i;
}
int inter_group_direct_mode(const float *feats, int c, int nn, int na, int ks);  // 0 = shape not covered
int launch_inter_group_direct(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, int k_blocks,
                              long long cols_per_z, int p_off, int p_cnt, int bc, int c, int p_in, int p, int nn, int na,
                              int ks, cudaStream_t s);
// steps = 1: the fused kernel's layout (every 16-k step contiguous, c_out <= 256) instead of split tiles
int launch_inter_w_tiles_kperm(const float *W, void *dst, int c_out, int ck, int trb, int mode, int steps, cudaStream_t s,
                               int fmt = 0);
// one input channel (feats NULL = occupancy ones), any row length up to 128: tiles of one K block in plain order
bool inter_group_occ_ok(int c, int nn, int na, int ks);
int launch_inter_group_occ(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, long long cols_per_z,
                           int p_off, int p_cnt, int bc, int p_in, int p, int nn, int na, int ks, cudaStream_t s, int fmt = 0);

// epn_group_tiles2.cu -- return 1 if the shape is unsupported (caller falls back)
int launch_intra_group_tiles(const float *feats, const int32_t *intra_idx, void *tiles, int mode, int p_off, int p_cnt,
                             int bc, int c, int p, int na, int kn, cudaStream_t s, const NormPrologue *pro = nullptr);
int launch_inter_scatter(const float *dG, long long stride_b, long long stride_ck, const int32_t *idx,
                         const InterGeom &g, float *dfeats, int p_off, int p_cnt, int bc, int c, int p_in, int p, int nn,
                         int na, int ks, cudaStream_t s);

// epn_gemm_umma.cu -- tcgen05 path.
// Source matrix of a split-tile conversion: element (row, k) lives at
//   ptr + (row / rows_per_z) * stride_rz + (row % rows_per_z) * stride_row
//       + (k / k_per_z) * stride_kz + (k % k_per_z) * stride_k
struct SplitSrc {
    const float *ptr;
    long long rows_per_z, stride_rz, stride_row;
    long long k_per_z, stride_kz, stride_k;
    NormPrologue pro;   // applied to element (row, k) with z = row / rows_per_z, ch = k (k_per_z must cover K)
};
// Where the GEMM writes D[row, col]: out + (row / rows_per_z) * stride_z + (row % rows_per_z) * stride_row
//                                        + col * stride_col      (atomic: RED add instead of store)
//   + (col / cols_per_z) * stride_cz for outputs whose COLUMN index runs over (cloud, point*anchor).
// When stride_col == 1 (and cols_per_z % 4 == 0) the epilogue writes 16-byte vectors.
struct GemmEpilogue {
    float *out;
    long long rows_per_z, stride_z, stride_row, stride_col;
    bool atomic;
    long long cols_per_z = 1LL << 60, stride_cz = 0;
};
size_t split_tiles_bytes(long long rows, long long K, int tr);
// fmt / scale: operand format of the tiles (epn_umma.cuh FMT_*) and an exact power-of-two factor applied first
int launch_split_tiles(const SplitSrc &src, void *dst, long long rows, long long K, int tr, cudaStream_t s, int fmt = 0,
                       float scale = 1.0f);
int umma_trb_for(int n_rows);
// D[m_rows, n_rows] = A[m_rows, K] * B[n_rows, K]^T on split tiles (A: 128-row tiles, B: trb-row tiles)
int launch_umma_gemm(const void *A_tiles, const void *B_tiles, int m_rows, int n_rows, long long K, int trb,
                     const GemmEpilogue &ep, int split_k, cudaStream_t s, int fmt = 0);

// Intra conv data gradient with the inverse-permutation reduction fused into the GEMM epilogue (epn_gemm_umma.cu);
// returns 1 if the shape is unsupported.
size_t intra_dx_wt_bytes(int c_rows, int c_k);
size_t intra_dx_dout_bytes(long long n, int c_k);
bool intra_dx_fused_ok(long long n_cols, int p, int na, int kn);
int launch_umma_intra_dx(const float *dout, long long dout_stride_z, long long dout_stride_o, const float *W,
                         const int32_t *intra_idx, float *dfeats, void *wt_tiles, void *dout_tiles, int bc, int c_in,
                         int c_out, int p, int forward, cudaStream_t s, int fmt = 0, const NormPrologue *pro = nullptr);

// epn_gemm_dw.cu -- dW[c_out, ck] += dout . G straight from the forward operand tiles of a slab
// (rows = n grouped columns, n % 128 == 0, K = ck), read as the MN-major M operand; B_tiles = dout tiles
// (rows = c_out in trb-row tiles, K = n).
int launch_umma_dw(const void *G_tiles, const void *B_tiles, int ck, int c_out, long long n, int trb, float *dW,
                   int kperm, cudaStream_t s);  // kperm: mode of the K'(c,k) order of the tiles (0 = plain)

// shapes the direct-to-tiles grouping kernels cover
bool inter_group_tiles_ok(int nn, int na, int ks);
bool intra_group_tiles_ok(int na, int kn);

// epn_gemm_simt.cu
int launch_sgemm(const GemmOperand &A, const GemmOperand &B, float *C, long long c_stride_z, long long ldc,
                 int M, int N, int K, int batch, int split_k, int accumulate, cudaStream_t s);

}  // namespace epn
