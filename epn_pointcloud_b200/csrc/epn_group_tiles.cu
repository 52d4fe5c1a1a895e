// Inter grouping stage that emits tensor-core operand tiles directly.
//
//   G[(c,k), (z,p,a)] = sum_n w(p,a,k,n) * feats[z, c, idx[z,p,n], a]      (spconv/functional.py:361-390,
//   w(p,a,k,n) = relu(1 - |x_idx - x_p - R_a kappa_k|^2 / sigma)             so3conv/functional.py:180-218)
//
// One CTA per output point.  lane <-> anchor a (feature rows [c, q, 0..na) are contiguous 4*na-byte
// segments), warp-pair <-> group of KG kernel points; a thread keeps its w[KG][NN] slice of the kernel
// weights in registers for the whole point.  Channels stream through in chunks of CCH: every neighbour
// feature row of a chunk is fetched with ONE bulk async copy (UBLKCP, 4*na bytes) that completes on the
// stage buffer's mbarrier (double buffered: the gather of chunk i+1 overlaps the FMAs of chunk i), the
// per-anchor spatial contraction writes fp32 results to a shared staging tile, and a conversion pass
// splits them into bf16 hi/lo and stores 16-byte (forward layout) or 8-byte (transposed layout) pieces of
// the canonical UMMA operand tiles (epn_umma.cuh) -- the grouped tensor never exists in fp32 in global
// memory and no separate conversion kernel runs.  All shared-memory strides are compile-time constants
// (NA = 60 anchors) so the hot loops are LDS/FFMA/STS with immediate offsets.
//   mode 0: tiles of A[rows = (z,p,a) columns, K = (c,k)]   -> forward GEMM  out = G^T-rows x W
//   mode 1: tiles of A[rows = (c,k),           K = columns] -> dW GEMM       dW^T = G x dout^T
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

constexpr int GT_LANES = 64;  // anchor lanes per kernel-point group
constexpr int GT_KS = 24;     // kernel points (kpsphere24)

struct TileOut {
    uint8_t *tiles;
    int k_blocks;          // K blocks (of 32) of the tile matrix
    int row_limit;         // mode 1: row_tiles * 128
    long long cols_per_z;  // flattened column of (z, pl, a) = z * cols_per_z + pl * na + a
    int mode;
};

template <int NN, int KG, int CCH, int NA, bool HAS_FEATS>
__global__ void __launch_bounds__(GT_LANES *(GT_KS / KG), NN == 16 ? 2 : 1)
inter_group_tiles_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx, InterGeom g, TileOut out,
                         int c, int p_in, int p, int nn, int p_off) {
    constexpr int NTHR = GT_LANES * (GT_KS / KG);
    constexpr int GSTR = NA + 1;       // odd staging stride: conflict-free for both conversion mappings
    constexpr int CKK = CCH * GT_KS;   // (c,k) rows produced per chunk (multiple of 32)
    extern __shared__ __align__(16) float s_dyn[];
    float *s_g = s_dyn;                                            // [NN][3]  unique neighbour offsets
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_dyn + NN * 3);  // [NN]     unique neighbour indices
    float *s_mult = s_dyn + NN * 4;                                // [NN]     their multiplicities
    int32_t *s_raw = reinterpret_cast<int32_t *>(s_dyn + NN * 5);  // [NN]     the ball-query row as stored
    float *Fs = s_dyn + NN * 6;                                    // [2][CCH][NN][NA]
    float *Gs = Fs + 2 * CCH * NN * NA;                            // [CKK][GSTR]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_nu;
    const int tid = threadIdx.x;
    const int a = tid % GT_LANES, grp = tid / GT_LANES;
    const int k0 = grp * KG;
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : a - 4;  // dead lanes shadow a live lane of their own warp (broadcast, no bank conflict)
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    const float *F = feats ? feats + (size_t)z * c * p_in * NA : nullptr;

    // The ball query repeat-fills short neighbour lists cyclically (grouping_cuda_kernel.cu:100-104), so a row
    // usually holds each neighbour several times.  Duplicates have identical offsets and therefore identical
    // kernel weights: keep each distinct neighbour once with its multiplicity folded into the weight
    // (sum_n w_n f_n == sum_u m_u w_u f_u) -- fewer rows to gather and fewer FMAs, same result.
    for (int n = tid; n < NN; n += NTHR) s_raw[n] = n < nn ? idx[((size_t)z * p + pi) * nn + n] : -1;
    __syncthreads();
    if (tid < 32) {  // nn <= 32: one warp de-duplicates the row, keeping first-occurrence order
        const int n = tid;
        const int q = n < nn ? s_raw[n] : -1;
        bool uniq = n < nn;
        for (int m = 0; m < n && uniq; ++m) uniq = s_raw[m] != q;
        int mult = 0;
        for (int m = n; m < nn; ++m) mult += (s_raw[m] == q) ? 1 : 0;
        const unsigned mask = __ballot_sync(0xffffffffu, uniq);
        const int pos = __popc(mask & ((1u << n) - 1u));
        if (uniq) {
            const float *X = g.xyz + (size_t)z * 3 * p_in;
            const float *Cn = g.centers + (size_t)z * 3 * p;
            s_idx[pos] = q;
            s_mult[pos] = (float)mult;
            s_g[pos * 3] = X[q] - Cn[pi];
            s_g[pos * 3 + 1] = X[p_in + q] - Cn[p + pi];
            s_g[pos * 3 + 2] = X[2 * p_in + q] - Cn[2 * p + pi];
        }
        const int cnt = __popc(mask);
        if (n >= cnt && n < NN) {
            s_idx[n] = 0; s_mult[n] = 0.f;
            s_g[n * 3] = 0.f; s_g[n * 3 + 1] = 0.f; s_g[n * 3 + 2] = 0.f;
        }
        if (n == 0) s_nu = cnt;
    }
    const uint32_t bar0 = smem_u32(&s_bar[0]);  // buffer b uses the barrier at bar0 + 8*b
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_barrier_init();
    }
    __syncthreads();
    nn = s_nu;  // from here on: number of DISTINCT neighbours
    // rows beyond the distinct neighbours are never copied: keep them zero in both buffers (the FMA loop runs
    // in groups of 4 neighbours and multiplies them by zero weights)
    for (int t = tid; t < 2 * CCH * (NN - nn) * NA; t += NTHR) {
        const int e = t % NA, r = t / NA, n = nn + r % (NN - nn), bc = r / (NN - nn);
        Fs[(bc * NN + n) * NA + e] = 0.f;
    }

    // kernel weights of this thread's (anchor, kernel-point group): registers for the whole point, packed as
    // (even neighbour, odd neighbour) pairs for the fp32x2 FMAs of the contraction
    uint64_t w2[KG][NN / 2];
    {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + aa * 9 + i);
#pragma unroll
        for (int i = 0; i < KG; ++i) {
            const float kx = __ldg(g.kernels + (k0 + i) * 3), ky = __ldg(g.kernels + (k0 + i) * 3 + 1),
                        kz = __ldg(g.kernels + (k0 + i) * 3 + 2);
            const float rx = R[0] * kx + R[1] * ky + R[2] * kz, ry = R[3] * kx + R[4] * ky + R[5] * kz,
                        rz = R[6] * kx + R[7] * ky + R[8] * kz;
#pragma unroll
            for (int n = 0; n < NN; n += 2) {
                float v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float t = kernel_weight_fast(s_g[(n + e) * 3], s_g[(n + e) * 3 + 1], s_g[(n + e) * 3 + 2], rx, ry, rz,
                                                       1.0f / g.sigma);
                    v[e] = (a_ok && n + e < nn) ? t * s_mult[n + e] : 0.f;
                }
                w2[i][n / 2] = pack_f32x2(v[0], v[1]);
            }
        }
    }

    const int nchunks = (c + CCH - 1) / CCH;
    const uint32_t fs_u32 = smem_u32(Fs);
    constexpr uint32_t ROW_BYTES = NA * 4;
    auto issue = [&](int chunk, int buf) {
        if (!HAS_FEATS) return;
        const int nch = min(CCH, c - chunk * CCH);
        const uint32_t bar = bar0 + 8u * (uint32_t)buf;
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nch * nn) * ROW_BYTES);
        for (int t = tid; t < nch * NN; t += NTHR) {
            const int cl = t / NN, n = t % NN;
            if (n < nn)
                bulk_g2s(fs_u32 + (uint32_t)(((buf * CCH + cl) * NN + n) * NA) * 4u,
                         F + ((size_t)(chunk * CCH + cl) * p_in + s_idx[n]) * NA, ROW_BYTES, bar);
        }
    };

    // conversion addressing that does not depend on the chunk
    const long long col0 = (long long)z * out.cols_per_z + (long long)pl * NA;
    const long long row_m0 = col0 + aa;  // mode 0: this thread's tile row
    uint8_t *m0_base = out.tiles + ((size_t)(row_m0 >> 7) * out.k_blocks) * tile_bytes(TR_A) + (size_t)(row_m0 & 127) * 16;

    uint32_t phase_bits = 0u;  // bit b = parity to wait for on buffer b
    issue(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
        if (HAS_FEATS) {
            mbar_wait(bar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
            phase_bits ^= 1u << buf;
        }
        __syncthreads();  // previous conversion finished reading Gs

        // ---- spatial contraction of CCH channels: 1 LDS + KG FFMA per neighbour, KG STS per channel
        const float *fbase = Fs + (size_t)(buf * CCH * NN) * NA + aa;
        float *gbase = Gs + (size_t)k0 * GSTR + aa;
#pragma unroll 2
        for (int cl = 0; cl < CCH; ++cl) {
            // acc2[i] = (sum over even neighbours, sum over odd neighbours) of w * f
            uint64_t acc2[KG];
#pragma unroll
            for (int i = 0; i < KG; ++i) acc2[i] = 0ull;
            if (chunk * CCH + cl < c) {
                const float *frow = fbase + cl * NN * NA;
#pragma unroll
                for (int n4 = 0; n4 < NN; n4 += 4) {
                    if (n4 < nn) {  // CTA-uniform: whole groups of 4 absent neighbours are skipped
#pragma unroll
                        for (int n = n4; n < n4 + 4; n += 2) {
                            // occupancy features == 1 when there is nothing to gather
                            const uint64_t f2 = HAS_FEATS ? pack_f32x2(frow[n * NA], frow[(n + 1) * NA]) : pack_f32x2(1.0f, 1.0f);
#pragma unroll
                            for (int i = 0; i < KG; ++i) acc2[i] = fma_f32x2(w2[i][n / 2], f2, acc2[i]);
                        }
                    }
                }
            }
            float acc[KG];
#pragma unroll
            for (int i = 0; i < KG; ++i) {
                float e, o;
                unpack_f32x2(acc2[i], e, o);
                acc[i] = e + o;
            }
            if (a_ok) {
#pragma unroll
                for (int i = 0; i < KG; ++i) gbase[(cl * GT_KS + i) * GSTR] = acc[i];
            }
        }
        __syncthreads();

        // ---- fp32 staging -> bf16 hi/lo operand tiles
        if (out.mode == 0) {
            // thread = (row a, 8-wide kk chunk kc): 8 LDS, split, two 16-byte stores; a warp stores 512 contiguous B
            if (a_ok) {
                for (int kc = grp; kc < CKK / 8; kc += GT_KS / KG) {
                    const int kkg = chunk * CKK + kc * 8;
                    if (kkg >= out.k_blocks * KB) break;
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = Gs[(kc * 8 + i) * GSTR + aa];
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    uint8_t *dst = m0_base + (size_t)(kkg >> 5) * tile_bytes(TR_A) + (size_t)((kkg & 31) >> 3) * (TR_A * 16);
                    *reinterpret_cast<uint4 *>(dst) = hi;
                    *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = lo;
                }
            }
        } else {
            // thread = (row kk, 4 columns): lanes along kk rows -> 16-byte-strided 8-byte stores
            for (int t = tid; t < CKK * (NA / 4); t += NTHR) {
                const int kkl = t % CKK, gq = t / CKK;
                const int row = chunk * CKK + kkl;
                if (row >= out.row_limit) continue;
                float x[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = Gs[kkl * GSTR + gq * 4 + i];
#pragma unroll
                for (int i = 4; i < 8; ++i) x[i] = 0.f;
                uint4 hi, lo;
                split8(x, hi, lo);
                const long long col = col0 + gq * 4;
                uint8_t *dst = out.tiles + ((size_t)(row >> 7) * out.k_blocks + (size_t)(col >> 5)) * tile_bytes(TR_A) +
                               (size_t)((col & 31) >> 3) * (TR_A * 16) + (size_t)(row & 127) * 16 + (size_t)((col & 7) >> 2) * 8;
                *reinterpret_cast<uint2 *>(dst) = make_uint2(hi.x, hi.y);
                *reinterpret_cast<uint2 *>(dst + part_bytes(TR_A)) = make_uint2(lo.x, lo.y);
            }
        }
        // the next iteration's top-of-loop __syncthreads orders these Gs reads before the next Gs writes
    }
}

template <int NN, int KG, int CCH, int NA, bool HAS_FEATS>
static int launch_variant(const float *feats, const int32_t *idx, const InterGeom &g, const TileOut &o, dim3 grid, int c,
                          int p_in, int p, int nn, int p_off, cudaStream_t s) {
    const size_t smem = (size_t)(NN * 6 + 2 * CCH * NN * NA + CCH * GT_KS * (NA + 1)) * sizeof(float);
    static DynSmemOnce once;  // one per template instantiation
    if (int rc = ensure_dyn_smem(once, inter_group_tiles_kernel<NN, KG, CCH, NA, HAS_FEATS>, (int)smem, "inter_group_tiles_kernel")) return rc;
    inter_group_tiles_kernel<NN, KG, CCH, NA, HAS_FEATS><<<grid, GT_LANES *(GT_KS / KG), smem, s>>>(feats, idx, g, o, c, p_in, p, nn, p_off);
    return check_launch("inter_group_tiles_kernel");
}

bool inter_group_tiles_ok(int nn, int na, int ks) { return ks == GT_KS && nn <= 32 && na == 60; }

// Returns 1 when the shape is not covered by the tile kernel (the caller uses the slab + split path).
int launch_inter_group_tiles(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, int k_blocks,
                             int row_limit, long long cols_per_z, int mode, int p_off, int p_cnt, int bc, int c,
                             int p_in, int p, int nn, int na, int ks, cudaStream_t s) {
    if (!inter_group_tiles_ok(nn, na, ks) || bc > 65535) return 1;
    TileOut o{static_cast<uint8_t *>(tiles), k_blocks, row_limit, cols_per_z, mode};
    dim3 grid(p_cnt, bc);
    ProfScope prof(s, KC_INTER_GROUP);
    if (feats == nullptr) {  // occupancy features (layer 0): nothing to gather
        if (nn <= 16) return launch_variant<16, 6, 8, 60, false>(feats, idx, g, o, grid, c, p_in, p, nn, p_off, s);
        return launch_variant<32, 3, 4, 60, false>(feats, idx, g, o, grid, c, p_in, p, nn, p_off, s);
    }
    if (nn <= 16) return launch_variant<16, 6, 8, 60, true>(feats, idx, g, o, grid, c, p_in, p, nn, p_off, s);
    return launch_variant<32, 3, 4, 60, true>(feats, idx, g, o, grid, c, p_in, p, nn, p_off, s);
}

}  // namespace epn
