// (1) Intra grouping straight into operand tiles:
//       G'[(c,k), (z,p,a)] = feats[z, c, p, intra_idx[a,k]]          (so3conv/functional.py:221-268)
//     a pure gather (no arithmetic): 8 gathered loads from L1-resident 4*na-byte feature rows -> bf16 hi/lo
//     split -> two 16-byte tile stores, lanes along tile rows (contiguous stores).
// (2) Inter grouping backward (scatter) with bulk-copy-staged gradient rows:
//       dfeats[z, c, idx[z,p,n], a] += sum_k w(p,a,k,n) * dG[(c,k), (z,p,a)]
//     thread = (anchor lane, 4 neighbours) holding w[24][4] in registers; the 24 rows of dG of every
//     channel of a chunk are staged in shared memory (UBLKCP on an mbarrier, double buffered) and each
//     (channel, neighbour) result goes out as one fp32 RED into the neighbour's 4*na-byte feature row.
// Both are specialised for the shipped geometry (60 anchors, 12 intra neighbours, 24 kernel points) so
// all shared-memory strides and index decodes are compile-time constants; other shapes fall back to the
// generic kernels of epn_group.cu.
#include "epn_dedup.cuh"
#include "epn_internal.cuh"
#include "epn_umma.cuh"

#ifndef EPN_INTRA_ROWS_BLOCKS
#define EPN_INTRA_ROWS_BLOCKS 4   // resident blocks per SM the gather kernel is compiled for (64 registers per thread)
#endif

namespace epn {
using namespace umma;

// ------------------------------------------------------------------ intra tiles, mode 0
// rows = (z,p,a) columns, K = (c,k).  Thread = one row; the 12 permuted anchor positions of the row sit in
// registers; channels are consumed in pairs (2 x 12 = 24 k = three 8-wide chunks).
template <int NA, int KN, int PAIRS>
__global__ void __launch_bounds__(256, EPN_INTRA_ROWS_BLOCKS)
intra_tiles_rows_kernel(const float *__restrict__ feats, const int32_t *__restrict__ intra_idx,
                        uint8_t *__restrict__ tiles, int k_blocks, int c, int p, int p_off, int p_cnt, int n_slab,
                        int rows_pad, NormPrologue pro) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_pad) return;
    const int cols = p_cnt * NA, kcgs = k_blocks * (KB / 8);
    const bool row_ok = row < n_slab;
    const int z = row_ok ? row / cols : 0, rem = row_ok ? row - z * cols : 0, pl = rem / NA, a = rem - pl * NA;
    int ix[KN];
#pragma unroll
    for (int k = 0; k < KN; ++k) ix[k] = __ldg(intra_idx + a * KN + k);
    const float *frow0 = feats + (((size_t)z * c) * p + p_off + pl) * NA;  // channel 0 of this (z, point)
    const size_t cstride = (size_t)p * NA;
    uint8_t *tbase = tiles + ((size_t)(row >> 7) * k_blocks) * tile_bytes(TR_A) + (size_t)(row & 127) * 16;
    for (int pr = blockIdx.y * PAIRS; pr < (blockIdx.y + 1) * PAIRS; ++pr) {
        if (pr * 3 >= kcgs) break;
        const int c0 = 2 * pr;
        float v[2 * KN];
#pragma unroll
        for (int k = 0; k < KN; ++k) {
            v[k] = (row_ok && c0 < c) ? __ldg(frow0 + (size_t)c0 * cstride + ix[k]) : 0.f;
            v[KN + k] = (row_ok && c0 + 1 < c) ? __ldg(frow0 + (size_t)(c0 + 1) * cstride + ix[k]) : 0.f;
        }
        if (pro.stats != nullptr && row_ok) {   // the producer's normalisation + leaky_relu, applied on the way in
#pragma unroll
            for (int k = 0; k < KN; ++k) {
                if (c0 < c) v[k] = pro.apply(v[k], z, c0);
                if (c0 + 1 < c) v[KN + k] = pro.apply(v[KN + k], z, c0 + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int kc = pr * 3 + j;
            if (kc < kcgs) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = v[j * 8 + i];
                uint4 hi, lo;
                split8(x, hi, lo);
                uint8_t *dst = tbase + (size_t)(kc >> 2) * tile_bytes(TR_A) + (size_t)(kc & 3) * (TR_A * 16);
                *reinterpret_cast<uint4 *>(dst) = hi;
                *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = lo;
            }
        }
    }
}

// ------------------------------------------------------------------ intra tiles, mode 1
// rows = (c,k), K = (z,p,a) columns.  Thread = one row and a run of CPT consecutive 8-column chunks; the
// column index is decoded once and then advanced incrementally.
template <int NA, int KN, int CPT>
__global__ void __launch_bounds__(256)
intra_tiles_cols_kernel(const float *__restrict__ feats, const int32_t *__restrict__ intra_idx,
                        uint8_t *__restrict__ tiles, int k_blocks, int c, int p, int p_off, int p_cnt, int n_slab,
                        int rows_pad) {
    __shared__ int32_t s_ix[NA * KN];
    for (int i = threadIdx.x; i < NA * KN; i += blockDim.x) s_ix[i] = intra_idx[i];
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_pad) return;
    const int ck = c * KN, cols = p_cnt * NA, kcgs = k_blocks * (KB / 8);
    const bool row_ok = row < ck;
    const int cc = row_ok ? row / KN : 0, k = row_ok ? row - cc * KN : 0;
    uint8_t *tbase = tiles + ((size_t)(row >> 7) * k_blocks) * tile_bytes(TR_A) + (size_t)(row & 127) * 16;
    int kcg = blockIdx.y * CPT;
    int col = kcg * 8;
    int z = col / cols, rem = col - z * cols, pl = rem / NA, a = rem - pl * NA;
    for (int j = 0; j < CPT && kcg < kcgs; ++j, ++kcg) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = (row_ok && col < n_slab)
                       ? __ldg(feats + (((size_t)z * c + cc) * p + p_off + pl) * NA + s_ix[a * KN + k]) : 0.f;
            ++col;
            if (++a == NA) { a = 0; if (++pl == p_cnt) { pl = 0; ++z; } }
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        uint8_t *dst = tbase + (size_t)(kcg >> 2) * tile_bytes(TR_A) + (size_t)(kcg & 3) * (TR_A * 16);
        *reinterpret_cast<uint4 *>(dst) = hi;
        *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = lo;
    }
}

bool intra_group_tiles_ok(int na, int kn) { return na == 60 && kn == 12; }

int launch_intra_group_tiles(const float *feats, const int32_t *intra_idx, void *tiles, int mode, int p_off, int p_cnt,
                             int bc, int c, int p, int na, int kn, cudaStream_t s, const NormPrologue *pro) {
    const long long n_slab = (long long)bc * p_cnt * na;
    const int ck = c * kn;
    if (n_slab >= (1LL << 31) - 4096 || !intra_group_tiles_ok(na, kn)) return 1;
    const long long rows = mode == 0 ? n_slab : ck, K = mode == 0 ? ck : n_slab;
    const int rows_pad = (int)((rows + 127) / 128 * 128), k_blocks = (int)((K + KB - 1) / KB);
    const int kcgs = k_blocks * (KB / 8);
    ProfScope prof(s, KC_INTRA_GROUP);
    if (mode == 0) {
        constexpr int PAIRS = 8;
        const int pairs = (kcgs + 2) / 3;
        dim3 grid((rows_pad + 255) / 256, (pairs + PAIRS - 1) / PAIRS);
        if (grid.y > 65535) return 1;
        intra_tiles_rows_kernel<60, 12, PAIRS><<<grid, 256, 0, s>>>(feats, intra_idx, static_cast<uint8_t *>(tiles), k_blocks,
                                                                  c, p, p_off, p_cnt, (int)n_slab, rows_pad,
                                                                  pro ? *pro : NormPrologue());
    } else {
        if (pro != nullptr) return 1;   // the column-major variant (backward re-gather) has no prologue
        constexpr int CPT = 8;
        dim3 grid((rows_pad + 255) / 256, (kcgs + CPT - 1) / CPT);
        if (grid.y > 65535) return 1;
        intra_tiles_cols_kernel<60, 12, CPT><<<grid, 256, 0, s>>>(feats, intra_idx, static_cast<uint8_t *>(tiles), k_blocks, c,
                                                                p, p_off, p_cnt, (int)n_slab, rows_pad);
    }
    return check_launch("intra_tiles_kernel");
}

// ------------------------------------------------------------------ inter scatter
constexpr int SC_LANES = 64, SC_KS = 24, SC_NB = 4, SC_CCH = 4, SC_BUFS = 3;

template <int NN, int NA>
__global__ void __launch_bounds__(SC_LANES *(NN / SC_NB), NN == 16 ? 2 : 1)
inter_scatter_kernel(const float *__restrict__ dG, long long stride_b, long long stride_ck, const int32_t *__restrict__ idx,
                     InterGeom g, float *__restrict__ dfeats, int c, int p_in, int p, int nn, int p_off) {
    constexpr int NTHR = SC_LANES * (NN / SC_NB);
    // NN == 16: rows of <= 16 slots, one CTA per point.  NN == 32: rows of up to DEDUP_MAX_RAW slots; the scatter is
    // additive over neighbours, so CTA blockIdx.z takes the distinct neighbours [32 z, 32 z + 32) of the point and
    // CTAs beyond the point's distinct count leave at once (the K = 64 layers of the rotation / 3DMatch models).
    constexpr int CAP = NN == 16 ? 16 : DEDUP_MAX_RAW;
    extern __shared__ __align__(16) float s_dyn[];
    float *Ds = s_dyn;                                             // [SC_BUFS][SC_CCH*24][NA]
    __shared__ NeighbourList<CAP> L;
    const int tid = threadIdx.x;
    const int a = tid % SC_LANES, grp = tid / SC_LANES;
    const int n0 = (int)blockIdx.z * NN + grp * SC_NB;
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : a - 4;  // dead lanes shadow a live lane of their own warp (broadcast, no bank conflict)
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    float *DF = dfeats + (size_t)z * c * p_in * NA;

    // distinct neighbours + multiplicities: one RED per distinct neighbour
    dedup_row(L, idx + ((size_t)z * p + pi) * nn, nn, g.xyz + (size_t)z * 3 * p_in, g.centers + (size_t)z * 3 * p, p_in, p, pi,
              tid, NTHR, [] { __syncthreads(); });
    const float *s_g = L.g, *s_mult = L.mult;
    const int32_t *s_idx = L.idx;
    nn = L.total;  // number of DISTINCT neighbours
    if ((int)blockIdx.z * NN >= nn) return;  // CTA-uniform
    const bool grp_active = n0 < nn;  // warp-uniform: this thread's 4 neighbours exist

    uint64_t w2[SC_KS][SC_NB / 2];  // (neighbour 2j, neighbour 2j+1) pairs for the fp32x2 FMAs
    {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + aa * 9 + i);
        const float inv_sigma = 1.0f / g.sigma;
        const uint64_t nis2 = pack_f32x2(-inv_sigma, -inv_sigma);
        uint64_t gx[SC_NB / 2], gy[SC_NB / 2], gz[SC_NB / 2], mm[SC_NB / 2];
#pragma unroll
        for (int j = 0; j < SC_NB; j += 2) {   // absent neighbours carry multiplicity 0; dead lanes never accumulate
            const int n = n0 + j;
            gx[j / 2] = pack_f32x2(s_g[n * 3], s_g[n * 3 + 3]);
            gy[j / 2] = pack_f32x2(s_g[n * 3 + 1], s_g[n * 3 + 4]);
            gz[j / 2] = pack_f32x2(s_g[n * 3 + 2], s_g[n * 3 + 5]);
            mm[j / 2] = pack_f32x2(s_mult[n], s_mult[n + 1]);
        }
#pragma unroll
        for (int k = 0; k < SC_KS; ++k) {
            const float kx = __ldg(g.kernels + k * 3), ky = __ldg(g.kernels + k * 3 + 1), kz = __ldg(g.kernels + k * 3 + 2);
            const KPoint2 rk = kpoint2(R[0] * kx + R[1] * ky + R[2] * kz, R[3] * kx + R[4] * ky + R[5] * kz,
                                       R[6] * kx + R[7] * ky + R[8] * kz);
#pragma unroll
            for (int j = 0; j < SC_NB / 2; ++j) w2[k][j] = kernel_weight_pair(gx[j], gy[j], gz[j], rk, nis2, mm[j]);
        }
    }
    // destination rows of this thread's 4 neighbours (element offset of [q, a] inside one channel plane)
    int qoff[SC_NB];
#pragma unroll
    for (int j = 0; j < SC_NB; ++j) qoff[j] = s_idx[n0 + j] * NA + aa;
    const size_t cplane = (size_t)p_in * NA;

    const int nchunks = (c + SC_CCH - 1) / SC_CCH;
    // staged gradient rows: 16-byte cp.async (LDGSTS) pieces, 15 per 240-byte row (one UBLKCP per row was
    // measured TMA-issue-bound at this size)
    constexpr int SEGS = NA / 4;
    const uint32_t ds_u32 = smem_u32(Ds);
    const float *src0 = dG + (size_t)z * stride_b + (size_t)pl * NA;
    auto issue = [&](int chunk, int buf) {
        const int rows = min(SC_CCH, c - chunk * SC_CCH) * SC_KS;
        for (int t = tid; t < rows * SEGS; t += NTHR) {
            const int rr = t / SEGS, seg = t - rr * SEGS;  // rr = cl*24 + k
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                         ::"r"(ds_u32 + (uint32_t)(((buf * SC_CCH * SC_KS + rr) * NA + seg * 4) * 4)),
                           "l"(src0 + (size_t)(chunk * SC_CCH * SC_KS + rr) * stride_ck + seg * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // three staging buffers, ONE barrier per chunk: the barrier that publishes chunk i also proves that every thread
    // has finished chunk i-1, whose buffer the copies of chunk i+2 (issued right after it) overwrite
    issue(0, 0);
    if (nchunks > 1) issue(1, 1);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk % SC_BUFS;
        if (chunk + 1 < nchunks) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // rows of this chunk visible to every thread; everybody is done with chunk - 1
        if (chunk + 2 < nchunks) issue(chunk + 2, (chunk + 2) % SC_BUFS);
        const float *dbase = Ds + (size_t)(buf * SC_CCH * SC_KS) * NA + aa;
#pragma unroll 2
        for (int cl = 0; cl < SC_CCH; ++cl) {
            const int cc = chunk * SC_CCH + cl;
            if (cc >= c || !grp_active) break;  // idle neighbour groups only keep the barriers company
            uint64_t t2[SC_NB / 2];
#pragma unroll
            for (int j = 0; j < SC_NB / 2; ++j) t2[j] = 0ull;
            const float *drow = dbase + cl * SC_KS * NA;
#pragma unroll
            for (int k = 0; k < SC_KS; ++k) {
                const float dv = drow[k * NA];
                const uint64_t d2 = pack_f32x2(dv, dv);
#pragma unroll
                for (int j = 0; j < SC_NB / 2; ++j) t2[j] = fma_f32x2(w2[k][j], d2, t2[j]);
            }
            float t[SC_NB];
#pragma unroll
            for (int j = 0; j < SC_NB / 2; ++j) unpack_f32x2(t2[j], t[2 * j], t[2 * j + 1]);
            float *dplane = DF + (size_t)cc * cplane;
#pragma unroll
            for (int j = 0; j < SC_NB; ++j)
                if (a_ok && n0 + j < nn) atomicAdd(dplane + qoff[j], t[j]);
        }
    }
}

// Returns 1 when the shape is not covered (caller falls back to inter_group_bwd_kernel).
int launch_inter_scatter(const float *dG, long long stride_b, long long stride_ck, const int32_t *idx,
                         const InterGeom &g, float *dfeats, int p_off, int p_cnt, int bc, int c, int p_in, int p, int nn,
                         int na, int ks, cudaStream_t s) {
    if (ks != SC_KS || nn > DEDUP_MAX_RAW || na != 60 || bc > 65535 || (stride_ck % 4) != 0 || (stride_b % 4) != 0 ||
        ((uintptr_t)dG & 15) != 0 || (long long)p_in * na >= (1LL << 31))
        return 1;
    dim3 grid(p_cnt, bc, nn <= 32 ? 1 : (nn + 31) / 32);
    ProfScope prof(s, KC_INTER_SCATTER);
    static DynSmemOnce once16, once32;
    if (int rc = ensure_dyn_smem(once16, inter_scatter_kernel<16, 60>, 80 * 1024, "inter_scatter_kernel<16>")) return rc;
    if (int rc = ensure_dyn_smem(once32, inter_scatter_kernel<32, 60>, 80 * 1024, "inter_scatter_kernel<32>")) return rc;
    if (nn <= 16) {
        const size_t smem = (size_t)(SC_BUFS * SC_CCH * SC_KS * 60) * sizeof(float);
        inter_scatter_kernel<16, 60><<<grid, SC_LANES * (16 / SC_NB), smem, s>>>(dG, stride_b, stride_ck, idx, g, dfeats, c,
                                                                               p_in, p, nn, p_off);
    } else {
        const size_t smem = (size_t)(SC_BUFS * SC_CCH * SC_KS * 60) * sizeof(float);
        inter_scatter_kernel<32, 60><<<grid, SC_LANES * (32 / SC_NB), smem, s>>>(dG, stride_b, stride_ck, idx, g, dfeats, c,
                                                                               p_in, p, nn, p_off);
    }
    return check_launch("inter_scatter_kernel");
}

}  // namespace epn
