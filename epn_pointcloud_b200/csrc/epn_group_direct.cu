// Inter grouping stage, K <= 16 neighbours, with the bf16 hi/lo split done IN REGISTERS.
//
//   G[(c,k), (z,p,a)] = sum_n w(p,a,k,n) * feats[z, c, idx[z,p,n], a]      (spconv/functional.py:361-390,
//   w(p,a,k,n) = relu(1 - |x_idx - x_p - R_a kappa_k|^2 / sigma)             so3conv/functional.py:180-218)
//
// Same mapping as inter_group_tiles_kernel<16,6,8,60> (one CTA per output point, lane <-> anchor, warp pair <->
// 6 kernel points, weights in registers as fp32x2 pairs, bulk-copy gather of the neighbour rows), but the fp32
// staging tile and the separate conversion pass are gone (they were 36 % of that kernel's instructions): the
// operand tiles use a PERMUTED K order chosen so that the 24 values a thread produces for 4 consecutive
// channels are three whole 8-wide K chunks,
//     K'(c, k) = (c/4)*96 + (k/6)*24 + (c%4)*6 + (k%6),
// which it converts to bf16 hi/lo pairs as they leave the FMA loop and stores as 16-byte pieces (a warp writes
// 512 contiguous bytes of a tile).  The order of the K dimension is internal to the engine: the weight tiles of
// the forward GEMM are built in the same order and the weight-gradient GEMM un-permutes its output rows
// (inter_kperm()).
#include "epn_dedup.cuh"
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

namespace {
constexpr int GD_LANES = 64, GD_KS = 24, GD_NN = 16, GD_KG = 6, GD_CCH = 8, GD_NA = 60;
constexpr int GD_THR = GD_LANES * (GD_KS / GD_KG);  // 256

__global__ void __launch_bounds__(GD_THR, 2)
inter_group_direct_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx, InterGeom g,
                          uint8_t *__restrict__ tiles, int k_blocks, long long cols_per_z, int c, int p_in, int p, int nn,
                          int p_off) {
    constexpr int NN = GD_NN, KG = GD_KG, CCH = GD_CCH, NA = GD_NA, NTHR = GD_THR;
    extern __shared__ __align__(16) float s_dyn[];
    float *Fs = s_dyn;                                             // [2][CCH][NN][NA]
    __shared__ NeighbourList<NN> L;
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x;
    const int a = tid % GD_LANES, grp = tid / GD_LANES;
    const int k0 = grp * KG;
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : a - 4;  // dead lanes shadow a live lane of their own warp
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    const float *F = feats + (size_t)z * c * p_in * NA;

    const uint32_t bar0 = smem_u32(&s_bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_barrier_init();
    }
    dedup_row(L, idx + ((size_t)z * p + pi) * nn, nn, g.xyz + (size_t)z * 3 * p_in, g.centers + (size_t)z * 3 * p, p_in, p, pi,
              tid, NTHR, [] { __syncthreads(); });
    const float *s_g = L.g, *s_mult = L.mult;
    const int32_t *s_idx = L.idx;
    nn = L.total;  // number of DISTINCT neighbours (<= the row's nn <= NN)
    for (int t = tid; t < 2 * CCH * (NN - nn) * NA; t += NTHR) {  // never-copied rows stay zero
        const int e = t % NA, r = t / NA, n = nn + r % (NN - nn), bc = r / (NN - nn);
        Fs[(bc * NN + n) * NA + e] = 0.f;
    }
    __syncthreads();

    uint64_t w2[KG][NN / 2];
    {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + aa * 9 + i);
        const float inv_sigma = 1.0f / g.sigma;
        const uint64_t nis2 = pack_f32x2(-inv_sigma, -inv_sigma);
#pragma unroll
        for (int i = 0; i < KG; ++i) {
            const float kx = __ldg(g.kernels + (k0 + i) * 3), ky = __ldg(g.kernels + (k0 + i) * 3 + 1),
                        kz = __ldg(g.kernels + (k0 + i) * 3 + 2);
            const KPoint2 rk = kpoint2(R[0] * kx + R[1] * ky + R[2] * kz, R[3] * kx + R[4] * ky + R[5] * kz,
                                       R[6] * kx + R[7] * ky + R[8] * kz);
#pragma unroll
            for (int n = 0; n < NN; n += 2)   // absent neighbours carry multiplicity 0; dead lanes never store
                w2[i][n / 2] = kernel_weight_pair(pack_f32x2(s_g[n * 3], s_g[n * 3 + 3]), pack_f32x2(s_g[n * 3 + 1], s_g[n * 3 + 4]),
                                                  pack_f32x2(s_g[n * 3 + 2], s_g[n * 3 + 5]), rk, nis2, pack_f32x2(s_mult[n], s_mult[n + 1]));
        }
    }

    const int nchunks = (c + CCH - 1) / CCH;
    const uint32_t fs_u32 = smem_u32(Fs);
    constexpr uint32_t ROW_BYTES = NA * 4;
    auto issue = [&](int chunk, int buf) {
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

    const long long row = (long long)z * cols_per_z + (long long)pl * NA + aa;  // this thread's tile row
    uint8_t *row_base = tiles + ((size_t)(row >> 7) * k_blocks) * tile_bytes(TR_A) + (size_t)(row & 127) * 16;

    uint32_t phase_bits = 0u;
    issue(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        __syncthreads();  // every thread is done with the buffer the next gather overwrites
        if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
        mbar_wait(bar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
        phase_bits ^= 1u << buf;
        const float *fbase = Fs + (size_t)(buf * CCH * NN) * NA + aa;
#pragma unroll 1
        for (int blk = 0; blk < CCH / 4; ++blk) {
            if ((chunk * CCH + blk * 4) >= c) break;  // c % 4 == 0: blocks of 4 channels are whole
            uint32_t hi[12], lo[12];
            const int kc0 = (chunk * (CCH / 4) + blk) * 12 + grp * 3;  // first of this thread's three K' chunks
#pragma unroll
            for (int cl4 = 0; cl4 < 4; ++cl4) {
                uint64_t acc2[KG];
#pragma unroll
                for (int i = 0; i < KG; ++i) acc2[i] = 0ull;
                const float *frow = fbase + (blk * 4 + cl4) * NN * NA;
#pragma unroll
                for (int n4 = 0; n4 < NN; n4 += 4) {
                    if (n4 < nn) {  // CTA-uniform
#pragma unroll
                        for (int n = n4; n < n4 + 4; n += 2) {
                            const uint64_t f2 = pack_f32x2(frow[n * NA], frow[(n + 1) * NA]);
#pragma unroll
                            for (int i = 0; i < KG; ++i) acc2[i] = fma_f32x2(w2[i][n / 2], f2, acc2[i]);
                        }
                    }
                }
#pragma unroll
                for (int ip = 0; ip < KG / 2; ++ip) {
                    float e0, o0, e1, o1;
                    unpack_f32x2(acc2[2 * ip], e0, o0);
                    unpack_f32x2(acc2[2 * ip + 1], e1, o1);
                    const float v0 = e0 + o0, v1 = e1 + o1;
                    const __nv_bfloat162 hp = __floats2bfloat162_rn(v0, v1);
                    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&hp);
                    const __nv_bfloat162 lp = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xffff0000u));
                    hi[cl4 * 3 + ip] = hb;
                    lo[cl4 * 3 + ip] = *reinterpret_cast<const uint32_t *>(&lp);
                }
                if (cl4 >= 1 && a_ok) {  // 6 (cl4 + 1) values so far: chunk j = cl4 - 1 (values 8j .. 8j+7) is complete
                    const int j = cl4 - 1, kc = kc0 + j;
                    uint8_t *dst = row_base + (size_t)(kc >> 2) * tile_bytes(TR_A) + (size_t)(kc & 3) * (TR_A * 16);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
        }
    }
}

// ---- up to 32 DISTINCT neighbours: 8 groups of 3 kernel points (512 threads, one CTA per SM), channels gathered 4 at
// a time.  A thread produces 3 values per channel, i.e. 24 values = three K chunks per EIGHT channels:
//     K'(c, k) = (c/8)*192 + (k/3)*24 + (c%8)*3 + (k%3)
// The row may have up to DEDUP_MAX_RAW slots (the rotation / 3DMatch models ask for K = 64 neighbours,
// SPConvNets/models/reg_so3net.py:100-171, inv_so3net_pn.py:108-113): what counts is the number of distinct ones.
// A point with more than 32 distinct neighbours is left to inter_group_direct64_kernel (launched next to this one
// whenever the row has more than 32 slots), so every point is produced by exactly one of the two kernels.
constexpr int G3_NN = 32, G3_KG = 3, G3_CCH = 4;
constexpr int G3_THR = GD_LANES * (GD_KS / G3_KG);  // 512

// weights of (anchor aa, kernel points k0..k0+KG) against neighbours n_first .. n_first+NN of the list, as
// (even, odd) pairs with the multiplicity folded in; neighbours >= nn and dead lanes get zero weights
template <int KG, int NN>
__device__ __forceinline__ void direct_weights(uint64_t (&w2)[KG][NN / 2], const InterGeom &g, const float *s_g,
                                               const float *s_mult, int aa, bool a_ok, int k0, int n_first, int nn) {
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + aa * 9 + i);
    const float inv_sigma = 1.0f / g.sigma;
    const uint64_t nis2 = pack_f32x2(-inv_sigma, -inv_sigma);
    (void)a_ok; (void)nn;  // absent neighbours carry multiplicity 0 in the list; dead lanes never store
#pragma unroll
    for (int i = 0; i < KG; ++i) {
        const float kx = __ldg(g.kernels + (k0 + i) * 3), ky = __ldg(g.kernels + (k0 + i) * 3 + 1),
                    kz = __ldg(g.kernels + (k0 + i) * 3 + 2);
        const KPoint2 rk = kpoint2(R[0] * kx + R[1] * ky + R[2] * kz, R[3] * kx + R[4] * ky + R[5] * kz,
                                   R[6] * kx + R[7] * ky + R[8] * kz);
#pragma unroll
        for (int n = 0; n < NN; n += 2) {
            const int m = n_first + n;
            w2[i][n / 2] = kernel_weight_pair(pack_f32x2(s_g[m * 3], s_g[m * 3 + 3]), pack_f32x2(s_g[m * 3 + 1], s_g[m * 3 + 4]),
                                              pack_f32x2(s_g[m * 3 + 2], s_g[m * 3 + 5]), rk, nis2, pack_f32x2(s_mult[m], s_mult[m + 1]));
        }
    }
}

__global__ void __launch_bounds__(G3_THR, 1)
inter_group_direct32_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx, InterGeom g,
                            uint8_t *__restrict__ tiles, int k_blocks, long long cols_per_z, int c, int p_in, int p, int nn,
                            int p_off) {
    constexpr int NN = G3_NN, KG = G3_KG, CCH = G3_CCH, NA = GD_NA, NTHR = G3_THR;
    extern __shared__ __align__(16) float s_dyn[];
    float *Fs = s_dyn;                                    // [2][CCH][NN][NA]
    __shared__ NeighbourList<NN> L;
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x;
    const int a = tid % GD_LANES, grp = tid / GD_LANES;
    const int k0 = grp * KG;
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : a - 4;
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    const float *F = feats + (size_t)z * c * p_in * NA;

    const uint32_t bar0 = smem_u32(&s_bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_barrier_init();
    }
    dedup_row(L, idx + ((size_t)z * p + pi) * nn, nn, g.xyz + (size_t)z * 3 * p_in, g.centers + (size_t)z * 3 * p, p_in, p, pi,
              tid, NTHR, [] { __syncthreads(); });
    const int32_t *s_idx = L.idx;
    nn = L.total;  // number of DISTINCT neighbours
    if (nn > NN) return;  // CTA-uniform: inter_group_direct64_kernel owns this point
    for (int t = tid; t < 2 * CCH * (NN - nn) * NA; t += NTHR) {
        const int e = t % NA, r = t / NA, n = nn + r % (NN - nn), bc = r / (NN - nn);
        Fs[(bc * NN + n) * NA + e] = 0.f;
    }
    __syncthreads();

    uint64_t w2[KG][NN / 2];
    direct_weights<KG, NN>(w2, g, L.g, L.mult, aa, a_ok, k0, 0, nn);

    const int nchunks = c / CCH;  // c % 8 == 0: an even number of whole chunks
    const uint32_t fs_u32 = smem_u32(Fs);
    constexpr uint32_t ROW_BYTES = NA * 4;
    auto issue = [&](int chunk, int buf) {
        const uint32_t bar = bar0 + 8u * (uint32_t)buf;
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(CCH * nn) * ROW_BYTES);
        for (int t = tid; t < CCH * NN; t += NTHR) {
            const int cl = t / NN, n = t % NN;
            if (n < nn)
                bulk_g2s(fs_u32 + (uint32_t)(((buf * CCH + cl) * NN + n) * NA) * 4u,
                         F + ((size_t)(chunk * CCH + cl) * p_in + s_idx[n]) * NA, ROW_BYTES, bar);
        }
    };

    const long long row = (long long)z * cols_per_z + (long long)pl * NA + aa;
    uint8_t *row_base = tiles + ((size_t)(row >> 7) * k_blocks) * tile_bytes(TR_A) + (size_t)(row & 127) * 16;
    auto store_chunk = [&](int kc, uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t l0, uint32_t l1, uint32_t l2,
                           uint32_t l3) {
        uint8_t *dst = row_base + (size_t)(kc >> 2) * tile_bytes(TR_A) + (size_t)(kc & 3) * (TR_A * 16);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(h0, h1, h2, h3);
        *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = make_uint4(l0, l1, l2, l3);
    };

    uint32_t phase_bits = 0u;
    issue(0, 0);
#pragma unroll 1
    for (int blk = 0; blk < nchunks / 2; ++blk) {
        uint32_t hi[12], lo[12];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int chunk = blk * 2 + half, buf = half;  // chunk parity == half
            __syncthreads();  // every thread is done with the buffer the next gather overwrites
            if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
            mbar_wait(bar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
            phase_bits ^= 1u << buf;
            const float *fbase = Fs + (size_t)(buf * CCH * NN) * NA + aa;
#pragma unroll
            for (int cp = 0; cp < CCH / 2; ++cp) {  // two channels -> 6 values -> 3 packed pairs
                uint64_t acc2[2][KG];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < KG; ++i) acc2[h][i] = 0ull;
#pragma unroll
                for (int n4 = 0; n4 < NN; n4 += 4) {
                    if (n4 < nn) {  // CTA-uniform
#pragma unroll
                        for (int n = n4; n < n4 + 4; n += 2) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float *frow = fbase + (cp * 2 + h) * NN * NA;
                                const uint64_t f2 = pack_f32x2(frow[n * NA], frow[(n + 1) * NA]);
#pragma unroll
                                for (int i = 0; i < KG; ++i) acc2[h][i] = fma_f32x2(w2[i][n / 2], f2, acc2[h][i]);
                            }
                        }
                    }
                }
                float v[6];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < KG; ++i) {
                        float e, o;
                        unpack_f32x2(acc2[h][i], e, o);
                        v[h * KG + i] = e + o;
                    }
#pragma unroll
                for (int ip = 0; ip < 3; ++ip) {
                    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * ip], v[2 * ip + 1]);
                    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&hp);
                    const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * ip] - __uint_as_float(hb << 16),
                                                                    v[2 * ip + 1] - __uint_as_float(hb & 0xffff0000u));
                    hi[half * 6 + cp * 3 + ip] = hb;
                    lo[half * 6 + cp * 3 + ip] = *reinterpret_cast<const uint32_t *>(&lp);
                }
            }
            const int kc0 = blk * 24 + grp * 3;  // first of this thread's three K' chunks of the 8-channel block
            if (a_ok) {
                if (half == 0) {
                    store_chunk(kc0, hi[0], hi[1], hi[2], hi[3], lo[0], lo[1], lo[2], lo[3]);
                } else {
                    store_chunk(kc0 + 1, hi[4], hi[5], hi[6], hi[7], lo[4], lo[5], lo[6], lo[7]);
                    store_chunk(kc0 + 2, hi[8], hi[9], hi[10], hi[11], lo[8], lo[9], lo[10], lo[11]);
                }
            }
        }
    }
}

// ---- 33 .. 64 DISTINCT neighbours.  The kernel weights of such a point (24 x 64 x 60 = 92 k values) no longer fit
// the register file of one SM, so TWO CTAs share a point: blockIdx.z owns 12 of the 24 kernel points (4 groups of
// 3).  Inside a CTA the freed thread slots split the neighbours instead: thread = (anchor lane, kernel-point group,
// neighbour half) keeps w[3][32] in registers exactly like the kernel above, the upper half hands its partial sums
// over through shared memory once per 4-channel chunk, and the lower half converts and stores.  Same K' order, same
// tile rows: the two kernels are interchangeable point by point.
constexpr int G6_NN = 64, G6_NH = 32;

__global__ void __launch_bounds__(G3_THR, 1)
inter_group_direct64_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx, InterGeom g,
                            uint8_t *__restrict__ tiles, int k_blocks, long long cols_per_z, int c, int p_in, int p, int nn,
                            int p_off) {
    constexpr int NN = G6_NN, NH = G6_NH, KG = G3_KG, CCH = G3_CCH, NA = GD_NA, NTHR = G3_THR;
    extern __shared__ __align__(16) float s_dyn[];
    float *Fs = s_dyn;                                    // [2][CCH][NN][NA]
    float *Rs = Fs + 2 * CCH * NN * NA;                   // [CCH*KG][4 groups][64 lanes] partial sums of the upper half
    __shared__ NeighbourList<NN> L;
    __shared__ __align__(8) uint64_t s_bar[2];
    const int tid = threadIdx.x;
    const int a = tid % GD_LANES, r8 = tid / GD_LANES;
    const int kgl = r8 & 3, nh = r8 >> 2;                 // local kernel-point group, neighbour half
    const int kgrp = (int)blockIdx.z * 4 + kgl;           // 0..7, as in the 32-neighbour kernel
    const int k0 = kgrp * KG;
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : a - 4;
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    const float *F = feats + (size_t)z * c * p_in * NA;

    const uint32_t bar0 = smem_u32(&s_bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_barrier_init();
    }
    dedup_row(L, idx + ((size_t)z * p + pi) * nn, nn, g.xyz + (size_t)z * 3 * p_in, g.centers + (size_t)z * 3 * p, p_in, p, pi,
              tid, NTHR, [] { __syncthreads(); });
    const int32_t *s_idx = L.idx;
    nn = L.total;
    if (nn <= G3_NN) return;  // CTA-uniform: inter_group_direct32_kernel owns this point
    for (int t = tid; t < 2 * CCH * (NN - nn) * NA; t += NTHR) {
        const int e = t % NA, r = t / NA, n = nn + r % (NN - nn), bc = r / (NN - nn);
        Fs[(bc * NN + n) * NA + e] = 0.f;
    }
    __syncthreads();

    uint64_t w2[KG][NH / 2];
    direct_weights<KG, NH>(w2, g, L.g, L.mult, aa, a_ok, k0, nh * NH, nn);
    const int nloc = nn - nh * NH;  // neighbours of this thread's half that exist (> 0 for the lower half)

    const int nchunks = c / CCH;
    const uint32_t fs_u32 = smem_u32(Fs);
    constexpr uint32_t ROW_BYTES = NA * 4;
    auto issue = [&](int chunk, int buf) {
        const uint32_t bar = bar0 + 8u * (uint32_t)buf;
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(CCH * nn) * ROW_BYTES);
        for (int t = tid; t < CCH * NN; t += NTHR) {
            const int cl = t / NN, n = t % NN;
            if (n < nn)
                bulk_g2s(fs_u32 + (uint32_t)(((buf * CCH + cl) * NN + n) * NA) * 4u,
                         F + ((size_t)(chunk * CCH + cl) * p_in + s_idx[n]) * NA, ROW_BYTES, bar);
        }
    };

    const long long row = (long long)z * cols_per_z + (long long)pl * NA + aa;
    uint8_t *row_base = tiles + ((size_t)(row >> 7) * k_blocks) * tile_bytes(TR_A) + (size_t)(row & 127) * 16;
    auto store_chunk = [&](int kc, uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t l0, uint32_t l1, uint32_t l2,
                           uint32_t l3) {
        uint8_t *dst = row_base + (size_t)(kc >> 2) * tile_bytes(TR_A) + (size_t)(kc & 3) * (TR_A * 16);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(h0, h1, h2, h3);
        *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = make_uint4(l0, l1, l2, l3);
    };
    float *rs_mine = Rs + kgl * GD_LANES + a;  // element i of this (group, lane): rs_mine[i * 256]

    uint32_t phase_bits = 0u;
    issue(0, 0);
#pragma unroll 1
    for (int blk = 0; blk < nchunks / 2; ++blk) {
        uint32_t hi[12], lo[12];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int chunk = blk * 2 + half, buf = half;
            __syncthreads();  // gather buffer and Rs of the previous chunk are free
            if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
            mbar_wait(bar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
            phase_bits ^= 1u << buf;
            const float *fbase = Fs + (size_t)(buf * CCH * NN + nh * NH) * NA + aa;
            float v[CCH * KG];
#pragma unroll
            for (int cp = 0; cp < CCH / 2; ++cp) {
                uint64_t acc2[2][KG];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < KG; ++i) acc2[h][i] = 0ull;
#pragma unroll
                for (int n4 = 0; n4 < NH; n4 += 4) {
                    if (n4 < nloc) {  // warp-uniform
#pragma unroll
                        for (int n = n4; n < n4 + 4; n += 2) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float *frow = fbase + (cp * 2 + h) * NN * NA;
                                const uint64_t f2 = pack_f32x2(frow[n * NA], frow[(n + 1) * NA]);
#pragma unroll
                                for (int i = 0; i < KG; ++i) acc2[h][i] = fma_f32x2(w2[i][n / 2], f2, acc2[h][i]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < KG; ++i) {
                        float e, o;
                        unpack_f32x2(acc2[h][i], e, o);
                        v[(cp * 2 + h) * KG + i] = e + o;
                    }
            }
            if (nh == 1) {
#pragma unroll
                for (int i = 0; i < CCH * KG; ++i) rs_mine[i * 256] = v[i];
            }
            __syncthreads();
            if (nh == 0) {
#pragma unroll
                for (int i = 0; i < CCH * KG; ++i) v[i] += rs_mine[i * 256];
#pragma unroll
                for (int ip = 0; ip < 6; ++ip) {
                    const __nv_bfloat162 hp = __floats2bfloat162_rn(v[2 * ip], v[2 * ip + 1]);
                    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&hp);
                    const __nv_bfloat162 lp = __floats2bfloat162_rn(v[2 * ip] - __uint_as_float(hb << 16),
                                                                    v[2 * ip + 1] - __uint_as_float(hb & 0xffff0000u));
                    hi[half * 6 + ip] = hb;
                    lo[half * 6 + ip] = *reinterpret_cast<const uint32_t *>(&lp);
                }
                const int kc0 = blk * 24 + kgrp * 3;
                if (a_ok) {
                    if (half == 0) {
                        store_chunk(kc0, hi[0], hi[1], hi[2], hi[3], lo[0], lo[1], lo[2], lo[3]);
                    } else {
                        store_chunk(kc0 + 1, hi[4], hi[5], hi[6], hi[7], lo[4], lo[5], lo[6], lo[7]);
                        store_chunk(kc0 + 2, hi[8], hi[9], hi[10], hi[11], lo[8], lo[9], lo[10], lo[11]);
                    }
                }
            }
        }
    }
}

// ---- one input channel (layer 0 of every model: occupancy features == 1, or a real single-channel tensor):
//     G[k, (z,p,a)] = sum_u m_u w(p,a,k,u) f[z,0,q_u,a]
// Nothing worth staging: thread = (anchor lane, 8 kernel points) walks the distinct neighbours, derives each weight
// once and accumulates; any row length up to DEDUP_MAX_RAW (the 3DMatch model's first layer asks for K = 128,
// SPConvNets/models/inv_so3net_pn.py:112-113).  Tiles: K = 24 in plain order, padded to one 32-wide block.
__global__ void __launch_bounds__(GD_LANES * 3)
inter_group_occ_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx, InterGeom g,
                       uint8_t *__restrict__ tiles, long long cols_per_z, int p_in, int p, int nn, int p_off, int fmt) {
    constexpr int NA = GD_NA, NTHR = GD_LANES * 3;
    __shared__ NeighbourList<DEDUP_MAX_RAW> L;
    const int tid = threadIdx.x;
    const int a = tid % GD_LANES, j = tid / GD_LANES;  // j: K chunk (kernel points 8j .. 8j+7)
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : a - 4;
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    // 192 threads = 6 warps >= ceil(128 / 32) chunks
    dedup_row(L, idx + ((size_t)z * p + pi) * nn, nn, g.xyz + (size_t)z * 3 * p_in, g.centers + (size_t)z * 3 * p, p_in, p, pi,
              tid, NTHR, [] { __syncthreads(); });
    nn = L.total;
    float rk[8][3];
    {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + aa * 9 + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = j * 8 + i;
            const float kx = __ldg(g.kernels + k * 3), ky = __ldg(g.kernels + k * 3 + 1), kz = __ldg(g.kernels + k * 3 + 2);
            rk[i][0] = R[0] * kx + R[1] * ky + R[2] * kz;
            rk[i][1] = R[3] * kx + R[4] * ky + R[5] * kz;
            rk[i][2] = R[6] * kx + R[7] * ky + R[8] * kz;
        }
    }
    const float inv_sigma = 1.0f / g.sigma;
    const float *F = feats ? feats + (size_t)z * p_in * NA + aa : nullptr;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int n = 0; n < nn; ++n) {
        const float gx = L.g[n * 3], gy = L.g[n * 3 + 1], gz = L.g[n * 3 + 2];
        float f = L.mult[n];
        if (F != nullptr) f *= __ldg(F + (size_t)L.idx[n] * NA);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(kernel_weight_fast(gx, gy, gz, rk[i][0], rk[i][1], rk[i][2], inv_sigma), f, acc[i]);
    }
    if (!a_ok) return;
    uint4 hi, lo;
    split8_fmt(acc, hi, lo, fmt);
    const long long row = (long long)z * cols_per_z + (long long)pl * NA + aa;
    uint8_t *dst = tiles + (size_t)(row >> 7) * tile_bytes(TR_A) + (size_t)(row & 127) * 16 + (size_t)j * (TR_A * 16);
    *reinterpret_cast<uint4 *>(dst) = hi;
    *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = lo;
    if (j == 0) {  // k = 24..31 of the block: zero padding
        *reinterpret_cast<uint4 *>(dst + 3 * (TR_A * 16)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(dst + 3 * (TR_A * 16) + part_bytes(TR_A)) = make_uint4(0u, 0u, 0u, 0u);
    }
}
}  // namespace

// Weight tiles of the forward GEMM (rows = c_out in trb-row tiles) with K in the permuted order K'(c,k).
__global__ void __launch_bounds__(256)
inter_w_tiles_kperm_kernel(const float *__restrict__ W, uint8_t *__restrict__ dst, int c_out, int ck, int trb, int k_blocks,
                           int mode, int steps, int fmt, float scale) {
    const int row = blockIdx.x * 32 + (threadIdx.x & 31), kcg = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int rows_pad = (c_out + trb - 1) / trb * trb;
    if (row >= rows_pad || kcg >= k_blocks * (KB / 8)) return;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int kp = kcg * 8 + i;
        x[i] = (row < c_out && kp < ck) ? __ldg(W + (size_t)row * ck + inter_kperm_inv(kp, mode)) : 0.f;
    }
    uint4 hi, lo;
    split8_fmt(x, hi, lo, fmt, scale);
    const int rt = row / trb, r = row - rt * trb;
    if (steps) {
        // "step" layout of the fused inter kernel (one row tile): every 16-wide k step is one contiguous block
        // [hi: 2 k-chunks x trb rows x 16 B][lo: same] so that ONE bulk copy moves a pipeline stage
        uint8_t *blk = dst + (size_t)(kcg >> 1) * trb * 64 + (size_t)(kcg & 1) * trb * 16 + (size_t)r * 16;
        *reinterpret_cast<uint4 *>(blk) = hi;
        *reinterpret_cast<uint4 *>(blk + (size_t)trb * 32) = lo;
        return;
    }
    uint8_t *tile = dst + ((size_t)rt * k_blocks + (kcg >> 2)) * tile_bytes(trb);
    *reinterpret_cast<uint4 *>(tile + (size_t)(kcg & 3) * trb * 16 + (size_t)r * 16) = hi;
    *reinterpret_cast<uint4 *>(tile + part_bytes(trb) + (size_t)(kcg & 3) * trb * 16 + (size_t)r * 16) = lo;
}

int launch_inter_w_tiles_kperm(const float *W, void *dst, int c_out, int ck, int trb, int mode, int steps, cudaStream_t s,
                               int fmt) {
    const int k_blocks = (ck + KB - 1) / KB, rows_pad = (c_out + trb - 1) / trb * trb;
    dim3 grid((rows_pad + 31) / 32, (k_blocks * (KB / 8) + 7) / 8);
    ProfScope prof(s, KC_SPLIT);
    inter_w_tiles_kperm_kernel<<<grid, 256, 0, s>>>(W, static_cast<uint8_t *>(dst), c_out, ck, trb, k_blocks, mode, steps, fmt,
                                                    fmt == FMT_F16 ? F16_W_SCALE : 1.0f);
    return check_launch("inter_w_tiles_kperm_kernel");
}

// 0: not covered; 1: row of <= 16 slots (K' over 4-channel blocks); 2: row of <= 64 slots (K' over 8-channel blocks)
int inter_group_direct_mode(const float *feats, int c, int nn, int na, int ks) {
    if (feats == nullptr || ks != GD_KS || na != GD_NA) return 0;
    if (nn <= GD_NN && c % 4 == 0) return 1;
    if (nn <= G6_NN && c % 8 == 0) return 2;
    return 0;
}

// Returns 1 when the shape is not covered.  Tiles: rows = (z, pl, a) columns, K in the permuted order K'(c,k).
int launch_inter_group_direct(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, int k_blocks,
                              long long cols_per_z, int p_off, int p_cnt, int bc, int c, int p_in, int p, int nn, int na,
                              int ks, cudaStream_t s) {
    const int mode = inter_group_direct_mode(feats, c, nn, na, ks);
    if (mode == 0 || bc > 65535) return 1;
    static DynSmemOnce once1, once2, once3;
    const size_t smem1 = (size_t)(2 * GD_CCH * GD_NN * GD_NA) * sizeof(float);
    const size_t smem2 = (size_t)(2 * G3_CCH * G3_NN * GD_NA) * sizeof(float);
    const size_t smem3 = (size_t)(2 * G3_CCH * G6_NN * GD_NA + G3_CCH * G3_KG * 4 * GD_LANES) * sizeof(float);
    if (int rc = ensure_dyn_smem(once1, inter_group_direct_kernel, (int)smem1, "inter_group_direct_kernel")) return rc;
    if (int rc = ensure_dyn_smem(once2, inter_group_direct32_kernel, (int)smem2, "inter_group_direct32_kernel")) return rc;
    if (int rc = ensure_dyn_smem(once3, inter_group_direct64_kernel, (int)smem3, "inter_group_direct64_kernel")) return rc;
    dim3 grid(p_cnt, bc);
    ProfScope prof(s, KC_INTER_GROUP);
    if (mode == 1) {
        inter_group_direct_kernel<<<grid, GD_THR, smem1, s>>>(feats, idx, g, static_cast<uint8_t *>(tiles), k_blocks, cols_per_z,
                                                           c, p_in, p, nn, p_off);
        return check_launch("inter_group_direct_kernel");
    }
    inter_group_direct32_kernel<<<grid, G3_THR, smem2, s>>>(feats, idx, g, static_cast<uint8_t *>(tiles), k_blocks,
                                                         cols_per_z, c, p_in, p, nn, p_off);
    if (int rc = check_launch("inter_group_direct32_kernel")) return rc;
    if (nn > G3_NN) {  // points with more than 32 distinct neighbours (the other kernel skipped them)
        dim3 grid2(p_cnt, bc, 2);
        inter_group_direct64_kernel<<<grid2, G3_THR, smem3, s>>>(feats, idx, g, static_cast<uint8_t *>(tiles), k_blocks,
                                                              cols_per_z, c, p_in, p, nn, p_off);
        return check_launch("inter_group_direct64_kernel");
    }
    return 0;
}

// One input channel (feats may be NULL = occupancy ones): tiles of one 32-wide K block in plain order.
bool inter_group_occ_ok(int c, int nn, int na, int ks) { return c == 1 && ks == GD_KS && na == GD_NA && nn <= DEDUP_MAX_RAW; }
int launch_inter_group_occ(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, long long cols_per_z,
                           int p_off, int p_cnt, int bc, int p_in, int p, int nn, int na, int ks, cudaStream_t s, int fmt) {
    if (!inter_group_occ_ok(1, nn, na, ks) || bc > 65535) return 1;
    dim3 grid(p_cnt, bc);
    ProfScope prof(s, KC_INTER_GROUP);
    inter_group_occ_kernel<<<grid, GD_LANES * 3, 0, s>>>(feats, idx, g, static_cast<uint8_t *>(tiles), cols_per_z, p_in, p, nn,
                                                        p_off, fmt);
    return check_launch("inter_group_occ_kernel");
}

}  // namespace epn
