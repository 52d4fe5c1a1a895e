// Fused data gradient of InterSO3Conv: channel GEMM (dG = dout . W^T) + transposed spatial contraction + scatter in ONE
// kernel; the 24x larger gradient of the grouped tensor (dG, 20 GB per training step of the BASELINE network) never
// leaves the SM.
//
//   dfeats[z, c, idx[z,p,n], a] += sum_k w(p,a,k,n) * dG[(c,k), (z,p,a)],   dG[(c,k), row] = sum_o W[o, c*24+k] * dout[z,o,p,a]
//   (autograd of vgtk/vgtk/so3conv/modules.py:48-55 w.r.t. its input, then of spconv/functional.py:361-390 + the
//    scatter_add of the gather, so3conv/functional.py:118-218)
//
// A CTA owns TWO consecutive output points of one cloud = 128 rows of the MMA M dimension (row = pt*64 + anchor,
// anchors 60..63 dead), rows of <= 16 neighbour slots (17..32 slots: two CTAs per point pair, 16 distinct neighbours
// each).  Its dout rows [128 x C_out] are split into bf16 hi/lo ONCE and
// stay in shared memory as the A operand for the whole kernel (128 KB at C_out = 256: affordable because this kernel
// needs neither gather buffers nor an A ring).  W^T streams through a ring in granules of 4 channels x 24 kernel points =
// 96 rows (B operand, N = 96, "step" layout: every 16-o step is one contiguous block), accumulating
//   D[128 rows, 96 columns] = dG of the granule
// into a DOUBLE-BUFFERED TMEM tile (hi*hi products and the cross terms in separate accumulators, see epn_inter_fused.cu).
// The consumers read dG straight from TMEM: a warp can only touch its own lane quarter, so warps w, w+4, w+8, w+12 see the
// same 32 rows -- which is exactly the scatter kernel's thread mapping: thread <-> (row = (point, anchor), group of four
// neighbours), the 24 x 4 kernel weights of that pair in registers, one fp32 RED per (channel, neighbour, anchor).
// No shared-memory staging of dG at all.
// A warp of neighbour group 3 (idle as a consumer unless its point has more than 12 distinct neighbours; the one on the
// point with fewer neighbours) is also the control warp: before it consumes granule g it issues the MMAs of granule g+1
// (and refills the W ring), so the tensor core works on the next granule while the 16 warps scatter the current one.
#include <stdlib.h>

#include "epn_dedup.cuh"
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

namespace {

constexpr int FB_NA = 60, FB_KS = 24, FB_GCH = 4;        // anchors, kernel points, channels per granule
constexpr int FB_GN = FB_GCH * FB_KS;                    // 96 = UMMA N = TMEM columns of one accumulator
constexpr int FB_THREADS = 512, FB_WARPS = 16;
constexpr uint32_t FB_A_LBO = 128 * 16, FB_A_PART = 4 * FB_A_LBO, FB_A_KB = 2 * FB_A_PART;   // 128-row K-major split tile
constexpr uint32_t FB_STEP_BYTES = FB_GN * 64;           // one 16-o step of a granule: [hi: 2 chunks x 96 rows x 16 B][lo]
constexpr uint32_t FB_ACC_STRIDE = 128;                  // TMEM columns between main / cross accumulator, 256 between buffers

struct BwdParams {
    const float *dout;       // element (z, o, pl, a) at dout + z*dout_sz + o*dout_so + pl*60 + a   (pl = point within the call)
    long long dout_sz, dout_so;
    const int32_t *idx;      // [b, p, nn]
    InterGeom g;
    const uint8_t *Wt;       // W^T step tiles: [granule][16-o step][FB_STEP_BYTES]
    float *dfeats;           // [b, c, p_in, 60], pre-zeroed
    int c, c_out, p_in, p, nn, p_off, nst, sps, ctrl_pick;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// predicated reduction (ptxas still emits a short branch around the REDG: it needs the address in uniform registers)
__device__ __forceinline__ void red_add_pred(float *p, float v, uint32_t ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.global.add.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"(ok) : "memory");
}

// PIPE: 1 = TMEM reads double-buffered in groups of 4 columns + predicated reductions; 0 = groups of 8, branches
template <int PIPE>
__global__ void __launch_bounds__(FB_THREADS, 1) inter_bwd_fused_kernel(BwdParams P) {
    constexpr int NA = FB_NA, KS = FB_KS, NB = 4;        // NB neighbours per thread
    extern __shared__ __align__(128) uint8_t smem[];
    const int k_blocks = P.c_out / 32;
    uint8_t *a_tiles = smem;                                        // [k_blocks] split tiles of 128 rows
    uint8_t *ring = smem + (size_t)k_blocks * FB_A_KB;              // W^T ring, nst stages of sps steps
    __shared__ NeighbourList<32> s_L[2];
    __shared__ __align__(8) uint64_t s_wfull[8], s_wempty[8], s_accfull[2], s_accempty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.y;
    const uint32_t stage_bytes = FB_STEP_BYTES * (uint32_t)P.sps;

    if (tid == 0) {
        for (int i = 0; i < P.nst; ++i) {
            mbar_init(smem_u32(&s_wfull[i]), 1);
            mbar_init(smem_u32(&s_wempty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&s_accfull[i]), 1);
            mbar_init(smem_u32(&s_accempty[i]), FB_WARPS);   // one arrival per consumer warp
        }
        fence_barrier_init();
    }
    // ---- distinct neighbours of the two points (one warp per point works, everybody takes part in the barriers)
    {
        const bool worker = tid < 64;
        const int dpt = worker ? tid >> 5 : 0;
        const int dpi = P.p_off + blockIdx.x * 2 + dpt;
        dedup_row(s_L[dpt], P.idx + ((size_t)z * P.p + dpi) * P.nn, worker ? P.nn : 0, P.g.xyz + (size_t)z * 3 * P.p_in,
                  P.g.centers + (size_t)z * 3 * P.p, P.p_in, P.p, dpi, worker ? tid - dpt * 32 : (1 << 20), worker ? 32 : 1,
                  [] { __syncthreads(); });
    }

    // Rows of 17..32 slots: CTA blockIdx.z takes the distinct neighbours [16 z, 16 z + 16) of both points (the scatter
    // is additive over neighbours; dG is recomputed by both CTAs); nothing to do past the count of both points
    const int nbase = blockIdx.z * 16;
    if (nbase > 0 && s_L[0].total <= nbase && s_L[1].total <= nbase) return;
    if (warp == FB_WARPS - 1) tmem_alloc(smem_u32(&s_tmem), 512);

    // ---- A operand: the dout rows of the two points, split once.  task = (8-wide o chunk, row); lanes run along rows
    //      (= along anchors: coalesced 4-byte loads of 60 consecutive floats per (o, point))
    {
        const float *D = P.dout + (size_t)z * P.dout_sz + (size_t)(blockIdx.x * 2) * NA;
        const int ntask = 128 * (P.c_out / 8);
        for (int t = tid; t < ntask; t += FB_THREADS) {
            const int row = t & 127, chunk = t >> 7;
            const int pt = row >> 6, a = row & 63;
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                x[i] = a < NA ? __ldg(D + (size_t)(chunk * 8 + i) * P.dout_so + pt * NA + a) : 0.f;
            uint4 hi, lo;
            split8(x, hi, lo);
            uint8_t *dst = a_tiles + (size_t)(chunk >> 2) * FB_A_KB + (size_t)(chunk & 3) * FB_A_LBO + (size_t)row * 16;
            *reinterpret_cast<uint4 *>(dst) = hi;
            *reinterpret_cast<uint4 *>(dst + FB_A_PART) = lo;
        }
    }
    fence_proxy_async_smem();   // generic-proxy writes of the A tile -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    // ---- this thread's (row, neighbour group): TMEM lane quarter q = warp % 4, group = warp / 4
    const int q = warp & 3, grp = warp >> 2;
    const int row = q * 32 + lane, pt = row >> 6, a = row & 63;
    const bool a_ok = a < NA;
    const int aa = a_ok ? a : NA - 1;
    const NeighbourList<32> &L = s_L[pt];
    const int n0 = nbase + grp * NB;
    const int nn_pt = L.total < nbase + 16 ? L.total : nbase + 16;
    const bool grp_active = n0 < nn_pt;                // warp-uniform (the lane quarter fixes the point)
    uint32_t red_mask = 0;                             // neighbour j of this thread exists (and the row is an anchor)
#pragma unroll
    for (int j = 0; j < 4; ++j) red_mask |= (a_ok && n0 + j < nn_pt) ? (1u << j) : 0u;
    uint64_t w2[KS][NB / 2];                            // (neighbour 2j, 2j+1) pairs
    int qoff[NB];
    if (grp_active) {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(P.g.anchors + aa * 9 + i);
        const float inv_sigma = 1.0f / P.g.sigma;
        const uint64_t nis2 = pack_f32x2(-inv_sigma, -inv_sigma);
        uint64_t gx[NB / 2], gy[NB / 2], gz[NB / 2], mm[NB / 2];
#pragma unroll
        for (int j = 0; j < NB; j += 2) {   // absent neighbours carry multiplicity 0 in the list
            const int n = n0 + j;
            gx[j / 2] = pack_f32x2(L.g[n * 3], L.g[n * 3 + 3]);
            gy[j / 2] = pack_f32x2(L.g[n * 3 + 1], L.g[n * 3 + 4]);
            gz[j / 2] = pack_f32x2(L.g[n * 3 + 2], L.g[n * 3 + 5]);
            mm[j / 2] = pack_f32x2(L.mult[n], L.mult[n + 1]);
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const float kx = __ldg(P.g.kernels + k * 3), ky = __ldg(P.g.kernels + k * 3 + 1), kz = __ldg(P.g.kernels + k * 3 + 2);
            const KPoint2 rk = kpoint2(R[0] * kx + R[1] * ky + R[2] * kz, R[3] * kx + R[4] * ky + R[5] * kz,
                                       R[6] * kx + R[7] * ky + R[8] * kz);
#pragma unroll
            for (int j = 0; j < NB / 2; ++j) w2[k][j] = kernel_weight_pair(gx[j], gy[j], gz[j], rk, nis2, mm[j]);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) qoff[j] = L.idx[n0 + j] * NA + aa;
    }
    float *DF = P.dfeats + (size_t)z * P.c * P.p_in * NA;
    const size_t cplane = (size_t)P.p_in * NA;

    // ---- control state.  Every warp is a consumer and whatever the control warp spends issuing delays the granule
    //      for all 16 (they meet at the accumulator-empty barrier), so the role goes to a warp of the LAST neighbour
    //      group (slots 12..15: idle as a consumer whenever its point has <= 12 distinct neighbours), on the point with
    //      fewer neighbours.  The whole warp runs the role converged, one elected lane issues.
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int left0 = min(max(s_L[0].total - nbase, 0), 16), left1 = min(max(s_L[1].total - nbase, 0), 16);
    const int ctrl_warp = ((P.ctrl_pick & 1) && left0 < left1) ? FB_WARPS - 4 : FB_WARPS - 1;
    const bool is_ctrl = warp_u == ctrl_warp;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int ngran = P.c / FB_GCH;
    const int ksteps = P.c_out / 16;
    const uint32_t sps = (uint32_t)P.sps, nst = (uint32_t)P.nst;
    const int stages_g = ksteps / (int)sps, total_stages = ngran * stages_g;
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t wfull0 = smem_u32(&s_wfull[0]), wempty0 = smem_u32(&s_wempty[0]);
    const uint32_t accfull0 = smem_u32(&s_accfull[0]), accempty0 = smem_u32(&s_accempty[0]);
    const uint8_t *wsrc = P.Wt;
    int loaded = 0, consumed = 0;
    uint32_t lslot = 0, lpar = 1, slot = 0, wpar = 0;
    const uint32_t idesc = instr_desc_bf16_m128(FB_GN);
    const uint64_t a_desc0 = smem_desc(smem_u32(a_tiles), FB_A_LBO, 128);
    const uint64_t b_desc0 = smem_desc(ring_u32, FB_GN * 16, 128);
    auto load_w = [&]() {
        bulk_g2s_expect_elect(ring_u32 + lslot * stage_bytes, wsrc, stage_bytes, wfull0 + 8u * lslot);
        wsrc += stage_bytes;
        if (++lslot == nst) { lslot = 0; lpar ^= 1u; }
        ++loaded;
    };
    // MMAs of granule gg into accumulator buffer gg & 1 (main at +0, cross terms at +FB_ACC_STRIDE)
    auto issue_granule = [&](int gg) {
        const uint32_t buf = (uint32_t)gg & 1u;
        if (P.ctrl_pick & 2) {
            // while the consumers drain the buffer: request every ring stage whose MMAs have completed meanwhile
            while (!mbar_test_wait(accempty0 + 8u * buf, (((uint32_t)gg >> 1) & 1u) ^ 1u))
                while (loaded < total_stages && loaded - consumed < (int)nst && mbar_test_wait(wempty0 + 8u * lslot, lpar)) load_w();
        }
        mbar_wait_q(accempty0 + 8u * buf, (((uint32_t)gg >> 1) & 1u) ^ 1u);   // every consumer has drained this buffer
        tc_fence_after();
        const uint32_t d_main = tmem_u + buf * 256u, d_cross = d_main + FB_ACC_STRIDE;
        uint32_t sub = 0;
        for (int ks = 0; ks < ksteps; ++ks) {
            if (sub == 0) {
                if (loaded == consumed) {   // ring ran dry: the stage to consume has not been requested yet
                    mbar_wait_q(wempty0 + 8u * lslot, lpar);
                    load_w();
                }
                mbar_wait_q(wfull0 + 8u * slot, wpar);
                tc_fence_after();
            }
            // A: k-block ks/2, 16-o half ks&1 (two 8-o chunks)
            const uint64_t a_hi = a_desc0 + (uint64_t)(((uint32_t)(ks >> 1) * FB_A_KB + (uint32_t)(ks & 1) * 2u * FB_A_LBO) >> 4);
            const uint64_t a_lo = a_hi + (uint64_t)(FB_A_PART >> 4);
            const uint64_t b_hi = b_desc0 + (uint64_t)((slot * stage_bytes + sub * FB_STEP_BYTES) >> 4);
            const uint64_t b_lo = b_hi + (uint64_t)((FB_GN * 32u) >> 4);
            const uint32_t acc = ks != 0;
            mma_bf16_ss_elect(d_main, a_hi, b_hi, idesc, acc);
            mma_bf16_ss_elect(d_cross, a_hi, b_lo, idesc, acc);
            mma_bf16_ss_elect(d_cross, a_lo, b_hi, idesc, 1);
            if (++sub == sps) {
                sub = 0;
                mma_commit_elect(wempty0 + 8u * slot);
                if (++slot == nst) { slot = 0; wpar ^= 1u; }
                ++consumed;
                while (loaded < total_stages && loaded - consumed < (int)nst && mbar_test_wait(wempty0 + 8u * lslot, lpar)) load_w();
            }
        }
        mma_commit_elect(accfull0 + 8u * buf);
    };
    if (is_ctrl) {
        while (loaded < total_stages && loaded < (int)nst) load_w();
        issue_granule(0);
    }

    for (int g = 0; g < ngran; ++g) {
        if (is_ctrl && g + 1 < ngran) issue_granule(g + 1);   // the tensor core works on g+1 while everybody scatters g
        const uint32_t buf = (uint32_t)g & 1u;
        mbar_wait_q(accfull0 + 8u * buf, ((uint32_t)g >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_main = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u, t_cross = t_main + FB_ACC_STRIDE;
#pragma unroll 1
        for (int cl = 0; cl < FB_GCH; ++cl) {
            if (grp_active) {
                uint64_t t2[NB / 2];
#pragma unroll
                for (int j = 0; j < NB / 2; ++j) t2[j] = 0ull;
                if constexpr (PIPE == 1) {
                    // 24 columns in 6 groups of 4, double-buffered: the load of group kg+1 is in flight while kg is used
                    uint32_t m[2][4], x[2][4];
                    tmem_ld4(t_main + (uint32_t)(cl * KS), m[0]);
                    tmem_ld4(t_cross + (uint32_t)(cl * KS), x[0]);
    #pragma unroll
                    for (int kg = 0; kg < KS / 4; ++kg) {
                        tmem_ld_wait();
                        if (kg + 1 < KS / 4) {
                            tmem_ld4(t_main + (uint32_t)(cl * KS + kg * 4 + 4), m[(kg + 1) & 1]);
                            tmem_ld4(t_cross + (uint32_t)(cl * KS + kg * 4 + 4), x[(kg + 1) & 1]);
                        }
    #pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float dv = __uint_as_float(m[kg & 1][i]) + __uint_as_float(x[kg & 1][i]);
                            const uint64_t d2 = pack_f32x2(dv, dv);
    #pragma unroll
                            for (int j = 0; j < NB / 2; ++j) t2[j] = fma_f32x2(w2[kg * 4 + i][j], d2, t2[j]);
                        }
                    }
                } else {
#pragma unroll
                    for (int kg = 0; kg < KS / 8; ++kg) {
                        uint32_t m[8], x[8];
                        tmem_ld8(t_main + (uint32_t)(cl * KS + kg * 8), m);
                        tmem_ld8(t_cross + (uint32_t)(cl * KS + kg * 8), x);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float dv = __uint_as_float(m[i]) + __uint_as_float(x[i]);
                            const uint64_t d2 = pack_f32x2(dv, dv);
#pragma unroll
                            for (int j = 0; j < NB / 2; ++j) t2[j] = fma_f32x2(w2[kg * 8 + i][j], d2, t2[j]);
                        }
                    }
                }
                float t[NB];
#pragma unroll
                for (int j = 0; j < NB / 2; ++j) unpack_f32x2(t2[j], t[2 * j], t[2 * j + 1]);
                float *dplane = DF + (size_t)(g * FB_GCH + cl) * cplane;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if constexpr (PIPE == 1) {
                        red_add_pred(dplane + qoff[j], t[j], red_mask & (1u << j));
                    } else {
                        if (red_mask & (1u << j)) atomicAdd(dplane + qoff[j], t[j]);
                    }
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accempty0 + 8u * buf);   // this warp is done with the buffer
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FB_WARPS - 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// W^T of the channel GEMM in the kernel's "step" layout: granule gr = 4 channels = rows (c,k) 96 gr .. 96 gr + 95 (plain
// order), 16-o step j: one block [hi: 2 o-chunks x 96 rows x 16 B][lo: same]
__global__ void __launch_bounds__(256)
inter_wt_steps_kernel(const float *__restrict__ W, uint8_t *__restrict__ dst, int c_out, int ck) {
    const int kp = blockIdx.x * 32 + (threadIdx.x & 31), oc = blockIdx.y * 8 + (threadIdx.x >> 5);   // (c,k) row, 8-o chunk
    if (kp >= ck || oc * 8 >= c_out) return;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ldg(W + (size_t)(oc * 8 + i) * ck + kp);
    uint4 hi, lo;
    split8(x, hi, lo);
    const int gr = kp / FB_GN, r = kp - gr * FB_GN, ksteps = c_out / 16;
    uint8_t *blk = dst + ((size_t)gr * ksteps + (oc >> 1)) * FB_STEP_BYTES + (size_t)(oc & 1) * FB_GN * 16 + (size_t)r * 16;
    *reinterpret_cast<uint4 *>(blk) = hi;
    *reinterpret_cast<uint4 *>(blk + FB_GN * 32) = lo;
}

}  // namespace

bool inter_bwd_fused_ok(int c, int c_out, int p_cnt, int nn, int na, int ks) {
    return ks == FB_KS && na == FB_NA && nn <= 32 && c % FB_GCH == 0 && c >= FB_GCH && c_out % 64 == 0 && c_out <= 256 &&
           p_cnt % 2 == 0;
}

// Returns 1 when the shape is not covered.  wt_scratch: >= 4 * c*24 * c_out bytes (the W^T tile buffer of the workspace).
int launch_inter_bwd_fused(const float *dout, long long dout_stride_z, long long dout_stride_o, const int32_t *idx,
                           const InterGeom &g, const float *W, void *wt_scratch, float *dfeats, int p_off, int p_cnt, int bc,
                           int c, int c_out, int p_in, int p, int nn, int na, int ks, cudaStream_t s) {
    if (!inter_bwd_fused_ok(c, c_out, p_cnt, nn, na, ks) || bc > 65535) return 1;
    const int ck = c * ks;
    {
        ProfScope prof(s, KC_SPLIT);
        dim3 grid(cdiv(ck, 32), cdiv(c_out / 8, 8));
        inter_wt_steps_kernel<<<grid, 256, 0, s>>>(W, static_cast<uint8_t *>(wt_scratch), c_out, ck);
        if (int rc = check_launch("inter_wt_steps_kernel")) return rc;
    }
    BwdParams P;
    P.dout = dout; P.dout_sz = dout_stride_z; P.dout_so = dout_stride_o;
    P.idx = idx; P.g = g; P.Wt = static_cast<const uint8_t *>(wt_scratch); P.dfeats = dfeats;
    P.c = c; P.c_out = c_out; P.p_in = p_in; P.p = p; P.nn = nn; P.p_off = p_off;
    const size_t a_bytes = (size_t)(c_out / 32) * FB_A_KB;
    const size_t budget = 227 * 1024 - 4 * 1024;          // static shared memory (neighbour lists, barriers) comes on top
    // 16-o steps per ring stage (c_out / 16 is a multiple of 4): 24 KB stages
    static const int sps_env = [] { const char *e = getenv("EPN_FB_SPS"); return e ? atoi(e) : 0; }();
    static const int ctrl_env = [] { const char *e = getenv("EPN_FB_CTRL"); return e ? atoi(e) : 3; }();
    P.ctrl_pick = ctrl_env;
    P.sps = (sps_env == 1 || sps_env == 2 || sps_env == 4) ? sps_env : 4;
    const size_t stage = (size_t)FB_STEP_BYTES * P.sps;
    if (a_bytes + 2 * stage > budget) return 1;
    int nst = (int)((budget - a_bytes) / stage);
    if (nst > 8) nst = 8;
    P.nst = nst;
    static const int pipe = [] { const char *e = getenv("EPN_FB_PIPE"); return e ? atoi(e) : 1; }();
    static DynSmemOnce once0, once1;
    dim3 grid(p_cnt / 2, bc, nn > 16 ? 2 : 1);
    ProfScope prof(s, KC_INTER_SCATTER);
    if (pipe == 1) {
        if (int rc = ensure_dyn_smem(once1, inter_bwd_fused_kernel<1>, (int)budget, "inter_bwd_fused_kernel")) return rc;
        inter_bwd_fused_kernel<1><<<grid, FB_THREADS, a_bytes + (size_t)nst * stage, s>>>(P);
    } else {
        if (int rc = ensure_dyn_smem(once0, inter_bwd_fused_kernel<0>, (int)budget, "inter_bwd_fused_kernel")) return rc;
        inter_bwd_fused_kernel<0><<<grid, FB_THREADS, a_bytes + (size_t)nst * stage, s>>>(P);
    }
    return check_launch("inter_bwd_fused_kernel");
}

}  // namespace epn
