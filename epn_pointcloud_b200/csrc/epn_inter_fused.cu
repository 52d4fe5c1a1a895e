// Fused InterSO3Conv forward: neighbour gather + kernel weights + spatial contraction + channel GEMM in ONE
// kernel; the grouped tensor G never leaves the SM.
//
//   out[z,o,p,a] = sum_{c,k} W[o, c*24+k] * G[(c,k),(z,p,a)],   G = sum_n w(p,a,k,n) * feats[z,c,idx[z,p,n],a]
//   (vgtk/vgtk/so3conv/modules.py:157-174 -> so3conv/functional.py:118-218 -> spconv/functional.py:361-390
//    -> so3conv/modules.py:48-55)
//
// A CTA owns PTS consecutive output points of one cloud (PTS*64 rows of the MMA M dimension, row = pt*64 + anchor,
// anchors 60..63 of a point are dead rows).  512 threads = 480 producers + one control warp: 16 warps is what lets
// every thread have 128 registers (4 warps per SM sub-partition; a 17th warp would cap the kernel at 96 and spill
// the weights), so the producers are packed 60 anchors per kernel-point group with no dead lanes.
// The PRODUCERS are the in-register-split grouping kernels of
// epn_group_direct.cu (thread <-> (anchor, group of KG kernel points), kernel weights in registers as
// fp32x2 pairs, bulk-copy gather of the distinct neighbours' feature rows, FFMA2 contraction, bf16 hi/lo split as the
// values leave the FMA loop) -- but their 16-byte operand pieces go to a double-buffered A tile in SHARED memory
// (canonical K-major UMMA layout, K in the permuted order K'(c,k) of epn_internal.cuh) instead of global memory.
// One control warp streams the weight tiles (bf16 hi/lo, rows = c_out, same K' order) from L2 through a ring of
// 16-k stages with bulk copies and issues three tcgen05.mma (hi*hi, hi*lo, lo*hi) per 16-wide k step into a
// [128 x c_out] fp32 accumulator in TMEM, one "granule" (4 or 8 channels) behind the producers.  The epilogue (all
// producer warps) reads TMEM and stores out[z,o,p,.] rows (128 contiguous bytes per warp and output channel).
//   MODE 1: rows of <= 16 slots, 6 kernel points per thread, TWO points per CTA, granule = 4 channels (96 K').
//   MODE 2: up to 32 distinct neighbours per pass, 3 kernel points per thread, ONE point per CTA (rows 64..127 of
//           the MMA are dead), granule = 8 channels (192 K').  A point with more distinct neighbours (rows of up to
//           128 slots: the K = 64 layers of the rotation / 3DMatch models) simply runs further passes over the next
//           32 neighbours, accumulating into the same TMEM tile (the GEMM is linear in G).
// In training the producers also write their pieces to global memory (the operand tiles the weight-gradient GEMM
// consumes, epn_gemm_dw.cu); that needs the complete G of a point in one pass, i.e. rows of <= 32 slots.
#include <stdlib.h>

#include "epn_dedup.cuh"
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

namespace {

constexpr int FU_LANES = 60;  // anchors per kernel-point group: no dead lanes (see the header comment)
constexpr int FU_KS = 24;     // kernel points (kpsphere24)
constexpr int FU_NA = 60;
constexpr int FU_CCH = 4;     // channels per gather chunk
constexpr int FU_CTRL = 32;   // control warp

template <int MODE> struct FuCfg;
template <> struct FuCfg<1> { static constexpr int NN = 16, KG = 6, PTS = 2, GCH = 4, CAP = 16; };
template <> struct FuCfg<2> { static constexpr int NN = 32, KG = 3, PTS = 1, GCH = 8, CAP = DEDUP_MAX_RAW; };
// MODE 3 ("halves"): ONE point per CTA whose distinct neighbours are dealt to two pseudo-points (rows 0..63 and
// 64..127 of the MMA); the producers are MODE 1's (6 kernel points per thread: one shared-memory load feeds twice
// the FMAs of MODE 2), the channel GEMM is linear in G so the two partial results are simply summed in the epilogue.
template <> struct FuCfg<3> { static constexpr int NN = 16, KG = 6, PTS = 2, GCH = 4, CAP = DEDUP_MAX_RAW; };

struct FusedParams {
    const float *feats;      // [b, c, p_in, 60]
    const int32_t *idx;      // [b, p, nn]
    InterGeom g;
    const uint8_t *Wt;       // weight tiles [k_blocks] of trb rows (split tiles, rows = c_out, K = K'(c,k))
    float *out;              // out + z*out_sz + o*out_so + pl*60 + a
    long long out_sz, out_so;
    uint8_t *keep;           // optional: forward operand tiles in global memory (rows = (z,pl,a), K = K'(c,k))
    int keep_k_blocks, keep_slab_clouds;   // clouds z are stored in slabs of keep_slab_clouds, each slab a tile matrix
    long long keep_cols_per_z;
    size_t keep_slab_bytes;
    int c, c_out, p_in, p, nn, p_off, trb, nst, sps;   // sps: 16-k steps per weight-ring stage (1, 2 or 3)
    int na;                  // anchors of the tensors in global memory (<= 60, multiple of 4); the kernel's lane / shared
                             // memory geometry is always the 60-anchor one, lanes >= na are dead (anchor subsets:
                             // so3conv/functional.py:281-289, the sweep's "first 12 anchors")
    uint32_t tmem_cols;      // columns of ONE accumulator; the kernel allocates two (see acc_cross)
    float out_scale;         // accumulator -> out factor (1 / F16_W_SCALE for fp16 operands, else 1)
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// GATHER: how the neighbours' 240-byte feature rows reach shared memory
//   0  one bulk copy (UBLKCP) per row, issued by the first 64 (MODE 1) / 128 (MODE 2) threads of the point
//   1  one bulk copy per row, rows dealt round-robin to ALL threads of the point (a warp issues its lanes' copies
//      one at a time, so spreading them shortens the slowest warp's issue phase)
//   2  15 cp.async of 16 bytes per row by all threads + mbarrier arrive.noinc
// FMT: operand format of the A pieces and the weight tiles (epn_umma.cuh: FMT_BF16, or FMT_F16 for inference forwards)
template <int MODE, int GATHER, int FMT>
__global__ void __launch_bounds__(FuCfg<MODE>::PTS *FU_LANES *(FU_KS / FuCfg<MODE>::KG) + FU_CTRL, 1)  // 480 + 32
inter_fused_kernel(FusedParams P) {
    using C = FuCfg<MODE>;
    constexpr int NN = C::NN, KG = C::KG, PTS = C::PTS, GCH = C::GCH, NA = FU_NA, CCH = FU_CCH;
    constexpr bool HALVES = MODE == 3;            // the PTS row blocks are neighbour subsets of ONE point
    constexpr int NLIST = HALVES ? 1 : PTS;       // neighbour lists (= real points) of the CTA
    constexpr int PASS_NN = HALVES ? 2 * NN : NN; // distinct neighbours one pass covers
    constexpr int GROUPS = FU_KS / KG;
    constexpr int PT_THR = FU_LANES * GROUPS;     // producer threads per point
    constexpr int NPROD = PTS * PT_THR;           // producer threads (480 in both modes)
    constexpr int NPWARPS = NPROD / 32;           // 15
    constexpr int NWARPS = NPWARPS + 1;
    static_assert(NPROD == 480, "15 producer warps + the control warp");
    constexpr int GK = GCH * FU_KS;               // K' values per granule
    constexpr int KBG = GK / 32;                  // k-blocks per granule
    constexpr int STEPS_G = KBG * 2;              // 16-k steps per granule
    constexpr int CH_G = GCH / CCH;               // gather chunks per granule
    constexpr int ROWS = PTS * 64;                // rows of the A tiles that exist in shared memory
    constexpr uint32_t A_LBO = ROWS * 16;         // bytes between 8-wide k chunks
    constexpr uint32_t A_PART = 4 * A_LBO;        // hi (or lo) part of one k-block
    constexpr uint32_t A_KB = 2 * A_PART;         // one k-block
    constexpr uint32_t A_BUF = KBG * A_KB;        // one granule
    constexpr uint32_t A_BYTES = 2 * A_BUF + (PTS == 1 ? 1024 : 0);  // PTS == 1: the M=128 MMA over-reads 64 dead rows

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *a_tiles = smem;                                               // [2][KBG] k-blocks
    float *Fs = reinterpret_cast<float *>(smem + A_BYTES);                 // [PTS][2][CCH][NN][NA]
    uint8_t *ring = reinterpret_cast<uint8_t *>(Fs + PTS * 2 * CCH * NN * NA);  // weight ring, nst stages
    __shared__ NeighbourList<C::CAP> s_L[NLIST];
    __shared__ __align__(8) uint64_t s_gbar[PTS][2];   // gather buffers: bytes landed
    __shared__ __align__(8) uint64_t s_wfull[8], s_wempty[8];
    __shared__ __align__(8) uint64_t s_afull[2], s_afree[2], s_accum;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const bool is_ctrl = tid >= NPROD;
    const int z = blockIdx.y;
    const uint32_t step_bytes = (uint32_t)P.trb * 64u;   // 16 k of hi + 16 k of lo
    const uint32_t stage_bytes = step_bytes * (uint32_t)P.sps;
    const int ngran = P.c / GCH;                          // granules per pass

    if (tid == 0) {
        for (int i = 0; i < PTS; ++i) {
            // GATHER 2: one (cp.async-tracking) arrival per thread of the point; else one arrive.expect_tx
            mbar_init(smem_u32(&s_gbar[i][0]), GATHER == 2 ? PT_THR : 1);
            mbar_init(smem_u32(&s_gbar[i][1]), GATHER == 2 ? PT_THR : 1);
        }
        for (int i = 0; i < P.nst; ++i) {
            mbar_init(smem_u32(&s_wfull[i]), 1);
            mbar_init(smem_u32(&s_wempty[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&s_afull[i]), NPWARPS);   // one arrival per producer warp
            mbar_init(smem_u32(&s_afree[i]), 1);
        }
        mbar_init(smem_u32(&s_accum), 1);
        fence_barrier_init();
    }
    // Two fp32 accumulators in TMEM: the hi*hi products go to the first, the two cross products (hi*lo, lo*hi, ~2^-9 /
    // 2^-11 of the former) to the second, and the epilogue adds them.  The tensor core TRUNCATES the accumulator after
    // every MMA (measured: the error of a K-long contraction grows linearly with the number of MMAs, 6e-6 at K = 1536,
    // 1.2e-5 at K = 3072, against 4e-7 at K = 64); a small addend costs the big accumulator as much as a large one, so
    // keeping the cross terms apart removes two of the three truncations per k step from the sum that matters.
    if (warp == NPWARPS) tmem_alloc(smem_u32(&s_tmem), 2 * P.tmem_cols);

    // ---- distinct neighbours of the PTS points: DW whole warps per point do the work, every thread of the CTA takes
    //      part in the three barriers of dedup_row
    {
        constexpr int DW = MODE == 1 ? 1 : 4;                      // warps per point (rows of <= 16 / <= 128 slots)
        const bool worker = tid < NLIST * DW * 32;
        const int dpt = worker ? tid / (DW * 32) : 0;
        const int dpi = P.p_off + blockIdx.x * NLIST + dpt;
        dedup_row(s_L[dpt], P.idx + ((size_t)z * P.p + dpi) * P.nn, worker ? P.nn : 0, P.g.xyz + (size_t)z * 3 * P.p_in,
                  P.g.centers + (size_t)z * 3 * P.p, P.p_in, P.p, dpi, worker ? tid - dpt * DW * 32 : (1 << 20),
                  worker ? DW * 32 : 1, [] { __syncthreads(); });
    }
    const int pt = is_ctrl ? 0 : tid / PT_THR;
    const int ptid = tid - pt * PT_THR;
    const int pl = HALVES ? blockIdx.x : blockIdx.x * PTS + pt;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    // passes over the distinct neighbours (MODE 1: always one; MODE 2: ceil(distinct / 32), CTA-uniform as PTS == 1)
    const int npass = MODE == 1 ? 1 : max(1, (s_L[0].total + PASS_NN - 1) / PASS_NN);
    const int total_gran = npass * ngran;

    // warp index as a value the compiler can prove warp-uniform (shuffle from lane 0): the control warp's loop below
    // then stays converged and its descriptors / barrier addresses live in uniform registers
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    if (warp_u == NPWARPS) {
        // ------------------------------------------------------------ control warp: W ring + MMA issue
        // The WHOLE warp runs this loop converged; one elected lane issues each MMA / commit / bulk copy
        // (epn_umma.cuh, "warp-converged issue").  The weight ring is refilled from the same loop: a slot is reloaded
        // as soon as a non-blocking test shows that the MMAs that read it have completed (blocking only when the
        // stage about to be consumed has not been requested yet), so up to nst - 1 loads stay in flight.
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t tmem_x = tmem_u + P.tmem_cols;   // accumulator of the cross terms
        const uint32_t sps = (uint32_t)P.sps;                      // 16-k steps per ring stage (divides STEPS_G)
        const int stages_g = STEPS_G / (int)sps;                   // ring stages per granule
        const int total_stages = total_gran * stages_g, stages_pass = ngran * stages_g;
        const uint32_t ring_u32 = smem_u32(ring);
        const uint32_t nst = (uint32_t)P.nst;
        const uint32_t wfull0 = smem_u32(&s_wfull[0]), wempty0 = smem_u32(&s_wempty[0]);
        // weight tiles in "step" layout (launch_inter_w_tiles_kperm, steps = 1): 16-k step jw is ONE contiguous
        // block [hi: 2 k-chunks x trb rows][lo: ...] of step_bytes, so a stage of `sps` steps is one bulk copy
        const uint8_t *wsrc = P.Wt;
        int jw = 0;                      // stage within the pass (every pass re-streams W)
        int loaded = 0;                  // stages requested so far
        uint32_t lslot = 0, lpar = 1;    // first round: the slots are empty (the test of the preceding phase passes)
        auto load_w = [&]() {
            bulk_g2s_expect_elect(ring_u32 + lslot * stage_bytes, wsrc, stage_bytes, wfull0 + 8u * lslot);
            wsrc += stage_bytes;
            if (++jw == stages_pass) { jw = 0; wsrc = P.Wt; }
            if (++lslot == nst) { lslot = 0; lpar ^= 1u; }
            ++loaded;
        };
        const uint32_t half_bytes = (uint32_t)P.trb * 32u;
        const uint32_t idesc = instr_desc_m128(P.trb, FMT);
        const uint32_t b_lbo = (uint32_t)P.trb * 16u;
        const uint64_t a_desc0 = smem_desc(smem_u32(a_tiles), A_LBO, 128);      // + (byte offset >> 4)
        const uint64_t b_desc0 = smem_desc(ring_u32, b_lbo, 128);
        const uint32_t stage16 = stage_bytes >> 4, step16 = step_bytes >> 4, half16 = half_bytes >> 4;
        uint32_t slot = 0, wpar = 0;     // slot / parity of the stage being consumed
        uint32_t accumulate = 0, sub = 0;
        int consumed = 0;                // stages whose MMAs have been issued
        const uint32_t afull0 = smem_u32(&s_afull[0]), afree0 = smem_u32(&s_afree[0]);
        while (loaded < total_stages && loaded < (int)nst) load_w();
        for (int gi = 0; gi < total_gran; ++gi) {
            const uint32_t ab = (uint32_t)gi & 1u;
            mbar_wait_q(afull0 + 8u * ab, ((uint32_t)gi >> 1) & 1u);
            tc_fence_after();
            const uint64_t a_g = a_desc0 + (uint64_t)(ab * (A_BUF >> 4));
#pragma unroll
            for (int s = 0; s < STEPS_G; ++s) {
                if (sub == 0) {
                    if (loaded == consumed) {   // ring ran dry: the stage to consume has not even been requested
                        mbar_wait_q(wempty0 + 8u * lslot, lpar);
                        load_w();
                    }
                    mbar_wait_q(wfull0 + 8u * slot, wpar);
                    tc_fence_after();
                }
                const uint64_t a_hi = a_g + (uint64_t)((uint32_t)(s >> 1) * (A_KB >> 4) + (uint32_t)(s & 1) * (2u * A_LBO >> 4));
                const uint64_t a_lo = a_hi + (uint64_t)(A_PART >> 4);
                const uint64_t b_hi = b_desc0 + (uint64_t)(slot * stage16 + sub * step16);
                const uint64_t b_lo = b_hi + (uint64_t)half16;
                mma_bf16_ss_elect(tmem_u, a_hi, b_hi, idesc, accumulate);
                mma_bf16_ss_elect(tmem_x, a_hi, b_lo, idesc, accumulate);
                accumulate = 1;
                mma_bf16_ss_elect(tmem_x, a_lo, b_hi, idesc, 1);
                if (++sub == sps) {
                    sub = 0;
                    mma_commit_elect(wempty0 + 8u * slot);   // the slot may be refilled once these MMAs have read it
                    if (++slot == nst) { slot = 0; wpar ^= 1u; }
                    ++consumed;
                    // refill every slot whose MMAs have completed (never more than nst requests ahead of consumption)
                    while (loaded < total_stages && loaded - consumed < (int)nst && mbar_test_wait(wempty0 + 8u * lslot, lpar)) load_w();
                }
            }
            mma_commit_elect(afree0 + 8u * ab);  // A tiles of this granule consumed
        }
        mma_commit_elect(smem_u32(&s_accum));
    } else {
        // ------------------------------------------------------------ producers
        const int aa = ptid % FU_LANES, grp = ptid / FU_LANES;
        const int k0 = grp * KG;
        constexpr bool a_ok = true;
        const float *F = P.feats + (size_t)z * P.c * P.p_in * P.na;
        const NeighbourList<C::CAP> &L = s_L[HALVES ? 0 : pt];
        float *Fp = Fs + (size_t)pt * 2 * CCH * NN * NA;
        const int total_nn = L.total;
        constexpr int bar_id = 1;   // named barrier of the producer threads (points advance in lockstep)

        const uint32_t fs_u32 = smem_u32(Fp);
        const uint32_t gbar0 = smem_u32(&s_gbar[pt][0]);
        const uint32_t afree0 = smem_u32(&s_afree[0]), afull0 = smem_u32(&s_afull[0]);

        // A-tile row of this thread: address of its 16-byte piece of K' chunk 0 of buffer 0
        const uint32_t a_row = smem_u32(a_tiles) + (uint32_t)(pt * 64 + aa) * 16u;
        uint8_t *keep_base = nullptr;
        if (P.keep != nullptr) {
            const int zs = z / P.keep_slab_clouds, zl = z - zs * P.keep_slab_clouds;
            const long long row = (long long)zl * P.keep_cols_per_z + (long long)pl * P.na + aa;
            keep_base = P.keep + (size_t)zs * P.keep_slab_bytes + ((size_t)(row >> 7) * P.keep_k_blocks) * tile_bytes(TR_A) +
                        (size_t)(row & 127) * 16;
        }
        // one 16-byte hi piece + one lo piece of K' chunk `kc` (global numbering) = granule-local chunk `kcl`
        auto put = [&](int ab, int kcl, int kc, uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3, uint32_t l0, uint32_t l1,
                       uint32_t l2, uint32_t l3) {
            if (!a_ok) return;
            const uint32_t dst = a_row + (uint32_t)ab * A_BUF + (uint32_t)(kcl >> 2) * A_KB + (uint32_t)(kcl & 3) * A_LBO;
            st_shared_v4(dst, h0, h1, h2, h3);
            st_shared_v4(dst + A_PART, l0, l1, l2, l3);
            if (keep_base != nullptr) {
                uint8_t *kd = keep_base + (size_t)(kc >> 2) * tile_bytes(TR_A) + (size_t)(kc & 3) * (TR_A * 16);
                *reinterpret_cast<uint4 *>(kd) = make_uint4(h0, h1, h2, h3);
                *reinterpret_cast<uint4 *>(kd + part_bytes(TR_A)) = make_uint4(l0, l1, l2, l3);
            }
        };

        int ci = 0;  // gather chunks issued so far by this point (buffer = ci & 1, parity = (ci >> 1) & 1)
        int gi = 0;  // granules produced so far by this CTA
        const int nchunks = P.c / CCH;
        for (int pass = 0; pass < npass; ++pass) {
            int n_first = pass * PASS_NN;
            int nn = total_nn - n_first;          // distinct neighbours of this pass
            nn = nn < 0 ? 0 : (nn > PASS_NN ? PASS_NN : nn);
            if (HALVES) {
                // half 0 takes the first ceil(groups / 2) groups of four neighbours, half 1 the rest: whole groups, so
                // that the FMA loop of half 0 (which skips groups past its count) never touches half 1's entries
                const int n0 = (((nn + 3) >> 2) + 1) >> 1 << 2;
                if (pt == 0) {
                    nn = nn < n0 ? nn : n0;
                } else {
                    n_first += n0;
                    nn = nn > n0 ? nn - n0 : 0;
                }
            }
            named_bar_sync(bar_id, NPROD);      // everybody is done with the previous pass's gather buffers
            for (int t = ptid; t < 2 * CCH * NN * NA; t += PT_THR)   // never-copied rows are zero
                if ((t / NA) % NN >= nn) Fp[t] = 0.f;
            // Kernel weights of this thread's (anchor, KG kernel points) against the NN neighbours of the pass, in the
            // packing the FMA loop below wants (96 registers either way):
            //   MODE 1 / 3: wk[j][n] = (w[k0+2j][n], w[k0+2j+1][n])  -- two kernel points per register pair, the feature
            //               enters the FFMA2 as a broadcast scalar, the accumulator pair is two finished G values;
            //   MODE 2:     ws[i][n] scalars, broadcast against the packed features of two channels.
            // (The first version packed two NEIGHBOURS per pair: every G value then was an (even, odd) pair that cost
            // an extra FADD before the hi/lo split -- 6 of ~105 instructions per thread and channel.)
            uint64_t wk[MODE != 2 ? KG / 2 : 1][MODE != 2 ? NN : 1];
            float ws[MODE == 2 ? KG : 1][MODE == 2 ? NN : 1];
            {
                float R[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) R[i] = __ldg(P.g.anchors + (aa < P.na ? aa : 0) * 9 + i);   // dead lanes: any valid anchor
                const float inv_sigma = 1.0f / P.g.sigma;
                const uint64_t nis2 = pack_f32x2(-inv_sigma, -inv_sigma);
                float rx[KG], ry[KG], rz[KG];
#pragma unroll
                for (int i = 0; i < KG; ++i) {
                    const float kx = __ldg(P.g.kernels + (k0 + i) * 3), ky = __ldg(P.g.kernels + (k0 + i) * 3 + 1),
                                kz = __ldg(P.g.kernels + (k0 + i) * 3 + 2);
                    rx[i] = R[0] * kx + R[1] * ky + R[2] * kz;
                    ry[i] = R[3] * kx + R[4] * ky + R[5] * kz;
                    rz[i] = R[6] * kx + R[7] * ky + R[8] * kz;
                }
                // absent neighbours (n >= nn) have multiplicity 0 in the list: their weights come out as 0
                const int m0 = MODE == 1 ? 0 : n_first;
                if (MODE != 2) {
#pragma unroll
                    for (int j = 0; j < KG / 2; ++j) {
                        const KPoint2 rk{pack_f32x2(rx[2 * j], rx[2 * j + 1]), pack_f32x2(ry[2 * j], ry[2 * j + 1]),
                                         pack_f32x2(rz[2 * j], rz[2 * j + 1])};
#pragma unroll
                        for (int n = 0; n < NN; ++n) {
                            const int m = m0 + n;
                            const float gx = L.g[m * 3], gy = L.g[m * 3 + 1], gz = L.g[m * 3 + 2], mu = L.mult[m];
                            wk[j][n] = kernel_weight_pair(pack_f32x2(gx, gx), pack_f32x2(gy, gy), pack_f32x2(gz, gz), rk, nis2,
                                                          pack_f32x2(mu, mu));
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < KG; ++i) {
                        const KPoint2 rk = kpoint2(rx[i], ry[i], rz[i]);
#pragma unroll
                        for (int n = 0; n < NN; n += 2) {
                            const int m = m0 + n;
                            const uint64_t w = kernel_weight_pair(pack_f32x2(L.g[m * 3], L.g[m * 3 + 3]), pack_f32x2(L.g[m * 3 + 1], L.g[m * 3 + 4]),
                                                                  pack_f32x2(L.g[m * 3 + 2], L.g[m * 3 + 5]), rk, nis2,
                                                                  pack_f32x2(L.mult[m], L.mult[m + 1]));
                            unpack_f32x2(w, ws[i][n], ws[i][n + 1]);
                        }
                    }
                }
            }
            // Gather chunk `chunk` of this pass into buffer ci & 1: every thread of the point copies its share of the
            // 16-byte pieces (15 per 240-byte feature row) with cp.async and then lets the buffer's mbarrier count
            // its copies (arrive.noinc: the arrival fires when this thread's copies have landed).  One UBLKCP per
            // row was issue-bound here: the copy instruction takes uniform registers, so a warp issues its lanes'
            // copies one at a time (~1000 cycles per warp and chunk, measured).
            auto issue = [&](int chunk) {
                const int buf = ci & 1;
                const uint32_t bar = gbar0 + 8u * (uint32_t)buf;
                if (GATHER == 2) {
                    constexpr int PIECES = CCH * NN * (NA / 4);
#pragma unroll
                    for (int it = 0; it < (PIECES + PT_THR - 1) / PT_THR; ++it) {
                        const int t = ptid + it * PT_THR;
                        const int piece = t % (NA / 4), slot = t / (NA / 4), cl = slot / NN, n = slot % NN;
                        if (t < PIECES && n < nn && piece * 4 < P.na) {
                            const float *src = F + ((size_t)(chunk * CCH + cl) * P.p_in + L.idx[n_first + n]) * P.na + piece * 4;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                         ::"r"(fs_u32 + (uint32_t)((((buf * CCH + cl) * NN + n) * NA + piece * 4) * 4)), "l"(src) : "memory");
                        }
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
                } else {
                    const uint32_t ROW_BYTES = (uint32_t)P.na * 4u;
                    constexpr int SLOTS = CCH * NN;                       // 64 / 128 rows
                    constexpr int SPREAD = GATHER == 1 ? PT_THR / SLOTS : 1;   // 3: every third thread owns a row
                    if (ptid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(CCH * nn) * ROW_BYTES);
                    const int t = ptid / SPREAD;
                    if (ptid % SPREAD == 0 && t < SLOTS) {
                        const int cl = t / NN, n = t % NN;
                        if (n < nn)
                            bulk_g2s(fs_u32 + (uint32_t)(((buf * CCH + cl) * NN + n) * NA) * 4u,
                                     F + ((size_t)(chunk * CCH + cl) * P.p_in + L.idx[n_first + n]) * P.na, ROW_BYTES, bar);
                    }
                }
                ++ci;
            };
            named_bar_sync(bar_id, NPROD);  // zero rows of both buffers are in place before the FMA loops read them
            issue(0);
            for (int g = 0; g < ngran; ++g, ++gi) {
                const int ab = gi & 1;
                uint32_t hi[12], lo[12];
#pragma unroll
                for (int h = 0; h < CH_G; ++h) {
                    const int chunk = g * CH_G + h;
                    const int cur = ci - 1;          // the chunk about to be consumed was issued last
                    const int buf = cur & 1;
                    named_bar_sync(bar_id, NPROD);  // every producer thread is done with the other gather buffer
                    if (chunk + 1 < nchunks) issue(chunk + 1);
                    mbar_wait_q(gbar0 + 8u * (uint32_t)buf, (uint32_t)(cur >> 1) & 1u);
                    if (h == 0 && gi >= 2)  // the MMAs of granule gi-2 have read A[ab]
                        mbar_wait_q(afree0 + 8u * (uint32_t)ab, (uint32_t)((gi >> 1) - 1) & 1u);
                    const float *fbase = Fp + (size_t)(buf * CCH * NN) * NA + aa;
                    if (MODE != 2) {
                        // 4 channels x 6 kernel points = 24 values = K' chunks 12 g + 3 grp + {0,1,2}
                        const int kcl0 = grp * 3, kc0 = g * 12 + grp * 3;
#pragma unroll
                        for (int cl4 = 0; cl4 < 4; ++cl4) {
                            uint64_t acc2[KG / 2];   // (G[k0+2j], G[k0+2j+1]) of this channel
                            const float *frow = fbase + cl4 * NN * NA;
                            {   // neighbour 0 initialises (absent neighbours have zero weights and zero feature rows)
                                const float f = frow[0];
                                const uint64_t f2 = pack_f32x2(f, f);
#pragma unroll
                                for (int j = 0; j < KG / 2; ++j) acc2[j] = mul_f32x2(wk[j][0], f2);
                            }
#pragma unroll
                            for (int n = 1; n < 4; ++n) {
                                const float f = frow[n * NA];
                                const uint64_t f2 = pack_f32x2(f, f);
#pragma unroll
                                for (int j = 0; j < KG / 2; ++j) acc2[j] = fma_f32x2(wk[j][n], f2, acc2[j]);
                            }
#pragma unroll
                            for (int n4 = 4; n4 < NN; n4 += 4) {
                                if (n4 < nn) {
#pragma unroll
                                    for (int n = n4; n < n4 + 4; ++n) {
                                        const float f = frow[n * NA];
                                        const uint64_t f2 = pack_f32x2(f, f);
#pragma unroll
                                        for (int j = 0; j < KG / 2; ++j) acc2[j] = fma_f32x2(wk[j][n], f2, acc2[j]);
                                    }
                                }
                            }
#pragma unroll
                            for (int ip = 0; ip < KG / 2; ++ip) {
                                float v0, v1;
                                unpack_f32x2(acc2[ip], v0, v1);
                                split2<FMT>(v0, v1, hi[cl4 * 3 + ip], lo[cl4 * 3 + ip]);
                            }
                            if (cl4 >= 1) {  // 6 (cl4 + 1) values so far: piece j = cl4 - 1 (values 8j .. 8j+7) is complete
                                const int jj = cl4 - 1;
                                put(ab, kcl0 + jj, kc0 + jj, hi[4 * jj], hi[4 * jj + 1], hi[4 * jj + 2], hi[4 * jj + 3], lo[4 * jj],
                                    lo[4 * jj + 1], lo[4 * jj + 2], lo[4 * jj + 3]);
                            }
                        }
                    } else {
                        // 8 channels (two chunks) x 3 kernel points = 24 values = K' chunks 24 g + 3 grp + {0,1,2}
#pragma unroll
                        for (int cp = 0; cp < CCH / 2; ++cp) {  // two channels -> 6 values -> 3 packed pairs
                            uint64_t acc2[KG];   // (G[c0][k0+i], G[c0+1][k0+i]): the two channels of the pair
                            const float *fr0 = fbase + (cp * 2) * NN * NA, *fr1 = fr0 + NN * NA;
                            {
                                const uint64_t f2 = pack_f32x2(fr0[0], fr1[0]);
#pragma unroll
                                for (int i = 0; i < KG; ++i) acc2[i] = mul_f32x2(pack_f32x2(ws[i][0], ws[i][0]), f2);
                            }
#pragma unroll
                            for (int n = 1; n < 4; ++n) {
                                const uint64_t f2 = pack_f32x2(fr0[n * NA], fr1[n * NA]);
#pragma unroll
                                for (int i = 0; i < KG; ++i) acc2[i] = fma_f32x2(pack_f32x2(ws[i][n], ws[i][n]), f2, acc2[i]);
                            }
#pragma unroll
                            for (int n4 = 4; n4 < NN; n4 += 4) {
                                if (n4 < nn) {
#pragma unroll
                                    for (int n = n4; n < n4 + 4; ++n) {
                                        const uint64_t f2 = pack_f32x2(fr0[n * NA], fr1[n * NA]);
#pragma unroll
                                        for (int i = 0; i < KG; ++i) acc2[i] = fma_f32x2(pack_f32x2(ws[i][n], ws[i][n]), f2, acc2[i]);
                                    }
                                }
                            }
                            float v[6];   // K' order: channel-major, three kernel points each
#pragma unroll
                            for (int i = 0; i < KG; ++i) unpack_f32x2(acc2[i], v[i], v[KG + i]);
#pragma unroll
                            for (int ip = 0; ip < 3; ++ip) {
                                split2<FMT>(v[2 * ip], v[2 * ip + 1], hi[h * 6 + cp * 3 + ip], lo[h * 6 + cp * 3 + ip]);
                            }
                        }
                        const int kcl0 = grp * 3, kc0 = g * 24 + grp * 3;
                        if (h == 0) {
                            put(ab, kcl0, kc0, hi[0], hi[1], hi[2], hi[3], lo[0], lo[1], lo[2], lo[3]);
                        } else {
                            put(ab, kcl0 + 1, kc0 + 1, hi[4], hi[5], hi[6], hi[7], lo[4], lo[5], lo[6], lo[7]);
                            put(ab, kcl0 + 2, kc0 + 2, hi[8], hi[9], hi[10], hi[11], lo[8], lo[9], lo[10], lo[11]);
                        }
                    }
                }
                // granule complete: publish this warp's pieces to the tensor core
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(afull0 + 8u * (uint32_t)ab);
            }
        }

    }
    {
        // ------------------------------------------------------------ epilogue: TMEM -> out[z, o, p, a], all 16 warps
        mbar_wait(smem_u32(&s_accum), 0);
        tc_fence_after();
        const int q = warp & 3;                          // TMEM lane quarter of this warp
        const int row = q * 32 + lane;                  // = pt*64 + anchor
        const int rpt = row >> 6, ra = row & 63;
        if (HALVES) {
            // rows r and r + 64 are the two partial results of (point, anchor r): quarters 2,3 hand theirs over through
            // shared memory (the A tiles are free: every MMA has completed), quarters 0,1 add and store
            float *stage = reinterpret_cast<float *>(a_tiles) + (size_t)(warp >> 2) * (32 * 64);   // [32 channels][64 rows]
            float *orow = P.out + (size_t)z * P.out_sz + (size_t)blockIdx.x * P.na + ra;
            for (int cg0 = 0; cg0 * 32 < P.c_out; cg0 += NWARPS / 4) {
                const int cg = cg0 + (warp >> 2);
                const bool active = cg * 32 < P.c_out;   // uniform over the four warps that share a staging area
                float v[32];
                if (active) tmem_ld_sum2(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32), P.tmem_cols, v);
                if (active && rpt == 1) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) stage[jj * 64 + ra] = v[jj];
                }
                __syncthreads();
                if (active && rpt == 0 && ra < P.na) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int o = cg * 32 + jj;
                        if (o < P.c_out) orow[(size_t)o * P.out_so] = (v[jj] + stage[jj * 64 + ra]) * P.out_scale;
                    }
                }
                __syncthreads();
            }
        } else if (rpt < PTS) {                         // warp-uniform (PTS == 1: quarters 2,3 hold dead rows)
            float *orow = P.out + (size_t)z * P.out_sz + (size_t)(blockIdx.x * PTS + rpt) * P.na + ra;
            for (int cg = warp >> 2; cg * 32 < P.c_out; cg += NWARPS / 4) {
                float v[32];
                tmem_ld_sum2(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32), P.tmem_cols, v);
                if (ra < P.na) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int o = cg * 32 + jj;
                        if (o < P.c_out) orow[(size_t)o * P.out_so] = v[jj] * P.out_scale;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPWARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * P.tmem_cols);
    }
}

template <int MODE, int GATHER, int FMT = 0>
int launch_fused_variant(FusedParams &P, int p_cnt, int bc, cudaStream_t s) {
    using C = FuCfg<MODE>;
    constexpr int NPROD = C::PTS * FU_LANES * (FU_KS / C::KG);
    constexpr int ROWS = C::PTS * 64;
    constexpr size_t A_BYTES = (size_t)2 * (C::GCH * FU_KS / 32) * 2 * 4 * ROWS * 16 + (C::PTS == 1 ? 1024 : 0);
    const size_t fixed = A_BYTES + sizeof(float) * (size_t)C::PTS * 2 * FU_CCH * C::NN * FU_NA;
    // ring stage = sps 16-k steps (one bulk copy, one full/empty barrier round trip): the narrower the MMA (small
    // c_out), the more steps per stage, so that the single control thread is not the bottleneck
    P.sps = P.trb <= 64 ? 3 : (P.trb <= 128 ? 2 : 1);
    if (const char *e = getenv("EPN_FU_SPS")) {   // tuning knob (tools/fused_sweep.py); must divide 6
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 3) P.sps = v;
    }
    const size_t stage = (size_t)P.trb * 64 * P.sps;
    const size_t budget = 227 * 1024 - 6 * 1024;  // static shared memory (neighbour lists, barriers) comes on top
    if (fixed + 2 * stage > budget) return 1;
    int nst = (int)((budget - fixed) / stage);
    if (nst > 8) nst = 8;
    P.nst = nst;
    const size_t smem_bytes = fixed + (size_t)nst * stage;
    static DynSmemOnce once;  // one per template instantiation
    if (int rc = ensure_dyn_smem(once, inter_fused_kernel<MODE, GATHER, FMT>, (int)budget, "inter_fused_kernel")) return rc;
    dim3 grid(MODE == 3 ? p_cnt : p_cnt / C::PTS, bc);
    inter_fused_kernel<MODE, GATHER, FMT><<<grid, NPROD + FU_CTRL, smem_bytes, s>>>(P);
    return check_launch("inter_fused_kernel");
}

}  // namespace

// EPN_FU_HALVES=0 sends rows of more than 16 slots to MODE 2 also in inference (tuning / A-B knob)
static bool fused_halves_enabled() {
    static const bool on = [] { const char *e = getenv("EPN_FU_HALVES"); return e == nullptr || atoi(e) != 0; }();
    return on;
}

// K' mode the fused kernel uses for this shape (0 = shape not covered).  keep = the caller wants the operand tiles.
int inter_fused_mode(int c, int c_out, int p_cnt, int nn, int na, int ks, bool keep) {
    if (ks != FU_KS || na > FU_NA || na < 4 || na % 4 != 0 || c_out > 256 || c < 4) return 0;
    if (keep && na != FU_NA) return 0;   // kept tiles (128-row tiles of 60-anchor rows): the 60-anchor group only
    if (nn <= 16 && c % 4 == 0 && p_cnt % 2 == 0) return 1;
    // inference (no kept tiles), longer rows: the "halves" variant of the kernel, which uses MODE 1's K' order
    // (measured: 3.10 vs 3.44 ms on the 64-channel K = 32 layer of the cls network, no gain from 128 channels on --
    // every G value is converted twice there, which costs what the better load : FMA ratio saves)
    if (!keep && nn <= DEDUP_MAX_RAW && c % 4 == 0 && c <= 64 && fused_halves_enabled()) return 1;
    if (nn <= (keep ? 32 : DEDUP_MAX_RAW) && c % 8 == 0) return 2;
    return 0;
}

// Returns 1 when the shape is not covered (the caller takes the grouping-kernel + GEMM route).
int launch_inter_fused(const float *feats, const int32_t *idx, const InterGeom &g, const void *w_tiles, float *out,
                       long long out_stride_z, long long out_stride_o, void *keep_tiles, int keep_k_blocks,
                       long long keep_cols_per_z, int keep_slab_clouds, size_t keep_slab_bytes, int p_off, int p_cnt, int bc, int c, int c_out, int p_in, int p, int nn,
                       int na, int ks, cudaStream_t s, int fmt) {
    const int mode = inter_fused_mode(c, c_out, p_cnt, nn, na, ks, keep_tiles != nullptr);
    if (fmt == FMT_F16 && keep_tiles != nullptr) return 1;   // kept tiles are bf16 (the weight-gradient GEMM's format)
    if (mode == 0 || bc > 65535 || feats == nullptr) return 1;
    FusedParams P;
    P.feats = feats;
    P.idx = idx;
    P.g = g;
    P.Wt = static_cast<const uint8_t *>(w_tiles);
    P.out = out;
    P.out_sz = out_stride_z;
    P.out_so = out_stride_o;
    P.keep = static_cast<uint8_t *>(keep_tiles);
    P.keep_k_blocks = keep_k_blocks;
    P.keep_cols_per_z = keep_cols_per_z;
    P.keep_slab_clouds = keep_slab_clouds > 0 ? keep_slab_clouds : 1;
    P.keep_slab_bytes = keep_slab_bytes;
    P.c = c; P.c_out = c_out; P.p_in = p_in; P.p = p; P.nn = nn; P.p_off = p_off; P.na = na;
    P.trb = umma_trb_for(c_out);
    uint32_t cols = 32;
    while ((int)cols < P.trb) cols *= 2;
    P.tmem_cols = cols;
    P.out_scale = fmt == FMT_F16 ? 1.0f / F16_W_SCALE : 1.0f;
    ProfScope prof(s, KC_INTER_FUSED);
    if (fmt == FMT_F16) {   // inference forward with fp16 operands: default gather variant only
        if (mode == 1 && (nn > 16 || p_cnt % 2 != 0)) return launch_fused_variant<3, 1, FMT_F16>(P, p_cnt, bc, s);
        if (mode == 1) return launch_fused_variant<1, 1, FMT_F16>(P, p_cnt, bc, s);
        return launch_fused_variant<2, 1, FMT_F16>(P, p_cnt, bc, s);
    }
    int gather = 1;
    if (const char *e = getenv("EPN_FU_GATHER")) gather = atoi(e);   // tuning knob (tools/fused_sweep.py)
    if (mode == 1 && (nn > 16 || p_cnt % 2 != 0)) {   // halves variant (inference only, see inter_fused_mode)
        if (gather == 0) return launch_fused_variant<3, 0>(P, p_cnt, bc, s);
        if (gather == 2) return launch_fused_variant<3, 2>(P, p_cnt, bc, s);
        return launch_fused_variant<3, 1>(P, p_cnt, bc, s);
    }
    if (mode == 1) {
        if (gather == 0) return launch_fused_variant<1, 0>(P, p_cnt, bc, s);
        if (gather == 2) return launch_fused_variant<1, 2>(P, p_cnt, bc, s);
        return launch_fused_variant<1, 1>(P, p_cnt, bc, s);
    }
    if (gather == 0) return launch_fused_variant<2, 0>(P, p_cnt, bc, s);
    if (gather == 2) return launch_fused_variant<2, 2>(P, p_cnt, bc, s);
    return launch_fused_variant<2, 1>(P, p_cnt, bc, s);
}

}  // namespace epn
