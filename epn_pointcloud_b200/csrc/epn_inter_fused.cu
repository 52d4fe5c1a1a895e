// Fused InterSO3Conv forward: neighbour gather + kernel weights + spatial contraction + channel GEMM in ONE
// kernel; the grouped tensor G never leaves the SM.
//
//   out[z,o,p,a] = sum_{c,k} W[o, c*24+k] * G[(c,k),(z,p,a)],   G = sum_n w(p,a,k,n) * feats[z,c,idx[z,p,n],a]
//   (vgtk/vgtk/so3conv/modules.py:157-174 -> so3conv/functional.py:118-218 -> spconv/functional.py:361-390
//    -> so3conv/modules.py:48-55)
//
// A CTA owns PTS consecutive output points of one cloud (PTS*64 rows of the MMA M dimension, row = pt*64 + anchor,
// anchors 60..63 of a point are dead rows).  Producer warps (PTS*64*(24/KG) threads, same mapping as
// inter_group_tiles_kernel: lane <-> anchor, warp pair <-> group of KG kernel points, kernel weights in
// registers as fp32x2 pairs) stream the input channels in chunks of CCH: bulk-copy gather of the neighbours'
// feature rows -> per-anchor spatial contraction (fp32 SIMT) -> fp32 staging -> bf16 hi/lo split written as
// canonical K-major UMMA operand tiles into shared memory.  One control thread streams the weight tiles
// (bf16 hi/lo, rows = c_out) from L2 through a ring of 16-k stages with bulk copies and issues three
// tcgen05.mma (hi*hi, hi*lo, lo*hi) per 16-wide k step into a [128 x c_out] fp32 accumulator in TMEM while the
// producers work on the next chunk.  The epilogue (all producer warps) reads TMEM and stores out[z,o,p,.] rows
// (128 contiguous bytes per warp and output channel).  In training the conversion pass also writes the operand
// tiles to global memory (what the weight-gradient GEMM consumes, see epn_gemm_dw.cu).
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

namespace {

constexpr int FU_LANES = 64;  // anchor lanes per kernel-point group
constexpr int FU_KS = 24;     // kernel points (kpsphere24)
constexpr int FU_NA = 60;

struct FusedParams {
    const float *feats;      // [b, c, p_in, 60] or null (occupancy features == 1, c == 1)
    const int32_t *idx;      // [b, p, nn]
    InterGeom g;
    const uint8_t *Wt;       // weight tiles [k_blocks] of trb rows (split tiles, rows = c_out, K = c*24)
    float *out;              // out + z*out_sz + o*out_so + pl*60 + a
    long long out_sz, out_so;
    uint8_t *keep;           // optional: forward operand tiles in global memory (rows = (z,pl,a), K = (c,k))
    int keep_k_blocks, keep_slab_clouds;   // clouds z are stored in slabs of keep_slab_clouds, each slab a tile matrix
    long long keep_cols_per_z;
    size_t keep_slab_bytes;
    int c, c_out, p_in, p, nn, p_off, trb, nst;
    uint32_t tmem_cols;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NN, int KG, int CCH, int PTS, bool HAS_FEATS>
__global__ void __launch_bounds__(PTS *FU_LANES *(FU_KS / KG) + 128, 1)
inter_fused_kernel(FusedParams P) {
    constexpr int NA = FU_NA;
    constexpr int GROUPS = FU_KS / KG;
    constexpr int PT_THR = FU_LANES * GROUPS;     // producer threads per point
    constexpr int NPROD = PTS * PT_THR;           // producer threads
    constexpr int GSTR = NA + 1;
    constexpr int CKK = CCH * FU_KS;              // (c,k) rows per chunk
    constexpr int KBC = CKK / 32;                 // k-blocks per chunk
    static_assert(CKK % 32 == 0, "chunk must be whole k-blocks");
    constexpr int ROWS = PTS * 64;                // valid rows of the A tiles
    constexpr uint32_t A_LBO = ROWS * 16;         // bytes between 8-wide k chunks
    constexpr uint32_t A_PART = 4 * A_LBO;        // hi (or lo) part of one k-block
    constexpr uint32_t A_KB = 2 * A_PART;         // one k-block
    constexpr uint32_t A_BYTES = KBC * A_KB + (PTS == 1 ? 1024 : 0);  // PTS == 1: the M=128 MMA over-reads 64 dead rows

    extern __shared__ __align__(128) uint8_t smem[];
    // carve-up (all offsets multiples of 128 bytes)
    uint8_t *a_tiles = smem;
    float *Fs = reinterpret_cast<float *>(smem + A_BYTES);                 // [PTS][2][CCH][NN][NA]
    float *Gs = Fs + PTS * 2 * CCH * NN * NA;                               // [PTS][CKK][GSTR]
    float *hdr = Gs + PTS * CKK * GSTR;                                     // [PTS][NN*6]
    uint8_t *ring = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(hdr + PTS * NN * 6) + 127) & ~(uintptr_t)127);
    __shared__ __align__(8) uint64_t s_gbar[PTS][2];   // gather buffers
    __shared__ __align__(8) uint64_t s_wfull[8], s_wempty[8];
    __shared__ __align__(8) uint64_t s_afull, s_afree, s_accum;
    __shared__ uint32_t s_tmem;
    __shared__ int s_nu[PTS];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const bool is_ctrl = tid >= NPROD;
    const int z = blockIdx.y;
    const uint32_t stage_bytes = (uint32_t)P.trb * 64u;  // 16 k of hi + 16 k of lo
    const int nchunks = P.c / CCH;
    const int total_steps = nchunks * KBC * 2;

    if (tid == 0) {
        for (int i = 0; i < PTS; ++i) {
            mbar_init(smem_u32(&s_gbar[i][0]), 1);
            mbar_init(smem_u32(&s_gbar[i][1]), 1);
        }
        for (int i = 0; i < P.nst; ++i) {
            mbar_init(smem_u32(&s_wfull[i]), 1);
            mbar_init(smem_u32(&s_wempty[i]), 1);
        }
        mbar_init(smem_u32(&s_afull), 1);
        mbar_init(smem_u32(&s_afree), 1);
        mbar_init(smem_u32(&s_accum), 1);
        fence_barrier_init();
    }
    if (warp == NPROD / 32) tmem_alloc(smem_u32(&s_tmem), P.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (is_ctrl) {
        // ------------------------------------------------------------ control warp: W ring + MMA issue
        // a whole warpgroup (setmaxnreg is a warpgroup-wide instruction) of which one thread works: it hands its
        // registers to the producers
        // (setmaxnreg rebalancing faulted on the B200 in this configuration; the kernel runs at 96 registers)
        if (tid == NPROD) {
            const uint32_t ring_u32 = smem_u32(ring);
            const uint32_t half_bytes = (uint32_t)P.trb * 32u;
            const size_t w_tile = tile_bytes(P.trb), w_part = part_bytes(P.trb);
            auto load_w = [&](int j) {  // 16-k step j -> ring slot j % nst
                const int slot = j % P.nst;
                const uint32_t bar = smem_u32(&s_wfull[slot]);
                const uint8_t *src = P.Wt + (size_t)(j >> 1) * w_tile + (size_t)(j & 1) * half_bytes;
                mbar_arrive_expect_tx(bar, stage_bytes);
                bulk_g2s(ring_u32 + slot * stage_bytes, src, half_bytes, bar);
                bulk_g2s(ring_u32 + slot * stage_bytes + half_bytes, src + w_part, half_bytes, bar);
            };
            for (int j = 0; j < P.nst && j < total_steps; ++j) load_w(j);
            const uint32_t idesc = instr_desc_bf16_m128(P.trb);
            const uint32_t b_lbo = (uint32_t)P.trb * 16u;
            const uint32_t a_u32 = smem_u32(a_tiles);
            int j = 0;
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                mbar_wait(smem_u32(&s_afull), (uint32_t)chunk & 1u);
                tc_fence_after();
                for (int s = 0; s < KBC * 2; ++s, ++j) {
                    const int slot = j % P.nst;
                    mbar_wait(smem_u32(&s_wfull[slot]), (uint32_t)(j / P.nst) & 1u);
                    tc_fence_after();
                    const uint32_t a0 = a_u32 + (uint32_t)(s >> 1) * A_KB + (uint32_t)(s & 1) * 2u * A_LBO;
                    const uint32_t b0 = ring_u32 + slot * stage_bytes;
                    const uint64_t a_hi = smem_desc(a0, A_LBO, 128), a_lo = smem_desc(a0 + A_PART, A_LBO, 128);
                    const uint64_t b_hi = smem_desc(b0, b_lbo, 128), b_lo = smem_desc(b0 + half_bytes, b_lbo, 128);
                    mma_bf16_ss(tmem_base, a_hi, b_hi, idesc, j != 0);
                    mma_bf16_ss(tmem_base, a_hi, b_lo, idesc, 1);
                    mma_bf16_ss(tmem_base, a_lo, b_hi, idesc, 1);
                    mma_commit(smem_u32(&s_wempty[slot]));
                    // refill the slot of the PREVIOUS step (its MMAs finish before this step's do, so this wait
                    // does not drain the tensor pipe)
                    const int r = j - 1;
                    if (r >= 0 && r + P.nst < total_steps) {
                        mbar_wait(smem_u32(&s_wempty[r % P.nst]), (uint32_t)(r / P.nst) & 1u);
                        load_w(r + P.nst);
                    }
                }
                mma_commit(smem_u32(&s_afree));  // A tiles of this chunk consumed
            }
            mma_commit(smem_u32(&s_accum));
        }
    } else {
        // ------------------------------------------------------------ producers

        const int pt = tid / PT_THR, ptid = tid - pt * PT_THR;
        const int a = ptid % FU_LANES, grp = ptid / FU_LANES;
        const int k0 = grp * KG;
        const bool a_ok = a < NA;
        const int aa = a_ok ? a : a - 4;  // dead lanes shadow a live lane of their own warp (broadcast, no bank conflict)
        const int pl = blockIdx.x * PTS + pt, pi = P.p_off + pl;
        const float *F = HAS_FEATS ? P.feats + (size_t)z * P.c * P.p_in * NA : nullptr;
        float *s_g = hdr + pt * NN * 6;
        int32_t *s_idx = reinterpret_cast<int32_t *>(s_g + NN * 3);
        float *s_mult = s_g + NN * 4;
        int32_t *s_raw = reinterpret_cast<int32_t *>(s_g + NN * 5);
        float *Fp = Fs + (size_t)pt * 2 * CCH * NN * NA;
        float *Gp = Gs + (size_t)pt * CKK * GSTR;

        // distinct neighbours + multiplicities (see inter_group_tiles_kernel)
        int nn = P.nn;
        for (int n = ptid; n < NN; n += PT_THR) s_raw[n] = n < nn ? P.idx[((size_t)z * P.p + pi) * nn + n] : -1;
        named_bar_sync(1, NPROD);
        if (ptid < 32) {
            const int n = ptid;
            const int q = n < nn ? s_raw[n] : -1;
            bool uniq = n < nn;
            for (int m = 0; m < n && uniq; ++m) uniq = s_raw[m] != q;
            int mult = 0;
            for (int m = n; m < nn; ++m) mult += (s_raw[m] == q) ? 1 : 0;
            const unsigned mask = __ballot_sync(0xffffffffu, uniq);
            const int pos = __popc(mask & ((1u << n) - 1u));
            if (uniq) {
                const float *X = P.g.xyz + (size_t)z * 3 * P.p_in;
                const float *Cn = P.g.centers + (size_t)z * 3 * P.p;
                s_idx[pos] = q;
                s_mult[pos] = (float)mult;
                s_g[pos * 3] = X[q] - Cn[pi];
                s_g[pos * 3 + 1] = X[P.p_in + q] - Cn[P.p + pi];
                s_g[pos * 3 + 2] = X[2 * P.p_in + q] - Cn[2 * P.p + pi];
            }
            const int cnt = __popc(mask);
            if (n >= cnt && n < NN) {
                s_idx[n] = 0; s_mult[n] = 0.f;
                s_g[n * 3] = 0.f; s_g[n * 3 + 1] = 0.f; s_g[n * 3 + 2] = 0.f;
            }
            if (n == 0) s_nu[pt] = cnt;
        }
        named_bar_sync(1, NPROD);
        nn = s_nu[pt];  // number of DISTINCT neighbours of this point
        for (int t = ptid; t < 2 * CCH * (NN - nn) * NA; t += PT_THR) {  // never-copied rows stay zero
            const int e = t % NA, r = t / NA, n = nn + r % (NN - nn), bc = r / (NN - nn);
            Fp[(bc * NN + n) * NA + e] = 0.f;
        }

        uint64_t w2[KG][NN / 2];
        {
            float R[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = __ldg(P.g.anchors + aa * 9 + i);
#pragma unroll
            for (int i = 0; i < KG; ++i) {
                const float kx = __ldg(P.g.kernels + (k0 + i) * 3), ky = __ldg(P.g.kernels + (k0 + i) * 3 + 1),
                            kz = __ldg(P.g.kernels + (k0 + i) * 3 + 2);
                const float rx = R[0] * kx + R[1] * ky + R[2] * kz, ry = R[3] * kx + R[4] * ky + R[5] * kz,
                            rz = R[6] * kx + R[7] * ky + R[8] * kz;
#pragma unroll
                for (int n = 0; n < NN; n += 2) {
                    float v[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float t = kernel_weight_fast(s_g[(n + e) * 3], s_g[(n + e) * 3 + 1], s_g[(n + e) * 3 + 2], rx, ry,
                                                           rz, 1.0f / P.g.sigma);
                        v[e] = (a_ok && n + e < nn) ? t * s_mult[n + e] : 0.f;
                    }
                    w2[i][n / 2] = pack_f32x2(v[0], v[1]);
                }
            }
        }

        const uint32_t fs_u32 = smem_u32(Fp);
        const uint32_t gbar0 = smem_u32(&s_gbar[pt][0]);
        constexpr uint32_t ROW_BYTES = NA * 4;
        auto issue = [&](int chunk, int buf) {
            if (!HAS_FEATS) return;
            const uint32_t bar = gbar0 + 8u * (uint32_t)buf;
            if (ptid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(CCH * nn) * ROW_BYTES);
            for (int t = ptid; t < CCH * NN; t += PT_THR) {
                const int cl = t / NN, n = t % NN;
                if (n < nn)
                    bulk_g2s(fs_u32 + (uint32_t)(((buf * CCH + cl) * NN + n) * NA) * 4u,
                             F + ((size_t)(chunk * CCH + cl) * P.p_in + s_idx[n]) * NA, ROW_BYTES, bar);
            }
        };

        // addressing of the conversion pass that does not depend on the chunk
        const uint32_t a_row = smem_u32(a_tiles) + (uint32_t)(pt * 64 + aa) * 16u;
        uint8_t *keep_base = nullptr;
        if (P.keep != nullptr) {
            const int zs = z / P.keep_slab_clouds, zl = z - zs * P.keep_slab_clouds;
            const long long row = (long long)zl * P.keep_cols_per_z + (long long)pl * NA + aa;
            keep_base = P.keep + (size_t)zs * P.keep_slab_bytes + ((size_t)(row >> 7) * P.keep_k_blocks) * tile_bytes(TR_A) +
                        (size_t)(row & 127) * 16;
        }

        uint32_t phase_bits = 0u;
        // the expect_tx of a buffer must be posted before any of its copies can complete: ptid 0 arms the barrier,
        // then everybody copies (named barrier keeps the order)
        issue(0, 0);
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int buf = chunk & 1;
            if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
            if (HAS_FEATS) {
                mbar_wait(gbar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
                phase_bits ^= 1u << buf;
            }
            // ---- spatial contraction of CCH channels
            const float *fbase = Fp + (size_t)(buf * CCH * NN) * NA + aa;
            float *gbase = Gp + (size_t)k0 * GSTR + aa;
#pragma unroll 2
            for (int cl = 0; cl < CCH; ++cl) {
                uint64_t acc2[KG];
#pragma unroll
                for (int i = 0; i < KG; ++i) acc2[i] = 0ull;
                const float *frow = fbase + cl * NN * NA;
#pragma unroll
                for (int n4 = 0; n4 < NN; n4 += 4) {
                    if (n4 < nn) {
#pragma unroll
                        for (int n = n4; n < n4 + 4; n += 2) {
                            const uint64_t f2 = HAS_FEATS ? pack_f32x2(frow[n * NA], frow[(n + 1) * NA]) : pack_f32x2(1.0f, 1.0f);
#pragma unroll
                            for (int i = 0; i < KG; ++i) acc2[i] = fma_f32x2(w2[i][n / 2], f2, acc2[i]);
                        }
                    }
                }
                if (a_ok) {
#pragma unroll
                    for (int i = 0; i < KG; ++i) {
                        float e, o;
                        unpack_f32x2(acc2[i], e, o);
                        gbase[(cl * FU_KS + i) * GSTR] = e + o;
                    }
                }
            }
            named_bar_sync(1, NPROD);  // staging complete; every thread is done with Fs[buf]
            if (chunk > 0) mbar_wait(smem_u32(&s_afree), (uint32_t)(chunk - 1) & 1u);  // MMAs of the previous chunk have read A
            // ---- fp32 staging -> bf16 hi/lo operand tiles in shared memory (+ global copy for the weight gradient)
            if (a_ok) {
                for (int kc = grp; kc < CKK / 8; kc += GROUPS) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = Gp[(kc * 8 + i) * GSTR + aa];
                    uint4 hi, lo;
                    split8(x, hi, lo);
                    const uint32_t dst = a_row + (uint32_t)(kc >> 2) * A_KB + (uint32_t)(kc & 3) * A_LBO;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + A_PART), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
                    if (keep_base != nullptr) {
                        const int kkg = chunk * CKK + kc * 8;
                        uint8_t *kd = keep_base + (size_t)(kkg >> 5) * tile_bytes(TR_A) + (size_t)((kkg & 31) >> 3) * (TR_A * 16);
                        *reinterpret_cast<uint4 *>(kd) = hi;
                        *reinterpret_cast<uint4 *>(kd + part_bytes(TR_A)) = lo;
                    }
                }
            }
            fence_proxy_async_smem();  // generic-proxy tile writes -> visible to the tensor core
            named_bar_sync(1, NPROD);
            if (tid == 0) mbar_arrive(smem_u32(&s_afull));
        }

        // ------------------------------------------------------------ epilogue: TMEM -> out[z, o, p, a]
        mbar_wait(smem_u32(&s_accum), 0);
        tc_fence_after();
        constexpr int NWARPS = NPROD / 32;
        const int lane = tid & 31, q = warp & 3;       // TMEM lane quarter of this warp
        const int row = q * 32 + lane;                  // = pt*64 + anchor
        const int rpt = row >> 6, ra = row & 63;
        if (rpt < PTS) {                                // warp-uniform (PTS == 1: quarters 2,3 hold dead rows)
            float *orow = P.out + (size_t)z * P.out_sz + (size_t)(blockIdx.x * PTS + rpt) * NA + ra;
            for (int cg = warp >> 2; cg * 32 < P.c_out; cg += NWARPS / 4) {
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32), v);
                if (ra < NA) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int o = cg * 32 + jj;
                        if (o < P.c_out) orow[(size_t)o * P.out_so] = v[jj];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPROD / 32) {
        tc_fence_after();
        tmem_dealloc(tmem_base, P.tmem_cols);
    }
}

template <int NN, int KG, int CCH, int PTS, bool HAS_FEATS>
int launch_fused_variant(FusedParams &P, int p_cnt, int bc, cudaStream_t s) {
    constexpr int NPROD = PTS * FU_LANES * (FU_KS / KG);
    constexpr int CKK = CCH * FU_KS;
    constexpr size_t A_BYTES = (size_t)(CKK / 32) * 2 * 4 * (PTS * 64) * 16 + (PTS == 1 ? 1024 : 0);
    const size_t fixed = A_BYTES + sizeof(float) * (size_t)PTS * (2 * CCH * NN * FU_NA + CKK * (FU_NA + 1) + NN * 6) + 128;
    const size_t stage = (size_t)P.trb * 64;
    const size_t budget = 227 * 1024 - 1024;  // static shared memory (barriers) comes on top
    if (fixed + 2 * stage > budget) return 1;
    int nst = (int)((budget - fixed) / stage);
    if (nst > 8) nst = 8;
    P.nst = nst;
    const size_t smem_bytes = fixed + (size_t)nst * stage;
    static DynSmemOnce once;  // one per template instantiation
    if (int rc = ensure_dyn_smem(once, inter_fused_kernel<NN, KG, CCH, PTS, HAS_FEATS>, 227 * 1024 - 1024, "inter_fused_kernel")) return rc;
    dim3 grid(p_cnt / PTS, bc);
    inter_fused_kernel<NN, KG, CCH, PTS, HAS_FEATS><<<grid, NPROD + 128, smem_bytes, s>>>(P);
    return check_launch("inter_fused_kernel");
}

}  // namespace

bool inter_fused_ok(int c, int c_out, int p_cnt, int nn, int na, int ks) {
    return ks == FU_KS && na == FU_NA && nn <= 32 && c >= 4 && c % 4 == 0 && c_out <= 256 && (nn > 16 || p_cnt % 2 == 0);
}

// Returns 1 when the shape is not covered (the caller takes the grouping-kernel + GEMM route).
int launch_inter_fused(const float *feats, const int32_t *idx, const InterGeom &g, const void *w_tiles, float *out,
                       long long out_stride_z, long long out_stride_o, void *keep_tiles, int keep_k_blocks,
                       long long keep_cols_per_z, int keep_slab_clouds, size_t keep_slab_bytes, int p_off, int p_cnt, int bc, int c, int c_out, int p_in, int p, int nn,
                       int na, int ks, cudaStream_t s) {
    if (!inter_fused_ok(c, c_out, p_cnt, nn, na, ks) || bc > 65535 || feats == nullptr) return 1;
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
    P.c = c; P.c_out = c_out; P.p_in = p_in; P.p = p; P.nn = nn; P.p_off = p_off;
    P.trb = umma_trb_for(c_out);
    uint32_t cols = 32;
    while ((int)cols < P.trb) cols *= 2;
    P.tmem_cols = cols;
    ProfScope prof(s, KC_INTER_FUSED);
    if (nn <= 16) return launch_fused_variant<16, 6, 4, 2, true>(P, p_cnt, bc, s);
    return launch_fused_variant<32, 3, 4, 1, true>(P, p_cnt, bc, s);
}

}  // namespace epn
