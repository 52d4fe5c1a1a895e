// Weight gradient straight from the operand tiles the forward pass kept.
//
//   dW[o, kk] += sum_n dout[o, n] * G[n, kk]          (so3conv/modules.py:48-55 under autograd)
//
// The forward GEMM consumed G as K-major tiles with rows = grouped columns n and K = kk = (c,k):
//     tile (rt, kb) = [part hi|lo][kc: 4][row: 128][8 x bf16 of kk]
// For dW the contraction runs over n, so G enters as the M operand (M = kk) and n is the MMA K dimension.
// Read that way the very same bytes are the canonical MN-MAJOR no-swizzle layout of tcgen05: 8 kk contiguous
// (16 bytes), consecutive n 16 bytes apart, i.e. one 8(k) x 8(mn) core matrix = 128 contiguous bytes.  So the
// backward pass neither recomputes the spatial contraction nor transposes anything: the producer moves the
// 1-KB pieces (64 n of one 8-kk group) of four adjacent forward tiles into a stage
//     A stage = [part][kk group: 16][n: 64][8 kk]     SBO (8-kk group stride) = 1 KB, LBO (8-n group stride) = 128 B
// and the MMA is issued with a_major = MN.  B = dout tiles (rows = o, K = n), K-major as everywhere else.
//
// The tile array is a dense 4-D tensor of 32-bit words  [k-block (over all row tiles)][part 2][k-chunk 4][512]
// (strides 16 KB / 8 KB / 2 KB / 4 B; the 128 rows x 16 B of one k-chunk are 2 KB contiguous), so ONE tensor-map
// TMA load (box 4 x 1 x 4 x 256 words = 16 KB, inner extent 1 KB) brings the 16 pieces of one part of a stage; two
// loads + the two dout tiles make a stage.  (A 5-D map with the 16-byte rows as the inner dimension was measured
// SLOWER than the bulk copies: the TMA unit works in units of the box's inner extent.)  (The first version issued the
// 32 pieces as 32 bulk copies: a warp issues those one lane at a time, ~26 GB/s per issuing warp, and the kernel
// ran at 3.3 TB/s; epn_set/EPN_DW_TMA=0 selects that path for comparison.)
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

namespace {

// UNIT = n per pipeline stage: 64 (A stage 32 KB + two dout tiles), or 32 for C_out > 128 -- a 64-n stage of a
// 256-row dout tile is 96 KB, i.e. ONE CTA per SM with two stages, which streamed at 3.1 TB/s (ncu launch list of the
// step, the 21 launches of the C_out = 256 layers); 32-n stages are 48 KB: two CTAs per SM like the narrow layers.
template <int UNIT> struct DwCfg {
    static constexpr uint32_t A_PART = 16 * UNIT * 16;   // 16 kk-groups x UNIT n x 16 B
    static constexpr uint32_t A_STAGE = 2 * A_PART;      // hi + lo
    static constexpr int NB = UNIT / 32;                 // 32-n dout tiles per stage
    static constexpr int UPT = 128 / UNIT;               // stages per 128-row forward tile
};

struct DwParams {
    const uint8_t *G;  // forward tiles [row_tiles][g_k_blocks]
    const uint8_t *B;  // dout tiles    [n_tiles][b_k_blocks] of trb rows
    int g_k_blocks, b_k_blocks, units, trb, stages, split_k;
    uint32_t tmem_cols;
    float *dW;
    int ck, c_out, kperm;
};

template <bool TMA, int UNIT>
__global__ void __launch_bounds__(192) umma_dw_kernel(DwParams p, const __grid_constant__ CUtensorMap g_map) {
    constexpr uint32_t A_PART = DwCfg<UNIT>::A_PART, A_STAGE = DwCfg<UNIT>::A_STAGE;
    constexpr int NB = DwCfg<UNIT>::NB, UPT = DwCfg<UNIT>::UPT;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (p.units + p.split_k - 1) / p.split_k;
    const int u0 = blockIdx.z * per;
    const int nu = min(p.units, u0 + per) - u0;
    if (nu <= 0) return;  // uniform for the CTA

    const uint32_t b_tile = (uint32_t)tile_bytes(p.trb);
    const uint32_t stage_bytes = A_STAGE + NB * b_tile;
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bars = base + p.stages * stage_bytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    const uint32_t accum_bar = bars + 16u * p.stages;
    const uint32_t tmem_slot = accum_bar + 8u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform
    if (warp_u != 5) {
        // producers: every lane owns one (forward k-block, part, 8-kk group) piece of the A stage.  A warp can issue
        // these 1-KB bulk copies only at ~26 GB/s (measured: one issuing warp per SM streams 3.8 TB/s, two 6.5 TB/s,
        // while 16-KB copies reach 7.5 TB/s from a single warp), so one warp PER PIPELINE STAGE issues them: warp w
        // owns ring slot w (iterations w, w + stages, ...), which also keeps every waiter at most one phase behind its
        // barrier (a warp roaming over slots could run two phases ahead and alias the parity).  Warps 0-3 turn into
        // the epilogue afterwards.
        const int NPW = p.stages;  // <= 4
        const int j = lane >> 3, part = (lane >> 2) & 1, kc = lane & 3;
        const int kb = 4 * blockIdx.x + j;
        const bool has = kb < p.g_k_blocks;
        const int n_kb = min(4, p.g_k_blocks - 4 * (int)blockIdx.x);
        const uint32_t tx = (uint32_t)n_kb * 8u * (UNIT * 16) + NB * b_tile;
        const uint32_t a_dst = (uint32_t)part * A_PART + (uint32_t)(j * 4 + kc) * (UNIT * 16);
        const size_t a_off = (size_t)kb * tile_bytes(TR_A) + (size_t)part * part_bytes(TR_A) + (size_t)kc * (TR_A * 16);
        const uint8_t *b_src = p.B + ((size_t)blockIdx.y * p.b_k_blocks) * b_tile;
        if (TMA && warp == 0 && lane == 0) tma_prefetch_desc(&g_map);
        for (int i = warp; i < nu && warp < NPW; i += NPW) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            const int u = u0 + i;
            mbar_wait(empty_bar(s), ph ^ 1u);
            const uint32_t st = base + s * stage_bytes;
            if (TMA) {
                // k-blocks past the end of this row tile's K range (ck not a multiple of 128) come from the next row
                // tile or are zero-filled: they only feed dW^T rows >= ck, which the epilogue drops
                if (lane == 0) mbar_arrive_expect_tx(full_bar(s), A_STAGE + NB * b_tile);
                __syncwarp();
                if (lane < 2)
                    tma_load_4d(st + lane * A_PART, &g_map, full_bar(s), (u % UPT) * (UNIT * 4), 0, lane,
                                (u / UPT) * p.g_k_blocks + 4 * (int)blockIdx.x);
            } else {
                if (lane == 0) mbar_arrive_expect_tx(full_bar(s), tx);
                __syncwarp();
                if (has)
                    bulk_g2s(st + a_dst, p.G + (size_t)(u / UPT) * p.g_k_blocks * tile_bytes(TR_A) + a_off + (size_t)(u % UPT) * (UNIT * 16),
                             UNIT * 16, full_bar(s));
            }
            if (lane < NB)
                bulk_g2s(st + A_STAGE + lane * b_tile, b_src + (size_t)(NB * u + lane) * b_tile, b_tile, full_bar(s));
        }
    } else {
        // the MMA warp runs converged and issues through elect.sync (epn_umma.cuh, "warp-converged issue")
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = instr_desc_bf16_m128(p.trb) | (1u << 15);  // A is MN-major
        const uint32_t b_lbo = (uint32_t)p.trb * 16;
        for (int i = 0; i < nu; ++i) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a0 = base + s * stage_bytes, b0 = a0 + A_STAGE;
#pragma unroll
            for (int ks = 0; ks < UNIT / 16; ++ks) {
                // 16 n per MMA = two 8-n groups 128 B apart (LBO); 8-kk groups UNIT*16 B apart (SBO)
                // (field order checked on the B200: the swapped assignment gives garbage)
                const uint64_t a_hi = smem_desc(a0 + ks * 256, 128, UNIT * 16);
                const uint64_t a_lo = smem_desc(a0 + A_PART + ks * 256, 128, UNIT * 16);
                const uint32_t bb = b0 + (ks >> 1) * b_tile + (ks & 1) * 2 * b_lbo;
                const uint64_t b_hi = smem_desc(bb, b_lbo, 128);
                const uint64_t b_lo = smem_desc(bb + (uint32_t)part_bytes(p.trb), b_lbo, 128);
                mma_bf16_ss_elect(tmem_u, a_hi, b_hi, idesc, (i | ks) != 0);
                mma_bf16_ss_elect(tmem_u, a_hi, b_lo, idesc, 1);
                mma_bf16_ss_elect(tmem_u, a_lo, b_hi, idesc, 1);
            }
            mma_commit_elect(empty_bar(s));
        }
        mma_commit_elect(accum_bar);
    }
    if (warp < 4) {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int kk_t = blockIdx.x * TR_A + warp * 32 + lane;  // lanes <-> consecutive tile rows
        const int kk = (p.kperm && kk_t < p.ck) ? inter_kperm_inv(kk_t, p.kperm) : kk_t;  // row of dW^T in the weight's own order
        for (int c0 = 0; c0 < p.trb; c0 += 32) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                const int o = blockIdx.y * p.trb + c0 + jj;
                if (kk < p.ck && c0 + jj < p.trb && o < p.c_out) atomicAdd(p.dW + (size_t)o * p.ck + kk, v[jj]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// Tensor map over the forward tile array (see the header comment): dims innermost first.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tile_map(CUtensorMap *map, const void *tiles, unsigned long long n_kblocks_total, int unit) {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    if (fn == nullptr) return 1;
    const cuuint64_t dims[4] = {512, 4, 2, n_kblocks_total};     // 32-bit words
    const cuuint64_t strides[3] = {2048, 8192, 16384};            // bytes, dims 1..3
    const cuuint32_t box[4] = {(cuuint32_t)unit * 4, 4, 1, 4};    // `unit` rows x 16 B (64 rows = 256 words)
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void *>(tiles), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace

// G_tiles: forward operand tiles of one slab (n rows, n % 128 == 0, K = ck); B_tiles: dout tiles (rows = c_out, K = n)
int launch_umma_dw(const void *G_tiles, const void *B_tiles, int ck, int c_out, long long n, int trb, float *dW,
                   int kperm, cudaStream_t s) {
    if (n % 128 != 0) {
        set_error("umma_dw: n must be a multiple of 128");
        return EPN_ERR_SHAPE;
    }
    static const int unit32_from = getenv("EPN_DW_UNIT32_FROM") ? atoi(getenv("EPN_DW_UNIT32_FROM")) : 129;   // tuning knob: dout-tile rows from which UNIT = 32
    const int unit = trb >= unit32_from ? 32 : 64;
    if (n / unit >= (1LL << 31)) {
        set_error("umma_dw: n too large");
        return EPN_ERR_SHAPE;
    }
    static DynSmemOnce once, once_tma, once32, once32_tma;
    if (int rc = ensure_dyn_smem(once, umma_dw_kernel<false, 64>, 220 * 1024, "umma_dw_kernel")) return rc;
    if (int rc = ensure_dyn_smem(once_tma, umma_dw_kernel<true, 64>, 220 * 1024, "umma_dw_kernel")) return rc;
    if (int rc = ensure_dyn_smem(once32, umma_dw_kernel<false, 32>, 220 * 1024, "umma_dw_kernel")) return rc;
    if (int rc = ensure_dyn_smem(once32_tma, umma_dw_kernel<true, 32>, 220 * 1024, "umma_dw_kernel")) return rc;
    DwParams p;
    p.G = static_cast<const uint8_t *>(G_tiles);
    p.B = static_cast<const uint8_t *>(B_tiles);
    p.g_k_blocks = (ck + KB - 1) / KB;
    p.b_k_blocks = (int)(n / KB);
    p.units = (int)(n / unit);
    p.trb = trb;
    const size_t stage = (unit == 64 ? DwCfg<64>::A_STAGE : DwCfg<32>::A_STAGE) + (size_t)(unit / 32) * tile_bytes(trb);
    int stages;
    if (2 * stage <= 110 * 1024) {
        stages = (int)((110 * 1024) / stage);  // two CTAs per SM: one's RED epilogue overlaps the other's main loop
    } else {
        stages = (int)((200 * 1024) / stage);
        if (stages > 4) stages = 4;
        if (stages < 2) stages = 2;
    }
    p.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < trb) cols *= 2;
    p.tmem_cols = cols;
    p.dW = dW;
    p.ck = ck;
    p.c_out = c_out;
    p.kperm = kperm;
    const int m_tiles = (ck + TR_A - 1) / TR_A, n_tiles = (c_out + trb - 1) / trb;
    static const int waves = getenv("EPN_DW_WAVES") ? atoi(getenv("EPN_DW_WAVES")) : 3;
    long long sk = (148LL * waves + (long long)m_tiles * n_tiles - 1) / ((long long)m_tiles * n_tiles);
    const long long maxk = p.units / 8 > 0 ? p.units / 8 : 1;
    if (sk > maxk) sk = maxk;
    if (sk > 65535) sk = 65535;
    p.split_k = (int)sk;
    dim3 grid(m_tiles, n_tiles, (unsigned)sk);
    const size_t smem = (size_t)stages * stage + 128 + 16 * stages + 32;
    ProfScope prof(s, KC_GEMM);
    static const int use_tma = (getenv("EPN_DW_TMA") && atoi(getenv("EPN_DW_TMA")) == 0) ? 0 : 1;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const bool tma = use_tma && encode_tile_map(&map, G_tiles, (unsigned long long)(n / TR_A) * p.g_k_blocks, unit) == 0;
    if (unit == 64) {
        if (tma) umma_dw_kernel<true, 64><<<grid, 192, smem, s>>>(p, map);
        else umma_dw_kernel<false, 64><<<grid, 192, smem, s>>>(p, map);
    } else {
        if (tma) umma_dw_kernel<true, 32><<<grid, 192, smem, s>>>(p, map);
        else umma_dw_kernel<false, 32><<<grid, 192, smem, s>>>(p, map);
    }
    return check_launch("umma_dw_kernel");
}

}  // namespace epn
