// Weight gradient straight from the operand tiles the forward pass kept.
//
//   dW[o, kk] += sum_n dout[o, n] * G[n, kk]          (so3conv/modules.py:48-55 under autograd)
//
// The forward GEMM consumed G as K-major tiles with rows = grouped columns n and K = kk = (c,k):
//     tile (rt, kb) = [part hi|lo][kc: 4][row: 128][8 x bf16 of kk]
// For dW the contraction runs over n, so G enters as the M operand (M = kk) and n is the MMA K dimension.
// Read that way the very same bytes are the canonical MN-MAJOR no-swizzle layout of tcgen05: 8 kk contiguous
// (16 bytes), consecutive n 16 bytes apart, i.e. one 8(k) x 8(mn) core matrix = 128 contiguous bytes.  So the
// backward pass neither recomputes the spatial contraction nor transposes anything: the producer copies
// 1-KB pieces (64 n of one 8-kk group) of four adjacent forward tiles into a stage
//     A stage = [part][kk group: 16][n: 64][8 kk]     SBO (8-kk group stride) = 1 KB, LBO (8-n group stride) = 128 B
// and the MMA is issued with a_major = MN.  B = dout tiles (rows = o, K = n), K-major as everywhere else.
#include <stdlib.h>

#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

namespace {

constexpr int UNIT = 64;                          // n per pipeline stage
constexpr uint32_t A_PART = 16 * UNIT * 16;       // 16 kk-groups x 64 n x 16 B
constexpr uint32_t A_STAGE = 2 * A_PART;          // hi + lo = 32 KB

struct DwParams {
    const uint8_t *G;  // forward tiles [row_tiles][g_k_blocks]
    const uint8_t *B;  // dout tiles    [n_tiles][b_k_blocks] of trb rows
    int g_k_blocks, b_k_blocks, units, trb, stages, split_k;
    uint32_t tmem_cols;
    float *dW;
    int ck, c_out, kperm;
};

__global__ void __launch_bounds__(192) umma_dw_kernel(DwParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (p.units + p.split_k - 1) / p.split_k;
    const int u0 = blockIdx.z * per;
    const int nu = min(p.units, u0 + per) - u0;
    if (nu <= 0) return;  // uniform for the CTA

    const uint32_t b_tile = (uint32_t)tile_bytes(p.trb);
    const uint32_t stage_bytes = A_STAGE + 2 * b_tile;
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

    if (warp != 5) {
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
        const uint32_t tx = (uint32_t)n_kb * 8u * (UNIT * 16) + 2 * b_tile;
        const uint32_t a_dst = (uint32_t)part * A_PART + (uint32_t)(j * 4 + kc) * (UNIT * 16);
        const size_t a_off = (size_t)kb * tile_bytes(TR_A) + (size_t)part * part_bytes(TR_A) + (size_t)kc * (TR_A * 16);
        const uint8_t *b_src = p.B + ((size_t)blockIdx.y * p.b_k_blocks) * b_tile;
        for (int i = warp; i < nu && warp < NPW; i += NPW) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            const int u = u0 + i;
            mbar_wait(empty_bar(s), ph ^ 1u);
            if (lane == 0) mbar_arrive_expect_tx(full_bar(s), tx);
            __syncwarp();
            const uint32_t st = base + s * stage_bytes;
            if (has)
                bulk_g2s(st + a_dst, p.G + (size_t)(u >> 1) * p.g_k_blocks * tile_bytes(TR_A) + a_off + (size_t)(u & 1) * (UNIT * 16),
                         UNIT * 16, full_bar(s));
            if (lane < 2)
                bulk_g2s(st + A_STAGE + lane * b_tile, b_src + (size_t)(2 * u + lane) * b_tile, b_tile, full_bar(s));
        }
    } else {
        if (lane == 0) {
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
                    mma_bf16_ss(tmem_base, a_hi, b_hi, idesc, (i | ks) != 0);
                    mma_bf16_ss(tmem_base, a_hi, b_lo, idesc, 1);
                    mma_bf16_ss(tmem_base, a_lo, b_hi, idesc, 1);
                }
                mma_commit(empty_bar(s));
            }
            mma_commit(accum_bar);
        }
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

}  // namespace

// G_tiles: forward operand tiles of one slab (n rows, n % 128 == 0, K = ck); B_tiles: dout tiles (rows = c_out, K = n)
int launch_umma_dw(const void *G_tiles, const void *B_tiles, int ck, int c_out, long long n, int trb, float *dW,
                   int kperm, cudaStream_t s) {
    if (n % 128 != 0 || n / UNIT >= (1LL << 31)) {
        set_error("umma_dw: n must be a multiple of 128");
        return EPN_ERR_SHAPE;
    }
    static DynSmemOnce once;
    if (int rc = ensure_dyn_smem(once, umma_dw_kernel, 220 * 1024, "umma_dw_kernel")) return rc;
    DwParams p;
    p.G = static_cast<const uint8_t *>(G_tiles);
    p.B = static_cast<const uint8_t *>(B_tiles);
    p.g_k_blocks = (ck + KB - 1) / KB;
    p.b_k_blocks = (int)(n / KB);
    p.units = (int)(n / UNIT);
    p.trb = trb;
    const size_t stage = A_STAGE + 2 * tile_bytes(trb);
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
    umma_dw_kernel<<<grid, 192, smem, s>>>(p);
    return check_launch("umma_dw_kernel");
}

}  // namespace epn
