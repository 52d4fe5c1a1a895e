// Inter grouping stage that emits tensor-core operand tiles directly.
//
//   G[(c,k), (z,p,a)] = sum_n w(p,a,k,n) * feats[z, c, idx[z,p,n], a]      (spconv/functional.py:361-390,
//   w(p,a,k,n) = relu(1 - |x_idx - x_p - R_a kappa_k|^2 / sigma)             so3conv/functional.py:180-218)
//
// One CTA per output point.  lane <-> anchor a (feature rows [c, q, 0..na) are contiguous 4*na-byte
// segments), warp-pair <-> group of KG kernel points; a thread keeps its w[KG][NN] slice of the kernel
// weights in registers for the whole point.  Channels stream through in chunks of CCH: the K neighbour
// rows of every channel of a chunk are staged in shared memory with 16-byte cp.async (double buffered,
// so the gather of chunk i+1 overlaps the FMAs of chunk i), the per-anchor spatial contraction writes
// fp32 results to a shared staging tile, and a conversion pass splits them into bf16 hi/lo and stores
// 16-byte (forward layout) or 8-byte (transposed layout) pieces of the canonical UMMA operand tiles
// (epn_umma.cuh) -- the grouped tensor never exists in fp32 in global memory and no separate
// conversion kernel runs.
//   mode 0: tiles of A[rows = (z,p,a) columns, K = (c,k)]   -> forward GEMM  out = G^T-rows x W
//   mode 1: tiles of A[rows = (c,k),           K = columns] -> dW GEMM       dW^T = G x dout^T
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

constexpr int GT_LANES = 64;   // anchor lanes per kernel-point group
// smem strides are runtime values: feature rows are na floats (na % 4 == 0 -> 16-byte aligned), staging rows
// na|1 floats (odd -> conflict-free for both conversion mappings)
constexpr int GT_KS = 24;

struct TileOut {
    uint8_t *tiles;
    int k_blocks;          // K blocks (of 32) of the tile matrix
    int row_limit;         // mode 1: row_tiles * 128
    long long cols_per_z;  // flattened column of (z, pl, a) = z * cols_per_z + pl * na + a
    int mode;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NN, int KG, int CCH>
__global__ void __launch_bounds__(GT_LANES *(GT_KS / KG), NN == 16 ? 2 : 1)
inter_group_tiles_kernel(const float *__restrict__ feats, const int32_t *__restrict__ idx, InterGeom g, TileOut out,
                         int c, int p_in, int p, int nn, int na, int p_off) {
    extern __shared__ __align__(16) float s_dyn[];
    float *s_g = s_dyn;                                             // [NN][3] (+pad)
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_dyn + NN * 3);   // [NN]
    const int GT_FROW = na, GT_GSTRIDE = na | 1;
    float *Fs = s_dyn + NN * 4;                                     // [2][CCH][NN][na]
    float *Gs = Fs + 2 * CCH * NN * GT_FROW;                        // [CCH*24][na|1]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int a = tid % GT_LANES, grp = tid / GT_LANES;
    const int k0 = grp * KG;
    const bool a_ok = a < na;
    const int aa = a_ok ? a : 0;
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    const float *F = feats ? feats + (size_t)z * c * p_in * na : nullptr;

    for (int n = tid; n < NN; n += nthr) {
        int q = 0;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (n < nn) {
            q = idx[((size_t)z * p + pi) * nn + n];
            const float *X = g.xyz + (size_t)z * 3 * p_in;
            const float *Cn = g.centers + (size_t)z * 3 * p;
            gx = X[q] - Cn[pi];
            gy = X[p_in + q] - Cn[p + pi];
            gz = X[2 * p_in + q] - Cn[2 * p + pi];
        }
        s_idx[n] = q;
        s_g[n * 3] = gx; s_g[n * 3 + 1] = gy; s_g[n * 3 + 2] = gz;
    }
    __syncthreads();

    // kernel weights of this thread's (anchor, kernel-point group): registers for the whole point
    float w[KG][NN];
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
            for (int n = 0; n < NN; ++n) {
                const float v = kernel_weight(s_g[n * 3], s_g[n * 3 + 1], s_g[n * 3 + 2], rx, ry, rz, g.sigma);
                w[i][n] = (a_ok && n < nn) ? v : 0.f;
            }
        }
    }

    const int nchunks = (c + CCH - 1) / CCH;
    // Gather = one bulk async copy (UBLKCP) per neighbour feature row (4*na contiguous bytes), completing
    // on the mbarrier of the stage buffer; thread t issues row (channel t / NN, neighbour t % NN).
    __shared__ __align__(8) uint64_t s_bar[2];
    const uint32_t fs_u32 = smem_u32(Fs);
    const uint32_t bar0 = smem_u32(&s_bar[0]);  // buffer b uses the barrier at bar0 + 8*b
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const uint32_t row_bytes = (uint32_t)na * 4u;
    auto issue = [&](int chunk, int buf) {
        if (F == nullptr) return;
        const int nch = min(CCH, c - chunk * CCH);
        const uint32_t bar = bar0 + 8u * (uint32_t)buf;
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nch * nn) * row_bytes);
        for (int t = tid; t < nch * NN; t += nthr) {
            const int cl = t / NN, n = t % NN;
            if (n < nn)
                bulk_g2s(fs_u32 + (uint32_t)(((buf * CCH + cl) * NN + n) * GT_FROW) * 4u,
                         F + ((size_t)(chunk * CCH + cl) * p_in + s_idx[n]) * na, row_bytes, bar);
        }
    };

    const long long col0 = (long long)z * out.cols_per_z + (long long)pl * na;
    uint32_t phase_bits = 0u;  // bit b = parity to wait for on buffer b
    issue(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
        if (F != nullptr) {
            mbar_wait(bar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
            phase_bits ^= 1u << buf;
        }
        __syncthreads();  // previous conversion finished reading Gs

        // ---- spatial contraction of CCH channels
        for (int cl = 0; cl < CCH; ++cl) {
            const int cc = chunk * CCH + cl;
            float acc[KG];
#pragma unroll
            for (int i = 0; i < KG; ++i) acc[i] = 0.f;
            if (cc < c) {
                const float *frow = Fs + ((buf * CCH + cl) * NN) * GT_FROW + aa;
#pragma unroll
                for (int n = 0; n < NN; ++n) {
                    const float f = (F != nullptr) ? ((n < nn) ? frow[n * GT_FROW] : 0.f) : 1.0f;
#pragma unroll
                    for (int i = 0; i < KG; ++i) acc[i] = fmaf(w[i][n], f, acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < KG; ++i)
                if (a_ok) Gs[(cl * GT_KS + k0 + i) * GT_GSTRIDE + a] = acc[i];
        }
        __syncthreads();

        // ---- fp32 staging -> bf16 hi/lo operand tiles
        const long long kk0 = (long long)chunk * CCH * GT_KS;
        if (out.mode == 0) {
            for (int t = tid; t < CCH * 3 * GT_LANES; t += nthr) {
                const int ra = t % GT_LANES, kc = t / GT_LANES;
                const long long kkg = kk0 + kc * 8;
                if (ra >= na || kkg >= (long long)out.k_blocks * KB) continue;
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = Gs[(kc * 8 + i) * GT_GSTRIDE + ra];
                uint4 hi, lo;
                split8(x, hi, lo);
                const long long row = col0 + ra;
                uint8_t *dst = out.tiles + ((size_t)(row >> 7) * out.k_blocks + (size_t)(kkg >> 5)) * tile_bytes(TR_A) +
                               (size_t)((kkg & 31) >> 3) * (TR_A * 16) + (size_t)(row & 127) * 16;
                *reinterpret_cast<uint4 *>(dst) = hi;
                *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = lo;
            }
        } else {
            const int quads = na / 4;
            for (int t = tid; t < CCH * GT_KS * quads; t += nthr) {
                const int kkl = t % (CCH * GT_KS), gq = t / (CCH * GT_KS);
                const long long row = kk0 + kkl;
                if (row >= out.row_limit) continue;
                float x[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = Gs[kkl * GT_GSTRIDE + gq * 4 + i];
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

// Returns 1 when the shape is not covered by the tile kernel (the caller uses the slab + split path).
int launch_inter_group_tiles(const float *feats, const int32_t *idx, const InterGeom &g, void *tiles, int k_blocks,
                             int row_limit, long long cols_per_z, int mode, int p_off, int p_cnt, int bc, int c,
                             int p_in, int p, int nn, int na, int ks, cudaStream_t s) {
    if (ks != GT_KS || nn > 32 || na > GT_LANES || (na % 4) != 0 || bc > 65535) return 1;
    TileOut o{static_cast<uint8_t *>(tiles), k_blocks, row_limit, cols_per_z, mode};
    dim3 grid(p_cnt, bc);
    ProfScope prof(s, KC_INTER_GROUP);
    if (nn <= 16) {
        constexpr int NN = 16, KG = 6, CCH = 8;
        const size_t smem = (size_t)(NN * 4 + 2 * CCH * NN * na + CCH * GT_KS * (na | 1)) * sizeof(float);
        static bool set = false;
        if (!set) {
            cudaFuncSetAttribute(inter_group_tiles_kernel<NN, KG, CCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
            set = true;
        }
        inter_group_tiles_kernel<NN, KG, CCH><<<grid, GT_LANES *(GT_KS / KG), smem, s>>>(feats, idx, g, o, c, p_in, p, nn, na, p_off);
    } else {
        constexpr int NN = 32, KG = 3, CCH = 4;
        const size_t smem = (size_t)(NN * 4 + 2 * CCH * NN * na + CCH * GT_KS * (na | 1)) * sizeof(float);
        static bool set = false;
        if (!set) {
            cudaFuncSetAttribute(inter_group_tiles_kernel<NN, KG, CCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
            set = true;
        }
        inter_group_tiles_kernel<NN, KG, CCH><<<grid, GT_LANES *(GT_KS / KG), smem, s>>>(feats, idx, g, o, c, p_in, p, nn, na, p_off);
    }
    return check_launch("inter_group_tiles_kernel");
}

}  // namespace epn
