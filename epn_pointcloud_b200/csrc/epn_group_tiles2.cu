// (1) Intra grouping straight into operand tiles:
//       G'[(c,k), (z,p,a)] = feats[z, c, p, intra_idx[a,k]]          (so3conv/functional.py:221-268)
//     a pure gather: thread = (tile row, 8-wide k chunk), 8 gathered loads (L1-resident 4*na-byte feature
//     rows) -> bf16 hi/lo split -> two 16-byte tile stores, lanes along rows (contiguous stores).
// (2) Inter grouping backward (scatter) with cp.async-staged gradient rows:
//       dfeats[z, c, idx[z,p,n], a] += sum_k w(p,a,k,n) * dG[(c,k), (z,p,a)]
//     thread = (anchor lane, 4 neighbours) holding w[24][4] in registers; the 24 rows of dG of every
//     channel of a chunk are staged in shared memory (double buffered) and each (channel, neighbour)
//     result goes out as one fp32 RED into the neighbour's 4*na-byte feature row.
#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

// ------------------------------------------------------------------ intra tiles
__global__ void __launch_bounds__(256)
intra_group_tiles_kernel(const float *__restrict__ feats, const int32_t *__restrict__ intra_idx,
                         uint8_t *__restrict__ tiles, int k_blocks, int mode, int c, int p, int na, int kn, int p_off,
                         int p_cnt, int n_slab, int rows_pad) {
    extern __shared__ int32_t s_ix[];  // [na*kn]
    for (int i = threadIdx.x; i < na * kn; i += blockDim.x) s_ix[i] = intra_idx[i];
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows_pad) return;
    const int ck = c * kn, cols = p_cnt * na, kcgs = k_blocks * (KB / 8);
    // the index decoded once per thread (row) ...
    int rz = 0, rpl = 0, ra = 0, rcc = 0, rk = 0;
    if (mode == 0) { rz = row / cols; const int rem = row - rz * cols; rpl = rem / na; ra = rem - rpl * na; }
    else { rcc = row / kn; rk = row - rcc * kn; }
    for (int kcg = blockIdx.y; kcg < kcgs; kcg += gridDim.y) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kidx = kcg * 8 + i;  // ... and per element for the k index
            int z, pl, a, cc, k;
            bool ok;
            if (mode == 0) {
                z = rz; pl = rpl; a = ra;
                cc = kidx / kn; k = kidx - cc * kn;
                ok = row < n_slab && kidx < ck;
            } else {
                cc = rcc; k = rk;
                z = kidx / cols; const int rem = kidx - z * cols; pl = rem / na; a = rem - pl * na;
                ok = row < ck && kidx < n_slab;
            }
            x[i] = ok ? __ldg(feats + (((size_t)z * c + cc) * p + p_off + pl) * na + s_ix[a * kn + k]) : 0.f;
        }
        uint4 hi, lo;
        split8(x, hi, lo);
        uint8_t *dst = tiles + ((size_t)(row >> 7) * k_blocks + (size_t)(kcg >> 2)) * tile_bytes(TR_A) +
                       (size_t)(kcg & 3) * (TR_A * 16) + (size_t)(row & 127) * 16;
        *reinterpret_cast<uint4 *>(dst) = hi;
        *reinterpret_cast<uint4 *>(dst + part_bytes(TR_A)) = lo;
    }
}

int launch_intra_group_tiles(const float *feats, const int32_t *intra_idx, void *tiles, int mode, int p_off, int p_cnt,
                             int bc, int c, int p, int na, int kn, cudaStream_t s) {
    const long long n_slab = (long long)bc * p_cnt * na;
    const int ck = c * kn;
    if (n_slab >= (1LL << 31) || (size_t)na * kn * 4 > 40 * 1024) return 1;
    const long long rows = mode == 0 ? n_slab : ck, K = mode == 0 ? ck : n_slab;
    const int rows_pad = (int)((rows + 127) / 128 * 128), k_blocks = (int)((K + KB - 1) / KB);
    const int kcgs = k_blocks * (KB / 8);
    dim3 grid((rows_pad + 255) / 256, kcgs < 65535 ? kcgs : 65535);
    ProfScope prof(s, KC_INTRA_GROUP);
    intra_group_tiles_kernel<<<grid, 256, (size_t)na * kn * sizeof(int32_t), s>>>(
        feats, intra_idx, static_cast<uint8_t *>(tiles), k_blocks, mode, c, p, na, kn, p_off, p_cnt, (int)n_slab, rows_pad);
    return check_launch("intra_group_tiles_kernel");
}

// ------------------------------------------------------------------ inter scatter
constexpr int SC_LANES = 64, SC_KS = 24, SC_NB = 4, SC_CCH = 4;

__device__ __forceinline__ void sc_cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int NN>
__global__ void __launch_bounds__(SC_LANES *(NN / SC_NB), NN == 16 ? 2 : 1)
inter_scatter_kernel(const float *__restrict__ dG, long long stride_b, long long stride_ck, const int32_t *__restrict__ idx,
                     InterGeom g, float *__restrict__ dfeats, int c, int p_in, int p, int nn, int na, int p_off) {
    extern __shared__ __align__(16) float s_dyn[];
    float *s_g = s_dyn;                                            // [NN][3]
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_dyn + NN * 3);  // [NN]
    float *Ds = s_dyn + NN * 4;                                    // [2][SC_CCH*24][na]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int a = tid % SC_LANES, grp = tid / SC_LANES;
    const int n0 = grp * SC_NB;
    const bool a_ok = a < na;
    const int aa = a_ok ? a : 0;
    const int z = blockIdx.y, pl = blockIdx.x, pi = p_off + pl;
    float *DF = dfeats + (size_t)z * c * p_in * na;

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

    float w[SC_KS][SC_NB];
    {
        float R[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(g.anchors + aa * 9 + i);
#pragma unroll
        for (int k = 0; k < SC_KS; ++k) {
            const float kx = __ldg(g.kernels + k * 3), ky = __ldg(g.kernels + k * 3 + 1), kz = __ldg(g.kernels + k * 3 + 2);
            const float rx = R[0] * kx + R[1] * ky + R[2] * kz, ry = R[3] * kx + R[4] * ky + R[5] * kz,
                        rz = R[6] * kx + R[7] * ky + R[8] * kz;
#pragma unroll
            for (int j = 0; j < SC_NB; ++j) {
                const int n = n0 + j;
                const float v = kernel_weight(s_g[n * 3], s_g[n * 3 + 1], s_g[n * 3 + 2], rx, ry, rz, g.sigma);
                w[k][j] = (a_ok && n < nn) ? v : 0.f;
            }
        }
    }
    int q[SC_NB];
#pragma unroll
    for (int j = 0; j < SC_NB; ++j) q[j] = s_idx[n0 + j];

    const int nchunks = (c + SC_CCH - 1) / SC_CCH;
    // one bulk async copy (UBLKCP) per staged gradient row (4*na contiguous bytes) on the buffer's mbarrier
    __shared__ __align__(8) uint64_t s_bar[2];
    const uint32_t ds_u32 = smem_u32(Ds);
    const uint32_t bar0 = smem_u32(&s_bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const uint32_t row_bytes = (uint32_t)na * 4u;
    const float *src0 = dG + (size_t)z * stride_b + (size_t)pl * na;
    auto issue = [&](int chunk, int buf) {
        const int rows = min(SC_CCH, c - chunk * SC_CCH) * SC_KS;
        const uint32_t bar = bar0 + 8u * (uint32_t)buf;
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)rows * row_bytes);
        for (int rr = tid; rr < rows; rr += nthr)  // rr = cl*24 + k
            bulk_g2s(ds_u32 + (uint32_t)((buf * SC_CCH * SC_KS + rr) * na) * 4u,
                     src0 + (size_t)(chunk * SC_CCH * SC_KS + rr) * stride_ck, row_bytes, bar);
    };

    uint32_t phase_bits = 0u;
    issue(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) issue(chunk + 1, buf ^ 1);
        mbar_wait(bar0 + 8u * (uint32_t)buf, (phase_bits >> buf) & 1u);
        phase_bits ^= 1u << buf;
        for (int cl = 0; cl < SC_CCH; ++cl) {
            const int cc = chunk * SC_CCH + cl;
            if (cc >= c) break;
            float t[SC_NB];
#pragma unroll
            for (int j = 0; j < SC_NB; ++j) t[j] = 0.f;
            const float *drow = Ds + (size_t)((buf * SC_CCH + cl) * SC_KS) * na + aa;
#pragma unroll
            for (int k = 0; k < SC_KS; ++k) {
                const float dv = drow[k * na];
#pragma unroll
                for (int j = 0; j < SC_NB; ++j) t[j] = fmaf(w[k][j], dv, t[j]);
            }
#pragma unroll
            for (int j = 0; j < SC_NB; ++j)
                if (a_ok && n0 + j < nn) atomicAdd(DF + ((size_t)cc * p_in + q[j]) * na + a, t[j]);
        }
        __syncthreads();  // all reads of Ds[buf] done before the loads of chunk+2 overwrite it
    }
}

// Returns 1 when the shape is not covered (caller falls back to inter_group_bwd_kernel).
int launch_inter_scatter(const float *dG, long long stride_b, long long stride_ck, const int32_t *idx,
                         const InterGeom &g, float *dfeats, int p_off, int p_cnt, int bc, int c, int p_in, int p, int nn,
                         int na, int ks, cudaStream_t s) {
    if (ks != SC_KS || nn > 32 || na > SC_LANES || (na % 4) != 0 || bc > 65535 || (stride_ck % 4) != 0 ||
        (stride_b % 4) != 0 || ((uintptr_t)dG & 15) != 0)
        return 1;
    dim3 grid(p_cnt, bc);
    ProfScope prof(s, KC_INTER_SCATTER);
    static bool set = false;
    if (!set) {
        cudaFuncSetAttribute(inter_scatter_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(inter_scatter_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        set = true;
    }
    if (nn <= 16) {
        const size_t smem = (size_t)(16 * 4 + 2 * SC_CCH * SC_KS * na) * sizeof(float);
        inter_scatter_kernel<16><<<grid, SC_LANES * (16 / SC_NB), smem, s>>>(dG, stride_b, stride_ck, idx, g, dfeats, c, p_in,
                                                                           p, nn, na, p_off);
    } else {
        const size_t smem = (size_t)(32 * 4 + 2 * SC_CCH * SC_KS * na) * sizeof(float);
        inter_scatter_kernel<32><<<grid, SC_LANES * (32 / SC_NB), smem, s>>>(dG, stride_b, stride_ck, idx, g, dfeats, c, p_in,
                                                                           p, nn, na, p_off);
    }
    return check_launch("inter_scatter_kernel");
}

}  // namespace epn
