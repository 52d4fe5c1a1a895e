// Tensor-core channel GEMM of the SPConv path on tcgen05 (sm_100a):
//     D[128-row tile, N] (+)= sum_kb  A_tile(rt, kb) * B_tile(nt, kb)^T
// Both operands are "split tiles" (bf16 hi/lo, canonical K-major no-swizzle layout, see
// epn_umma.cuh) living in global memory / L2; a stage is moved with two cp.async.bulk
// (UBLKCP) copies that complete on an mbarrier, three tcgen05.mma (hi*hi, hi*lo, lo*hi) per
// 16-wide k step accumulate in fp32 in TMEM, and four epilogue warps read the accumulator with
// tcgen05.ld and write rows (= point/anchor columns of the conv) with coalesced stores or REDs.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 bulk-copy
// producer (one elected lane), warp 5 TMEM allocator + MMA issuer (one elected lane).
#include <stdlib.h>

#include "epn_internal.cuh"
#include "epn_umma.cuh"

namespace epn {
using namespace umma;

// ------------------------------------------------------------------ fp32 -> split tiles
// thread <-> (row, 8-wide k chunk); lanes run along rows so the 16-byte tile stores of a warp are
// contiguous; source reads are coalesced when rows are contiguous in memory and whole 32-byte
// sectors per lane when k is contiguous.
// K_LANES = false: lanes run along rows (use when rows are contiguous in the source: coalesced loads and
// contiguous 16-byte tile stores).  K_LANES = true: lanes run along 8-wide k chunks (use when k is
// contiguous in the source: every lane reads one full 32-byte sector, a warp 1 KB).
// Elements (row, 8 kcg .. 8 kcg + 7) of a SplitSrc matrix (zeros outside it), with the producer's normalisation applied
// when the source carries one.  Written for instruction count -- the conversion kernels are ISSUE-bound, not HBM-bound
// (ncu: 76 % issue-active at 39 % DRAM): the generic form did two 64-bit divisions per thread and a 64-bit multiply +
// wrap test per element; here the z-splits are 32-bit divisions taken only when the matrix really is split (rows and K
// are < 2^31), and the common unsplit-K case walks a pointer.
// row part: pointer to element (row, 0) and the row's z index
__device__ __forceinline__ const float *split_row_base(const SplitSrc &src, int row, int rows, int &z) {
    z = 0;
    if (src.rows_per_z < rows) {
        const uint32_t rpz = (uint32_t)src.rows_per_z;
        const uint32_t zr = (uint32_t)row / rpz, rr = (uint32_t)row - zr * rpz;
        z = (int)zr;
        return src.ptr + (long long)zr * src.stride_rz + (long long)rr * src.stride_row;
    }
    return src.ptr + (long long)row * src.stride_row;
}
// K part: elements 8 kcg .. 8 kcg + 7 of the row at `base`
__device__ __forceinline__ void split_load8_from(const SplitSrc &src, const float *base, int z, int kcg, int K, float (&x)[8]) {
    const int k0 = kcg * 8;
    if (src.k_per_z >= K) {
        const float *p = base + (long long)k0 * src.stride_k;
        if (k0 + 8 <= K) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = __ldg(p);
                p += src.stride_k;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x[i] = (k0 + i < K) ? __ldg(p) : 0.f;
                p += src.stride_k;
            }
        }
    } else {
        const uint32_t kpz = (uint32_t)src.k_per_z;
        uint32_t kz = (uint32_t)k0 / kpz, kj = (uint32_t)k0 - kz * kpz;
        const float *p = base + (long long)kz * src.stride_kz + (long long)kj * src.stride_k;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = (k0 + i < K) ? __ldg(p) : 0.f;
            p += src.stride_k;
            if (++kj == kpz) { kj = 0; ++kz; p = base + (long long)kz * src.stride_kz; }
        }
    }
    if (src.pro.stats != nullptr) {   // normalisation + leaky_relu of the producer, applied on the way in
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (k0 + i < K) x[i] = src.pro.apply(x[i], z, k0 + i);
    }
}
__device__ __forceinline__ void split_load8(const SplitSrc &src, int row, int kcg, int rows, int K, float (&x)[8]) {
    if (row >= rows) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
        return;
    }
    int z;
    const float *base = split_row_base(src, row, rows, z);
    split_load8_from(src, base, z, kcg, K, x);
}

template <bool K_LANES>
__global__ void __launch_bounds__(256)
split_tiles_kernel(SplitSrc src, uint8_t *__restrict__ dst, int rows, int K, int tr, int row_tiles, int k_blocks, int fmt,
                   float scale) {
    const int rows_pad = row_tiles * tr, kcgs = k_blocks * (KB / 8);
    const int x_idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (x_idx >= (K_LANES ? kcgs : rows_pad)) return;
    // lanes along rows: the row part of the address (a division when the matrix is z-split) is computed once per thread and
    // the thread walks over several 8-wide k chunks (the launch gives it ~4), instead of paying it for every chunk
    int z_row = 0;
    const float *row_base = (!K_LANES && x_idx < rows) ? split_row_base(src, x_idx, rows, z_row) : nullptr;
    for (int y = blockIdx.y; y < (K_LANES ? rows_pad : kcgs); y += gridDim.y) {
        const int row = K_LANES ? y : x_idx, kcg = K_LANES ? x_idx : y;
        float x[8];
        if (K_LANES) {
            split_load8(src, row, kcg, rows, K, x);
        } else if (row_base != nullptr) {
            split_load8_from(src, row_base, z_row, kcg, K, x);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 0.f;
        }
        uint4 hi, lo;
        split8_fmt(x, hi, lo, fmt, scale);
        const int rt = row / tr, r = row - rt * tr, kb = kcg / (KB / 8), kc = kcg % (KB / 8);
        uint8_t *tile = dst + ((size_t)rt * k_blocks + kb) * tile_bytes(tr);
        *reinterpret_cast<uint4 *>(tile + (size_t)kc * tr * 16 + (size_t)r * 16) = hi;
        *reinterpret_cast<uint4 *>(tile + part_bytes(tr) + (size_t)kc * tr * 16 + (size_t)r * 16) = lo;
    }
}

// k-contiguous sources (dout as the dW operand, W): a block converts 32 rows x 64 k.  Loads run along k
// (8 lanes x 32 B contiguous per row), the hi/lo 16-byte chunks are transposed through shared memory, and
// the stores run along rows (32 lanes x 16 B = 512 contiguous bytes of the tile) instead of isolated 16-byte
// pieces.
__global__ void __launch_bounds__(256)
split_tiles_kcontig_kernel(SplitSrc src, uint8_t *__restrict__ dst, int rows, int K, int tr, int k_blocks, int fmt, float scale) {
    __shared__ uint4 s_hi[8][33], s_lo[8][33];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.y * 32;
    const int kcgs = k_blocks * (KB / 8), rows_pad = (rows + tr - 1) / tr * tr;
    // the row part of the source address (a division when the matrix is z-split) and of the tile address is computed once;
    // the block then walks over several 64-wide k chunks (the launch gives it ~4)
    const int lr = tid >> 3, lkq = tid & 7;
    int z_row = 0;
    const float *row_base = row0 + lr < rows ? split_row_base(src, row0 + lr, rows, z_row) : nullptr;
    const int skq = tid >> 5, sr = tid & 31;
    const int srow = row0 + sr;
    const int rt = srow / tr, rr = srow - rt * tr;
    uint8_t *tile_row = dst + (size_t)rt * k_blocks * tile_bytes(tr) + (size_t)rr * 16;
    for (int kx = blockIdx.x; kx * 8 < kcgs; kx += gridDim.x) {
        const int kcg0 = kx * 8;
        {
            float x[8];
            if (row_base != nullptr) {
                split_load8_from(src, row_base, z_row, kcg0 + lkq, K, x);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = 0.f;
            }
            split8_fmt(x, s_hi[lkq][lr], s_lo[lkq][lr], fmt, scale);
        }
        __syncthreads();
        {
            const int kcg = kcg0 + skq;
            if (kcg < kcgs && srow < rows_pad) {
                uint8_t *tile = tile_row + (size_t)(kcg >> 2) * tile_bytes(tr) + (size_t)(kcg & 3) * tr * 16;
                *reinterpret_cast<uint4 *>(tile) = s_hi[skq][sr];
                *reinterpret_cast<uint4 *>(tile + part_bytes(tr)) = s_lo[skq][sr];
            }
        }
        __syncthreads();
    }
}

size_t split_tiles_bytes(long long rows, long long K, int tr) {
    const long long row_tiles = (rows + tr - 1) / tr, k_blocks = (K + KB - 1) / KB;
    return (size_t)row_tiles * k_blocks * tile_bytes(tr);
}

int launch_split_tiles(const SplitSrc &src, void *dst, long long rows, long long K, int tr, cudaStream_t s, int fmt, float scale) {
    const int row_tiles = (int)((rows + tr - 1) / tr), k_blocks = (int)((K + KB - 1) / KB);
    const int rows_pad = row_tiles * tr, kcgs = k_blocks * (KB / 8);
    ProfScope prof(s, KC_SPLIT);
    if (src.pro.stats == nullptr && src.stride_k == 1 && src.stride_row != 1 && (rows_pad + 31) / 32 <= 65535) {
        const int kx = (kcgs + 7) / 8;
        dim3 grid((kx + 3) / 4, (rows_pad + 31) / 32);   // ~4 k chunks of 64 per block
        split_tiles_kcontig_kernel<<<grid, 256, 0, s>>>(src, static_cast<uint8_t *>(dst), (int)rows, (int)K, tr, k_blocks, fmt,
                                                        scale);
    } else if (src.stride_k == 1 && src.stride_row != 1) {
        dim3 grid((kcgs + 255) / 256, rows_pad < 65535 ? rows_pad : 65535);
        split_tiles_kernel<true><<<grid, 256, 0, s>>>(src, static_cast<uint8_t *>(dst), (int)rows, (int)K, tr, row_tiles,
                                                      k_blocks, fmt, scale);
    } else {
        const int ychunks = (kcgs + 3) / 4;   // ~4 k chunks per thread: amortises the per-thread set-up (see the kernel)
        dim3 grid((rows_pad + 255) / 256, ychunks < 65535 ? ychunks : 65535);
        split_tiles_kernel<false><<<grid, 256, 0, s>>>(src, static_cast<uint8_t *>(dst), (int)rows, (int)K, tr, row_tiles,
                                                       k_blocks, fmt, scale);
    }
    return check_launch("split_tiles_kernel");
}

// ------------------------------------------------------------------ GEMM
struct UmmaGemmParams {
    const uint8_t *A;  // [m_tiles][k_blocks] tiles of 128 rows
    const uint8_t *B;  // [n_tiles][k_blocks] tiles of trb rows
    int k_blocks, trb, stages;
    uint32_t tmem_cols;
    float *out;
    long long rows_per_z, stride_z, stride_row, stride_col, cols_per_z, stride_cz;
    int m_valid, n_valid, mode, split_k, vec;
    // ep_kind 2 (intra conv data gradient): the tile is reduced through the inverse anchor permutations
    // before it leaves the SM (see launch_umma_intra_dx)
    int ep_kind = 0;
    const int32_t *intra_idx = nullptr;  // [60][12]
    float *dfeats = nullptr;             // [z][ch_total][pts_per_z][60]
    int ch_total = 0, pts_per_z = 0;
    int fmt = 0;               // operand format of both tile arrays (FMT_BF16 / FMT_F16)
    float out_scale = 1.0f;    // the accumulator is multiplied by this as it leaves TMEM (1 / F16_W_SCALE for FMT_F16)
};

__global__ void __launch_bounds__(192)
umma_gemm_kernel(UmmaGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (p.k_blocks + p.split_k - 1) / p.split_k;
    const int kb0 = blockIdx.z * per;
    const int nkb = min(p.k_blocks, kb0 + per) - kb0;
    if (nkb <= 0) return;  // uniform for the CTA

    const uint32_t a_bytes = (uint32_t)tile_bytes(TR_A), b_bytes = (uint32_t)tile_bytes(p.trb);
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bars = base + p.stages * stage_bytes;  // full[stages], empty[stages], accum, tmem slot
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
    // the loader and the MMA warp stay CONVERGED and issue through elect.sync (epn_umma.cuh, "warp-converged issue"):
    // their branch conditions and operands are made provably warp-uniform with shuffles from lane 0
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);

    if (warp_u == 4) {
        const uint8_t *a_src = p.A + ((size_t)blockIdx.x * p.k_blocks + kb0) * a_bytes;
        const uint8_t *b_src = p.B + ((size_t)blockIdx.y * p.k_blocks + kb0) * b_bytes;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            bulk_g2s2_expect_elect(base + s * stage_bytes, a_src + (size_t)i * a_bytes, a_bytes, base + s * stage_bytes + a_bytes,
                                   b_src + (size_t)i * b_bytes, b_bytes, full_bar(s));
        }
    } else if (warp_u == 5) {
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = instr_desc_m128(p.trb, p.fmt);
        const uint32_t a_lbo = TR_A * 16, b_lbo = (uint32_t)p.trb * 16;
        for (int i = 0; i < nkb; ++i) {
            const int s = i % p.stages;
            const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t a0 = base + s * stage_bytes, b0 = a0 + a_bytes;
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks) {
                const uint64_t a_hi = smem_desc(a0 + ks * 2 * a_lbo, a_lbo, 128);
                const uint64_t a_lo = smem_desc(a0 + (uint32_t)part_bytes(TR_A) + ks * 2 * a_lbo, a_lbo, 128);
                const uint64_t b_hi = smem_desc(b0 + ks * 2 * b_lbo, b_lbo, 128);
                const uint64_t b_lo = smem_desc(b0 + (uint32_t)part_bytes(p.trb) + ks * 2 * b_lbo, b_lbo, 128);
                mma_bf16_ss_elect(tmem_u, a_hi, b_hi, idesc, (i | ks) != 0);
                mma_bf16_ss_elect(tmem_u, a_hi, b_lo, idesc, 1);
                mma_bf16_ss_elect(tmem_u, a_lo, b_hi, idesc, 1);
            }
            mma_commit_elect(empty_bar(s));  // frees the stage when these MMAs have read it
        }
        mma_commit_elect(accum_bar);
    } else {
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const long long row = (long long)blockIdx.x * TR_A + warp * 32 + lane;
        const bool row_ok = row < p.m_valid;
        float *dst_row = p.out;
        if (row_ok) dst_row += (row / p.rows_per_z) * p.stride_z + (row % p.rows_per_z) * p.stride_row;  // once per thread
        const bool col_split = p.cols_per_z < (long long)p.n_valid;  // columns run over (cloud, point*anchor)
        const uint32_t cpz = col_split ? (uint32_t)p.cols_per_z : 1u;
        auto col_offset = [&](uint32_t col, long long stride_col) -> long long {
            if (!col_split) return (long long)col * stride_col;
            const uint32_t cz = col / cpz;
            return (long long)cz * p.stride_cz + (long long)(col - cz * cpz) * stride_col;
        };
        const bool row_split = p.rows_per_z < (long long)p.m_valid;
        auto row_offset = [&](long long r) -> long long {
            if (!row_split) return r * p.stride_row;
            const uint32_t rz = (uint32_t)r / (uint32_t)p.rows_per_z;
            return (long long)rz * p.stride_z + (long long)((uint32_t)r - rz * (uint32_t)p.rows_per_z) * p.stride_row;
        };
        // Column-contiguous outputs (dX): transpose through shared memory (the pipeline stages are free once
        // the accumulator is complete) so that every warp store is one 512-byte row segment instead of 32
        // scattered 16-byte pieces.
        const int gw = p.trb < 128 ? p.trb : 128;  // columns per staged group
        if (p.vec && (size_t)4 * 32 * (gw + 4) * sizeof(float) <= (size_t)p.stages * stage_bytes) {
            float *stg = reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw))) + (size_t)warp * 32 * (gw + 4);
            const long long row0 = (long long)blockIdx.x * TR_A + warp * 32;
            for (int g0 = 0; g0 < p.trb; g0 += gw) {
                for (int cc = 0; cc < gw; cc += 32) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g0 + cc), v);
                    if (p.out_scale != 1.0f) {
                    #pragma unroll
                        for (int jq = 0; jq < 32; ++jq) v[jq] *= p.out_scale;
                    }
                    if (cc == 0 && g0 > 0) {  // the previous group's rows have left the staging buffer
                        bulk_wait_read0();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (cc + j < gw)
                            *reinterpret_cast<float4 *>(stg + (size_t)lane * (gw + 4) + cc + j) =
                                make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                fence_proxy_async_smem();  // this lane's staging writes -> visible to the bulk-copy engine
                __syncwarp();
                // one asynchronous bulk store per staged row (<= 512 contiguous bytes): the warp goes straight
                // back to TMEM while the copy engine streams the rows out
                const long long colg = (long long)blockIdx.y * p.trb + g0;
                const long long r = row0 + lane;
                if (r < p.m_valid && colg < p.n_valid) {
                    const int ncols = (int)min((long long)min(gw, p.trb - g0), (long long)p.n_valid - colg);
                    float *dst = p.out + row_offset(r) + col_offset((uint32_t)colg, 1);
                    const float *src = stg + (size_t)lane * (gw + 4);
                    const bool whole = !col_split || ((uint32_t)colg / cpz == (uint32_t)(colg + ncols - 1) / cpz);
                    if (whole && (ncols & 3) == 0) {
                        bulk_s2g(dst, smem_u32(src), (uint32_t)ncols * 4u);
                    } else {
                        for (int e = 0; e < ncols; ++e) p.out[row_offset(r) + col_offset((uint32_t)(colg + e), 1)] = src[e];
                    }
                }
                bulk_commit();
            }
            bulk_wait_read0();  // shared memory must stay valid until the copy engine has read it
        } else
        for (int c0 = 0; c0 < p.trb; c0 += 32) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (p.out_scale != 1.0f) {
            #pragma unroll
                for (int jq = 0; jq < 32; ++jq) v[jq] *= p.out_scale;
            }
            const long long colb = (long long)blockIdx.y * p.trb + c0;
            if (p.vec) {
                // output columns are contiguous in memory: the thread's 32 values go out as 8 x 16 bytes
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const long long col = colb + j;
                    if (row_ok && c0 + j < p.trb && col < p.n_valid) {
                        float *dst = dst_row + col_offset((uint32_t)col, 1);
                        if (col + 3 < p.n_valid) {
                            *reinterpret_cast<float4 *>(dst) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
                            for (int t = 0; t < 4 && col + t < p.n_valid; ++t) dst[t] = v[j + t];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const long long col = colb + j;
                    if (row_ok && c0 + j < p.trb && col < p.n_valid) {
                        float *dst = dst_row + col_offset((uint32_t)col, p.stride_col);
                        if (p.mode) atomicAdd(dst, v[j]);
                        else *dst = v[j];
                    }
                }
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

constexpr int IDX_STR = 244;  // row stride (floats) of the whole-tile staging of the intra data-gradient epilogue
// warps 0-3 epilogue (TMEM readers), 4 loader, 5 MMA, 6-9 helpers of the permuted reduction of the intra conv epilogues
// (ep_kind >= 2): with 120 of the 128 epilogue threads reducing 20 outputs each, the epilogue of a C <= 128 tile took longer
// than its MMAs; the reduction now runs on 240 threads (one tile column each)
constexpr int PERSISTENT_THREADS = 320;

// ------------------------------------------------------------------ persistent GEMM
// Same math and operand format as umma_gemm_kernel, scheduled differently: ONE CTA per SM walks over the output
// tiles (m tile fastest, so neighbouring CTAs share the B tile in L2), the bulk-copy pipeline runs ahead across
// tile boundaries, and the fp32 accumulator is double buffered in TMEM (2 x tmem_cols <= 512 columns) so the
// epilogue of tile t (TMEM -> registers -> [shared staging ->] global) overlaps the main loop of tile t+1.
// Per-tile fixed costs of the one-tile-per-CTA kernel (launch, TMEM allocation, barrier set-up, pipeline fill
// and drain, store drain at exit) are paid once per SM; they dominated GEMMs with a short K loop and a large
// output (dX = W^T dout: K = c_out, 128 KB of output per tile).
__global__ void __launch_bounds__(PERSISTENT_THREADS, 1)
umma_gemm_persistent_kernel(UmmaGemmParams p, int m_tiles, int n_tiles, uint32_t stg_off) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.k_blocks;
    const int total = m_tiles * n_tiles;

    const uint32_t a_bytes = (uint32_t)tile_bytes(TR_A), b_bytes = (uint32_t)tile_bytes(p.trb);
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bars = base + p.stages * stage_bytes;  // full[stages], empty[stages], acc_full[2], acc_empty[2], slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (p.stages + s); };
    auto acc_full = [&](int b) { return bars + 16u * p.stages + 8u * b; };
    auto acc_empty = [&](int b) { return bars + 16u * p.stages + 16u + 8u * b; };
    const uint32_t tmem_slot = bars + 16u * p.stages + 32u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), 4);  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 2 * p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform: loader / MMA warps stay converged
    if (warp_u == 4) {
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            const int tm = tile % m_tiles, tn = tile / m_tiles;
            const uint8_t *a_src = p.A + (size_t)tm * nkb * a_bytes;
            const uint8_t *b_src = p.B + (size_t)tn * nkb * b_bytes;
            for (int i = 0; i < nkb; ++i, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                bulk_g2s2_expect_elect(base + s * stage_bytes, a_src + (size_t)i * a_bytes, a_bytes,
                                       base + s * stage_bytes + a_bytes, b_src + (size_t)i * b_bytes, b_bytes, full_bar(s));
            }
        }
    } else if (warp_u == 5) {
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc = instr_desc_m128(p.trb, p.fmt);
        const uint32_t a_lbo = TR_A * 16, b_lbo = (uint32_t)p.trb * 16;
        int it = 0, t = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++t) {
            const int buf = t & 1;
            mbar_wait(acc_empty(buf), ((uint32_t)(t >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + (uint32_t)buf * p.tmem_cols;
            for (int i = 0; i < nkb; ++i, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a0 = base + s * stage_bytes, b0 = a0 + a_bytes;
#pragma unroll
                for (int ks = 0; ks < KB / 16; ++ks) {
                    const uint64_t a_hi = smem_desc(a0 + ks * 2 * a_lbo, a_lbo, 128);
                    const uint64_t a_lo = smem_desc(a0 + (uint32_t)part_bytes(TR_A) + ks * 2 * a_lbo, a_lbo, 128);
                    const uint64_t b_hi = smem_desc(b0 + ks * 2 * b_lbo, b_lbo, 128);
                    const uint64_t b_lo = smem_desc(b0 + (uint32_t)part_bytes(p.trb) + ks * 2 * b_lbo, b_lbo, 128);
                    mma_bf16_ss_elect(d_tmem, a_hi, b_hi, idesc, (i | ks) != 0);
                    mma_bf16_ss_elect(d_tmem, a_hi, b_lo, idesc, 1);
                    mma_bf16_ss_elect(d_tmem, a_lo, b_hi, idesc, 1);
                }
                mma_commit_elect(empty_bar(s));
            }
            mma_commit_elect(acc_full(buf));
        }
    } else {
        const bool col_split = p.cols_per_z < (long long)p.n_valid;
        const uint32_t cpz = col_split ? (uint32_t)p.cols_per_z : 1u;
        auto col_offset = [&](uint32_t col, long long stride_col) -> long long {
            if (!col_split) return (long long)col * stride_col;
            const uint32_t cz = col / cpz;
            return (long long)cz * p.stride_cz + (long long)(col - cz * cpz) * stride_col;
        };
        const bool row_split = p.rows_per_z < (long long)p.m_valid;
        auto row_offset = [&](long long r) -> long long {
            if (!row_split) return r * p.stride_row;
            const uint32_t rz = (uint32_t)r / (uint32_t)p.rows_per_z;
            return (long long)rz * p.stride_z + (long long)((uint32_t)r - rz * (uint32_t)p.rows_per_z) * p.stride_row;
        };
        const int gw = p.trb < 128 ? p.trb : 128;  // columns per staged group
        float *stg = reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)) + stg_off) + (size_t)warp * 32 * (gw + 4);
        // ep_kind 2: whole-tile staging S[128][IDX_STR] + the inverse anchor permutations inv[k][a']
        float *S = reinterpret_cast<float *>(smem_raw + (base - smem_u32(smem_raw)) + stg_off);
        int32_t *s_inv = reinterpret_cast<int32_t *>(S + 128 * IDX_STR);
        const bool helper = warp >= 6;                                   // reduction helpers: never touch TMEM
        const int ridx = helper ? (int)threadIdx.x - 64 : (int)threadIdx.x;   // 0..255 over the 8 reducing warps
        if (p.ep_kind >= 2) {
            for (int i = ridx; i < 60 * 12; i += 256) {
                const int a = i / 12, k = i - a * 12;
                if (p.ep_kind == 2) s_inv[k * 60 + p.intra_idx[i]] = a;  // source anchor of (k, a') under the inverse
                else s_inv[k * 60 + a] = p.intra_idx[i];                 // forward: source anchor of (k, a)
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // one tile column (point, anchor) per reducing thread: its 12 permuted source columns sit in registers, the 10
        // channels of the tile stream through (12 LDS + 12 FADD per output)
        auto reduce_tile = [&](int tm, int tn) {
            if (ridx < 240) {
                const int col = ridx, pt = col / 60, a2 = col - pt * 60;
                const long long gp = (long long)tn * 4 + pt;  // point index over (z, pt) of this launch
                if (gp * 60 < p.n_valid) {
                    int src[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) src[k] = k * IDX_STR + pt * 60 + s_inv[k * 60 + a2];
                    const long long zz = gp / p.pts_per_z, pp = gp - zz * p.pts_per_z;
                    float *drow = p.dfeats + ((zz * p.ch_total + (long long)tm * 10) * p.pts_per_z + pp) * 60 + a2;
                    const int nch = min(10, p.ch_total - tm * 10);
                    for (int cl = 0; cl < nch; ++cl) {
                        const float *sblk = S + (size_t)(cl * 12) * IDX_STR;
                        float acc_v = 0.f;
#pragma unroll
                        for (int k = 0; k < 12; ++k) acc_v += sblk[src[k]];
                        drow[(size_t)cl * p.pts_per_z * 60] = acc_v;
                    }
                }
            }
        };
        if (helper) {
            if (p.ep_kind >= 2) {
                for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                    asm volatile("bar.sync 1, 256;" ::: "memory");   // the epilogue warps have staged the tile
                    reduce_tile(tile % m_tiles, tile / m_tiles);
                    asm volatile("bar.sync 1, 256;" ::: "memory");   // S is free for the next tile
                }
            }
        } else {
        int t = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++t) {
            const int tm = tile % m_tiles, tn = tile / m_tiles;
            const int buf = t & 1;
            mbar_wait(acc_full(buf), (uint32_t)(t >> 1) & 1u);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)buf * p.tmem_cols + ((uint32_t)(warp * 32) << 16);
            const long long row0 = (long long)tm * TR_A + warp * 32;
            if (p.ep_kind >= 2) {
                // dfeats[z, ch, pt, a'] = sum_k dG[(ch,k), (pt, inv_k(a'))]: tile rows = 10 channels x 12 k (+8 dead),
                // tile columns = 4 points x 60 anchors, so the whole reduction is tile-local
                const int tid = threadIdx.x;  // = TMEM lane = tile row
                for (int c0 = 0; c0 < 240; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(acc + (uint32_t)c0, v);
                    if (p.out_scale != 1.0f) {
                    #pragma unroll
                        for (int jq = 0; jq < 32; ++jq) v[jq] *= p.out_scale;
                    }
                    if (c0 + 32 >= 240) {  // accumulator fully read: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(buf));
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (c0 + j < 240)
                            *reinterpret_cast<float4 *>(S + (size_t)tid * IDX_STR + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                reduce_tile(tm, tn);
                asm volatile("bar.sync 1, 256;" ::: "memory");  // S is free for the next tile
            } else if (stg_off != 0u) {
                // column-contiguous output (dX): stage 32 rows x gw columns, one asynchronous bulk store per row
                for (int g0 = 0; g0 < p.trb; g0 += gw) {
                    for (int cc = 0; cc < gw; cc += 32) {
                        float v[32];
                        tmem_ld_32x32(acc + (uint32_t)(g0 + cc), v);
                        if (p.out_scale != 1.0f) {
                        #pragma unroll
                            for (int jq = 0; jq < 32; ++jq) v[jq] *= p.out_scale;
                        }
                        if (cc == 0) {  // the previous group's rows have left the staging buffer
                            bulk_wait_read0();
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (cc + j < gw)
                                *reinterpret_cast<float4 *>(stg + (size_t)lane * (gw + 4) + cc + j) =
                                    make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                    if (g0 + gw >= p.trb) {  // accumulator fully read: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(buf));
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    const long long colg = (long long)tn * p.trb + g0;
                    const long long r = row0 + lane;
                    if (r < p.m_valid && colg < p.n_valid) {
                        const int ncols = (int)min((long long)min(gw, p.trb - g0), (long long)p.n_valid - colg);
                        float *dst = p.out + row_offset(r) + col_offset((uint32_t)colg, 1);
                        const float *src = stg + (size_t)lane * (gw + 4);
                        const bool whole = !col_split || ((uint32_t)colg / cpz == (uint32_t)(colg + ncols - 1) / cpz);
                        if (whole && (ncols & 3) == 0) {
                            bulk_s2g(dst, smem_u32(src), (uint32_t)ncols * 4u);
                        } else {
                            for (int e = 0; e < ncols; ++e) p.out[row_offset(r) + col_offset((uint32_t)(colg + e), 1)] = src[e];
                        }
                    }
                    bulk_commit();
                }
            } else {
                const long long row = row0 + lane;
                const bool row_ok = row < p.m_valid;
                float *dst_row = p.out;
                if (row_ok) dst_row += (row / p.rows_per_z) * p.stride_z + (row % p.rows_per_z) * p.stride_row;
                for (int c0 = 0; c0 < p.trb; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(acc + (uint32_t)c0, v);
                    if (p.out_scale != 1.0f) {
                    #pragma unroll
                        for (int jq = 0; jq < 32; ++jq) v[jq] *= p.out_scale;
                    }
                    if (c0 + 32 >= p.trb) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty(buf));
                    }
                    const long long colb = (long long)tn * p.trb + c0;
                    if (p.vec) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const long long col = colb + j;
                            if (row_ok && c0 + j < p.trb && col < p.n_valid) {
                                float *dst = dst_row + col_offset((uint32_t)col, 1);
                                if (col + 3 < p.n_valid) {
                                    *reinterpret_cast<float4 *>(dst) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                                } else {
                                    for (int e = 0; e < 4 && col + e < p.n_valid; ++e) dst[e] = v[j + e];
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const long long col = colb + j;
                            if (row_ok && c0 + j < p.trb && col < p.n_valid) {
                                float *dst = dst_row + col_offset((uint32_t)col, p.stride_col);
                                if (p.mode) atomicAdd(dst, v[j]);
                                else *dst = v[j];
                            }
                        }
                    }
                }
            }
        }
        if (stg_off != 0u && p.ep_kind < 2) bulk_wait_read0();
        }   // !helper
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * p.tmem_cols);
    }
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int umma_trb_for(int n_rows) {  // rows per B tile = UMMA N: multiple of 16, at most 256
    int t = (n_rows + 15) / 16 * 16;
    return t > 256 ? 256 : t;
}

int launch_umma_gemm(const void *A_tiles, const void *B_tiles, int m_rows, int n_rows, long long K, int trb,
                     const GemmEpilogue &ep, int split_k, cudaStream_t s, int fmt) {
    static DynSmemOnce once;
    if (int rc = ensure_dyn_smem(once, umma_gemm_kernel, 220 * 1024, "umma_gemm_kernel")) return rc;
    UmmaGemmParams p;
    p.A = static_cast<const uint8_t *>(A_tiles);
    p.B = static_cast<const uint8_t *>(B_tiles);
    p.k_blocks = (int)((K + KB - 1) / KB);
    p.trb = trb;
    p.fmt = fmt;
    p.out_scale = fmt == FMT_F16 ? 1.0f / F16_W_SCALE : 1.0f;
    const size_t stage = tile_bytes(TR_A) + tile_bytes(trb);
    int stages = (int)((200 * 1024) / stage);
    if (stages > 4) stages = 4;
    if (trb <= 128 && stages > 3) stages = 3;  // 2 CTAs per SM
    if (trb > 128) stages = 2;                 // 96 KB per CTA: 2 CTAs per SM (2 x 256 TMEM columns), the
                                               // epilogue of one overlaps the main loop of the other
    if (stages < 2) stages = 2;
    p.stages = stages;
    uint32_t cols = 32;
    while ((int)cols < trb) cols *= 2;
    p.tmem_cols = cols;
    p.out = ep.out;
    p.rows_per_z = ep.rows_per_z;
    p.stride_z = ep.stride_z;
    p.stride_row = ep.stride_row;
    p.stride_col = ep.stride_col;
    p.cols_per_z = ep.cols_per_z;
    p.stride_cz = ep.stride_cz;
    if (ep.stride_col == 1 && ep.stride_cz == ep.cols_per_z) {  // (z, j) columns are contiguous after all
        p.cols_per_z = 1LL << 60;
        p.stride_cz = 0;
    }
    p.vec = (!ep.atomic && ep.stride_col == 1 && (ep.cols_per_z % 4) == 0 && (ep.stride_cz % 4) == 0 &&
             (ep.stride_row % 4) == 0 && (ep.stride_z % 4) == 0 && ((uintptr_t)ep.out & 15) == 0) ? 1 : 0;
    p.m_valid = m_rows;
    p.n_valid = n_rows;
    p.mode = ep.atomic ? 1 : 0;
    if (split_k < 1) split_k = 1;
    if (split_k > p.k_blocks) split_k = p.k_blocks;
    p.split_k = split_k;
    dim3 grid((m_rows + TR_A - 1) / TR_A, (n_rows + trb - 1) / trb, split_k);
    if (grid.y > 65535 || grid.z > 65535) {
        set_error("umma_gemm: grid too large");
        return EPN_ERR_SHAPE;
    }
    ProfScope prof(s, KC_GEMM);
    static const int persistent_on = (getenv("EPN_GEMM_PERSISTENT") && atoi(getenv("EPN_GEMM_PERSISTENT")) == 0) ? 0 : 1;
    const long long total_tiles = (long long)grid.x * grid.y;
    // column-contiguous outputs (dX: short K loop, 128 KB of output per tile) run persistently: measured
    // 264 -> 167 us per launch; the K-long forward GEMM streams better as two independent CTAs per SM
    if (persistent_on && p.vec && split_k == 1 && total_tiles >= 2LL * sm_count() && total_tiles < (1LL << 31)) {
        static DynSmemOnce once2;
        if (int rc = ensure_dyn_smem(once2, umma_gemm_persistent_kernel, 226 * 1024, "umma_gemm_persistent_kernel")) return rc;
        const int gw = trb < 128 ? trb : 128;
        const size_t stg_bytes = p.vec ? (size_t)4 * 32 * (gw + 4) * sizeof(float) : 0;  // staged bulk-store epilogue (dX)
        const size_t avail = (size_t)226 * 1024 - 512 - stg_bytes;
        int pst = (int)(avail / stage);
        if (pst > 8) pst = 8;
        if (pst >= 2) {
            p.stages = pst;
            const size_t pipe = (((size_t)pst * stage + 16 * pst + 64) + 127) & ~(size_t)127;
            const size_t smem_p = 128 + pipe + stg_bytes;
            const int ctas = (int)(total_tiles < sm_count() ? total_tiles : sm_count());
            umma_gemm_persistent_kernel<<<ctas, PERSISTENT_THREADS, smem_p, s>>>(p, (int)grid.x, (int)grid.y, stg_bytes ? (uint32_t)pipe : 0u);
            return check_launch("umma_gemm_persistent_kernel");
        }
    }
    const size_t smem = (size_t)stages * stage + 128 + 16 * stages + 32;
    umma_gemm_kernel<<<grid, 192, smem, s>>>(p);
    return check_launch("umma_gemm_kernel");
}

// ------------------------------------------------------------------ intra conv data gradient, fused
// Weight operand with padded rows: tile t holds rows (cl*12 + k), cl < 10 channels (ch = t*10 + cl), + 8 zero rows;
// element (row (ch,k), K index j) = W[ch*s_ch + j*s_j + k].  Data gradient: ch = input channel, j = output channel
// (s_ch = 12, s_j = c_in*12); forward: ch = output channel, j = input channel (s_ch = c_in*12, s_j = 12).
__global__ void __launch_bounds__(256)
intra_wt_tiles_kernel(const float *__restrict__ W, uint8_t *__restrict__ dst, int n_ch, int n_j, long long s_ch,
                      long long s_j, int k_blocks, int fmt, float scale) {
    const int rt = blockIdx.x, kcg = blockIdx.y * 2 + (threadIdx.x >> 7), r = threadIdx.x & 127;
    if (kcg >= k_blocks * (KB / 8)) return;
    const int cl = r / 12, k = r - cl * 12, ch = rt * 10 + cl;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int j = kcg * 8 + i;
        x[i] = (r < 120 && ch < n_ch && j < n_j) ? __ldg(W + (size_t)ch * s_ch + (size_t)j * s_j + k) : 0.f;
    }
    uint4 hi, lo;
    split8_fmt(x, hi, lo, fmt, scale);
    uint8_t *tile = dst + ((size_t)rt * k_blocks + (kcg >> 2)) * tile_bytes(TR_A);
    *reinterpret_cast<uint4 *>(tile + (size_t)(kcg & 3) * TR_A * 16 + (size_t)r * 16) = hi;
    *reinterpret_cast<uint4 *>(tile + part_bytes(TR_A) + (size_t)(kcg & 3) * TR_A * 16 + (size_t)r * 16) = lo;
}

size_t intra_dx_wt_bytes(int c_rows, int c_k) { return (size_t)cdiv(c_rows, 10) * cdiv(c_k, KB) * tile_bytes(TR_A); }
size_t intra_dx_dout_bytes(long long n, int c_k) { return split_tiles_bytes(n, c_k, 240); }

bool intra_dx_fused_ok(long long n_cols, int p, int na, int kn) { return na == 60 && kn == 12 && p % 4 == 0 && n_cols % 240 == 0; }

// Anchor-permuted channel GEMM of the intra convolution, both directions, for bc clouds:
//   forward == 0 (data gradient):  res[z, ch, pt, a'] = sum_{o,k} W[o, ch*12+k] * x[z, o, pt, inv_k(a')]
//       (autograd of IntraSO3Conv w.r.t. its input; x = dout, c_rows = c_in, c_k = c_out)
//   forward == 1:                  res[z, o, pt, a]   = sum_{c,k} W[o, c*12+k]  * x[z, c, pt, intra_idx[a,k]]
//       (IntraSO3Conv.forward, vgtk/vgtk/so3conv/modules.py:197-200 + so3conv/functional.py:221-268;
//        x = feats, c_rows = c_out, c_k = c_in)
// wt_tiles / x_tiles: scratch of intra_dx_wt_bytes(c_rows, c_k) / intra_dx_dout_bytes(n, c_k).  The 12x larger
// grouped tensor (G resp. dG) only ever exists as one TMEM accumulator tile per SM.
int launch_umma_intra_dx(const float *dout, long long dout_stride_z, long long dout_stride_o, const float *W,
                         const int32_t *intra_idx, float *dfeats, void *wt_tiles, void *dout_tiles, int bc, int c_in,
                         int c_out, int p, int forward, cudaStream_t s, int fmt, const NormPrologue *pro) {
    const long long n = (long long)bc * p * 60;
    if (!intra_dx_fused_ok(n, p, 60, 12) || n >= (1LL << 31)) return 1;
    const int c_rows = forward ? c_out : c_in, c_k = forward ? c_in : c_out;
    const int k_blocks = cdiv(c_k, KB), m_tiles = cdiv(c_rows, 10);
    {
        ProfScope prof(s, KC_SPLIT);
        dim3 grid(m_tiles, cdiv(k_blocks * (KB / 8), 2));
        intra_wt_tiles_kernel<<<grid, 256, 0, s>>>(W, static_cast<uint8_t *>(wt_tiles), c_rows, c_k,
                                                   forward ? (long long)c_in * 12 : 12LL, forward ? 12LL : (long long)c_in * 12,
                                                   k_blocks, fmt, fmt == FMT_F16 ? F16_W_SCALE : 1.0f);
        int rc = check_launch("intra_wt_tiles_kernel");
        if (rc) return rc;
    }
    SplitSrc src{dout, (long long)p * 60, dout_stride_z, 1, 1LL << 60, 0, dout_stride_o};
    if (pro != nullptr) src.pro = *pro;
    int rc = launch_split_tiles(src, dout_tiles, n, c_k, 240, s, fmt);
    if (rc) return rc;
    static DynSmemOnce once3;
    if (int rc3 = ensure_dyn_smem(once3, umma_gemm_persistent_kernel, 226 * 1024, "umma_gemm_persistent_kernel")) return rc3;
    UmmaGemmParams q;
    q.A = static_cast<const uint8_t *>(wt_tiles);
    q.B = static_cast<const uint8_t *>(dout_tiles);
    q.k_blocks = k_blocks;
    q.trb = 240;
    q.tmem_cols = 256;
    q.out = nullptr;
    q.rows_per_z = 1LL << 60; q.stride_z = 0; q.stride_row = 0; q.stride_col = 1; q.cols_per_z = 1LL << 60; q.stride_cz = 0;
    q.m_valid = m_tiles * TR_A;
    q.n_valid = (int)n;
    q.mode = 0; q.split_k = 1; q.vec = 0;
    q.ep_kind = forward ? 3 : 2;  // 3: the permutations themselves, 2: their inverses
    q.fmt = fmt;
    q.out_scale = fmt == FMT_F16 ? 1.0f / F16_W_SCALE : 1.0f;
    q.intra_idx = intra_idx;
    q.dfeats = dfeats;
    q.ch_total = c_rows;
    q.pts_per_z = p;
    const size_t stage = tile_bytes(TR_A) + tile_bytes(240);
    const size_t stg_bytes = (size_t)128 * IDX_STR * sizeof(float) + 60 * 12 * sizeof(int32_t);
    int pst = (int)(((size_t)226 * 1024 - 512 - stg_bytes) / stage);
    if (pst > 8) pst = 8;
    if (pst < 2) return 1;
    q.stages = pst;
    const size_t pipe = (((size_t)pst * stage + 16 * pst + 64) + 127) & ~(size_t)127;
    const long long n_tiles = n / 240, total = (long long)m_tiles * n_tiles;
    const int ctas = (int)(total < sm_count() ? total : sm_count());
    ProfScope prof(s, KC_GEMM);
    umma_gemm_persistent_kernel<<<ctas, PERSISTENT_THREADS, 128 + pipe + stg_bytes, s>>>(q, m_tiles, (int)n_tiles, (uint32_t)pipe);
    return check_launch("umma_gemm_persistent_kernel(intra dX)");
}

}  // namespace epn
