// sm_100a primitives for the tensor-core path: mbarrier, bulk async copy (UBLKCP), TMEM
// allocation, tcgen05.mma / commit / ld, and the shared-memory matrix descriptor of the
// canonical K-major no-swizzle layout used by every operand tile of this library.
//
// "Split tile" operand format (global and shared memory are identical, so one
// cp.async.bulk moves a whole stage):
//   an fp32 matrix X[rows, K] is stored as bf16 hi/lo pairs, X ~= hi + lo, |X - hi - lo| <= 2^-17 |X|,
//   cut into tiles of TR rows x KB=32 k;  tile (rt, kb) is contiguous:
//       [part: hi, lo][kc: KB/8][r: TR][8 x bf16]           (2 * TR * KB * 2 bytes)
//   i.e. 8x(8 bf16) core matrices of 128 contiguous bytes, 8-row groups 128 B apart (SBO),
//   k-chunks TR*16 B apart (LBO).  fp32-faithful product: a*b ~= ah*bh + ah*bl + al*bh
//   (three kind::f16 MMAs per 16-wide k step, fp32 accumulation in TMEM, error ~1e-5 relative).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace epn {
namespace umma {

constexpr int KB = 32;        // k extent of one tile / pipeline stage
constexpr int TR_A = 128;     // rows of an A tile = UMMA M

__host__ __device__ constexpr size_t tile_bytes(int tr) { return (size_t)2 * tr * KB * 2; }
__host__ __device__ constexpr size_t part_bytes(int tr) { return (size_t)tr * KB * 2; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// Bounded wait: a lost arrival traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("epn: mbarrier wait timed out (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// Same bounded wait without the diagnostic printf (no call, no stack frame): for waits inside hot loops.
__device__ __forceinline__ void mbar_wait_q(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / bulk copy)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- bulk copy (UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}

// shared -> global bulk store (bulk_group completion): the issuing thread's earlier generic-proxy writes to
// the source must be made visible with fence_proxy_async_smem() by their writers first.
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- TMA (tensor-map) tile load
// 4-D tiled load global -> shared, completion on an mbarrier (SASS: UTMALDG).  `tmap` is the address of a CUtensorMap
// kernel parameter (__grid_constant__).
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const void *tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// K-major, no swizzle: start address, LBO (k-chunk stride), SBO (8-row group stride), version 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16, A = B = BF16 (K-major), D = F32, M = 128, N = n.
__host__ __device__ constexpr uint32_t instr_desc_bf16_m128(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// Operand format of the split tiles.  FMT_BF16 (default everywhere): hi/lo are bf16 -- fp32's exponent range, 16
// significand bits (2^-18 relative per operand).  FMT_F16 (inference forwards that ask for it, see
// epn_set_forward_operands): hi/lo are fp16 -- 22 significand bits where |x| >= 2^-3 (below that the lo part is
// subnormal: absolute error <= 2^-25), at the same MMA rate and the same bytes, but only fp16's range (|x| > 65504
// becomes inf and the result NaN: loud, not silently wrong).  Weights are multiplied by the exact power of two
// F16_W_SCALE before the split (|W| >= 1.2e-4 then has a normal lo part; |W| < 64 is required) and the epilogue
// multiplies the accumulator by 1 / F16_W_SCALE.
enum { FMT_BF16 = 0, FMT_F16 = 1 };
constexpr float F16_W_SCALE = 1024.0f;
// same descriptor with A = B = F16 (format code 0) when fmt == FMT_F16
__host__ __device__ constexpr uint32_t instr_desc_m128(int n, int fmt) {
    return (1u << 4) | (fmt == FMT_F16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier when all MMAs issued so far by this thread have completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- warp-converged issue (elect.sync)
// The variants below are executed by a WHOLE converged warp; one elected lane issues the instruction.  Keeping the
// issuing warp converged lets the compiler hold descriptors / addresses in uniform registers: under a divergent
// `if (lane == 0)` it wraps every tcgen05 / bulk-copy instruction (whose operands are uniform registers) in an
// ELECT + R2UR + BRA.U.ANY loop over the active lanes, ~25 dependent instructions per MMA (measured ~1100 cycles per
// 16-k step of three MMAs in the fused inter conv, which made the issuing thread the kernel's bottleneck).
__device__ __forceinline__ void mma_bf16_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
// arrive.expect_tx + one bulk copy global -> shared, by one elected lane
__device__ __forceinline__ void bulk_g2s_expect_elect(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
        ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
// arrive.expect_tx(bytes_a + bytes_b) + two bulk copies, by one elected lane
__device__ __forceinline__ void bulk_g2s2_expect_elect(uint32_t dst_a, const void *src_a, uint32_t bytes_a, uint32_t dst_b,
                                                       const void *src_b, uint32_t bytes_b, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b32 t;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "add.u32 t, %2, %5;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%6], t;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%6];\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%3], [%4], %5, [%6];\n\t}"
        ::"r"(dst_a), "l"(src_a), "r"(bytes_a), "r"(dst_b), "l"(src_b), "r"(bytes_b), "r"(bar) : "memory");
}
// non-blocking phase test
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <-> TMEM lane base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// main accumulator + cross-term accumulator `cross_cols` columns further on (see epn_inter_fused.cu): v = main + cross
__device__ __forceinline__ void tmem_ld_sum2(uint32_t taddr, uint32_t cross_cols, float (&v)[32]) {
    float x[32];
    tmem_ld_32x32(taddr, v);
    tmem_ld_32x32(taddr + cross_cols, x);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += x[i];
}

// ---------------------------------------------------------------- bf16 hi/lo split
// x ~= hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi); returns 8 elements as two 16-byte chunks.
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // one packed cvt.rn.bf16x2.f32 per pair (element 2i in the low half), bf16 -> f32 is a shift / mask
        const __nv_bfloat162 hp = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t *>(&hp);
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        const __nv_bfloat162 lp = __floats2bfloat162_rn(x[2 * i] - h0, x[2 * i + 1] - h1);
        l[i] = *reinterpret_cast<const uint32_t *>(&lp);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp16 hi/lo split of two values (low half = v0): hi = f16_rn(v), lo = f16_rn(v - hi)
__device__ __forceinline__ void split2_f16(float v0, float v1, uint32_t &h, uint32_t &l) {
    const __half2 hp = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(hp);
    const __half2 lp = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    h = *reinterpret_cast<const uint32_t *>(&hp);
    l = *reinterpret_cast<const uint32_t *>(&lp);
}
__device__ __forceinline__ void split2_bf16(float v0, float v1, uint32_t &h, uint32_t &l) {
    const __nv_bfloat162 hp = __floats2bfloat162_rn(v0, v1);
    h = *reinterpret_cast<const uint32_t *>(&hp);
    const __nv_bfloat162 lp = __floats2bfloat162_rn(v0 - __uint_as_float(h << 16), v1 - __uint_as_float(h & 0xffff0000u));
    l = *reinterpret_cast<const uint32_t *>(&lp);
}
template <int FMT>
__device__ __forceinline__ void split2(float v0, float v1, uint32_t &h, uint32_t &l) {
    if (FMT == FMT_F16) split2_f16(v0, v1, h, l);
    else split2_bf16(v0, v1, h, l);
}
// 8 elements in the format chosen at run time (`scale`: exact power of two applied first -- F16_W_SCALE for weights)
__device__ __forceinline__ void split8_fmt(const float (&x)[8], uint4 &hi, uint4 &lo, int fmt, float scale = 1.0f) {
    if (fmt == FMT_F16) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split2_f16(x[2 * i] * scale, x[2 * i + 1] * scale, h[i], l[i]);
        hi = make_uint4(h[0], h[1], h[2], h[3]);
        lo = make_uint4(l[0], l[1], l[2], l[3]);
    } else {
        split8(x, hi, lo);
    }
}

}  // namespace umma
}  // namespace epn
