// Index ops of the hot path: ball query, furthest point sampling, point gather.
// Replaces vgtk.cuda.grouping.{ball_query,furthest_point_sampling} and
// vgtk.cuda.gathering.gather_points_{forward,backward} of the reference.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "epn_common.cuh"

namespace epn {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-kernel-class timing (off by default; bench.py turns it on for one pass)
struct ProfSlot { cudaEvent_t a, b; int kclass; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfSlot *> g_prof_slots;

ProfScope::ProfScope(cudaStream_t stream, int kclass) : s(stream), slot(nullptr) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfSlot *p = new ProfSlot;
    p->kclass = kclass;
    cudaEventCreate(&p->a);
    cudaEventCreate(&p->b);
    cudaEventRecord(p->a, s);
    slot = p;
}

ProfScope::~ProfScope() {
    if (slot == nullptr) return;
    ProfSlot *p = static_cast<ProfSlot *>(slot);
    cudaEventRecord(p->b, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_slots.push_back(p);
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ ball query
// One warp per query.  The warp sweeps the supports 32 at a time (coalesced
// 128-B loads per coordinate, L1/L2 resident: a cloud's xyz is 12-192 KB),
// compacts hits in ascending order with ballot + popc, stops as soon as
// `nsample` hits are found, applies the reference fill rule and writes the
// whole idx row with coalesced stores.
// Reference semantics: vgtk/vgtk/cuda/grouping_cuda_kernel.cu:83-112.
constexpr int BQ_WARPS = 4;

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                  int32_t *__restrict__ idx, int n, int m, float r2, int nsample) {
    extern __shared__ int32_t s_hits[];  // [BQ_WARPS][nsample]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int j = blockIdx.x * BQ_WARPS + warp;
    const int b = blockIdx.y;
    if (j >= m) return;
    const float *q = new_xyz + (size_t)b * 3 * m;
    const float *sx = xyz + (size_t)b * 3 * n;
    const float *sy = sx + n;
    const float *sz = sy + n;
    const float qx = __ldg(q + j), qy = __ldg(q + m + j), qz = __ldg(q + 2 * m + j);
    int32_t *hits = s_hits + warp * nsample;

    int cnt = 0;
    for (int base = 0; base < n && cnt < nsample; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < n) {
            const float d2 = sqdist3(qx - __ldg(sx + k), qy - __ldg(sy + k), qz - __ldg(sz + k));
            hit = d2 < r2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < nsample) hits[pos] = k;
        cnt += __popc(mask);
    }
    cnt = min(cnt, nsample);
    __syncwarp();
    int32_t *row = idx + ((size_t)b * m + j) * nsample;
    const bool repeat = cnt > 0 && cnt < nsample - 1;
    for (int t = lane; t < nsample; t += 32) {
        int32_t v = 0;
        if (t < cnt) v = hits[t];
        else if (repeat) v = hits[t % cnt];
        row[t] = v;
    }
}

// ------------------------------------------------------ furthest point sampling
// One CTA per cloud; every thread keeps its points (xyz + running min distance)
// in registers, so a round is: 3 broadcast loads of the last pick, PPT distance
// updates, two REDUX per warp, one 8-byte smem slot per warp, ONE __syncthreads,
// and a redundant 32-slot warp reduction.  No temp[] traffic, no smem tree.
//
// The reference's result depends on its thread layout (grouping_cuda_kernel.cu:
// 340-466): thread t owns points t, t+T, ... (T = 2^floor(log2 n) <= 1024), keeps
// the first strict maximum, and the halving smem tree keeps the lower slot on
// ties, which orders ties by the BIT-REVERSED thread id.  That order is encoded
// in the reduction key so the picks are bit-identical:
//   key = [ float_bits(best)+1 (0 = no candidate) | ~bitrev(t) (10 b) | besti (22 b) ]
__device__ __forceinline__ void fps_warp_max(unsigned &hi, unsigned &lo) {
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    hi = mhi;
    lo = mlo;
}

template <int PPT>
__global__ void __launch_bounds__(1024)
fps_kernel(const float *__restrict__ xyz, float *__restrict__ temp_ws, int32_t *__restrict__ idx,
           int n, int m, int T, int logT) {
    __shared__ unsigned long long s_key[2][32];
    const int b = blockIdx.x;
    const float *px = xyz + (size_t)b * 3 * n;
    const float *py = px + n;
    const float *pz = py + n;
    int32_t *out = idx + (size_t)b * m;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    const bool owner = t < T;

    // PPT > 0: points live in registers.  PPT == 0: generic path, points are
    // re-read from global (L2) and the running minimum lives in temp_ws.
    constexpr int R = PPT > 0 ? PPT : 1;
    float x[R], y[R], z[R], tmp[R];
    unsigned valid = 0;
    float *temp = PPT > 0 ? nullptr : temp_ws + (size_t)b * n;
    if (PPT > 0) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int k = t + i * T;
            x[i] = y[i] = z[i] = 0.f;
            tmp[i] = 1e10f;
            if (owner && k < n) {
                x[i] = __ldg(px + k); y[i] = __ldg(py + k); z[i] = __ldg(pz + k);
                const float mag = __fmaf_rn(z[i], z[i], __fmaf_rn(x[i], x[i], __fmul_rn(y[i], y[i])));
                if (!((double)mag <= 1e-3)) valid |= 1u << i;
            }
        }
    } else if (owner) {
        for (int k = t; k < n; k += T) temp[k] = 1e10f;
    }
    const unsigned prio = logT > 0 ? (__brev((unsigned)t) >> (32 - logT)) : 0u;
    const unsigned lo_tag = owner ? ((~prio) & 0x3ffu) << 22 : 0u;

    int old = 0;
    if (t == 0) out[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = __ldg(px + old), y1 = __ldg(py + old), z1 = __ldg(pz + old);
        float best = -1.f;
        int besti = 0;
        if (PPT > 0) {
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if (valid & (1u << i)) {
                    const float d = sqdist3(x[i] - x1, y[i] - y1, z[i] - z1);
                    const float d2 = fminf(d, tmp[i]);
                    tmp[i] = d2;
                    if (d2 > best) { best = d2; besti = t + i * T; }
                }
            }
        } else if (owner) {
            for (int k = t; k < n; k += T) {
                const float x2 = __ldg(px + k), y2 = __ldg(py + k), z2 = __ldg(pz + k);
                const float mag = __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
                if ((double)mag <= 1e-3) continue;
                const float d = sqdist3(x2 - x1, y2 - y1, z2 - z1);
                const float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                if (d2 > best) { best = d2; besti = k; }
            }
        }
        unsigned hi = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
        unsigned lo = owner ? (lo_tag | (unsigned)besti) : 0u;
        fps_warp_max(hi, lo);
        if (lane == 0) s_key[j & 1][warp] = ((unsigned long long)hi << 32) | lo;
        __syncthreads();
        const unsigned long long kk = lane < nwarps ? s_key[j & 1][lane] : 0ull;
        hi = (unsigned)(kk >> 32);
        lo = (unsigned)kk;
        fps_warp_max(hi, lo);
        old = (int)(lo & 0x3fffffu);
        if (t == 0) out[j] = old;
    }
}

// ----------------------------------------------------------------------- gather
// out[b,c,j] = points[b,c,idx[b,j]]; j is the fastest thread index so idx reads
// and out writes are coalesced (the reference maps channel-fastest).
__global__ void gather_fwd_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx,
                                  float *__restrict__ out, int c, int n, int m) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    if (j >= m) return;
    const int a = __ldg(idx + (size_t)b * m + j);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        out[((size_t)b * c + ci) * m + j] = __ldg(points + ((size_t)b * c + ci) * n + a);
}

__global__ void gather_bwd_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                                  float *__restrict__ grad_points, int c, int n, int m) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    if (j >= m) return;
    const int a = __ldg(idx + (size_t)b * m + j);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y)
        atomicAdd(grad_points + ((size_t)b * c + ci) * n + a, __ldg(grad_out + ((size_t)b * c + ci) * m + j));
}

}  // namespace epn

using namespace epn;

EPN_API int epn_version(void) { return EPN_B200_VERSION; }

EPN_API const char *epn_last_error(void) { return g_err; }

EPN_API unsigned long long epn_launch_count(void) { return g_launches.load(); }

EPN_API void epn_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }

EPN_API int epn_profile_read(double *ms_per_class, long long *scopes_per_class, int n_class) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < n_class; ++i) {
        if (ms_per_class) ms_per_class[i] = 0.0;
        if (scopes_per_class) scopes_per_class[i] = 0;
    }
    int rc = 0;
    for (ProfSlot *p : g_prof_slots) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(p->b);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, p->a, p->b);
        if (e != cudaSuccess) { set_error("epn_profile_read: %s", cudaGetErrorString(e)); rc = (int)e; }
        if (p->kclass >= 0 && p->kclass < n_class) {
            if (ms_per_class) ms_per_class[p->kclass] += ms;
            if (scopes_per_class) scopes_per_class[p->kclass] += 1;
        }
        cudaEventDestroy(p->a);
        cudaEventDestroy(p->b);
        delete p;
    }
    g_prof_slots.clear();
    return rc;
}

EPN_API int epn_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    return prop.major == 10 ? 1 : 0;
}

EPN_API int epn_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int b, int n,
                               int m, float radius, int nsample, void *stream) {
    EPN_REQUIRE_PTR(new_xyz); EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(idx);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(n); EPN_REQUIRE_POS(m); EPN_REQUIRE_POS(nsample);
    EPN_REQUIRE(b <= 65535, EPN_ERR_SHAPE, "batch > 65535");
    EPN_REQUIRE(nsample <= 2048, EPN_ERR_SHAPE, "nsample > 2048");
    const float r2 = radius * radius;  // fp32, as the reference (grouping_cuda_kernel.cu:83)
    dim3 grid(cdiv(m, BQ_WARPS), b);
    const size_t smem = (size_t)BQ_WARPS * nsample * sizeof(int32_t);
    ProfScope prof(as_stream(stream), KC_INDEX);
    ball_query_kernel<<<grid, BQ_WARPS * 32, smem, as_stream(stream)>>>(new_xyz, xyz, idx, n, m, r2, nsample);
    return check_launch("ball_query_kernel");
}

EPN_API size_t epn_fps_workspace_bytes(int b, int n) {
    if (n <= 16384) return 0;
    return (size_t)b * (size_t)n * sizeof(float);
}

EPN_API int epn_fps_f32(const float *xyz, void *temp, int32_t *idx, int b, int n, int m, void *stream) {
    EPN_REQUIRE_PTR(xyz); EPN_REQUIRE_PTR(idx);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(n); EPN_REQUIRE_POS(m);
    EPN_REQUIRE(n < (1 << 22), EPN_ERR_SHAPE, "n >= 2^22");
    int T = 1, logT = 0;
    while (T * 2 <= n && T < 1024) { T *= 2; ++logT; }  // grouping_cuda_kernel.cu:29-33
    const int threads = T < 32 ? 32 : T;
    const int ppt = cdiv(n, T);
    cudaStream_t s = as_stream(stream);
    ProfScope prof(s, KC_INDEX);
    float *tw = static_cast<float *>(temp);
#define EPN_FPS(P) fps_kernel<P><<<b, threads, 0, s>>>(xyz, tw, idx, n, m, T, logT)
    if (ppt <= 1) EPN_FPS(1);
    else if (ppt <= 2) EPN_FPS(2);
    else if (ppt <= 4) EPN_FPS(4);
    else if (ppt <= 8) EPN_FPS(8);
    else if (ppt <= 16) EPN_FPS(16);
    else {
        EPN_REQUIRE(temp != nullptr, EPN_ERR_WORKSPACE, "n > 16384 needs the temp workspace");
        EPN_FPS(0);
    }
#undef EPN_FPS
    return check_launch("fps_kernel");
}

EPN_API int epn_gather_fwd_f32(const float *points, const int32_t *idx, float *out, int b, int c, int n,
                               int m, void *stream) {
    EPN_REQUIRE_PTR(points); EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(out);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(n); EPN_REQUIRE_POS(m);
    EPN_REQUIRE(b <= 65535, EPN_ERR_SHAPE, "batch > 65535");
    dim3 grid(cdiv(m, 256), c < 64 ? c : 64, b);
    ProfScope prof(as_stream(stream), KC_INDEX);
    gather_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, idx, out, c, n, m);
    return check_launch("gather_fwd_kernel");
}

EPN_API int epn_gather_bwd_f32(const float *grad_out, const int32_t *idx, float *grad_points, int b,
                               int c, int n, int m, void *stream) {
    EPN_REQUIRE_PTR(grad_out); EPN_REQUIRE_PTR(idx); EPN_REQUIRE_PTR(grad_points);
    EPN_REQUIRE_POS(b); EPN_REQUIRE_POS(c); EPN_REQUIRE_POS(n); EPN_REQUIRE_POS(m);
    EPN_REQUIRE(b <= 65535, EPN_ERR_SHAPE, "batch > 65535");
    dim3 grid(cdiv(m, 256), c < 64 ? c : 64, b);
    gather_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, grad_points, c, n, m);
    return check_launch("gather_bwd_kernel");
}
