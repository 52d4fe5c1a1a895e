// De-duplication of one ball-query row, shared by every kernel of the inter conv (forward grouping, fused forward,
// backward scatter).
//
// The ball query repeat-fills short neighbour lists cyclically (grouping_cuda_kernel.cu:100-104), so a row of
// `nn_raw` slots usually holds each neighbour several times.  Duplicates have identical offsets and therefore
// identical kernel weights: each distinct neighbour is kept once with its multiplicity folded into the weight
// (sum_n w_n f_n == sum_u m_u w_u f_u) -- fewer rows to gather, fewer FMAs, same result.  Nothing is assumed about
// the row's structure (a caller may pass its own inter_idx): every slot is compared with every earlier slot.
#pragma once
#include <stdint.h>

namespace epn {

constexpr int DEDUP_MAX_RAW = 128;  // slots of a row the kernels built on this helper accept

// Shared-memory record of one point's distinct neighbours (CAP = capacity of the list).
template <int CAP>
struct NeighbourList {
    float g[CAP * 3];                // offsets x_q - x_p of the distinct neighbours, first-occurrence order
    int32_t idx[CAP];                // their indices
    float mult[CAP];                 // their multiplicities
    int32_t raw[DEDUP_MAX_RAW];      // the row as stored
    int32_t cnt[DEDUP_MAX_RAW / 32]; // distinct entries first seen in each 32-slot chunk
    int32_t total;                   // number of distinct neighbours of the row (may exceed CAP)
};

// Called by ALL `nthr` threads (tid = 0..nthr-1, nthr >= 32 * ceil(nn_raw / 32), whole warps) of the group that
// owns `L`; `sync()` is that group's barrier.  On return (after a final sync) every thread may read L: entries
// [0, min(total, CAP)) are valid, the rest of the list is zero (index 0, multiplicity 0, offset 0).
template <int CAP, class Sync>
__device__ __forceinline__ void dedup_row(NeighbourList<CAP> &L, const int32_t *__restrict__ row, int nn_raw,
                                          const float *__restrict__ X, const float *__restrict__ Cn, int p_in, int p,
                                          int pi, int tid, int nthr, Sync sync) {
    for (int n = tid; n < DEDUP_MAX_RAW; n += nthr) L.raw[n] = n < nn_raw ? row[n] : -1;
    sync();
    const int chunks = (nn_raw + 31) >> 5;
    const int w = tid >> 5, lane = tid & 31;
    bool uniq = false;
    int mult = 0, q = -1;
    unsigned mask = 0u;
    if (w < chunks) {  // warp w examines slots 32w .. 32w+31
        const int r = w * 32 + lane;
        q = r < nn_raw ? L.raw[r] : -1;
        uniq = r < nn_raw;
        for (int m = 0; m < r && uniq; ++m) uniq = L.raw[m] != q;
        if (uniq)
            for (int m = r; m < nn_raw; ++m) mult += (L.raw[m] == q) ? 1 : 0;
        mask = __ballot_sync(0xffffffffu, uniq);
        if (lane == 0) L.cnt[w] = __popc(mask);
    }
    sync();
    int total = 0;
    for (int i = 0; i < chunks; ++i) total += L.cnt[i];
    if (w < chunks && uniq) {
        int pos = __popc(mask & ((1u << lane) - 1u));
        for (int i = 0; i < w; ++i) pos += L.cnt[i];
        if (pos < CAP) {
            L.idx[pos] = q;
            L.mult[pos] = (float)mult;
            L.g[pos * 3] = X[q] - Cn[pi];
            L.g[pos * 3 + 1] = X[p_in + q] - Cn[p + pi];
            L.g[pos * 3 + 2] = X[2 * p_in + q] - Cn[2 * p + pi];
        }
    }
    for (int n = (total < CAP ? total : CAP) + tid; n < CAP; n += nthr) {
        L.idx[n] = 0;
        L.mult[n] = 0.f;
        L.g[n * 3] = 0.f; L.g[n * 3 + 1] = 0.f; L.g[n * 3 + 2] = 0.f;
    }
    if (tid == 0) L.total = total;
    sync();
}

}  // namespace epn
