// Building blocks shared by the persistent recurrence kernels (gru_seq.cu, attgru_seq.cu): a CTA of 512 threads owns
// BT = 4 batch rows; matrix-vector products against weights streamed from L2 are split over (k-range group) x (column quad)
// and reduced through shared memory.
#pragma once
#include "common.cuh"
#include <math.h>

namespace caphn {

constexpr int AT_THREADS = 512;
constexpr int AT_WARPS = AT_THREADS / 32;
constexpr int AT_BT = 4;
constexpr int AT_MAXSLOT = 8;  // (BT * ceil(H/32)) / 16 warps <= 8  =>  H <= 1024

__host__ __device__ inline int at_pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
__host__ __device__ inline int at_cqt(int ld) {
    int c = at_pow2_ceil(ld >> 2);
    if (c > AT_THREADS) c = AT_THREADS;
    if (c < 32) c = 32;
    return c;
}

// tanh through exp: 1 - 2/(e^{2x}+1) with ex2.approx / rcp.approx.  Absolute error ~1e-7 (<< the 1e-4 parity budget of
// the attention weights), ~6 instructions instead of ~25 for tanhf; saturates correctly for |x| large.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}

// part[(g*BT+b)*ldw + 4cq + c] (= | +=) sum_{k in chunk g} Wt[k*ldw + 4cq + c] * xs[k*BT + b]
template <int BT, bool ADD>
__device__ __forceinline__ void block_matvec(const float* __restrict__ Wt, int ldw, int Kdim, const float* xs,
                                             float* part, int CQT, int tid) {
    const int NCQ = ldw >> 2;
    const int KG = AT_THREADS / CQT;
    const int kg = tid / CQT, cq0 = tid - kg * CQT;
    const int kchunk = (Kdim + KG - 1) / KG;
    const int k0 = kg * kchunk, k1 = min(Kdim, k0 + kchunk);
    for (int cq = cq0; cq < NCQ; cq += CQT) {
        float acc[BT][4];
#pragma unroll
        for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[b][c] = 0.f;
        const float* wp = Wt + (long)k0 * ldw + 4 * cq;
#pragma unroll 4
        for (int k = k0; k < k1; ++k, wp += ldw) {
            const float4 w = *reinterpret_cast<const float4*>(wp);
            const float4 x4 = *reinterpret_cast<const float4*>(xs + k * BT);
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                acc[b][0] = fmaf(w.x, xv[b], acc[b][0]);
                acc[b][1] = fmaf(w.y, xv[b], acc[b][1]);
                acc[b][2] = fmaf(w.z, xv[b], acc[b][2]);
                acc[b][3] = fmaf(w.w, xv[b], acc[b][3]);
            }
        }
#pragma unroll
        for (int b = 0; b < BT; ++b) {
            float4* o = reinterpret_cast<float4*>(part + ((long)(kg * BT + b)) * ldw + 4 * cq);
            float4 v = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
            if (ADD) {
                const float4 old = *o;
                v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
            }
            *o = v;
        }
    }
}

__device__ __forceinline__ float part_sum(const float* part, int KG, int BT, int ld, int b, int j) {
    float s = 0.f;
    for (int g = 0; g < KG; ++g) s += part[((long)(g * BT + b)) * ld + j];
    return s;
}

}  // namespace caphn
