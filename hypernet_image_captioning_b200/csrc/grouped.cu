// Helpers of the many-style ("grouped") path: a batch whose rows use G different generated weight sets (BASELINE.json
// configs[3]: Conceptual-Captions domains; reference semantics = one HyperNet.forward per sample's style,
// train_cc.py:90-123).  The recurrence keeps the time-major layout [t][b] with the batch sorted by group (so a row tile of
// one group is contiguous at every step); the time-batched grouped GEMMs (x-projection, dX, dW_ih, dW_hh) want each
// group's rows of ALL steps contiguous ("group-major").  These kernels convert between the two while producing the bf16
// hi/lo tensor-core operands, so the permutation costs no extra pass:
//   split_gather   hi/lo[i, :] = split(src[rowmap[i], :])   (rowmap[i] < 0: a zero row -- the padding of a group's K range)
//   split_batched  the W_ih [3H, E+F] block of every row of Theta [G, theta] -> one operand array [G*3H, Kp]
//   group_colsum   bias gradients: out[g, n] = sum over the rows of group g (all steps) of X[row, n]
#include "common.cuh"
#include <cuda_bf16.h>

namespace caphn {

__device__ __forceinline__ void split_store(float x0, float x1, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    __nv_bfloat162 hv; hv.x = h0; hv.y = h1;
    *reinterpret_cast<__nv_bfloat162*>(hi) = hv;
    if (lo) {
        __nv_bfloat162 lv;
        lv.x = __float2bfloat16_rn(x0 - __bfloat162float(h0));
        lv.y = __float2bfloat16_rn(x1 - __bfloat162float(h1));
        *reinterpret_cast<__nv_bfloat162*>(lo) = lv;
    }
}

// one thread per 8 destination columns; columns [C, Kp) and rows with rowmap < 0 are zero
__global__ void split_gather_kernel(const float* __restrict__ src, long lds, const int* __restrict__ rowmap, long R, int C,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long Kp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long oct = Kp >> 3;
    if (i >= R * oct) return;
    const long r = i / oct;
    const int c = (int)(i - r * oct) * 8;
    const long sr = rowmap ? rowmap[r] : r;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
    if (sr >= 0) {
        const float* p = src + sr * lds + c;
        if (c + 8 <= C && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4 a = ldg_stream4(p), b = ldg_stream4(p + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (c + e < C) x[e] = p[e];
        }
    }
#pragma unroll
    for (int e = 0; e < 8; e += 2) split_store(x[e], x[e + 1], hi + r * Kp + c + e, lo ? lo + r * Kp + c + e : nullptr);
}

// blockIdx.y = batch b: src_b = src + b * sstride (floats), [R, C] with row pitch lds; dst rows b*R .. b*R+R-1
__global__ void split_batched_kernel(const float* __restrict__ src, long sstride, long lds, int R, int C,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long Kp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long half = Kp >> 1;
    if (i >= (long)R * half) return;
    const long r = i / half;
    const int c = (int)(i - r * half) * 2;
    const float* s = src + (long)blockIdx.y * sstride + r * lds;
    const float x0 = c < C ? s[c] : 0.f, x1 = c + 1 < C ? s[c + 1] : 0.f;
    const long dr = (long)blockIdx.y * R + r;
    split_store(x0, x1, hi + dr * Kp + c, lo ? lo + dr * Kp + c : nullptr);
}

// grid (G, ceil(N / 128), T): out[g*ldo + n] += sum_{b in group g} X[(t*B + b)*ldx + n]   (out zero-initialised)
__global__ void __launch_bounds__(128) group_colsum_kernel(const float* __restrict__ X, long ldx, const int* __restrict__ goff,
                                                            int B, int N, float* __restrict__ out, long ldo) {
    const int g = blockIdx.x, n = blockIdx.y * 128 + threadIdx.x, t = blockIdx.z;
    if (n >= N) return;
    const int b0 = goff[g], b1 = goff[g + 1];
    if (b0 >= b1) return;
    float s0 = 0.f, s1 = 0.f;
    const float* x = X + ((long)t * B) * ldx + n;
    int b = b0;
    for (; b + 1 < b1; b += 2) { s0 += x[(long)b * ldx]; s1 += x[(long)(b + 1) * ldx]; }
    if (b < b1) s0 += x[(long)b * ldx];
    atomicAdd(out + (long)g * ldo + n, s0 + s1);
}

// y = y > 0 ? y : slope * y (in place); backward: dy = y > 0 ? dy : slope * dy (y = the activation OUTPUT; slope > 0)
__global__ void leaky_relu_kernel(float* __restrict__ y, long n, float slope) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = y[i]; y[i] = v > 0.f ? v : slope * v; }
}
__global__ void leaky_relu_bwd_kernel(const float* __restrict__ y, float* __restrict__ dy, long n, float slope) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(y[i] > 0.f)) dy[i] *= slope;
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// hi/lo [R, Kp] bf16 (Kp % 64 == 0, Kp >= C): row i = split of src row rowmap[i] (fp32, row pitch lds), zero when
// rowmap[i] < 0; rowmap == NULL: identity (same as caphn_split_bf16).  lo may be NULL (plain bf16 mode).
int caphn_split_bf16_gather(const float* src, long lds, const int* rowmap, long R, int C, void* hi, void* lo, long Kp,
                            void* stream) {
    if (R <= 0 || C <= 0 || Kp < C || (Kp & 63)) return CAPHN_EINVAL;
    const long n = R * (Kp >> 3);
    split_gather_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, rowmap, R, C, (__nv_bfloat16*)hi,
                                                                           (__nv_bfloat16*)lo, Kp);
    CAPHN_RETURN_LAST();
}

// nb matrices [R, C] (row pitch lds), matrix b starting sstride floats after matrix b-1 -> one operand array hi/lo [nb*R, Kp].
int caphn_split_bf16_batched(const float* src, long sstride, long lds, int nb, int R, int C, void* hi, void* lo, long Kp,
                             void* stream) {
    if (nb <= 0 || R <= 0 || C <= 0 || Kp < C || (Kp & 63)) return CAPHN_EINVAL;
    const long n = (long)R * (Kp >> 1);
    split_batched_kernel<<<dim3(ceil_div(n, 256), nb), 256, 0, (cudaStream_t)stream>>>(src, sstride, lds, R, C,
                                                                                      (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Kp);
    CAPHN_RETURN_LAST();
}

// out[g*ldo + n] += sum over t < T and batch rows goff[g] <= b < goff[g+1] of X[(t*B + b)*ldx + n]; goff has G + 1 entries
// (the batch is sorted by group).  out must be zero-initialised by the caller (atomic accumulation over t).
int caphn_group_colsum(const float* X, long ldx, const int* goff, int G, int B, int T, int N, float* out, long ldo,
                       void* stream) {
    if (G <= 0 || B <= 0 || T <= 0 || N <= 0 || !goff) return CAPHN_EINVAL;
    group_colsum_kernel<<<dim3(G, ceil_div(N, 128), T), 128, 0, (cudaStream_t)stream>>>(X, ldx, goff, B, N, out, ldo);
    CAPHN_RETURN_LAST();
}

// nn.LeakyReLU(slope) of the hypernet layers (hypernet_attention.py:64,66,87,93) for the many-group path, where the
// layers themselves run as dense GEMMs: in place on the pre-activation / on the incoming gradient.
int caphn_leaky_relu(float* y, long n, float slope, void* stream) {
    if (n <= 0) return CAPHN_EINVAL;
    leaky_relu_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(y, n, slope);
    CAPHN_RETURN_LAST();
}
int caphn_leaky_relu_bwd(const float* y, float* dy, long n, float slope, void* stream) {
    if (n <= 0) return CAPHN_EINVAL;
    leaky_relu_bwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(y, dy, n, slope);
    CAPHN_RETURN_LAST();
}

}  // extern "C"
