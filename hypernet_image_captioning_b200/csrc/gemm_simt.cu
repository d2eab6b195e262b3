// General fp32 GEMM on the CUDA cores (exact-fp32 path): C[M,N] = op(A) op(B)^T (+ bias[n]) (ReLU).
//
// Used for every dense contraction of the decoder whose tensor-core (tcgen05, bf16x3-split) kernel is not (yet) wired
// in, and for the small / oddly shaped ones (dW_gen with split-K, init_h, ...).  Replaces the addmm calls torch makes
// for nn.Linear / GRUCell in reference models/decoderlstm.py:61,100,105, later.py:411,418,442 and their backward.
//
//   A(m,k) = a_kmajor ? A[m*lda + k] : A[k*lda + m]
//   B(n,k) = b_kmajor ? B[n*ldb + k] : B[k*ldb + n]
// 128x128x16 tiles, 256 threads, 8x8 register tile per thread, register-prefetch double buffering.
// grid.z > 1 => split-K with fp32 atomics into a pre-zeroed (or to-be-accumulated-into) C.
#include "common.cuh"

namespace caphn {

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

template <bool AK, bool BK>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, long lda,
                                                       const float* __restrict__ B, long ldb, float* __restrict__ C,
                                                       long ldc, const float* __restrict__ bias, int M, int N, int K,
                                                       int kchunk, int relu, int atomic) {
    __shared__ __align__(16) float As[2][GBK][GBM + GPAD];
    __shared__ __align__(16) float Bs[2][GBK][GBN + GPAD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto gload = [&](int kt) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (AK) { m = idx >> 4; k = idx & 15; } else { k = idx >> 7; m = idx & 127; }
            const int gm = m0 + m, gk = kt + k;
            float v = 0.f;
            if (gm < M && gk < kend) v = AK ? A[(long)gm * lda + gk] : A[(long)gk * lda + gm];
            ra[i] = v;
            int n, k2;
            if (BK) { n = idx >> 4; k2 = idx & 15; } else { k2 = idx >> 7; n = idx & 127; }
            const int gn = n0 + n, gk2 = kt + k2;
            float w = 0.f;
            if (gn < N && gk2 < kend) w = BK ? B[(long)gn * ldb + gk2] : B[(long)gk2 * ldb + gn];
            rb[i] = w;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (AK) { m = idx >> 4; k = idx & 15; } else { k = idx >> 7; m = idx & 127; }
            As[buf][k][m] = ra[i];
            int n, k2;
            if (BK) { n = idx >> 4; k2 = idx & 15; } else { k2 = idx >> 7; n = idx & 127; }
            Bs[buf][k2][n] = rb[i];
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        gload(kbeg);
        sstore(0);
    }
    __syncthreads();
    for (int kt = kbeg; kt < kend; kt += GBK) {
        const bool more = kt + GBK < kend;
        if (more) gload(kt + GBK);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    const bool add_bias = bias != nullptr && blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= N) continue;
            float v = acc[i][j];
            if (add_bias) v += bias[gn];
            if (relu) v = fmaxf(v, 0.f);
            float* c = C + (long)gm * ldc + gn;
            if (atomic) atomicAdd(c, v); else *c = v;
        }
    }
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// C[M,N] (ldc) = A' B'^T (+bias[n]) (relu).  splitk > 1 or accumulate != 0 => atomicAdd into C (caller zeroes C when
// it wants a plain product).  relu is not allowed together with atomics.
int caphn_gemm_f32(const float* A, long lda, int a_kmajor, const float* B, long ldb, int b_kmajor, float* C, long ldc,
                   const float* bias, int M, int N, int K, int relu, int splitk, int accumulate, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (M <= 0 || N <= 0 || K <= 0) return CAPHN_EINVAL;
    if (splitk < 1) splitk = 1;
    int kchunk = ((K + splitk - 1) / splitk + GBK - 1) / GBK * GBK;
    splitk = (K + kchunk - 1) / kchunk;
    const int atomic = (splitk > 1 || accumulate) ? 1 : 0;
    if (atomic && relu) return CAPHN_EINVAL;
    dim3 grid(ceil_div(N, GBN), ceil_div(M, GBM), splitk);
    if (a_kmajor && b_kmajor)
        gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, M, N, K, kchunk, relu, atomic);
    else if (a_kmajor && !b_kmajor)
        gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, M, N, K, kchunk, relu, atomic);
    else if (!a_kmajor && b_kmajor)
        gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, M, N, K, kchunk, relu, atomic);
    else
        gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, bias, M, N, K, kchunk, relu, atomic);
    CAPHN_RETURN_LAST();
}

}  // extern "C"
