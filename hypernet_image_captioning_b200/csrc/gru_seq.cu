// Persistent GRU recurrence (pooled-feature decoder, "Variant A"): all T time steps in one launch.
//
// Replaces the per-step nn.GRUCell calls of DecoderGRU.forward / infer (reference later.py:411,418,471-477; cell
// formula = torch GRUCell, gates r,z,n) and their autograd backward (BPTT).  The input projection
// gi = x W_ih^T + b_ih does not depend on the recurrence in teacher-forced mode and arrives precomputed for all steps
// (time-major [T,B,3H]); this kernel does gh = h W_hh^T + b_hh, the gate non-linearities and the state update.
//
// Round-1 decomposition: one CTA owns BT batch rows for the whole sequence and keeps h on-chip (shared memory); the
// generated W_hh (transposed + padded to 16-byte rows by caphn_transpose_pad) is re-read from L2 every step.
// (A cluster version that keeps W_hh slices resident in shared memory is the next step; see DESIGN.md.)
//
// Sequence tensors are time-major: row index = t*B + b.  Hall is [T+1, B, H] with Hall[0] = h0, so that
// Hall[1:] = h_t and Hall[:-1] = h_{t-1} are both plain [T*B, H] matrices for the dW_hh GEMM.
#include "common.cuh"

namespace caphn {

constexpr int SEQ_THREADS = 512;

__host__ __device__ inline int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <int BT>
__global__ void __launch_bounds__(SEQ_THREADS) gru_seq_fwd_kernel(
    const float* __restrict__ GI,    // [T,B,3H]
    const float* __restrict__ WhhT,  // [H, ld3]
    int ld3, const float* __restrict__ bhh, float* __restrict__ Hall,  // [T+1,B,H], Hall[0] = h0 on entry
    float* __restrict__ Hbm,   // [B,T,H] batch-major copy (may be null)
    float* __restrict__ R, float* __restrict__ Z, float* __restrict__ Nn, float* __restrict__ GHN,  // [T,B,H] or null
    int B, int T, int H, int CQT) {
    extern __shared__ __align__(16) float smem[];
    float* hs = smem;                 // [H][BT]
    float* part = smem + H * BT;      // [KG][BT][ld3]   (H*BT is a multiple of 4 floats because BT % 4 == 0)
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * BT;
    const int NCQ = ld3 >> 2;
    const int KG = SEQ_THREADS / CQT;
    const int kg = tid / CQT, cq0 = tid - kg * CQT;
    const int kchunk = (H + KG - 1) / KG;
    const int k0 = kg * kchunk, k1 = min(H, k0 + kchunk);
    const int H3 = 3 * H;

    for (int i = tid; i < H * BT; i += SEQ_THREADS) {
        const int k = i / BT, b = i - k * BT;
        hs[i] = (b0 + b < B) ? Hall[(long)(b0 + b) * H + k] : 0.f;
    }
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        // ---- phase 1: partial products over this thread group's k range ----
        for (int cq = cq0; cq < NCQ; cq += CQT) {
            float acc[BT][4];
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[b][c] = 0.f;
            const float* wp = WhhT + (long)k0 * ld3 + 4 * cq;
#pragma unroll 4
            for (int k = k0; k < k1; ++k, wp += ld3) {
                const float4 w = *reinterpret_cast<const float4*>(wp);
                float hv[BT];
#pragma unroll
                for (int b4 = 0; b4 < BT; b4 += 4) {
                    const float4 h4 = *reinterpret_cast<const float4*>(hs + k * BT + b4);
                    hv[b4] = h4.x; hv[b4 + 1] = h4.y; hv[b4 + 2] = h4.z; hv[b4 + 3] = h4.w;
                }
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    acc[b][0] = fmaf(w.x, hv[b], acc[b][0]);
                    acc[b][1] = fmaf(w.y, hv[b], acc[b][1]);
                    acc[b][2] = fmaf(w.z, hv[b], acc[b][2]);
                    acc[b][3] = fmaf(w.w, hv[b], acc[b][3]);
                }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b)
                *reinterpret_cast<float4*>(part + ((long)(kg * BT + b)) * ld3 + 4 * cq) =
                    make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
        }
        __syncthreads();
        // ---- phase 2: gates + state update, one (row, unit) per thread-iteration ----
        for (int i = tid; i < BT * H; i += SEQ_THREADS) {
            const int b = i / H, j = i - b * H;
            const int gb = b0 + b;
            if (gb < B) {
                float ghr = bhh[j], ghz = bhh[H + j], ghn = bhh[2 * H + j];
                for (int g = 0; g < KG; ++g) {
                    const float* pp = part + ((long)(g * BT + b)) * ld3;
                    ghr += pp[j]; ghz += pp[H + j]; ghn += pp[2 * H + j];
                }
                const float* gi = GI + ((long)t * B + gb) * H3;
                const float r = sigmoidf_acc(gi[j] + ghr);
                const float z = sigmoidf_acc(gi[H + j] + ghz);
                const float n = tanhf(gi[2 * H + j] + r * ghn);
                const float hp = hs[j * BT + b];
                const float hn = (1.f - z) * n + z * hp;
                hs[j * BT + b] = hn;
                const long o = ((long)t * B + gb) * H + j;
                Hall[o + (long)B * H] = hn;
                if (Hbm) Hbm[((long)gb * T + t) * H + j] = hn;
                if (R) { R[o] = r; Z[o] = z; Nn[o] = n; GHN[o] = ghn; }
            }
        }
        __syncthreads();
    }
}

// BPTT.  dHbm[b,t,:] is the gradient flowing into h_t from the vocabulary projection (batch-major);
// outputs dGI/dGH (time-major [T,B,3H]) feed the time-batched dW_ih / dW_hh / dx GEMMs; dh0 [B,H].
template <int BT>
__global__ void __launch_bounds__(SEQ_THREADS) gru_seq_bwd_kernel(
    const float* __restrict__ dHbm, const float* __restrict__ R, const float* __restrict__ Z,
    const float* __restrict__ Nn, const float* __restrict__ GHN, const float* __restrict__ Hall,
    const float* __restrict__ Whh,  // [3H, ldh]
    int ldh, float* __restrict__ dGI, float* __restrict__ dGH, float* __restrict__ dh0, int B, int T, int H,
    int CQT) {
    extern __shared__ __align__(16) float smem[];
    const int H3 = 3 * H;
    float* dg = smem;                     // [3H][BT]
    float* dhd = dg + H3 * BT;            // [BT][H]   direct term dh_t * z
    float* part = dhd + ((BT * H + 3) & ~3);  // [JG][BT][ldh]
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * BT;
    const int NCQ = ldh >> 2;
    const int JG = SEQ_THREADS / CQT;
    const int jg = tid / CQT, cq0 = tid - jg * CQT;
    const int jchunk = (H3 + JG - 1) / JG;
    const int j0 = jg * jchunk, j1 = min(H3, j0 + jchunk);

    for (int i = tid; i < BT * H; i += SEQ_THREADS) dhd[i] = 0.f;
    for (int i = tid; i < JG * BT * ldh; i += SEQ_THREADS) part[i] = 0.f;
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        for (int i = tid; i < BT * H; i += SEQ_THREADS) {
            const int b = i / H, j = i - b * H;
            const int gb = b0 + b;
            float dar = 0.f, daz = 0.f, dan = 0.f, danr = 0.f, keep = 0.f;
            if (gb < B) {
                float dht = dhd[i] + dHbm[((long)gb * T + t) * H + j];
                for (int g = 0; g < JG; ++g) dht += part[((long)(g * BT + b)) * ldh + j];
                const long o = ((long)t * B + gb) * H + j;
                const float r = R[o], z = Z[o], n = Nn[o], ghn = GHN[o];
                const float hp = Hall[o];  // Hall[t] = h_{t-1}
                const float dn = dht * (1.f - z);
                const float dz = dht * (hp - n);
                dan = dn * (1.f - n * n);
                dar = dan * ghn * r * (1.f - r);
                daz = dz * z * (1.f - z);
                danr = dan * r;
                keep = dht * z;
                float* gi = dGI + ((long)t * B + gb) * H3;
                float* gh = dGH + ((long)t * B + gb) * H3;
                gi[j] = dar; gi[H + j] = daz; gi[2 * H + j] = dan;
                gh[j] = dar; gh[H + j] = daz; gh[2 * H + j] = danr;
            }
            dhd[i] = keep;
            dg[j * BT + b] = dar;
            dg[(H + j) * BT + b] = daz;
            dg[(2 * H + j) * BT + b] = danr;
        }
        __syncthreads();
        for (int cq = cq0; cq < NCQ; cq += CQT) {
            float acc[BT][4];
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[b][c] = 0.f;
            const float* wp = Whh + (long)j0 * ldh + 4 * cq;
#pragma unroll 4
            for (int j = j0; j < j1; ++j, wp += ldh) {
                const float4 w = *reinterpret_cast<const float4*>(wp);
                float dv[BT];
#pragma unroll
                for (int b4 = 0; b4 < BT; b4 += 4) {
                    const float4 d4 = *reinterpret_cast<const float4*>(dg + j * BT + b4);
                    dv[b4] = d4.x; dv[b4 + 1] = d4.y; dv[b4 + 2] = d4.z; dv[b4 + 3] = d4.w;
                }
#pragma unroll
                for (int b = 0; b < BT; ++b) {
                    acc[b][0] = fmaf(w.x, dv[b], acc[b][0]);
                    acc[b][1] = fmaf(w.y, dv[b], acc[b][1]);
                    acc[b][2] = fmaf(w.z, dv[b], acc[b][2]);
                    acc[b][3] = fmaf(w.w, dv[b], acc[b][3]);
                }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b)
                *reinterpret_cast<float4*>(part + ((long)(jg * BT + b)) * ldh + 4 * cq) =
                    make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
        }
        __syncthreads();
    }
    for (int i = tid; i < BT * H; i += SEQ_THREADS) {
        const int b = i / H, j = i - b * H;
        if (b0 + b < B) {
            float v = dhd[i];
            for (int g = 0; g < JG; ++g) v += part[((long)(g * BT + b)) * ldh + j];
            dh0[(long)(b0 + b) * H + j] = v;
        }
    }
}

// dst[c*ldd + r] = src[r*lds + c]  (r < R, c < C); pad columns [R, ldd) of dst are zeroed.
__global__ void transpose_pad_kernel(const float* __restrict__ src, long lds, float* __restrict__ dst, long ldd, int R,
                                     int C) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < R && c < C) ? src[(long)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < ldd) dst[(long)c * ldd + r] = tile[threadIdx.x][i];
    }
}

// dst[r*ldd + c] = c < C ? src[r*lds + c] : 0
__global__ void copy_pad_kernel(const float* __restrict__ src, long lds, float* __restrict__ dst, long ldd, long R,
                                int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * ldd) return;
    const long r = i / ldd;
    const int c = (int)(i - r * ldd);
    dst[i] = c < C ? src[r * lds + c] : 0.f;
}

}  // namespace caphn

using namespace caphn;

extern "C" {

int caphn_transpose_pad(const float* src, long lds, float* dst, long ldd, int R, int C, void* stream) {
    if (R <= 0 || C <= 0 || ldd < R) return CAPHN_EINVAL;
    dim3 grid(ceil_div(C, 32), ceil_div(ldd, 32));
    transpose_pad_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, R, C);
    CAPHN_RETURN_LAST();
}

int caphn_copy_pad(const float* src, long lds, float* dst, long ldd, long R, int C, void* stream) {
    if (R <= 0 || C <= 0 || ldd < C) return CAPHN_EINVAL;
    copy_pad_kernel<<<ceil_div(R * ldd, 256), 256, 0, (cudaStream_t)stream>>>(src, lds, dst, ldd, R, C);
    CAPHN_RETURN_LAST();
}

// GRU recurrence over T steps.  GI [T,B,3H] (x-projection incl. b_ih); WhhT [H, ld3] (ld3 % 4 == 0, 16B aligned);
// Hall [T+1,B,H] with Hall[0] = h0 filled by the caller; optional Hbm [B,T,H]; optional saved gates R,Z,Nn,GHN [T,B,H].
int caphn_gru_seq_fwd(const float* GI, const float* WhhT, int ld3, const float* bhh, float* Hall, float* Hbm, float* R,
                      float* Z, float* Nn, float* GHN, int B, int T, int H, void* stream) {
    if (B <= 0 || T <= 0 || H <= 0 || (ld3 & 3) || ld3 < 3 * H || ((uintptr_t)WhhT & 15)) return CAPHN_EINVAL;
    if (R && !(Z && Nn && GHN)) return CAPHN_EINVAL;
    constexpr int BT = 4;
    int CQT = pow2_ceil(ld3 >> 2);
    if (CQT > SEQ_THREADS) CQT = SEQ_THREADS;
    if (CQT < 32) CQT = 32;
    const int KG = SEQ_THREADS / CQT;
    const size_t smem = ((size_t)H * BT + (size_t)KG * BT * ld3) * sizeof(float);
    if (smem > 227 * 1024) return CAPHN_EINVAL;
    CAPHN_CHECK(cudaFuncSetAttribute(gru_seq_fwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gru_seq_fwd_kernel<BT><<<ceil_div(B, BT), SEQ_THREADS, smem, (cudaStream_t)stream>>>(
        GI, WhhT, ld3, bhh, Hall, Hbm, R, Z, Nn, GHN, B, T, H, CQT);
    CAPHN_RETURN_LAST();
}

// BPTT of caphn_gru_seq_fwd.  Whh [3H, ldh] (ldh % 4 == 0).  dGI, dGH [T,B,3H]; dh0 [B,H].
int caphn_gru_seq_bwd(const float* dHbm, const float* R, const float* Z, const float* Nn, const float* GHN,
                      const float* Hall, const float* Whh, int ldh, float* dGI, float* dGH, float* dh0, int B, int T,
                      int H, void* stream) {
    if (B <= 0 || T <= 0 || H <= 0 || (ldh & 3) || ldh < H || ((uintptr_t)Whh & 15)) return CAPHN_EINVAL;
    constexpr int BT = 4;
    int CQT = pow2_ceil(ldh >> 2);
    if (CQT > SEQ_THREADS) CQT = SEQ_THREADS;
    if (CQT < 32) CQT = 32;
    const int JG = SEQ_THREADS / CQT;
    const size_t smem = ((size_t)3 * H * BT + (size_t)((BT * H + 3) & ~3) + (size_t)JG * BT * ldh) * sizeof(float);
    if (smem > 227 * 1024) return CAPHN_EINVAL;
    CAPHN_CHECK(cudaFuncSetAttribute(gru_seq_bwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gru_seq_bwd_kernel<BT><<<ceil_div(B, BT), SEQ_THREADS, smem, (cudaStream_t)stream>>>(
        dHbm, R, Z, Nn, GHN, Hall, Whh, ldh, dGI, dGH, dh0, B, T, H, CQT);
    CAPHN_RETURN_LAST();
}

}  // extern "C"
