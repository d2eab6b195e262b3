// Weights-resident GRU recurrence: the generated W_hh lives in SHARED MEMORY for all T steps, split across the CTAs of a
// thread-block cluster by hidden unit; the hidden state is exchanged through distributed shared memory (DSMEM) once per
// step.  This is the north-star design of BASELINE.json ("persistent kernel that keeps each style group's generated
// W_hh resident in shared memory across timesteps, gates / state update / BPTT gate gradients fused in").
//
// Replaces the same reference calls as gru_seq.cu (later.py:411,418 nn.GRUCell per step + autograd BPTT) for the single
// layer case; gru_seq.cu (weights streamed from L2) remains the fallback for num_layers > 1 or when the slice does not
// fit in shared memory.
//
// Cluster of CS CTAs owns BT = 8 batch rows.  CTA c owns hidden units [c*HS, (c+1)*HS) and keeps the 3*HS rows
// {r_j, z_j, n_j} of W_hh ([3H, H] row-major, fp32) in smem with an odd row stride (H | 1) so that both access patterns
// are bank-conflict free: forward  (thread = weight row, loop over k)  and backward (thread = k, loop over rows).
//   forward  step: gh = W_slice h (registers) -> smem -> gates for own units -> h' written to every CTA's next-step
//                  state buffer (DSMEM) -> cluster barrier.
//   backward step: gate gradients for own units -> partial dh = dgh_slice^T W_slice for ALL k -> reduce-scattered to the
//                  owners of k through DSMEM -> cluster barrier.
#include "common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace caphn {

constexpr int CL_THREADS = 512;
constexpr int CL_BT = 8;

struct GruClArgs {
    const float* GI;    // [T,B,3H]
    const float* Whh;   // [3H,H] row-major (dense)
    const float* bhh;   // [3H]
    float* Hall;        // [T+1,B,H]
    float* Hbm;         // [B,T,H] or null
    float* saved;       // [4][T,B,H] or null
    int B, T, H, HS, ldw;   // HS = hidden units per CTA, ldw = smem row stride (odd)
    // many-style batch (rows sorted by style group): cluster i owns rows [tiles[i].x, +tiles[i].y <= 8) of group tiles[i].z and
    // keeps THAT group's W_hh slice resident; group g's W_hh / b_hh start wstride / bstride floats after group g-1's.
    const int4* tiles;
    long wstride, bstride;
};

__device__ __forceinline__ void load_w_slice(float* Ws, const float* __restrict__ Whh, int H, int HS, int ldw, int c) {
    // local row lr = g*HS + jl  <->  global row g*H + c*HS + jl   (g = gate r,z,n);  rows of units >= H are zero
    const int rows = 3 * HS;
    for (int i = threadIdx.x; i < rows * H; i += CL_THREADS) {
        const int lr = i / H, k = i - lr * H;
        const int g = lr / HS, jl = lr - g * HS;
        const int j = c * HS + jl;
        Ws[lr * ldw + k] = (j < H) ? Whh[((long)g * H + j) * H + k] : 0.f;
    }
}

__global__ void __launch_bounds__(CL_THREADS, 1) gru_cluster_fwd_kernel(const GruClArgs a) {
    constexpr int BT = CL_BT;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = cluster.num_blocks();
    const int c = cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, HS = a.HS, ldw = a.ldw, B = a.B, T = a.T, H3 = 3 * a.H;
    const int rows = 3 * HS;
    float* Ws = smem;                                  // [3*HS][ldw]
    float* hs = Ws + ((rows * ldw + 3) & ~3);          // [2][H][BT]  double-buffered state (k-major, BT contiguous)
    float* ghs = hs + 2 * H * BT;                      // [3*HS][BT]
    const int tid = threadIdx.x;
    int b0 = (blockIdx.x / CS) * BT, nb = min(BT, B - b0), grp = 0;
    if (a.tiles) { const int4 tl = a.tiles[blockIdx.x / CS]; b0 = tl.x; nb = tl.y; grp = tl.z; }
    const float* bhh = a.bhh + (long)grp * a.bstride;
    const long TBH = (long)T * B * H;

    load_w_slice(Ws, a.Whh + (long)grp * a.wstride, H, HS, ldw, c);
    for (int i = tid; i < H * BT; i += CL_THREADS) {
        const int k = i / BT, b = i - k * BT;
        hs[i] = (b < nb) ? a.Hall[(long)(b0 + b) * H + k] : 0.f;
    }
    __syncthreads();
    cluster.sync();

    const int RW = (rows + 31) & ~31;                  // threads per k-range group (warp multiple)
    const int NG = max(1, CL_THREADS / RW);            // k-range groups
    const int kg = tid / RW, lr = tid - kg * RW;
    const int kchunk = (H + NG - 1) / NG;
    const int k0 = min(H, kg * kchunk), k1 = (kg < NG) ? min(H, k0 + kchunk) : k0;
    float* part = ghs + rows * BT;                     // [NG][rows][BT]

    // per-thread gate items (fixed over time): up to 2 x (unit jl, row b); biases hoisted, GI prefetched one step ahead
    constexpr int MAXI = 2;
    int it_j[MAXI], it_b[MAXI];
    bool it_ok[MAXI];
    float bh[MAXI][3], gi_next[MAXI][3];
#pragma unroll
    for (int q = 0; q < MAXI; ++q) {
        const int i = tid + q * CL_THREADS;
        const int jl = i / BT, b = i - jl * BT;
        it_j[q] = c * HS + jl; it_b[q] = b;
        it_ok[q] = (i < HS * BT) && (it_j[q] < H);
        const bool live = it_ok[q] && (b < nb);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            bh[q][g] = live ? bhh[g * H + it_j[q]] : 0.f;
            gi_next[q][g] = live ? a.GI[((long)0 * B + b0 + b) * H3 + g * H + it_j[q]] : 0.f;
        }
    }

    for (int t = 0; t < T; ++t) {
        const float* hc = hs + (t & 1) * H * BT;
        float* hn = hs + ((t + 1) & 1) * H * BT;
        float gi_cur[MAXI][3];
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            const bool live = it_ok[q] && (it_b[q] < nb) && (t + 1 < T);
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                gi_cur[q][g] = gi_next[q][g];
                if (live) gi_next[q][g] = a.GI[((long)(t + 1) * B + b0 + it_b[q]) * H3 + g * H + it_j[q]];
            }
        }
        // ---- gh partial: thread = (k range, local weight row) ----
        if (kg < NG) {
            for (int r = lr; r < rows; r += RW) {
                float acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = 0.f;
                const float* wr = Ws + r * ldw;
#pragma unroll 4
                for (int k = k0; k < k1; ++k) {
                    const float w = wr[k];
                    const float4 h0 = *reinterpret_cast<const float4*>(hc + k * BT);
                    const float4 h1 = *reinterpret_cast<const float4*>(hc + k * BT + 4);
                    acc[0] = fmaf(w, h0.x, acc[0]); acc[1] = fmaf(w, h0.y, acc[1]);
                    acc[2] = fmaf(w, h0.z, acc[2]); acc[3] = fmaf(w, h0.w, acc[3]);
                    acc[4] = fmaf(w, h1.x, acc[4]); acc[5] = fmaf(w, h1.y, acc[5]);
                    acc[6] = fmaf(w, h1.z, acc[6]); acc[7] = fmaf(w, h1.w, acc[7]);
                }
                float* pp = part + ((long)kg * rows + r) * BT;
                *reinterpret_cast<float4*>(pp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                *reinterpret_cast<float4*>(pp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
        __syncthreads();
        // ---- gates for own units; broadcast h' to every CTA of the cluster; THEN the global stores ----
        float o_r[MAXI], o_z[MAXI], o_n[MAXI], o_g[MAXI], o_h[MAXI];
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            o_r[q] = o_z[q] = o_n[q] = o_g[q] = o_h[q] = 0.f;
            if (it_ok[q]) {
                const int j = it_j[q], b = it_b[q], jl = j - c * HS;
                if (b < nb) {
                    float ghr = bh[q][0], ghz = bh[q][1], ghn = bh[q][2];
                    for (int g = 0; g < NG; ++g) {
                        const float* pp = part + (long)g * rows * BT;
                        ghr += pp[(jl) * BT + b]; ghz += pp[(HS + jl) * BT + b]; ghn += pp[(2 * HS + jl) * BT + b];
                    }
                    const float r = sigmoidf_acc(gi_cur[q][0] + ghr);
                    const float z = sigmoidf_acc(gi_cur[q][1] + ghz);
                    const float n = tanhf(gi_cur[q][2] + r * ghn);
                    const float hp = hc[j * BT + b];
                    o_r[q] = r; o_z[q] = z; o_n[q] = n; o_g[q] = ghn;
                    o_h[q] = (1.f - z) * n + z * hp;
                }
                for (int rk = 0; rk < CS; ++rk) cluster.map_shared_rank(hn, rk)[j * BT + b] = o_h[q];
            }
        }
        // release the DSMEM writes now; the (slow) global stores below are not covered by this barrier phase
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            const int gb = b0 + it_b[q];
            if (it_ok[q] && it_b[q] < nb) {
                const int j = it_j[q];
                const long o = ((long)t * B + gb) * H + j;
                a.Hall[o + (long)B * H] = o_h[q];
                if (a.Hbm) a.Hbm[((long)gb * T + t) * H + j] = o_h[q];
                if (a.saved) {
                    a.saved[o] = o_r[q]; a.saved[TBH + o] = o_z[q]; a.saved[2 * TBH + o] = o_n[q];
                    a.saved[3 * TBH + o] = o_g[q];
                }
            }
        }
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

struct GruClBwdArgs {
    const float* dHbm;   // [B,T,H]
    const float* saved;  // [4][T,B,H]
    const float* Hall;   // [T+1,B,H]
    const float* Whh;    // [3H,H]
    float* dGI; float* dGH;   // [T,B,3H]
    float* dh0;          // [B,H]
    int B, T, H, HS, ldw;
    const int4* tiles;   // many-style batch: see GruClArgs
    long wstride;
};

__global__ void __launch_bounds__(CL_THREADS, 1) gru_cluster_bwd_kernel(const GruClBwdArgs a) {
    constexpr int BT = CL_BT;
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = cluster.num_blocks();
    const int c = cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, HS = a.HS, ldw = a.ldw, B = a.B, T = a.T, H3 = 3 * a.H;
    const int rows = 3 * HS;
    const int HP = CS * HS;                            // padded hidden size (>= H)
    float* Ws = smem;                                  // [3*HS][ldw]
    float* dgs = Ws + ((rows * ldw + 3) & ~3);         // [3*HS][BT]   dgh of own units
    float* dhd = dgs + rows * BT;                      // [HS][BT]     direct term dh_t * z of own units
    float* recv = dhd + HS * BT;                       // [2][CS][HS][BT]  partial dh for own units from every CTA
    float* part = recv + 2 * CS * HS * BT;             // [NG][HP][BT]
    const int tid = threadIdx.x;
    int b0 = (blockIdx.x / CS) * BT, nb = min(BT, B - b0), grp = 0;
    if (a.tiles) { const int4 tl = a.tiles[blockIdx.x / CS]; b0 = tl.x; nb = tl.y; grp = tl.z; }
    const long TBH = (long)T * B * H;

    load_w_slice(Ws, a.Whh + (long)grp * a.wstride, H, HS, ldw, c);
    for (int i = tid; i < HS * BT; i += CL_THREADS) dhd[i] = 0.f;
    for (int i = tid; i < 2 * CS * HS * BT; i += CL_THREADS) recv[i] = 0.f;
    __syncthreads();
    cluster.sync();

    const int KW = (H + 31) & ~31;                     // threads per row-range group
    const int NG = max(1, CL_THREADS / KW);
    const int jg = tid / KW, kk = tid - jg * KW;
    const int jchunk = (rows + NG - 1) / NG;
    const int j0 = min(rows, jg * jchunk), j1 = (jg < NG) ? min(rows, j0 + jchunk) : j0;

    constexpr int MAXI = 2;
    int it_j[MAXI], it_b[MAXI];
    bool it_live[MAXI];
    float nx[MAXI][6];   // prefetched r, z, n, ghn, h_prev, dHbm of the step about to be processed
#pragma unroll
    for (int q = 0; q < MAXI; ++q) {
        const int i = tid + q * CL_THREADS;
        const int jl = i / BT, b = i - jl * BT;
        it_j[q] = c * HS + jl; it_b[q] = b;
        it_live[q] = (i < HS * BT) && (it_j[q] < H) && (b < nb);
        if (it_live[q]) {
            const long o = ((long)(T - 1) * B + b0 + b) * H + it_j[q];
            nx[q][0] = a.saved[o]; nx[q][1] = a.saved[TBH + o]; nx[q][2] = a.saved[2 * TBH + o];
            nx[q][3] = a.saved[3 * TBH + o]; nx[q][4] = a.Hall[o];
            nx[q][5] = a.dHbm[((long)(b0 + b) * T + (T - 1)) * H + it_j[q]];
        }
    }

    for (int t = T - 1; t >= 0; --t) {
        const int par = (T - 1 - t) & 1;
        const float* rc = recv + par * CS * HS * BT;           // written during the previous iteration
        float* rn_local = recv + (par ^ 1) * CS * HS * BT;     // to be written now (for the next iteration)
        // ---- gate gradients for own units ----
#pragma unroll
        for (int q = 0; q < MAXI; ++q) {
            const int i = tid + q * CL_THREADS;
            if (i < HS * BT) {
                const int jl = i / BT, b = it_b[q], j = it_j[q];
                float dar = 0.f, daz = 0.f, danr = 0.f, keep = 0.f;
                if (it_live[q]) {
                    float dht = dhd[i] + nx[q][5];
                    for (int sidx = 0; sidx < CS; ++sidx) dht += rc[(sidx * HS + jl) * BT + b];
                    const float r = nx[q][0], z = nx[q][1], n = nx[q][2], ghn = nx[q][3], hp = nx[q][4];
                    const float dn = dht * (1.f - z);
                    const float dz = dht * (hp - n);
                    const float dan = dn * (1.f - n * n);
                    dar = dan * ghn * r * (1.f - r);
                    daz = dz * z * (1.f - z);
                    danr = dan * r;
                    keep = dht * z;
                    const int gb = b0 + b;
                    float* gi = a.dGI + ((long)t * B + gb) * H3;
                    float* gh = a.dGH + ((long)t * B + gb) * H3;
                    gi[j] = dar; gi[H + j] = daz; gi[2 * H + j] = dan;
                    gh[j] = dar; gh[H + j] = daz; gh[2 * H + j] = danr;
                    if (t > 0) {   // prefetch the next (earlier) step while the matvec below runs
                        const long o = ((long)(t - 1) * B + gb) * H + j;
                        nx[q][0] = a.saved[o]; nx[q][1] = a.saved[TBH + o]; nx[q][2] = a.saved[2 * TBH + o];
                        nx[q][3] = a.saved[3 * TBH + o]; nx[q][4] = a.Hall[o];
                        nx[q][5] = a.dHbm[((long)gb * T + (t - 1)) * H + j];
                    }
                }
                dhd[i] = keep;
                dgs[jl * BT + b] = dar; dgs[(HS + jl) * BT + b] = daz; dgs[(2 * HS + jl) * BT + b] = danr;
            }
        }
        __syncthreads();
        // ---- partial dh[b][k] = sum over own rows of dgh[row][b] * W[row][k], for all k ----
        if (jg < NG) {
            for (int k = kk; k < H; k += KW) {
                float acc[BT];
#pragma unroll
                for (int b = 0; b < BT; ++b) acc[b] = 0.f;
#pragma unroll 4
                for (int r = j0; r < j1; ++r) {
                    const float w = Ws[r * ldw + k];
                    const float4 d0 = *reinterpret_cast<const float4*>(dgs + r * BT);
                    const float4 d1 = *reinterpret_cast<const float4*>(dgs + r * BT + 4);
                    acc[0] = fmaf(w, d0.x, acc[0]); acc[1] = fmaf(w, d0.y, acc[1]);
                    acc[2] = fmaf(w, d0.z, acc[2]); acc[3] = fmaf(w, d0.w, acc[3]);
                    acc[4] = fmaf(w, d1.x, acc[4]); acc[5] = fmaf(w, d1.y, acc[5]);
                    acc[6] = fmaf(w, d1.z, acc[6]); acc[7] = fmaf(w, d1.w, acc[7]);
                }
                float* pp = part + ((long)jg * HP + k) * BT;
                *reinterpret_cast<float4*>(pp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                *reinterpret_cast<float4*>(pp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
        __syncthreads();
        // ---- reduce-scatter: send the partial for unit k to its owner (slot = my rank) ----
        for (int i = tid; i < H * BT; i += CL_THREADS) {
            const int k = i / BT, b = i - k * BT;
            float v = 0.f;
            for (int g = 0; g < NG; ++g) v += part[((long)g * HP + k) * BT + b];
            const int owner = k / HS, kl = k - owner * HS;
            float* dst = cluster.map_shared_rank(rn_local, owner);
            dst[(c * HS + kl) * BT + b] = v;
        }
        cluster.sync();
    }
    const float* rc = recv + (T & 1) * CS * HS * BT;
    for (int i = tid; i < HS * BT; i += CL_THREADS) {
        const int jl = i / BT, b = i - jl * BT;
        const int j = c * HS + jl, gb = b0 + b;
        if (j < H && b < nb) {
            float v = dhd[i];
            for (int s = 0; s < CS; ++s) v += rc[(s * HS + jl) * BT + b];
            a.dh0[(long)gb * H + j] = v;
        }
    }
}

static inline int cl_groups(int n) {   // k / j range groups for a work width of n threads (rounded to a warp)
    const int w = (n + 31) & ~31;
    const int g = CL_THREADS / w;
    return g < 1 ? 1 : g;
}

static int pick_cluster(int H, int* HS, int* ldw, size_t* smem, bool bwd) {
    const int ld = H | 1;
    for (int CS = 2; CS <= 8; CS *= 2) {
        const int hs = (H + CS - 1) / CS;
        const size_t rows = 3 * (size_t)hs;
        size_t fl = ((rows * ld + 3) & ~(size_t)3);
        if (!bwd) fl += 2 * (size_t)H * CL_BT + rows * CL_BT + (size_t)cl_groups((int)rows) * rows * CL_BT;
        else fl += rows * CL_BT + (size_t)hs * CL_BT + 2 * (size_t)CS * hs * CL_BT +
                   (size_t)cl_groups(H) * CS * hs * CL_BT;
        if (rows > CL_THREADS || (size_t)hs * CL_BT > 2 * CL_THREADS) continue;   // thread-mapping limits
        const size_t bytes = fl * sizeof(float);
        if (bytes <= 200 * 1024) { *HS = hs; *ldw = ld; *smem = bytes; return CS; }
    }
    return 0;
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// Returns (via *cs) the cluster size the weights-resident kernels would use for hidden size H, 0 if W_hh does not fit.
int caphn_gru_cluster_plan(int H, int* cs) {
    int hs, ld; size_t sm, sm2;
    const int c1 = pick_cluster(H, &hs, &ld, &sm, false);
    const int c2 = pick_cluster(H, &hs, &ld, &sm2, true);
    *cs = (c1 && c2) ? (c1 > c2 ? c1 : c2) : 0;
    return CAPHN_OK;
}

// Weights-resident single-layer GRU recurrence (see file header).  Same tensors as caphn_gru_seq_fwd with NL = 1, but
// Whh is the plain row-major [3H,H] generated matrix (no transposed / padded copy).  saved = [4][T,B,H] or NULL.
static int gru_cluster_fwd_impl(const float* GI, const float* Whh, const float* bhh, float* Hall, float* Hbm, float* saved,
                                int B, int T, int H, const int* tiles, int ntiles, long wstride, long bstride, void* stream) {
    if (B <= 0 || T <= 0 || H <= 0) return CAPHN_EINVAL;
    GruClArgs a{GI, Whh, bhh, Hall, Hbm, saved, B, T, H, 0, 0, (const int4*)tiles, wstride, bstride};
    size_t smem;
    int cs1, cs;
    caphn_gru_cluster_plan(H, &cs);
    if (!cs) return CAPHN_EINVAL;
    cs1 = cs;
    a.HS = (H + cs1 - 1) / cs1;
    a.ldw = H | 1;
    const size_t rows = 3 * (size_t)a.HS;
    smem = (((rows * a.ldw + 3) & ~(size_t)3) + 2 * (size_t)H * CL_BT + rows * CL_BT +
            (size_t)cl_groups((int)rows) * rows * CL_BT) * sizeof(float);
    CAPHN_CHECK(cudaFuncSetAttribute(gru_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((tiles ? ntiles : ceil_div(B, CL_BT)) * cs1));
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CAPHN_CHECK(cudaLaunchKernelEx(&cfg, gru_cluster_fwd_kernel, a));
    CAPHN_RETURN_LAST();
}

int caphn_gru_cluster_fwd(const float* GI, const float* Whh, const float* bhh, float* Hall, float* Hbm, float* saved,
                          int B, int T, int H, void* stream) {
    return gru_cluster_fwd_impl(GI, Whh, bhh, Hall, Hbm, saved, B, T, H, nullptr, 0, 0, 0, stream);
}

// Many-style batch (rows sorted by style group): cluster i runs rows [tiles[i].x, +tiles[i].y <= 8) with the W_hh / b_hh of
// group tiles[i].z resident in its shared memory for all T steps (Whh / bhh point at group 0; strides in floats).
int caphn_gru_cluster_fwd_grouped(const float* GI, const float* Whh, const float* bhh, float* Hall, float* Hbm,
                                  float* saved, int B, int T, int H, const int* tiles, int ntiles, long wstride,
                                  long bstride, void* stream) {
    if (!tiles || ntiles < 1 || ((uintptr_t)tiles & 15)) return CAPHN_EINVAL;
    return gru_cluster_fwd_impl(GI, Whh, bhh, Hall, Hbm, saved, B, T, H, tiles, ntiles, wstride, bstride, stream);
}

// BPTT of caphn_gru_cluster_fwd: dGI, dGH [T,B,3H], dh0 [B,H].
static int gru_cluster_bwd_impl(const float* dHbm, const float* saved, const float* Hall, const float* Whh, float* dGI,
                                float* dGH, float* dh0, int B, int T, int H, const int* tiles, int ntiles, long wstride,
                                void* stream) {
    if (B <= 0 || T <= 0 || H <= 0) return CAPHN_EINVAL;
    int cs;
    caphn_gru_cluster_plan(H, &cs);
    if (!cs) return CAPHN_EINVAL;
    GruClBwdArgs a{dHbm, saved, Hall, Whh, dGI, dGH, dh0, B, T, H, (H + cs - 1) / cs, H | 1, (const int4*)tiles, wstride};
    const size_t rows = 3 * (size_t)a.HS;
    const size_t smem = (((rows * a.ldw + 3) & ~(size_t)3) + rows * CL_BT + (size_t)a.HS * CL_BT +
                         2 * (size_t)cs * a.HS * CL_BT + (size_t)cl_groups(H) * cs * a.HS * CL_BT) * sizeof(float);
    CAPHN_CHECK(cudaFuncSetAttribute(gru_cluster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((tiles ? ntiles : ceil_div(B, CL_BT)) * cs));
    cfg.blockDim = dim3(CL_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CAPHN_CHECK(cudaLaunchKernelEx(&cfg, gru_cluster_bwd_kernel, a));
    CAPHN_RETURN_LAST();
}

int caphn_gru_cluster_bwd(const float* dHbm, const float* saved, const float* Hall, const float* Whh, float* dGI,
                          float* dGH, float* dh0, int B, int T, int H, void* stream) {
    return gru_cluster_bwd_impl(dHbm, saved, Hall, Whh, dGI, dGH, dh0, B, T, H, nullptr, 0, 0, stream);
}

int caphn_gru_cluster_bwd_grouped(const float* dHbm, const float* saved, const float* Hall, const float* Whh, float* dGI,
                                  float* dGH, float* dh0, int B, int T, int H, const int* tiles, int ntiles, long wstride,
                                  void* stream) {
    if (!tiles || ntiles < 1 || ((uintptr_t)tiles & 15)) return CAPHN_EINVAL;
    return gru_cluster_bwd_impl(dHbm, saved, Hall, Whh, dGI, dGH, dh0, B, T, H, tiles, ntiles, wstride, stream);
}

}  // extern "C"
