// Weights-resident attention-GRU recurrence (forward): the generated W_hh, W_ih[:,E:] and the attention's U_a stay ON CHIP
// for all time steps, split by hidden unit over a thread-block cluster of 8 CTAs.
//
// Same reference calls as attgru_seq.cu (models/decoderlstm.py:97-100 + models/attention.py:33-45 per step); that
// L2-streaming kernel remains the fallback (odd sizes) and still provides the backward pass.
//
// Decomposition (cluster = 8 CTAs = 16 batch rows; CTA c owns hidden units [c*HS, (c+1)*HS), HS = ceil(H/8)):
//   * the CTA's weight rows -- U_a[j,:], W_hh[{r,z,n} j, :] (inputs: h) and W_ih[{r,z,n} j, E:] (input: ctx) for its
//     units j -- are held as bf16 hi/lo A-fragments of warp-level MMAs (mma.sync m16n8k16) in REGISTERS: one 16-row tile
//     per warp, loaded once, reused for every step.  fp32 accuracy comes from the same bf16x3 scheme as the GEMMs
//     (hi*hi + hi*lo + lo*hi, fp32 accumulate).  The per-step products are tiny (112 x 32 x 208): warp MMAs with
//     register-resident weights have far lower latency than a tcgen05/TMEM round trip, which is why they are used here.
//   * the attention itself is partitioned by batch row: CTA c scores / soft-maxes / forms the context for rows
//     2c, 2c+1 of the cluster.  K = W_a f + b_a and f of those two rows (157 KB fp32 at P=49, H=F=200) are loaded into
//     shared memory ONCE and stay there for all steps: the per-step attention touches no global memory.
//   * three DSMEM exchanges per step: u (all-to-all, so every CTA has the full u of its 4 rows), ctx (all-gather),
//     h' (all-gather), each followed by a cluster barrier.
// Per step: [MMA: u_slice, gh_slice = W.h] -> exchange u -> scores, softmax, ctx for own rows -> exchange ctx ->
//           [MMA: gi_ctx_slice = W_ihc.ctx] -> gates, state update for own units -> exchange h'.
#include "seq_common.cuh"
#include "mma_common.cuh"
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace caphn {

constexpr int AC_WARPS = 12;
constexpr int AC_THREADS = AC_WARPS * 32;
constexpr int AC_CS = 8;        // cluster size
constexpr int AC_BT = 16;       // batch rows per cluster (2 MMA n-tiles): K and f of these rows fit in the cluster's smem
constexpr int AC_NT = AC_BT / 8;
constexpr int AC_RPC = AC_BT / AC_CS;   // attention rows per CTA (2)
constexpr int AC_KT = 13;       // max k-tiles of 16 (H, F <= 208)
constexpr int AC_KP = AC_KT * 16 + 8;   // bf16 row pitch of the B-operand arrays (216: conflict-free fragment loads)
constexpr int AC_MAXI = 2;      // gate items per thread: HS * 16 <= 2 * 384  (HS <= 48)

struct AttClArgs {
    const float* Kp;    // [B,P,H]
    const float* f;     // [B,P,F]
    const float* GIw;   // [T,B,3H]
    const float* Ua;    // [H,H]    row-major (plain)
    const float* bu;    // [H]
    const float* va;    // [H]
    const float* bv;    // [1]
    const float* Wih;   // [3H, E+F] row-major (plain); the context half is columns E..E+F
    const float* Whh;   // [3H, H]
    const float* bhh;   // [3H]
    float* Hall;        // [T+1,B,H]
    float* Hbm;         // [B,T,H] or null
    float* attn;        // [B,T,P]
    float* ctx;         // ctx[t,b,:] at ctx + (t*B+b)*ldctx
    long ldctx;
    float* Upre; float* R; float* Z; float* Nn; float* GHN;   // [T,B,H] or null
    int B, T, P, H, F, E, HS, t0, t1;
};

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// local weight row lr of this CTA's slice.  group 0 (input h): [U_a | W_hh r | W_hh z | W_hh n]; group 1 (input ctx):
// [W_ih r | z | n][:, E:].  Returns the row pointer (nullptr: padding row) and its length.
__device__ __forceinline__ const float* slice_row(const AttClArgs& a, int group, int c, int lr, int& klen) {
    const int HS = a.HS, H = a.H;
    const int blk = lr / HS, jl = lr - blk * HS;
    const int j = c * HS + jl;
    klen = group == 0 ? H : a.F;
    if (j >= H || blk > (group == 0 ? 3 : 2)) return nullptr;
    if (group == 0) return blk == 0 ? a.Ua + (long)j * H : a.Whh + ((long)(blk - 1) * H + j) * H;
    return a.Wih + ((long)blk * H + j) * (a.E + a.F) + a.E;
}
__device__ __forceinline__ float row_at(const float* row, int k, int klen) { return (row && k < klen) ? __ldg(row + k) : 0.f; }

#ifdef CAPHN_ATTCL_TIMING
__device__ long long g_attcl_ts[32];
#define TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && t == a.t0 + 3) g_attcl_ts[i] = clock64(); } while (0)
#define TSP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_attcl_ts[i] = clock64(); } while (0)
#else
#define TS(i)
#define TSP(i)
#endif

// r(x) = 1 / (exp(2x) + 1)  (tanh x = 1 - 2 r(x); sigmoid(x) = r(-x/2)) with ex2.approx / rcp.approx: ~1e-7 absolute
__device__ __forceinline__ float recip_exp2x_p1(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));   // 2 * log2(e)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return recip_exp2x_p1(-0.5f * x); }

// push `n4` float4 from my shared memory to the same offset in every other CTA of the cluster (coalesced DSMEM stores)
__device__ __forceinline__ void cluster_push(cg::cluster_group& cluster, float* buf, int off_floats, int n4, int my_rank) {
    const float4* src = reinterpret_cast<const float4*>(buf + off_floats);
    for (int i = threadIdx.x; i < n4 * (AC_CS - 1); i += AC_THREADS) {
        int rk = i / n4;
        const int e = i - rk * n4;
        rk += (rk >= my_rank);
        reinterpret_cast<float4*>(cluster.map_shared_rank(buf, rk) + off_floats)[e] = src[e];
    }
}

__global__ void __launch_bounds__(AC_THREADS, 1) attgru_cluster_fwd_kernel(const AttClArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = cluster.block_rank();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int H = a.H, F = a.F, P = a.P, B = a.B, T = a.T, HS = a.HS, H3 = 3 * a.H;
    const int PS = (P + 3) & ~3;
    // ---- shared memory carve-up ----
    float* Ks = reinterpret_cast<float*>(smem_raw);                             // [RPC][P][H]  keys of my attention rows
    float* fs = Ks + AC_RPC * P * H;                                            // [RPC][P][F]  features of my rows
    float* stage_c = fs + AC_RPC * P * F;                                       // [BT][F]       ctx of all rows   (DSMEM target)
    float* stage_h = stage_c + AC_BT * F;                                       // [CS][BT][HS]  h' by owner slice (DSMEM target)
    float* us = stage_h + AC_CS * AC_BT * HS;                                   // [RPC][H]  u of my attention rows (DSMEM target)
    float* res_h = us + AC_RPC * H;                                             // [4*HS][BT] u | gh_r | gh_z | gh_n
    float* res_c = res_h + 4 * HS * AC_BT;                                      // [3*HS][BT] gi_ctx r | z | n
    float* hown = res_c + 3 * HS * AC_BT;                                       // [HS][BT]  fp32 state of my units
    float* sc = hown + HS * AC_BT;                                              // [RPC][PS]
    int* hmap = reinterpret_cast<int*>(sc + AC_RPC * PS);                       // [H] offset of unit j inside stage_h (row 0)
    __nv_bfloat16* hb_hi = reinterpret_cast<__nv_bfloat16*>(hmap + ((H + 3) & ~3));   // [BT][KP]  h as MMA B operand
    __nv_bfloat16* hb_lo = hb_hi + AC_BT * AC_KP;
    __nv_bfloat16* cb_hi = hb_lo + AC_BT * AC_KP;                               // [BT][KP]  ctx as MMA B operand
    __nv_bfloat16* cb_lo = cb_hi + AC_BT * AC_KP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = (blockIdx.x / AC_CS) * AC_BT;
    const int MTH = (4 * HS + 15) >> 4;                 // m-tiles of the h group; the ctx group uses the remaining warps
    const int group = warp < MTH ? 0 : 1;
    const int lr0 = (group == 0 ? warp : warp - MTH) * 16;

    TSP(20);
    // ---- weight fragments: loaded once, live in registers for all steps ----
    uint32_t Ahi[AC_KT][4], Alo[AC_KT][4];
    {
        const int kc = (lane & 3) * 2;
        int klen;
        const float* rowa = slice_row(a, group, c, lr0 + (lane >> 2), klen);
        const float* rowb = slice_row(a, group, c, lr0 + (lane >> 2) + 8, klen);
#pragma unroll
        for (int kt = 0; kt < AC_KT; ++kt) {
            const int k = kt * 16 + kc;
            split2(row_at(rowa, k, klen), row_at(rowa, k + 1, klen), Ahi[kt][0], Alo[kt][0]);
            split2(row_at(rowb, k, klen), row_at(rowb, k + 1, klen), Ahi[kt][1], Alo[kt][1]);
            split2(row_at(rowa, k + 8, klen), row_at(rowa, k + 9, klen), Ahi[kt][2], Alo[kt][2]);
            split2(row_at(rowb, k + 8, klen), row_at(rowb, k + 9, klen), Ahi[kt][3], Alo[kt][3]);
        }
    }
    TSP(21);
    // ---- K and f of my attention rows: read ONCE, resident for all steps ----
    for (int bl = 0; bl < AC_RPC; ++bl) {
        const int gb = b0 + c * AC_RPC + bl;
        const float4* ksrc = reinterpret_cast<const float4*>(a.Kp + (long)gb * P * H);
        const float4* fsrc = reinterpret_cast<const float4*>(a.f + (long)gb * P * F);
        float4* kd = reinterpret_cast<float4*>(Ks + bl * P * H);
        float4* fd = reinterpret_cast<float4*>(fs + bl * P * F);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < (P * H) >> 2; i += AC_THREADS) kd[i] = gb < B ? ksrc[i] : z4;
        for (int i = tid; i < (P * F) >> 2; i += AC_THREADS) fd[i] = gb < B ? fsrc[i] : z4;
    }
    TSP(22);
    // ---- initial state ----
    for (int j = tid; j < H; j += AC_THREADS) hmap[j] = (j / HS) * AC_BT * HS + (j % HS);
    for (int b = warp; b < AC_BT; b += AC_WARPS) {
        for (int k = lane; k < AC_KP; k += 32) {
            float v = 0.f;
            if (k < H && b0 + b < B) v = a.Hall[((long)a.t0 * B + b0 + b) * H + k];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hb_hi[b * AC_KP + k] = h;
            hb_lo[b * AC_KP + k] = __float2bfloat16_rn(v - __bfloat162float(h));
            cb_hi[b * AC_KP + k] = __float2bfloat16_rn(0.f);
            cb_lo[b * AC_KP + k] = __float2bfloat16_rn(0.f);
        }
    }
    // gate items of this thread (fixed over time): (unit jl, row b), lanes run over b
    int it_jl[AC_MAXI], it_b[AC_MAXI];
    bool it_in[AC_MAXI], it_live[AC_MAXI];
    float bh[AC_MAXI][3], bu_r[AC_MAXI];
#pragma unroll
    for (int q = 0; q < AC_MAXI; ++q) {
        const int i = tid + q * AC_THREADS;
        it_jl[q] = i / AC_BT; it_b[q] = i - it_jl[q] * AC_BT;
        const int j = c * HS + it_jl[q];
        it_in[q] = (i < HS * AC_BT) && (j < H);
        it_live[q] = it_in[q] && (b0 + it_b[q] < B);
#pragma unroll
        for (int g = 0; g < 3; ++g) bh[q][g] = it_live[q] ? a.bhh[g * H + j] : 0.f;
        bu_r[q] = it_in[q] ? a.bu[j] : 0.f;
        if (i < HS * AC_BT) hown[i] = it_live[q] ? a.Hall[((long)a.t0 * B + b0 + it_b[q]) * H + j] : 0.f;
    }
    // v_a slice of this lane (scores: lanes run over j); tanh = 1 - 2r  =>  score = b_v + sum(v) - 2 sum(v r)
    constexpr int JI = (AC_KT * 16 + 31) / 32;
    float va_r[JI];
    int jc[JI];
#pragma unroll
    for (int i = 0; i < JI; ++i) jc[i] = min(lane + 32 * i, H - 1);
    float sumva = 0.f;
#pragma unroll
    for (int i = 0; i < JI; ++i) { va_r[i] = (lane + 32 * i < H) ? a.va[lane + 32 * i] : 0.f; sumva += va_r[i]; }
    sumva = warp_sum(sumva) + a.bv[0];
    __syncthreads();
    cluster.sync();
    TSP(23);

    for (int t = a.t0; t < a.t1; ++t) {
        // prefetch the word half of the input projection for the gate phase of this step
        float giw[AC_MAXI][3];
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q)
#pragma unroll
            for (int g = 0; g < 3; ++g)
                giw[q][g] = it_live[q] ? a.GIw[((long)t * B + b0 + it_b[q]) * H3 + g * H + c * HS + it_jl[q]] : 0.f;

        TS(0);
        // ---- P1: [u | gh] slice = W_h-group . h   (warp MMA, weights from registers; 4 independent accumulator chains) ----
        if (group == 0) {
            float acc[2][AC_NT][4];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int nt = 0; nt < AC_NT; ++nt) { acc[e][nt][0] = acc[e][nt][1] = acc[e][nt][2] = acc[e][nt][3] = 0.f; }
            const int nrow = lane >> 2, kc = (lane & 3) * 2;
#pragma unroll
            for (int kt = 0; kt < AC_KT; ++kt) {     // unconditional: fragments / operand columns beyond the real K are zero
#pragma unroll
                for (int nt = 0; nt < AC_NT; ++nt) {
                    const uint32_t* bh_p = reinterpret_cast<const uint32_t*>(hb_hi + (nt * 8 + nrow) * AC_KP + kt * 16 + kc);
                    const uint32_t* bl_p = reinterpret_cast<const uint32_t*>(hb_lo + (nt * 8 + nrow) * AC_KP + kt * 16 + kc);
                    const uint32_t h0 = bh_p[0], h1 = bh_p[4], l0 = bl_p[0], l1 = bl_p[4];
                    mma_bf16(acc[kt & 1][nt], Ahi[kt], h0, h1);
                    mma_bf16(acc[kt & 1][nt], Ahi[kt], l0, l1);
                    mma_bf16(acc[kt & 1][nt], Alo[kt], h0, h1);
                }
            }
            const int ra = lr0 + (lane >> 2);
#pragma unroll
            for (int nt = 0; nt < AC_NT; ++nt) {
                const int col = nt * 8 + (lane & 3) * 2;
                if (ra < 4 * HS) {
                    res_h[ra * AC_BT + col] = acc[0][nt][0] + acc[1][nt][0];
                    res_h[ra * AC_BT + col + 1] = acc[0][nt][1] + acc[1][nt][1];
                }
                if (ra + 8 < 4 * HS) {
                    res_h[(ra + 8) * AC_BT + col] = acc[0][nt][2] + acc[1][nt][2];
                    res_h[(ra + 8) * AC_BT + col + 1] = acc[0][nt][3] + acc[1][nt][3];
                }
            }
        }
        __syncthreads();
        TS(1);
        // ---- exchange u: element (unit jl, row b) goes to the CTA that attends row b ----
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q) {
            if (it_in[q]) {
                const int jl = it_jl[q], b = it_b[q], j = c * HS + jl;
                const float u = res_h[jl * AC_BT + b] + bu_r[q];
                cluster.map_shared_rank(us, b / AC_RPC)[(b % AC_RPC) * H + j] = u;
                if (a.Upre && it_live[q]) a.Upre[((long)t * B + b0 + b) * H + j] = u;
            }
        }
        TS(2);
        cl_arrive();
        cl_wait();
        TS(3);
        // ---- P2: attention for my rows, everything from shared memory: scores -> softmax -> context ----
        for (int pair = warp; pair < AC_RPC * P; pair += 2 * AC_WARPS) {     // two (row, position) pairs per iteration
            const int pair2 = pair + AC_WARPS;
            const bool two = pair2 < AC_RPC * P;
            const int bl1 = pair / P, p1 = pair - bl1 * P;
            const int bl2 = two ? pair2 / P : bl1, p2 = two ? pair2 - bl2 * P : p1;
            const float* k1 = Ks + (bl1 * P + p1) * H;
            const float* k2 = Ks + (bl2 * P + p2) * H;
            const float* u1 = us + bl1 * H;
            const float* u2 = us + bl2 * H;
            float kv1[JI], kv2[JI], uv1[JI], uv2[JI];
#pragma unroll
            for (int i = 0; i < JI; ++i) {      // clamped index: branch-free, all loads issued before the math (v_a is 0 beyond H)
                kv1[i] = k1[jc[i]]; uv1[i] = u1[jc[i]];
                kv2[i] = k2[jc[i]]; uv2[i] = u2[jc[i]];
            }
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < JI; ++i) {
                s1 = fmaf(va_r[i], recip_exp2x_p1(kv1[i] + uv1[i]), s1);
                s2 = fmaf(va_r[i], recip_exp2x_p1(kv2[i] + uv2[i]), s2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) {
                sc[bl1 * PS + p1] = sumva - 2.f * s1;
                if (two) sc[bl2 * PS + p2] = sumva - 2.f * s2;
            }
        }
        __syncthreads();
        TS(4);
        if (warp < AC_RPC) {
            const int bl = warp, gb = b0 + c * AC_RPC + bl;
            const float x0 = lane < P ? sc[bl * PS + lane] : -INFINITY;
            const float x1 = lane + 32 < P ? sc[bl * PS + lane + 32] : -INFINITY;
            const float mx = warp_max(fmaxf(x0, x1));
            const float e0 = lane < P ? __expf(x0 - mx) : 0.f;
            const float e1 = lane + 32 < P ? __expf(x1 - mx) : 0.f;
            const float inv = 1.f / warp_sum(e0 + e1);
            if (lane < P) { sc[bl * PS + lane] = e0 * inv; if (gb < B) a.attn[((long)gb * T + t) * P + lane] = e0 * inv; }
            if (lane + 32 < P) { sc[bl * PS + lane + 32] = e1 * inv; if (gb < B) a.attn[((long)gb * T + t) * P + lane + 32] = e1 * inv; }
            for (int p = lane + 64; p < P; p += 32) {   // P > 64: generic tail (not the 7x7 case)
                const float al = __expf(sc[bl * PS + p] - mx) * inv;
                sc[bl * PS + p] = al;
                if (gb < B) a.attn[((long)gb * T + t) * P + p] = al;
            }
        }
        __syncthreads();
        TS(5);
        for (int i = tid; i < AC_RPC * F; i += AC_THREADS) {
            const int bl = i / F, fi = i - bl * F;
            const int brow = c * AC_RPC + bl, gb = b0 + brow;
            const float* fp = fs + bl * P * F + fi;
            const float* al = sc + bl * PS;
            float cv0 = 0.f, cv1 = 0.f;
            int p = 0;
            for (; p + 8 <= P; p += 8) {
                float fv[8], av[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { fv[e] = fp[(p + e) * F]; av[e] = al[p + e]; }
#pragma unroll
                for (int e = 0; e < 8; e += 2) { cv0 = fmaf(av[e], fv[e], cv0); cv1 = fmaf(av[e + 1], fv[e + 1], cv1); }
            }
            for (; p < P; ++p) cv0 = fmaf(al[p], fp[p * F], cv0);
            const float cv = cv0 + cv1;
            if (gb < B) a.ctx[((long)t * B + gb) * a.ldctx + fi] = cv;
            stage_c[brow * F + fi] = cv;
        }
        __syncthreads();
        cluster_push(cluster, stage_c, c * AC_RPC * F, (AC_RPC * F) >> 2, c);
        TS(6);
        cl_arrive();
        cl_wait();
        TS(7);
        // ---- ctx -> bf16 hi/lo B operand (local), then P3: gi_ctx slice = W_ihc-group . ctx ----
        {
            const int halfF = F >> 1;
            for (int i = tid; i < AC_BT * halfF; i += AC_THREADS) {
                const int b = i / halfF, k = (i - b * halfF) * 2;
                const float2 v = *reinterpret_cast<const float2*>(stage_c + b * F + k);
                uint32_t hi, lo;
                split2(v.x, v.y, hi, lo);
                *reinterpret_cast<uint32_t*>(cb_hi + b * AC_KP + k) = hi;
                *reinterpret_cast<uint32_t*>(cb_lo + b * AC_KP + k) = lo;
            }
        }
        __syncthreads();
        TS(8);
        if (group == 1 && lr0 < 3 * HS) {
            float acc[2][AC_NT][4];
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int nt = 0; nt < AC_NT; ++nt) { acc[e][nt][0] = acc[e][nt][1] = acc[e][nt][2] = acc[e][nt][3] = 0.f; }
            const int nrow = lane >> 2, kc = (lane & 3) * 2;
#pragma unroll
            for (int kt = 0; kt < AC_KT; ++kt) {
#pragma unroll
                for (int nt = 0; nt < AC_NT; ++nt) {
                    const uint32_t* bh_p = reinterpret_cast<const uint32_t*>(cb_hi + (nt * 8 + nrow) * AC_KP + kt * 16 + kc);
                    const uint32_t* bl_p = reinterpret_cast<const uint32_t*>(cb_lo + (nt * 8 + nrow) * AC_KP + kt * 16 + kc);
                    const uint32_t h0 = bh_p[0], h1 = bh_p[4], l0 = bl_p[0], l1 = bl_p[4];
                    mma_bf16(acc[kt & 1][nt], Ahi[kt], h0, h1);
                    mma_bf16(acc[kt & 1][nt], Ahi[kt], l0, l1);
                    mma_bf16(acc[kt & 1][nt], Alo[kt], h0, h1);
                }
            }
            const int ra = lr0 + (lane >> 2);
#pragma unroll
            for (int nt = 0; nt < AC_NT; ++nt) {
                const int col = nt * 8 + (lane & 3) * 2;
                if (ra < 3 * HS) {
                    res_c[ra * AC_BT + col] = acc[0][nt][0] + acc[1][nt][0];
                    res_c[ra * AC_BT + col + 1] = acc[0][nt][1] + acc[1][nt][1];
                }
                if (ra + 8 < 3 * HS) {
                    res_c[(ra + 8) * AC_BT + col] = acc[0][nt][2] + acc[1][nt][2];
                    res_c[(ra + 8) * AC_BT + col + 1] = acc[0][nt][3] + acc[1][nt][3];
                }
            }
        }
        __syncthreads();
        TS(9);
        // ---- P4: gates + state update for my units (written into my slice of stage_h), then pushed to every CTA ----
        float o_r[AC_MAXI], o_z[AC_MAXI], o_n[AC_MAXI], o_g[AC_MAXI], o_h[AC_MAXI];
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q) {
            o_r[q] = o_z[q] = o_n[q] = o_g[q] = o_h[q] = 0.f;
            const int i = tid + q * AC_THREADS;
            if (i < HS * AC_BT) {
                const int jl = it_jl[q], b = it_b[q];
                if (it_live[q]) {
                    const float ghr = res_h[(HS + jl) * AC_BT + b] + bh[q][0];
                    const float ghz = res_h[(2 * HS + jl) * AC_BT + b] + bh[q][1];
                    const float ghn = res_h[(3 * HS + jl) * AC_BT + b] + bh[q][2];
                    const float gir = giw[q][0] + res_c[jl * AC_BT + b];
                    const float giz = giw[q][1] + res_c[(HS + jl) * AC_BT + b];
                    const float gin = giw[q][2] + res_c[(2 * HS + jl) * AC_BT + b];
                    const float r = sigmoid_fast(gir + ghr);
                    const float z = sigmoid_fast(giz + ghz);
                    const float n = 1.f - 2.f * recip_exp2x_p1(gin + r * ghn);
                    const float hp = hown[i];
                    o_r[q] = r; o_z[q] = z; o_n[q] = n; o_g[q] = ghn;
                    o_h[q] = (1.f - z) * n + z * hp;
                    hown[i] = o_h[q];
                }
                stage_h[(c * AC_BT + b) * HS + jl] = o_h[q];
            }
        }
        __syncthreads();
        cluster_push(cluster, stage_h, c * AC_BT * HS, (AC_BT * HS) >> 2, c);
        TS(10);
        cl_arrive();
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q) {
            if (it_live[q]) {
                const int j = c * HS + it_jl[q], gb = b0 + it_b[q];
                const long o = ((long)t * B + gb) * H + j;
                a.Hall[o + (long)B * H] = o_h[q];
                if (a.Hbm) a.Hbm[((long)gb * T + t) * H + j] = o_h[q];
                if (a.R) { a.R[o] = o_r[q]; a.Z[o] = o_z[q]; a.Nn[o] = o_n[q]; a.GHN[o] = o_g[q]; }
            }
        }
        cl_wait();
        TS(11);
        // ---- h' -> bf16 hi/lo B operand (local) for the next step ----
        {
            const int halfH = (H + 1) >> 1;
            for (int i = tid; i < AC_BT * halfH; i += AC_THREADS) {
                const int b = i / halfH, j = (i - b * halfH) * 2;
                const float v0 = stage_h[hmap[j] + b * HS];
                const float v1 = (j + 1 < H) ? stage_h[hmap[j + 1] + b * HS] : 0.f;
                uint32_t hi, lo;
                split2(v0, v1, hi, lo);
                *reinterpret_cast<uint32_t*>(hb_hi + b * AC_KP + j) = hi;
                *reinterpret_cast<uint32_t*>(hb_lo + b * AC_KP + j) = lo;
            }
        }
        __syncthreads();
        TS(12);
        // (stage_c / stage_h / us are single-buffered: each is rewritten by peers only after a cluster barrier that every
        //  CTA reaches after its last read of the previous contents -- see the phase order above.)
    }
    cluster.sync();   // nobody exits while peers may still address its shared memory
    TSP(24);
}

static size_t attcl_smem(int H, int F, int P, int HS) {
    const int PS = (P + 3) & ~3;
    return ((size_t)AC_RPC * P * H + (size_t)AC_RPC * P * F + (size_t)AC_BT * F + (size_t)AC_CS * AC_BT * HS +
            (size_t)AC_RPC * H + (size_t)7 * HS * AC_BT + (size_t)HS * AC_BT + (size_t)AC_RPC * PS + (size_t)((H + 3) & ~3)) *
               sizeof(float) +
           (size_t)4 * AC_BT * AC_KP * 2;
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// *ok = 1 when the weights-resident attention recurrence supports these sizes (else use caphn_attgru_seq_fwd).
int caphn_attgru_cluster_plan(int H, int F, int P, int* ok) {
    const int HS = (H + AC_CS - 1) / AC_CS;
    const int MTH = (4 * HS + 15) / 16, MTC = (3 * HS + 15) / 16;
    *ok = (H >= 8 && H <= AC_KT * 16 && F >= 1 && F <= AC_KT * 16 && MTH + MTC <= AC_WARPS &&
           HS * AC_BT <= AC_MAXI * AC_THREADS && P >= 1 && attcl_smem(H, F, P, HS) <= 226 * 1024 && (P * H) % 4 == 0 && (P * F) % 4 == 0 && F % 4 == 0 &&
           (AC_BT * HS) % 4 == 0) ? 1 : 0;
    return CAPHN_OK;
}

// Steps [t0,t1) of the attention-GRU recurrence with the weights resident on chip (see file header).  Same tensors as
// caphn_attgru_seq_fwd, except that the weights are the PLAIN row-major matrices: Ua [H,H], Wih [3H,E+F], Whh [3H,H].
int caphn_attgru_cluster_fwd(const float* Kp, const float* f, const float* GIw, const float* Ua, const float* bu,
                             const float* va, const float* bv, const float* Wih, const float* Whh, const float* bhh,
                             float* Hall, float* Hbm, float* attn, float* ctx, long ldctx, float* Upre, float* R, float* Z,
                             float* Nn, float* GHN, int B, int T, int P, int H, int F, int E, int t0, int t1,
                             void* stream) {
    int ok = 0;
    caphn_attgru_cluster_plan(H, F, P, &ok);
    if (!ok || B <= 0 || T <= 0 || t0 < 0 || t1 > T || t0 >= t1) return CAPHN_EINVAL;
    if (R && !(Z && Nn && GHN && Upre)) return CAPHN_EINVAL;
    const int HS = (H + AC_CS - 1) / AC_CS;
    AttClArgs a{Kp, f, GIw, Ua, bu, va, bv, Wih, Whh, bhh, Hall, Hbm, attn, ctx, ldctx, Upre, R, Z, Nn, GHN,
                B, T, P, H, F, E, HS, t0, t1};
    const size_t smem = attcl_smem(H, F, P, HS);
    CAPHN_CHECK(cudaFuncSetAttribute(attgru_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ceil_div(B, AC_BT) * AC_CS));
    cfg.blockDim = dim3(AC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = AC_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CAPHN_CHECK(cudaLaunchKernelEx(&cfg, attgru_cluster_fwd_kernel, a));
    CAPHN_RETURN_LAST();
}

#ifdef CAPHN_ATTCL_TIMING
int caphn_attcl_timestamps(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, caphn::g_attcl_ts, sizeof(long long) * 32);
}
int caphn_attcl_max_clusters(int H, int F, int P) {
    const int HS = (H + caphn::AC_CS - 1) / caphn::AC_CS;
    const size_t smem = caphn::attcl_smem(H, F, P, HS);
    cudaFuncSetAttribute(caphn::attgru_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(caphn::AC_CS * 64);
    cfg.blockDim = dim3(caphn::AC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = caphn::AC_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaOccupancyMaxActiveClusters(&n, caphn::attgru_cluster_fwd_kernel, &cfg);
    return n;
}
#endif

}  // extern "C"
