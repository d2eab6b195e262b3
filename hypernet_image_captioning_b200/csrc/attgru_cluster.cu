// Weights-resident attention-GRU recurrence (forward): the generated W_hh, W_ih[:,E:] and the attention's U_a stay ON CHIP
// for all time steps, split by hidden unit over a thread-block cluster of 8 CTAs.
//
// Same reference calls as attgru_seq.cu (models/decoderlstm.py:97-100 + models/attention.py:33-45 per step); that
// L2-streaming kernel remains the fallback (odd sizes) and still provides the backward pass.
//
// Decomposition (cluster = 8 CTAs = 32 batch rows; CTA c owns hidden units [c*HS, (c+1)*HS), HS = ceil(H/8)):
//   * the CTA's weight rows -- U_a[j,:], W_hh[{r,z,n} j, :] (inputs: h) and W_ih[{r,z,n} j, E:] (input: ctx) for its
//     units j -- are held as bf16 hi/lo A-fragments of warp-level MMAs (mma.sync m16n8k16) in REGISTERS: one 16-row tile
//     per warp, loaded once, reused for every step.  fp32 accuracy comes from the same bf16x3 scheme as the GEMMs
//     (hi*hi + hi*lo + lo*hi, fp32 accumulate).  The per-step products are tiny (112 x 32 x 208): warp MMAs with
//     register-resident weights have far lower latency than a tcgen05/TMEM round trip, which is why they are used here.
//   * the attention itself is partitioned by batch row: CTA c scores / soft-maxes / forms the context for rows
//     4c..4c+3 of the cluster, reading K = W_a f + b_a and f for those rows (coalesced, L2 resident).
//   * three DSMEM exchanges per step: u (all-to-all, so every CTA has the full u of its 4 rows), ctx (all-gather),
//     h' (all-gather), each followed by a cluster barrier.
// Per step: [MMA: u_slice, gh_slice = W.h] -> exchange u -> scores, softmax, ctx for own rows -> exchange ctx ->
//           [MMA: gi_ctx_slice = W_ihc.ctx] -> gates, state update for own units -> exchange h'.
#include "seq_common.cuh"
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace caphn {

constexpr int AC_WARPS = 12;
constexpr int AC_THREADS = AC_WARPS * 32;
constexpr int AC_CS = 8;        // cluster size
constexpr int AC_BT = 32;       // batch rows per cluster (4 MMA n-tiles)
constexpr int AC_RPC = AC_BT / AC_CS;   // attention rows per CTA (4)
constexpr int AC_KT = 13;       // max k-tiles of 16 (H, F <= 208)
constexpr int AC_KP = AC_KT * 16 + 8;   // bf16 row pitch of the B-operand arrays (216: conflict-free fragment loads)
constexpr int AC_MAXI = 3;      // gate items per thread: HS * 32 <= 3 * 384  (HS <= 36)

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

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    hi = pack_bf16(h0, h1);
    lo = pack_bf16(__float2bfloat16_rn(x0 - __bfloat162float(h0)), __float2bfloat16_rn(x1 - __bfloat162float(h1)));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// value of local weight row lr, column k, of this CTA's slice.  group 0 (input h): [U_a | W_hh r | W_hh z | W_hh n];
// group 1 (input ctx): [W_ih r | z | n][:, E + k].
__device__ __forceinline__ float slice_w(const AttClArgs& a, int group, int c, int lr, int k) {
    const int HS = a.HS, H = a.H;
    const int blk = lr / HS, jl = lr - blk * HS;
    const int j = c * HS + jl;
    if (j >= H) return 0.f;
    if (group == 0) {
        if (blk > 3 || k >= H) return 0.f;
        return blk == 0 ? a.Ua[(long)j * H + k] : a.Whh[((long)(blk - 1) * H + j) * H + k];
    }
    if (blk > 2 || k >= a.F) return 0.f;
    return a.Wih[((long)blk * H + j) * (a.E + a.F) + a.E + k];
}

__global__ void __launch_bounds__(AC_THREADS, 1) attgru_cluster_fwd_kernel(const AttClArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int c = cluster.block_rank();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int H = a.H, F = a.F, P = a.P, B = a.B, T = a.T, HS = a.HS, H3 = 3 * a.H;
    const int PS = (P + 3) & ~3;
    // ---- shared memory carve-up ----
    __nv_bfloat16* hb_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [32][KP]  h as MMA B operand
    __nv_bfloat16* hb_lo = hb_hi + AC_BT * AC_KP;
    __nv_bfloat16* cb_hi = hb_lo + AC_BT * AC_KP;                               // [32][KP]  ctx as MMA B operand
    __nv_bfloat16* cb_lo = cb_hi + AC_BT * AC_KP;
    float* hx = reinterpret_cast<float*>(cb_lo + AC_BT * AC_KP);                // [32][H]   fp32 h' staging (DSMEM target)
    float* cxs = hx + AC_BT * H;                                                // [32][F]   fp32 ctx staging (DSMEM target)
    float* us = cxs + AC_BT * F;                                                // [RPC][H]  u of my attention rows (DSMEM)
    float* res_h = us + AC_RPC * H;                                             // [4*HS][32] u | gh_r | gh_z | gh_n
    float* res_c = res_h + 4 * HS * AC_BT;                                      // [3*HS][32] gi_ctx r | z | n
    float* hown = res_c + 3 * HS * AC_BT;                                       // [HS][32]  fp32 state of my units
    float* sc = hown + HS * AC_BT;                                              // [RPC][PS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = (blockIdx.x / AC_CS) * AC_BT;
    const int KHT = (H + 15) >> 4, KFT = (F + 15) >> 4;
    const int MTH = (4 * HS + 15) >> 4;                 // m-tiles of the h group; the ctx group uses the remaining warps
    const int group = warp < MTH ? 0 : 1;
    const int lr0 = (group == 0 ? warp : warp - MTH) * 16;
    const int nkt = group == 0 ? KHT : KFT;

    // ---- weight fragments: loaded once, live in registers for all steps ----
    uint32_t Ahi[AC_KT][4], Alo[AC_KT][4];
    {
        const int ra = lr0 + (lane >> 2), rb = ra + 8, kc = (lane & 3) * 2;
#pragma unroll
        for (int kt = 0; kt < AC_KT; ++kt) {
            if (kt < nkt) {
                const int k = kt * 16 + kc;
                split2(slice_w(a, group, c, ra, k), slice_w(a, group, c, ra, k + 1), Ahi[kt][0], Alo[kt][0]);
                split2(slice_w(a, group, c, rb, k), slice_w(a, group, c, rb, k + 1), Ahi[kt][1], Alo[kt][1]);
                split2(slice_w(a, group, c, ra, k + 8), slice_w(a, group, c, ra, k + 9), Ahi[kt][2], Alo[kt][2]);
                split2(slice_w(a, group, c, rb, k + 8), slice_w(a, group, c, rb, k + 9), Ahi[kt][3], Alo[kt][3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) { Ahi[kt][i] = 0u; Alo[kt][i] = 0u; }
            }
        }
    }
    // ---- initial state ----
    for (int i = tid; i < AC_BT * AC_KP; i += AC_THREADS) {
        const int b = i / AC_KP, k = i - b * AC_KP;
        float v = 0.f;
        if (k < H && b0 + b < B) v = a.Hall[((long)a.t0 * B + b0 + b) * H + k];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hb_hi[i] = h;
        hb_lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
        cb_hi[i] = __float2bfloat16_rn(0.f);
        cb_lo[i] = __float2bfloat16_rn(0.f);
    }
    for (int i = tid; i < HS * AC_BT; i += AC_THREADS) {
        const int jl = i / AC_BT, b = i - jl * AC_BT;
        const int j = c * HS + jl;
        hown[i] = (j < H && b0 + b < B) ? a.Hall[((long)a.t0 * B + b0 + b) * H + j] : 0.f;
    }
    // gate items of this thread (fixed over time): (unit jl, row b), lanes run over b
    int it_jl[AC_MAXI], it_b[AC_MAXI];
    bool it_live[AC_MAXI];
    float bh[AC_MAXI][3], bu_r[AC_MAXI];
#pragma unroll
    for (int q = 0; q < AC_MAXI; ++q) {
        const int i = tid + q * AC_THREADS;
        it_jl[q] = i / AC_BT; it_b[q] = i - it_jl[q] * AC_BT;
        const int j = c * HS + it_jl[q];
        it_live[q] = (i < HS * AC_BT) && (j < H) && (b0 + it_b[q] < B);
#pragma unroll
        for (int g = 0; g < 3; ++g) bh[q][g] = it_live[q] ? a.bhh[g * H + j] : 0.f;
        bu_r[q] = it_live[q] ? a.bu[j] : 0.f;
    }
    const float bv = a.bv[0];
    __syncthreads();
    cluster.sync();

    for (int t = a.t0; t < a.t1; ++t) {
        // prefetch the word half of the input projection for the gate phase of this step
        float giw[AC_MAXI][3];
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q)
#pragma unroll
            for (int g = 0; g < 3; ++g)
                giw[q][g] = it_live[q] ? a.GIw[((long)t * B + b0 + it_b[q]) * H3 + g * H + c * HS + it_jl[q]] : 0.f;

        // ---- P1: [u | gh] slice = W_h-group . h   (warp MMA, weights from registers) ----
        if (group == 0) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const int n = nt * 8 + (lane >> 2);
                const uint32_t* bh_p = reinterpret_cast<const uint32_t*>(hb_hi + n * AC_KP + (lane & 3) * 2);
                const uint32_t* bl_p = reinterpret_cast<const uint32_t*>(hb_lo + n * AC_KP + (lane & 3) * 2);
#pragma unroll
                for (int kt = 0; kt < AC_KT; ++kt) {
                    if (kt < nkt) {
                        const uint32_t h0 = bh_p[kt * 8], h1 = bh_p[kt * 8 + 4];
                        const uint32_t l0 = bl_p[kt * 8], l1 = bl_p[kt * 8 + 4];
                        mma_bf16(acc, Ahi[kt], h0, h1);
                        mma_bf16(acc, Ahi[kt], l0, l1);
                        mma_bf16(acc, Alo[kt], h0, h1);
                    }
                }
                const int ra = lr0 + (lane >> 2), col = nt * 8 + (lane & 3) * 2;
                if (ra < 4 * HS) { res_h[ra * AC_BT + col] = acc[0]; res_h[ra * AC_BT + col + 1] = acc[1]; }
                if (ra + 8 < 4 * HS) { res_h[(ra + 8) * AC_BT + col] = acc[2]; res_h[(ra + 8) * AC_BT + col + 1] = acc[3]; }
            }
        }
        __syncthreads();
        // ---- exchange u: element (unit jl, row b) goes to the CTA that attends row b ----
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q) {
            const int i = tid + q * AC_THREADS;
            if (i < HS * AC_BT) {
                const int jl = it_jl[q], b = it_b[q], j = c * HS + jl;
                if (j < H) {
                    const float u = res_h[jl * AC_BT + b] + bu_r[q];
                    float* dst = cluster.map_shared_rank(us, b / AC_RPC);
                    dst[(b % AC_RPC) * H + j] = u;
                    if (a.Upre && it_live[q]) a.Upre[((long)t * B + b0 + b) * H + j] = u;
                }
            }
        }
        cl_arrive();
        cl_wait();
        // ---- P2: attention for my rows: scores -> softmax -> context ----
        for (int pair = warp; pair < AC_RPC * P; pair += AC_WARPS) {
            const int bl = pair / P, p = pair - bl * P;
            const int gb = b0 + c * AC_RPC + bl;
            float s = 0.f;
            if (gb < B) {
                const float* kp = a.Kp + ((long)gb * P + p) * H;
                for (int j = lane; j < H; j += 32) s = fmaf(a.va[j], tanh_fast(kp[j] + us[bl * H + j]), s);
            }
            s = warp_sum(s);
            if (lane == 0) sc[bl * PS + p] = s + bv;
        }
        __syncthreads();
        if (warp < AC_RPC) {
            const int bl = warp, gb = b0 + c * AC_RPC + bl;
            float mx = -INFINITY;
            for (int p = lane; p < P; p += 32) mx = fmaxf(mx, sc[bl * PS + p]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int p = lane; p < P; p += 32) sum += expf(sc[bl * PS + p] - mx);
            sum = warp_sum(sum);
            for (int p = lane; p < P; p += 32) {
                const float al = expf(sc[bl * PS + p] - mx) / sum;
                sc[bl * PS + p] = al;
                if (gb < B) a.attn[((long)gb * T + t) * P + p] = al;
            }
        }
        __syncthreads();
        for (int i = tid; i < AC_RPC * F; i += AC_THREADS) {
            const int bl = i / F, fi = i - bl * F;
            const int brow = c * AC_RPC + bl, gb = b0 + brow;
            float cv = 0.f;
            if (gb < B) {
                const float* fp = a.f + (long)gb * P * F + fi;
                for (int p = 0; p < P; ++p) cv = fmaf(sc[bl * PS + p], fp[(long)p * F], cv);
                a.ctx[((long)t * B + gb) * a.ldctx + fi] = cv;
            }
#pragma unroll
            for (int rk = 0; rk < AC_CS; ++rk) cluster.map_shared_rank(cxs, rk)[brow * F + fi] = cv;
        }
        cl_arrive();
        cl_wait();
        // ---- ctx -> bf16 hi/lo B operand (local), then P3: gi_ctx slice = W_ihc-group . ctx ----
        for (int i = tid; i < AC_BT * F; i += AC_THREADS) {
            const int b = i / F, k = i - b * F;
            const float v = cxs[i];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            cb_hi[b * AC_KP + k] = h;
            cb_lo[b * AC_KP + k] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
        __syncthreads();
        if (group == 1 && lr0 < 3 * HS) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const int n = nt * 8 + (lane >> 2);
                const uint32_t* bh_p = reinterpret_cast<const uint32_t*>(cb_hi + n * AC_KP + (lane & 3) * 2);
                const uint32_t* bl_p = reinterpret_cast<const uint32_t*>(cb_lo + n * AC_KP + (lane & 3) * 2);
#pragma unroll
                for (int kt = 0; kt < AC_KT; ++kt) {
                    if (kt < nkt) {
                        const uint32_t h0 = bh_p[kt * 8], h1 = bh_p[kt * 8 + 4];
                        const uint32_t l0 = bl_p[kt * 8], l1 = bl_p[kt * 8 + 4];
                        mma_bf16(acc, Ahi[kt], h0, h1);
                        mma_bf16(acc, Ahi[kt], l0, l1);
                        mma_bf16(acc, Alo[kt], h0, h1);
                    }
                }
                const int ra = lr0 + (lane >> 2), col = nt * 8 + (lane & 3) * 2;
                if (ra < 3 * HS) { res_c[ra * AC_BT + col] = acc[0]; res_c[ra * AC_BT + col + 1] = acc[1]; }
                if (ra + 8 < 3 * HS) { res_c[(ra + 8) * AC_BT + col] = acc[2]; res_c[(ra + 8) * AC_BT + col + 1] = acc[3]; }
            }
        }
        __syncthreads();
        // ---- P4: gates + state update for my units; broadcast h' (fp32) to every CTA ----
        float o_r[AC_MAXI], o_z[AC_MAXI], o_n[AC_MAXI], o_g[AC_MAXI], o_h[AC_MAXI];
#pragma unroll
        for (int q = 0; q < AC_MAXI; ++q) {
            const int i = tid + q * AC_THREADS;
            o_r[q] = o_z[q] = o_n[q] = o_g[q] = o_h[q] = 0.f;
            if (i < HS * AC_BT) {
                const int jl = it_jl[q], b = it_b[q], j = c * HS + jl;
                if (it_live[q]) {
                    const float ghr = res_h[(HS + jl) * AC_BT + b] + bh[q][0];
                    const float ghz = res_h[(2 * HS + jl) * AC_BT + b] + bh[q][1];
                    const float ghn = res_h[(3 * HS + jl) * AC_BT + b] + bh[q][2];
                    const float gir = giw[q][0] + res_c[jl * AC_BT + b];
                    const float giz = giw[q][1] + res_c[(HS + jl) * AC_BT + b];
                    const float gin = giw[q][2] + res_c[(2 * HS + jl) * AC_BT + b];
                    const float r = sigmoidf_acc(gir + ghr);
                    const float z = sigmoidf_acc(giz + ghz);
                    const float n = tanhf(gin + r * ghn);
                    const float hp = hown[i];
                    o_r[q] = r; o_z[q] = z; o_n[q] = n; o_g[q] = ghn;
                    o_h[q] = (1.f - z) * n + z * hp;
                    hown[i] = o_h[q];
                }
                if (j < H) {
#pragma unroll
                    for (int rk = 0; rk < AC_CS; ++rk) cluster.map_shared_rank(hx, rk)[b * H + j] = o_h[q];
                }
            }
        }
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
        // ---- h' -> bf16 hi/lo B operand (local) for the next step ----
        for (int i = tid; i < AC_BT * H; i += AC_THREADS) {
            const int b = i / H, k = i - b * H;
            const float v = hx[i];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hb_hi[b * AC_KP + k] = h;
            hb_lo[b * AC_KP + k] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
        __syncthreads();
    }
    cluster.sync();   // nobody exits while peers may still address its shared memory
}

static size_t attcl_smem(int H, int F, int P, int HS) {
    const int PS = (P + 3) & ~3;
    return (size_t)4 * AC_BT * AC_KP * 2 +
           ((size_t)AC_BT * H + (size_t)AC_BT * F + (size_t)AC_RPC * H + (size_t)7 * HS * AC_BT + (size_t)HS * AC_BT +
            (size_t)AC_RPC * PS) * sizeof(float);
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// *ok = 1 when the weights-resident attention recurrence supports these sizes (else use caphn_attgru_seq_fwd).
int caphn_attgru_cluster_plan(int H, int F, int P, int* ok) {
    const int HS = (H + AC_CS - 1) / AC_CS;
    const int MTH = (4 * HS + 15) / 16, MTC = (3 * HS + 15) / 16;
    *ok = (H >= 8 && H <= AC_KT * 16 && F >= 1 && F <= AC_KT * 16 && MTH + MTC <= AC_WARPS &&
           HS * AC_BT <= AC_MAXI * AC_THREADS && P >= 1 && attcl_smem(H, F, P, HS) <= 200 * 1024) ? 1 : 0;
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

}  // extern "C"
