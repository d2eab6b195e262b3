// Persistent attention-GRU recurrence ("Variant B": AttentionGru + BahdanauAttention), forward and BPTT.
//
// Replaces, per time step, the torch calls of reference models/decoderlstm.py:97-100 (attention + GRUCell on
// cat[word_embed, context]) and models/attention.py:33-45 (U_a h, tanh, v_a, softmax over the P image positions,
// weighted sum), and their autograd backward.  Hoisted out of the loop (they do not
// depend on h): the keys K = W_a f + b_a (the reference recomputes them every step, attention.py:34) and the word half
// of the input projection GIw = x_t W_ih[:, :E]^T + b_ih.
//
// One CTA owns BT batch rows for all steps [t0, t1) and keeps h on-chip; the generated W_hh / W_ih[:, E:] and U_a
// (transposed / padded to 16-byte rows on the host side of the C-ABI) stream from L2 every step, K and f likewise.
// Forward per step:  u = U_a h + b_u;  s_p = v_a . tanh(K_p + u) + b_v;  alpha = softmax_p(s);  ctx = sum_p alpha_p f_p;
//                    gi = GIw_t + W_ih[:,E:] ctx;  gh = W_hh h + b_hh;  r,z,n gates;  h' = (1-z) n + z h.
// Greedy decode calls the same kernel one step at a time (t1 = t0 + 1) between vocabulary projections.
#include "seq_common.cuh"

namespace caphn {

struct AttFwdArgs {
    const float* Kp;     // [B,P,H]  keys  W_a f + b_a
    const float* f;      // [B,P,F]
    const float* GIw;    // [T,B,3H] word half of the input projection (+ b_ih)
    const float* UaT;    // [H, ldh]
    const float* bu;     // [H]
    const float* va;     // [H]
    const float* bv;     // [1]
    const float* WihcT;  // [F, ld3]
    const float* WhhT;   // [H, ld3]
    const float* bhh;    // [3H]
    float* Hall;         // [T+1,B,H]
    float* Hbm;          // [B,T,H] or null
    float* attn;         // [B,T,P]
    float* ctx;          // ctx[t,b,:] at ctx + (t*B+b)*ldctx      (the E.. columns of the [T*B, E+F] input matrix)
    long ldctx;
    float* Upre;         // [T,B,H] or null
    float* R; float* Z; float* Nn; float* GHN;  // [T,B,H] or null
    int B, T, P, H, F, ldh, ld3, t0, t1;
};

__global__ void __launch_bounds__(AT_THREADS) attgru_seq_fwd_kernel(const AttFwdArgs a) {
    constexpr int BT = AT_BT;
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, F = a.F, P = a.P, B = a.B, T = a.T, ldh = a.ldh, ld3 = a.ld3;
    const int PS = (P + 3) & ~3;
    const int CQTh = at_cqt(ldh), CQT3 = at_cqt(ld3);
    const int KGh = AT_THREADS / CQTh, KG3 = AT_THREADS / CQT3;
    float* hs = smem;                               // [H][BT]
    float* cx = hs + H * BT;                        // [F][BT]
    float* us = cx + F * BT;                        // [BT][H]
    float* sc = us + BT * H;                        // [BT][PS]
    float* part_u = sc + BT * PS;                   // [KGh][BT][ldh]
    float* part_gh = part_u + KGh * BT * ldh;       // [KG3][BT][ld3]
    float* part_gi = part_gh + KG3 * BT * ld3;      // [KG3][BT][ld3]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * BT;
    const int H3 = 3 * H;
    const float bv = a.bv[0];

    for (int i = tid; i < H * BT; i += AT_THREADS) {
        const int k = i / BT, b = i - k * BT;
        hs[i] = (b0 + b < B) ? a.Hall[((long)a.t0 * B + b0 + b) * H + k] : 0.f;
    }
    __syncthreads();

    for (int t = a.t0; t < a.t1; ++t) {
        block_matvec<BT, false>(a.UaT, ldh, H, hs, part_u, CQTh, tid);
        block_matvec<BT, false>(a.WhhT, ld3, H, hs, part_gh, CQT3, tid);
        __syncthreads();
        for (int i = tid; i < BT * H; i += AT_THREADS) {
            const int b = i / H, j = i - b * H;
            const float u = a.bu[j] + part_sum(part_u, KGh, BT, ldh, b, j);
            us[i] = u;
            if (a.Upre && b0 + b < B) a.Upre[((long)t * B + b0 + b) * H + j] = u;
        }
        __syncthreads();
        // scores: one warp per (row, group of 4 positions); v_a and u stay in registers across the 4 positions
        {
            const int PG = (P + 3) >> 2;
            for (int task = warp; task < BT * PG; task += AT_WARPS) {
                const int b = task / PG, p0 = (task - b * PG) * 4;
                const int gb = b0 + b;
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
                if (gb < B) {
                    const float* kp = a.Kp + ((long)gb * P + p0) * H;
                    for (int j = lane; j < H; j += 32) {
                        const float vj = a.va[j], uj = us[b * H + j];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (p0 + q < P) s4[q] = fmaf(vj, tanh_fast(kp[(long)q * H + j] + uj), s4[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float s = warp_sum(s4[q]);
                    if (lane == 0 && p0 + q < P) sc[b * PS + p0 + q] = s + bv;
                }
            }
        }
        __syncthreads();
        if (warp < BT) {
            const int b = warp, gb = b0 + b;
            float mx = -INFINITY;
            for (int p = lane; p < P; p += 32) mx = fmaxf(mx, sc[b * PS + p]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int p = lane; p < P; p += 32) sum += expf(sc[b * PS + p] - mx);
            sum = warp_sum(sum);
            for (int p = lane; p < P; p += 32) {
                const float al = expf(sc[b * PS + p] - mx) / sum;
                sc[b * PS + p] = al;
                if (gb < B) a.attn[((long)gb * T + t) * P + p] = al;
            }
        }
        __syncthreads();
        for (int i = tid; i < BT * F; i += AT_THREADS) {
            const int b = i / F, fi = i - b * F;
            const int gb = b0 + b;
            float c = 0.f;
            if (gb < B) {
                const float* fp = a.f + (long)gb * P * F + fi;
                for (int p = 0; p < P; ++p) c = fmaf(sc[b * PS + p], fp[(long)p * F], c);
                a.ctx[((long)t * B + gb) * a.ldctx + fi] = c;
            }
            cx[fi * BT + b] = c;
        }
        __syncthreads();
        block_matvec<BT, false>(a.WihcT, ld3, F, cx, part_gi, CQT3, tid);
        __syncthreads();
        for (int i = tid; i < BT * H; i += AT_THREADS) {
            const int b = i / H, j = i - b * H;
            const int gb = b0 + b;
            if (gb < B) {
                const float ghr = a.bhh[j] + part_sum(part_gh, KG3, BT, ld3, b, j);
                const float ghz = a.bhh[H + j] + part_sum(part_gh, KG3, BT, ld3, b, H + j);
                const float ghn = a.bhh[2 * H + j] + part_sum(part_gh, KG3, BT, ld3, b, 2 * H + j);
                const float* gw = a.GIw + ((long)t * B + gb) * H3;
                const float gir = gw[j] + part_sum(part_gi, KG3, BT, ld3, b, j);
                const float giz = gw[H + j] + part_sum(part_gi, KG3, BT, ld3, b, H + j);
                const float gin = gw[2 * H + j] + part_sum(part_gi, KG3, BT, ld3, b, 2 * H + j);
                const float r = sigmoidf_acc(gir + ghr);
                const float z = sigmoidf_acc(giz + ghz);
                const float n = tanhf(gin + r * ghn);
                const float hp = hs[j * BT + b];
                const float hn = (1.f - z) * n + z * hp;
                hs[j * BT + b] = hn;
                const long o = ((long)t * B + gb) * H + j;
                a.Hall[o + (long)B * H] = hn;
                if (a.Hbm) a.Hbm[((long)gb * T + t) * H + j] = hn;
                if (a.R) { a.R[o] = r; a.Z[o] = z; a.Nn[o] = n; a.GHN[o] = ghn; }
            }
        }
        __syncthreads();
    }
}

struct AttBwdArgs {
    const float* dHbm;   // [B,T,H]
    const float* dattn;  // [B,T,P] or null
    const float* Kp; const float* f;          // [B,P,H], [B,P,F]
    const float* attn;   // [B,T,P]
    const float* Upre; const float* R; const float* Z; const float* Nn; const float* GHN;  // [T,B,H]
    const float* Hall;   // [T+1,B,H]
    const float* Ua;     // [H, ldh]    (row j, col k)
    const float* va;     // [H]
    const float* Wihc;   // [3H, ldf]   (= W_ih[:, E:])
    const float* Whh;    // [3H, ldh]
    float* dGI; float* dGH;   // [T,B,3H]
    float* dU;           // [T,B,H]
    float* dCTX;         // [T,B,F]
    float* dK;           // [B,P,H]   accumulated over t (zero-initialised by the caller)
    float* dva;          // [H]       atomically accumulated (zero-initialised)
    float* dbv;          // [1]       atomically accumulated (zero-initialised)
    float* dh0;          // [B,H]
    int B, T, P, H, F, ldh, ldf;
};

__global__ void __launch_bounds__(AT_THREADS) attgru_seq_bwd_kernel(const AttBwdArgs a) {
    constexpr int BT = AT_BT;
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, F = a.F, P = a.P, B = a.B, T = a.T, ldh = a.ldh, ldf = a.ldf;
    const int H3 = 3 * H;
    const int PS = (P + 3) & ~3;
    const int CQTh = at_cqt(ldh), CQTf = at_cqt(ldf);
    const int KGh = AT_THREADS / CQTh, KGf = AT_THREADS / CQTf;
    float* dgi = smem;                          // [3H][BT]
    float* dgh = dgi + H3 * BT;                 // [3H][BT]
    float* dus = dgh + H3 * BT;                 // [H][BT]
    float* dhd = dus + H * BT;                  // [BT][H]
    float* us = dhd + BT * H;                   // [BT][H]
    float* dcx = us + BT * H;                   // [BT][F]
    float* al = dcx + BT * F;                   // [BT][PS]
    float* dal = al + BT * PS;                  // [BT][PS]
    float* part_dh = dal + BT * PS;             // [KGh][BT][ldh]
    float* part_dc = part_dh + KGh * BT * ldh;  // [KGf][BT][ldf]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * BT;
    const int NJC = (H + 31) >> 5;
    const int ntasks = BT * NJC;

    float dva_acc[AT_MAXSLOT];
#pragma unroll
    for (int s = 0; s < AT_MAXSLOT; ++s) dva_acc[s] = 0.f;
    float dbv_acc = 0.f;

    for (int i = tid; i < BT * H; i += AT_THREADS) dhd[i] = 0.f;
    for (int i = tid; i < KGh * BT * ldh; i += AT_THREADS) part_dh[i] = 0.f;
    __syncthreads();

    for (int t = T - 1; t >= 0; --t) {
        // ---- gate gradients ----
        for (int i = tid; i < BT * H; i += AT_THREADS) {
            const int b = i / H, j = i - b * H;
            const int gb = b0 + b;
            float dar = 0.f, daz = 0.f, dan = 0.f, danr = 0.f, keep = 0.f, u = 0.f;
            if (gb < B) {
                const float dht = dhd[i] + a.dHbm[((long)gb * T + t) * H + j] + part_sum(part_dh, KGh, BT, ldh, b, j);
                const long o = ((long)t * B + gb) * H + j;
                const float r = a.R[o], z = a.Z[o], n = a.Nn[o], ghn = a.GHN[o];
                const float hp = a.Hall[o];
                u = a.Upre[o];
                const float dn = dht * (1.f - z);
                const float dz = dht * (hp - n);
                dan = dn * (1.f - n * n);
                dar = dan * ghn * r * (1.f - r);
                daz = dz * z * (1.f - z);
                danr = dan * r;
                keep = dht * z;
                float* gi = a.dGI + ((long)t * B + gb) * H3;
                float* gh = a.dGH + ((long)t * B + gb) * H3;
                gi[j] = dar; gi[H + j] = daz; gi[2 * H + j] = dan;
                gh[j] = dar; gh[H + j] = daz; gh[2 * H + j] = danr;
            }
            dhd[i] = keep;
            us[i] = u;
            dgi[j * BT + b] = dar; dgi[(H + j) * BT + b] = daz; dgi[(2 * H + j) * BT + b] = dan;
            dgh[j * BT + b] = dar; dgh[(H + j) * BT + b] = daz; dgh[(2 * H + j) * BT + b] = danr;
        }
        for (int i = tid; i < BT * P; i += AT_THREADS) {
            const int b = i / P, p = i - b * P;
            al[b * PS + p] = (b0 + b < B) ? a.attn[((long)(b0 + b) * T + t) * P + p] : 0.f;
        }
        __syncthreads();
        // ---- dctx = dgi W_ih[:,E:]   and   dh += dgh W_hh ----
        block_matvec<BT, false>(a.Wihc, ldf, H3, dgi, part_dc, CQTf, tid);
        block_matvec<BT, false>(a.Whh, ldh, H3, dgh, part_dh, CQTh, tid);
        __syncthreads();
        for (int i = tid; i < BT * F; i += AT_THREADS) {
            const int b = i / F, fi = i - b * F;
            const float d = part_sum(part_dc, KGf, BT, ldf, b, fi);
            dcx[i] = d;
            if (b0 + b < B) a.dCTX[((long)t * B + b0 + b) * F + fi] = d;
        }
        __syncthreads();
        // ---- d alpha_p = <dctx, f_p> (+ external gradient of the returned attention weights) ----
        for (int pair = warp; pair < BT * P; pair += AT_WARPS) {
            const int b = pair / P, p = pair - b * P;
            const int gb = b0 + b;
            float s = 0.f;
            if (gb < B) {
                const float* fp = a.f + ((long)gb * P + p) * F;
                for (int fi = lane; fi < F; fi += 32) s = fmaf(dcx[b * F + fi], fp[fi], s);
            }
            s = warp_sum(s);
            if (lane == 0) {
                if (a.dattn && gb < B) s += a.dattn[((long)gb * T + t) * P + p];
                dal[b * PS + p] = s;
            }
        }
        __syncthreads();
        // ---- softmax backward: ds_p = alpha_p (dalpha_p - sum_q alpha_q dalpha_q) ----
        if (warp < BT) {
            const int b = warp;
            float c = 0.f;
            for (int p = lane; p < P; p += 32) c = fmaf(al[b * PS + p], dal[b * PS + p], c);
            c = warp_sum(c);
            for (int p = lane; p < P; p += 32) {
                const float ds = al[b * PS + p] * (dal[b * PS + p] - c);
                dal[b * PS + p] = ds;
                dbv_acc += ds;
            }
        }
        __syncthreads();
        // ---- score backward: q = tanh(K_p + u); dv_a, dK, du ----
#pragma unroll
        for (int slot = 0; slot < AT_MAXSLOT; ++slot) {
            const int task = warp + slot * AT_WARPS;
            if (task < ntasks) {
                const int b = task / NJC, jc = task - b * NJC;
                const int j = jc * 32 + lane;
                const int gb = b0 + b;
                float du = 0.f;
                if (j < H && gb < B) {
                    const float uj = us[b * H + j], vj = a.va[j];
                    const float* kp = a.Kp + (long)gb * P * H + j;
                    float* dkp = a.dK + (long)gb * P * H + j;
                    for (int p = 0; p < P; ++p) {
                        const float q = tanh_fast(kp[(long)p * H] + uj);
                        const float ds = dal[b * PS + p];
                        dva_acc[slot] = fmaf(ds, q, dva_acc[slot]);
                        const float dpre = ds * vj * (1.f - q * q);
                        dkp[(long)p * H] += dpre;
                        du += dpre;
                    }
                    a.dU[((long)t * B + gb) * H + j] = du;
                }
                if (j < H) dus[j * BT + b] = du;
            }
        }
        __syncthreads();
        // ---- dh += du U_a ----
        block_matvec<BT, true>(a.Ua, ldh, H, dus, part_dh, CQTh, tid);
        __syncthreads();
    }
    for (int i = tid; i < BT * H; i += AT_THREADS) {
        const int b = i / H, j = i - b * H;
        if (b0 + b < B) a.dh0[(long)(b0 + b) * H + j] = dhd[i] + part_sum(part_dh, KGh, BT, ldh, b, j);
    }
#pragma unroll
    for (int slot = 0; slot < AT_MAXSLOT; ++slot) {
        const int task = warp + slot * AT_WARPS;
        if (task < ntasks) {
            const int b = task / NJC, jc = task - b * NJC;
            const int j = jc * 32 + lane;
            if (j < H && dva_acc[slot] != 0.f) atomicAdd(a.dva + j, dva_acc[slot]);
        }
    }
    dbv_acc = warp_sum(dbv_acc);
    if (lane == 0 && warp < BT && dbv_acc != 0.f) atomicAdd(a.dbv, dbv_acc);
}

// df[b,p,fi] (+)= sum_t alpha[b,t,p] * dCTX[t,b,fi]     (context backward, deferred out of the BPTT loop)
__global__ void __launch_bounds__(256) attn_df_kernel(const float* __restrict__ attn, const float* __restrict__ dCTX,
                                                      float* __restrict__ df, int B, int T, int P, int F) {
    extern __shared__ float sm[];
    float* al = sm;           // [T][P]
    float* dc = sm + T * P;   // [T][F]
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < T * P; i += 256) al[i] = attn[(long)b * T * P + i];
    for (int i = threadIdx.x; i < T * F; i += 256) {
        const int t = i / F, fi = i - t * F;
        dc[i] = dCTX[((long)t * B + b) * F + fi];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * F; i += 256) {
        const int p = i / F, fi = i - p * F;
        float s = 0.f;
        for (int t = 0; t < T; ++t) s = fmaf(al[t * P + p], dc[t * F + fi], s);
        df[(long)b * P * F + i] += s;
    }
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// Steps [t0, t1) of the attention-GRU recurrence (see file header).  Hall[t0] must hold h_{t0-1} (h0 for t0 = 0).
// ctx rows are written at ctx + (t*B+b)*ldctx.  Upre/R/Z/Nn/GHN (all or none) are saved for the backward.
int caphn_attgru_seq_fwd(const float* Kp, const float* f, const float* GIw, const float* UaT, const float* bu,
                         const float* va, const float* bv, const float* WihcT, const float* WhhT, const float* bhh,
                         float* Hall, float* Hbm, float* attn, float* ctx, long ldctx, float* Upre, float* R, float* Z,
                         float* Nn, float* GHN, int B, int T, int P, int H, int F, int ldh, int ld3, int t0, int t1,
                         void* stream) {
    if (B <= 0 || T <= 0 || P <= 0 || H <= 0 || F <= 0 || (ldh & 3) || (ld3 & 3) || ldh < H || ld3 < 3 * H ||
        t0 < 0 || t1 > T || t0 >= t1 || ((uintptr_t)UaT & 15) || ((uintptr_t)WihcT & 15) || ((uintptr_t)WhhT & 15))
        return CAPHN_EINVAL;
    if (R && !(Z && Nn && GHN && Upre)) return CAPHN_EINVAL;
    AttFwdArgs a{Kp, f, GIw, UaT, bu, va, bv, WihcT, WhhT, bhh, Hall, Hbm, attn, ctx, ldctx, Upre, R, Z, Nn, GHN,
                 B, T, P, H, F, ldh, ld3, t0, t1};
    const int PS = (P + 3) & ~3;
    const int KGh = AT_THREADS / at_cqt(ldh), KG3 = AT_THREADS / at_cqt(ld3);
    const size_t smem = ((size_t)H * AT_BT + (size_t)F * AT_BT + (size_t)AT_BT * H + (size_t)AT_BT * PS +
                         (size_t)KGh * AT_BT * ldh + 2 * (size_t)KG3 * AT_BT * ld3) * sizeof(float);
    if (smem > 227 * 1024) return CAPHN_EINVAL;
    CAPHN_CHECK(cudaFuncSetAttribute(attgru_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attgru_seq_fwd_kernel<<<ceil_div(B, AT_BT), AT_THREADS, smem, (cudaStream_t)stream>>>(a);
    CAPHN_RETURN_LAST();
}

// BPTT of caphn_attgru_seq_fwd over all T steps.  dK, dva, dbv must be zero-initialised by the caller.
int caphn_attgru_seq_bwd(const float* dHbm, const float* dattn, const float* Kp, const float* f, const float* attn,
                         const float* Upre, const float* R, const float* Z, const float* Nn, const float* GHN,
                         const float* Hall, const float* Ua, const float* va, const float* Wihc, const float* Whh,
                         float* dGI, float* dGH, float* dU, float* dCTX, float* dK, float* dva, float* dbv, float* dh0,
                         int B, int T, int P, int H, int F, int ldh, int ldf, void* stream) {
    if (B <= 0 || T <= 0 || P <= 0 || H <= 0 || F <= 0 || (ldh & 3) || (ldf & 3) || ldh < H || ldf < F ||
        ((uintptr_t)Ua & 15) || ((uintptr_t)Wihc & 15) || ((uintptr_t)Whh & 15))
        return CAPHN_EINVAL;
    if (AT_BT * ((H + 31) / 32) > AT_MAXSLOT * AT_WARPS) return CAPHN_EINVAL;
    AttBwdArgs a{dHbm, dattn, Kp, f, attn, Upre, R, Z, Nn, GHN, Hall, Ua, va, Wihc, Whh, dGI, dGH, dU, dCTX, dK, dva,
                 dbv, dh0, B, T, P, H, F, ldh, ldf};
    const int PS = (P + 3) & ~3;
    const int KGh = AT_THREADS / at_cqt(ldh), KGf = AT_THREADS / at_cqt(ldf);
    const size_t smem = (2 * (size_t)3 * H * AT_BT + (size_t)H * AT_BT + 2 * (size_t)AT_BT * H + (size_t)AT_BT * F +
                         2 * (size_t)AT_BT * PS + (size_t)KGh * AT_BT * ldh + (size_t)KGf * AT_BT * ldf) * sizeof(float);
    if (smem > 227 * 1024) return CAPHN_EINVAL;
    CAPHN_CHECK(cudaFuncSetAttribute(attgru_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attgru_seq_bwd_kernel<<<ceil_div(B, AT_BT), AT_THREADS, smem, (cudaStream_t)stream>>>(a);
    CAPHN_RETURN_LAST();
}

// df[b,p,:] += sum_t attn[b,t,p] * dCTX[t,b,:]
int caphn_attn_df(const float* attn, const float* dCTX, float* df, int B, int T, int P, int F, void* stream) {
    if (B <= 0 || T <= 0 || P <= 0 || F <= 0) return CAPHN_EINVAL;
    const size_t smem = ((size_t)T * P + (size_t)T * F) * sizeof(float);
    if (smem > 200 * 1024) return CAPHN_EINVAL;
    CAPHN_CHECK(cudaFuncSetAttribute(attn_df_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_df_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(attn, dCTX, df, B, T, P, F);
    CAPHN_RETURN_LAST();
}

}  // extern "C"
