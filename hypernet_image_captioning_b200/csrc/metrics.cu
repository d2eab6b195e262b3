// On-device caption post-processing for the metric step of the training loops.
//
// The reference turns every predicted caption into text with one `.item()` per token (utils.py:161-174 cap_to_text,
// called twice per caption from metric_score utils.py:229-262, inside every training_step cc_train_hypernet.py:154):
// B*T device->host synchronisations per step.  Here the token-level part runs on the device and the batch leaves in
// one copy:
//   caption_compact : cap_to_text's filter on token ids -- drop <pad>/<s>, stop at the first </s> -- for all captions
//   bleu_counts     : clipped n-gram matches / candidates per order + corpus lengths, i.e. the sufficient statistics of
//                     the corpus BLEU-1..4 that metric_score requests (utils.py:250-258, `datasets` "bleu" metric ==
//                     compute_bleu of tensorflow/nmt: Counter intersection per caption, summed over the corpus).
// Token-id n-grams equal word n-grams because Vocab.i2w is injective (build_vocab.py:18-24).
#include "common.cuh"

namespace caphn {

// One thread per caption (T <= a few dozen tokens; the batch gives the parallelism).
__global__ void caption_compact_kernel(const long long* __restrict__ tok, long ldt, int B, int T, long long pad,
                                       long long start, long long end, long long* __restrict__ out,
                                       int* __restrict__ len) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long long* x = tok + (long)b * ldt;
    long long* y = out + (long)b * T;
    int n = 0;
    bool stopped = false;
    for (int t = 0; t < T; ++t) {
        const long long w = x[t];
        if (!stopped) {
            if (w == end) stopped = true;
            else if (w != pad && w != start) y[n++] = w;
        }
    }
    for (int t = n; t < T; ++t) y[t] = pad;
    len[b] = n;
}

constexpr int BLEU_THREADS = 128;
constexpr int BLEU_MAX_ORDER = 4;

// One CTA per caption.  counts[0..max_order) += clipped matches per order, counts[max_order..2*max_order) += candidate
// n-grams per order, counts[2*max_order] += hypothesis length, counts[2*max_order+1] += reference length.
__global__ void __launch_bounds__(BLEU_THREADS) bleu_counts_kernel(
    const long long* __restrict__ hyp, const int* __restrict__ hyp_len, int Th, const long long* __restrict__ ref,
    const int* __restrict__ ref_len, int Tr, int max_order, unsigned long long* __restrict__ counts) {
    extern __shared__ long long toks[];   // [Th + Tr]
    __shared__ unsigned int match[BLEU_MAX_ORDER];
    const int b = blockIdx.x;
    long long* h = toks;
    long long* r = toks + Th;
    const int hl = min(hyp_len[b], Th), rl = min(ref_len[b], Tr);
    for (int i = threadIdx.x; i < hl; i += BLEU_THREADS) h[i] = hyp[(long)b * Th + i];
    for (int i = threadIdx.x; i < rl; i += BLEU_THREADS) r[i] = ref[(long)b * Tr + i];
    if (threadIdx.x < BLEU_MAX_ORDER) match[threadIdx.x] = 0u;
    __syncthreads();
    // work item = (order n, start i) of a hypothesis n-gram; it contributes min(#in hyp, #in ref) once, at its first
    // occurrence in the hypothesis (Counter & Counter, summed over distinct n-grams)
    for (int item = threadIdx.x; item < max_order * hl; item += BLEU_THREADS) {
        const int n = item / hl + 1, i = item - (n - 1) * hl;
        if (i + n > hl) continue;
        bool first = true;
        for (int j = 0; j < i && first; ++j) {
            bool same = true;
            for (int k = 0; k < n; ++k) same = same && (h[j + k] == h[i + k]);
            if (same) first = false;
        }
        if (!first) continue;
        unsigned int ch = 0, cr = 0;
        for (int j = i; j + n <= hl; ++j) {
            bool same = true;
            for (int k = 0; k < n; ++k) same = same && (h[j + k] == h[i + k]);
            ch += same ? 1u : 0u;
        }
        for (int j = 0; j + n <= rl; ++j) {
            bool same = true;
            for (int k = 0; k < n; ++k) same = same && (r[j + k] == h[i + k]);
            cr += same ? 1u : 0u;
        }
        const unsigned int m = ch < cr ? ch : cr;
        if (m) atomicAdd(&match[n - 1], m);
    }
    __syncthreads();
    if (threadIdx.x < max_order) {
        const int n = threadIdx.x + 1;
        if (match[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)match[threadIdx.x]);
        if (hl - n + 1 > 0) atomicAdd(counts + max_order + threadIdx.x, (unsigned long long)(hl - n + 1));
    }
    if (threadIdx.x == 0) {
        if (hl) atomicAdd(counts + 2 * max_order, (unsigned long long)hl);
        if (rl) atomicAdd(counts + 2 * max_order + 1, (unsigned long long)rl);
    }
}

}  // namespace caphn

using namespace caphn;

extern "C" {

// cap_to_text (utils.py:161-174) / cap_to_text_gt (:177-190) on token ids: out[b, :len[b]] = the tokens of tok[b, :]
// that are neither pad nor start, up to (excluding) the first end token; out[b, len[b]:] = pad.
int caphn_caption_compact(const long long* tok, long ldt, int B, int T, long long pad, long long start, long long end,
                          long long* out, int* len, void* stream) {
    if (B <= 0 || T <= 0 || ldt < T) return CAPHN_EINVAL;
    caption_compact_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(tok, ldt, B, T, pad, start, end, out, len);
    CAPHN_RETURN_LAST();
}

// Corpus-BLEU sufficient statistics of B (hypothesis, single reference) pairs of compacted captions, ADDED into
// counts[2*max_order + 2] (uint64; the caller zeroes it): matches per order, candidates per order, hyp length, ref length.
int caphn_bleu_counts(const long long* hyp, const int* hyp_len, int Th, const long long* ref, const int* ref_len, int Tr,
                      int B, int max_order, unsigned long long* counts, void* stream) {
    if (B <= 0 || Th <= 0 || Tr <= 0 || max_order < 1 || max_order > BLEU_MAX_ORDER) return CAPHN_EINVAL;
    const size_t smem = (size_t)(Th + Tr) * sizeof(long long);
    if (smem > 48 * 1024) return CAPHN_EINVAL;
    bleu_counts_kernel<<<(unsigned)B, BLEU_THREADS, smem, (cudaStream_t)stream>>>(hyp, hyp_len, Th, ref, ref_len, Tr,
                                                                                 max_order, counts);
    CAPHN_RETURN_LAST();
}

}  // extern "C"
