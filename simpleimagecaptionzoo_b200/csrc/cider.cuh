// CIDEr-D self-critical reward on the device: the SCST step's reward
//   Utils.get_self_critical_reward (Utils.py:319-367) -> CiderD.compute_score (cider/pyciderevalcap/ciderD/ciderD.py:32-55)
//   -> CiderScorer.compute_cider (ciderD_scorer.py:127-206)
// computed from word ids instead of word strings (an n-gram of ids <-> an n-gram of words under the vocabulary
// bijection).  One CTA per image, one warp per hypothesis (the image's sampled rollouts + its greedy rollout); the
// image's reference captions are turned into tf-idf vectors by the warps in turn and matched from shared memory.
// Integer / hash work plus a handful of fp64 transcendentals per sentence: latency-bound, not a tensor-core shape.
#pragma once
#include <cstdint>

namespace capdec {

constexpr int CIDER_N = 4;               // 1..4-grams (ciderD.py:23)
constexpr int CIDER_MAX_TOKENS = 65;     // longest sentence handled: 4*L - 6 <= CIDER_CAP
constexpr int CIDER_CAP = 256;           // n-gram slots per sentence vector
constexpr int CIDER_REF_SLOTS = 8;       // reference vectors resident in shared memory at a time
constexpr int CIDER_MAX_HYPS = 9;        // MAX_ROWS sampled rollouts + the greedy one

// 64-bit key of an n-gram of word ids (k = its length).  Same function in simpleimagecaptionzoo_b200/scst.py (numpy).
__host__ __device__ __forceinline__ uint64_t cider_mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ uint64_t cider_ngram_key(const int* ids, int k) {
    uint64_t h = static_cast<uint64_t>(k) * 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < k; ++i) h = cider_mix64(h ^ (static_cast<uint64_t>(static_cast<uint32_t>(ids[i])) + 1ull) * 0xBF58476D1CE4E5B9ull);
    return h ? h : 1ull;
}

struct CiderTable {              // open-addressing table of document frequencies (ciderD_scorer.py:80-83), 0 = empty slot
    const uint64_t* keys;
    const float* df;
    uint64_t mask;               // slots - 1 (power of two)
    double log_ref_len;          // np.log(float(ref_len))
};

__device__ __forceinline__ double cider_df(const CiderTable& t, uint64_t key) {
    if (!t.keys) return 0.0;
    uint64_t s = cider_mix64(key) & t.mask;
    for (;;) {
        const uint64_t k = t.keys[s];
        if (k == key) return static_cast<double>(t.df[s]);
        if (k == 0) return 0.0;  // "give word count 1 if it doesn't appear in reference corpus": max(1, 0)
        s = (s + 1) & t.mask;
    }
}

struct CiderVec {                // tf-idf vector of one sentence, in shared memory
    uint64_t key[CIDER_CAP];
    double w[CIDER_CAP];         // tf * (ref_len - log(max(1, df))) for the first occurrence of a key, else unused
    short tf[CIDER_CAP];         // term frequency at the first occurrence, 0 at repeats
    short order[CIDER_CAP];      // n-gram length - 1
    double norm[CIDER_N];
    int count;                   // n-gram slots in use
    int length;                  // "length": sum of bigram term frequencies (ciderD_scorer.py:150-151)
};

// counts2vec (ciderD_scorer.py:128-153) of the sentence ids[0..L) by one warp.
__device__ __forceinline__ void cider_build(CiderVec& v, const int* __restrict__ ids, int L, const CiderTable& tab, int lane) {
    int off[CIDER_N + 1];
    off[0] = 0;
#pragma unroll
    for (int k = 1; k <= CIDER_N; ++k) off[k] = off[k - 1] + (L - k + 1 > 0 ? L - k + 1 : 0);
    const int N = off[CIDER_N];
    for (int i = lane; i < N; i += 32) {
        int k = 1;
        while (i >= off[k]) ++k;
        const int pos = i - off[k - 1];
        v.key[i] = cider_ngram_key(ids + pos, k);
        v.order[i] = static_cast<short>(k - 1);
    }
    __syncwarp();
    double nsq[CIDER_N] = {0.0, 0.0, 0.0, 0.0};
    int len = 0;
    for (int i = lane; i < N; i += 32) {
        const uint64_t key = v.key[i];
        int tf = 0;
        bool first = true;
        for (int j = 0; j < N; ++j) {
            const bool same = v.key[j] == key;
            tf += same;
            first = first && !(same && j < i);
        }
        double w = 0.0;
        if (first) {
            const double df = cider_df(tab, key);
            w = static_cast<double>(tf) * (tab.log_ref_len - log(df > 1.0 ? df : 1.0));
            const int o = v.order[i];
#pragma unroll
            for (int q = 0; q < CIDER_N; ++q)
                if (q == o) nsq[q] += w * w;
            if (o == 1) len += tf;
        }
        v.w[i] = w;
        v.tf[i] = static_cast<short>(first ? tf : 0);
    }
#pragma unroll
    for (int q = 0; q < CIDER_N; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nsq[q] += __shfl_xor_sync(0xffffffffu, nsq[q], o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < CIDER_N; ++q) v.norm[q] = sqrt(nsq[q]);
        v.count = N;
        v.length = len;
    }
    __syncwarp();
}

// gen     [B*n, T] sampled rollouts as sample_rl stores them (<end> and everything after it = 0): the caption is the
//         words up to the last non-zero id, at least one word (Utils.py:338-347)
// greedy  [B, T] greedy rollouts: the words before the first <end> (Utils.py:349-357)
// refs    ref_tok [n_refs_total, ref_ld] word ids, ref_len [n_refs_total], image b owns refs [ref_off[b], ref_off[b+1])
// out     rewards [B*n] = weight * (CIDEr-D(sample) - CIDEr-D(greedy))  (Utils.py:362-363), scores [B, n+1] or null
__global__ void __launch_bounds__(32 * CIDER_MAX_HYPS)
cider_reward_kernel(CiderTable tab, const int* __restrict__ gen, int n, const int* __restrict__ greedy, int T,
                    const int* __restrict__ ref_tok, const int* __restrict__ ref_len, const int* __restrict__ ref_off, int ref_ld,
                    double sigma, double weight, float* __restrict__ rewards, float* __restrict__ scores) {
    extern __shared__ __align__(16) unsigned char cider_smem[];
    CiderVec* hyp = reinterpret_cast<CiderVec*>(cider_smem);  // [n + 1]
    CiderVec* ref = hyp + (n + 1);                            // [CIDER_REF_SLOTS]
    __shared__ double s_score[CIDER_MAX_HYPS];
    const int img = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;  // nw == n + 1
    {
        const bool is_greedy = warp == n;
        const int* ids = is_greedy ? greedy + static_cast<size_t>(img) * T : gen + (static_cast<size_t>(img) * n + warp) * T;
        int L;
        if (is_greedy) {
            L = T;
            for (int i = T - 1; i >= 0; --i)
                if (ids[i] == TOK_END) L = i;
        } else {
            L = 1;
            for (int i = T - 1; i >= 1; --i)
                if (ids[i] != 0) {
                    L = i + 1;
                    break;
                }
        }
        if (L > CIDER_MAX_TOKENS) L = CIDER_MAX_TOKENS;
        cider_build(hyp[warp], ids, L, tab, lane);
    }
    double acc[CIDER_N] = {0.0, 0.0, 0.0, 0.0};  // lane 0: sum over references of the per-order similarities
    const int r0 = ref_off[img], r1 = ref_off[img + 1];
    for (int base = r0; base < r1; base += CIDER_REF_SLOTS) {
        const int nslots = r1 - base < CIDER_REF_SLOTS ? r1 - base : CIDER_REF_SLOTS;
        __syncthreads();  // previous chunk fully consumed
        for (int s = warp; s < nslots; s += nw) {
            int L = ref_len[base + s];
            if (L > CIDER_MAX_TOKENS) L = CIDER_MAX_TOKENS;
            cider_build(ref[s], ref_tok + static_cast<size_t>(base + s) * ref_ld, L, tab, lane);
        }
        __syncthreads();
        const CiderVec& h = hyp[warp];
        for (int s = 0; s < nslots; ++s) {  // sim (ciderD_scorer.py:155-183)
            const CiderVec& r = ref[s];
            double val[CIDER_N] = {0.0, 0.0, 0.0, 0.0};
            for (int i = lane; i < h.count; i += 32) {
                if (h.tf[i] == 0) continue;
                const uint64_t key = h.key[i];
                double wr = 0.0;  // vec_ref[n][ngram] of a defaultdict: 0 when the reference lacks the n-gram
                for (int j = 0; j < r.count; ++j)
                    if (r.tf[j] != 0 && r.key[j] == key) {
                        wr = r.w[j];
                        break;
                    }
                const double wh = h.w[i];
                const double c = (wh < wr ? wh : wr) * wr;
                const int o = h.order[i];
#pragma unroll
                for (int q = 0; q < CIDER_N; ++q)
                    if (q == o) val[q] += c;
            }
#pragma unroll
            for (int q = 0; q < CIDER_N; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val[q] += __shfl_xor_sync(0xffffffffu, val[q], o);
            }
            if (lane == 0) {
                const double delta = static_cast<double>(h.length - r.length);
                const double pen = exp(-(delta * delta) / (2.0 * sigma * sigma));
#pragma unroll
                for (int q = 0; q < CIDER_N; ++q) {
                    double x = val[q];
                    if (h.norm[q] != 0.0 && r.norm[q] != 0.0) x /= h.norm[q] * r.norm[q];
                    acc[q] += x * pen;
                }
            }
        }
    }
    if (lane == 0) {
        const int nrefs = r1 - r0;
        double sc = (acc[0] + acc[1] + acc[2] + acc[3]) / CIDER_N;
        sc = nrefs > 0 ? sc / nrefs * 10.0 : 0.0;
        s_score[warp] = sc;
        if (scores) scores[static_cast<size_t>(img) * (n + 1) + warp] = static_cast<float>(sc);
    }
    __syncthreads();
    if (lane == 0 && warp < n) rewards[static_cast<size_t>(img) * n + warp] = static_cast<float>(weight * (s_score[warp] - s_score[n]));
}

}  // namespace capdec
