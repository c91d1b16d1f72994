// Persistent warp-specialised tcgen05 GEMM for sm_100a with fused decode-step epilogues.
//
//   D[M,N] = A[M,K] * B[N,K]^T      A, B fp16 K-major (row-major, K contiguous), fp32 accumulate in TMEM
//
// Roles (320 threads, one CTA per SM):  warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer,
// warps 2-9 = epilogue (thread == output row == TMEM lane; two warps share a 32-lane quarter and split the columns).
// Pipelines: smem ring full/empty (TMA <-> MMA) and a 2-deep TMEM accumulator ring (MMA <-> epilogue), so
// the epilogue of tile i overlaps the MMAs of tile i+1.
//
// "Split" math mode (fp32-grade): every operand is stored as fp16 hi | fp16 lo (lo at +lo_off columns) and
// the K loop runs three passes hi*hi + hi*lo + lo*hi into the same fp32 accumulator.
//
// Epilogues (what the reference computes after each Linear / LSTMCell):
//   EPI_STORE  bias add -> fp32 (+fp16) store                      (enc_att / dec_att / Q / K,V projections)
//   EPI_LSTM   gate-interleaved columns -> LSTMCell pointwise      (BUTD_Model.py:265,268; NIC_Model.py:173)
//   EPI_GLU    (a,gate)-interleaved columns -> a*sigmoid(gate)     (AoA_Model.py:118)
//   EPI_TOPK   bias add -> per-row running (max, sum-exp) + top-K  (predict + log_softmax + topk, :270-276)
//   EPI_SAMPLE bias add -> per-row (max, sum-exp) + Gumbel-max     (sample / sample_rl, :183, :221-224)
#pragma once
#include "ptx.cuh"

namespace capdec {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 fp16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                      // 2 warps per TMEM lane quarter, each owns half of the tile's columns
constexpr int EPI_SPLIT = EPI_WARPS / 4;
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int LOGIT_EPI_WARPS = 12;               // pair kernel, sampling epilogue (see gemm2_kernel)
constexpr int LOGIT_EPI_SPLIT_MAX = LOGIT_EPI_WARPS / 4;

enum EpiKind { EPI_STORE = 0, EPI_LSTM = 1, EPI_GLU = 2, EPI_TOPK = 3, EPI_SAMPLE = 4 };

struct EpiParams {
    // common
    const float* bias;      // [N] (in the packed column order) or null
    // STORE / GLU
    float* out32;           // fp32 output or null
    int ld32;
    __half* out16;          // fp16 output (hi at col, lo at col+lo16 when lo16 > 0) or null
    int ld16;
    int lo16;
    int relu;               // STORE: max(x, 0) after the bias (img_feats_porjection, AoA_Model.py:661-665)
    const float* row_keep;  // STORE: per-row {0,1} flags or null; rows flagged 0 are stored as zeros (pack_wrapper's padding)
    const float* resid;     // GLU: residual added to the gated output (SublayerConnection, AoA_Model.py:38); may alias out32
    int ld_resid;
    // LSTM
    const float* rowadd;    // additive per-row-group term [M/rows_per_group, N] or null (hoisted, step-invariant part)
    int rowadd_ld;
    int rows_per_group;
    const float* gather;    // additive per-row term gathered by index: gather[gather_idx[row], N] (embedding-table
    const int* gather_idx;  //   contribution of the fed-back word: precomputed embed(word) * W_ih[:, emb]^T), or null
    int gather_ld;
    const float* c_in;      // [.., H] previous cell state (null = zeros)
    float* c_out;           // [M, H]
    int ldc;
    const int* parent;      // c_in row indirection (beam reorder folded into the read) or null
    float* h32;             // optional fp32 copy of h
    int ldh32;
    // TOPK / SAMPLE
    float* part;            // [M, n_tiles, PS] per-(row, N-tile) partials
    int n_tiles;
    uint32_t seed;          // SAMPLE: counter-based Gumbel noise (0 noise when use_noise == 0)
    const uint32_t* seed_ptr;  // SAMPLE: when non-null the seed is read from device memory (captured rollouts)
    int step;
    int use_noise;
    int scst_n;             // SAMPLE, > 0: rows come in groups of scst_n sampled rollouts + ONE greedy rollout per image (the
                            //   two rollouts of an SCST step in one pass); sample rows keep the noise stream of row img*scst_n+j
    const int* forced;      // SAMPLE: teacher forcing -- the word of row r at this step is forced[r * forced_ld + step]
    int forced_ld;          //   (capdec_score); null = draw / arg-max
};

struct GemmParams {
    int M, N;
    int k_blocks;           // K / 64 (per pass)
    int passes;             // 1 (fp16) or 3 (split)
    int a_lo_off, b_lo_off; // column offset of the lo halves (elements)
    int num_m_blocks, num_n_blocks;
    int runs;               // > 0: (row block, vocabulary-tile run) work items, see TileIter
    int m_group;            // > 0: output tiles are walked in groups of m_group row blocks, see TileIter
    EpiParams epi;
};

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N >= 256) ? 4 : 6;
    static constexpr int TMEM_COLS = (2 * BLOCK_N >= 512) ? 512 : ((2 * BLOCK_N >= 256) ? 256 : 128);
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

__host__ __device__ constexpr int topk_part_stride(int ktop) { return 2 + 2 * ktop; }
constexpr int SAMPLE_PART_STRIDE = 5;  // max, sumexp, best perturbed, best index, best raw logit

// ------------------------------------------------------------------------------------------------ math helpers
// sigmoid / tanh from ex2.approx + fast division: absolute error ~1e-7 (fp32 noise level), ~6 instructions each.
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
// MUFU.EX2 / MUFU.LG2 alone.  exp2f / __expf / __log2f wrap the MUFU in a range test and two multiplies for subnormal
// results / arguments (3 of the ~10 instructions per logit in the log-softmax epilogues); a flushed subnormal cannot change any
// of the sums formed here (each holds a term >= 1, or is 1 + e), and the logarithms only see normal arguments.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float exp_ftz(float x) { return ex2_ftz(x * LOG2E); }
__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdividef(1.0f, 1.0f + exp_ftz(-x)); }
__device__ __forceinline__ float tanhf_acc(float x) { return 1.0f - __fdividef(2.0f, 1.0f + exp_ftz(2.0f * x)); }

__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
// Gumbel(0,1) noise g = -log(-log(u)) on the counter-based uniforms of oracle/capdec_oracle.py:gumbel_noise, computed
// with two MUFU.LG2 instead of two libm logf calls (the draw runs once per vocabulary entry per row-step inside the
// logit GEMM's epilogue).  lg2.approx has ~2^-22 ABSOLUTE error, which is useless for -log(u) when u -> 1 (exactly the
// draws that win the arg-max), so that range uses the series of -log(1 - t), t = 1 - u (exact by Sterbenz):
// |g - exact| <= ~1e-5 everywhere, far below the 1e-3 tie tolerance of the sampling parity tests.
__device__ __forceinline__ uint32_t gumbel_bits(uint32_t row_step_hash, uint32_t v) {
    return fmix32(row_step_hash ^ (v * 0xC2B2AE3Du));
}
__device__ __forceinline__ float gumbel_from_bits(uint32_t x) {
    const float u = (static_cast<float>(x >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
    const float t = 1.0f - u;
    const float e_series = t * fmaf(t, fmaf(t, fmaf(t, 0.25f, 0.33333334f), 0.5f), 1.0f);  // -log(1-t), t < 2^-6: rel err < 1e-8
    const float e_lg2 = -LN2 * lg2_ftz(u);
    const float e = t < 0.015625f ? e_series : e_lg2;                                       // E = -log(u) ~ Exp(1)
    return -LN2 * lg2_ftz(e);
}
// gumbel_bits without fmix32's closing h ^= h >> 16 (which leaves the top 16 bits as they are): enough for a conservative
// compare of the top bits, two instructions less per column in the sampling epilogue's filter.
__device__ __forceinline__ uint32_t gumbel_bits_open(uint32_t row_step_hash, uint32_t v) {
    uint32_t h = row_step_hash ^ (v * 0xC2B2AE3Du);
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    return h;
}
__device__ __forceinline__ float gumbel_from_hash(uint32_t row_step_hash, uint32_t v) {
    return gumbel_from_bits(gumbel_bits(row_step_hash, v));
}
// Smallest 24-bit uniform (x >> 8) whose noise can exceed tau:  g > tau  <=>  u > exp(-exp(-tau)).  Conservative by two
// units (and by the caller's margin on tau), so that no possible winner is ever skipped; tau = -inf gives 0 (all pass).
__device__ __forceinline__ uint32_t gumbel_pass_threshold(float tau) {
    const float U = exp_ftz(-exp_ftz(-tau));
    const int thr = static_cast<int>(U * 16777216.0f) - 2;
    return thr > 0 ? static_cast<uint32_t>(thr) : 0u;
}
__device__ __forceinline__ uint32_t gumbel_row_step_hash(uint32_t seed, uint32_t row, uint32_t t) {
    uint32_t h = fmix32(seed ^ (row * 0x9E3779B1u));
    return fmix32(h ^ (t * 0x85EBCA77u));
}

__device__ __forceinline__ void store_h16x8(__half* dst, int lo_off, const float (&h)[8]) {
    __align__(16) __half hi[8];
    __align__(16) __half lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) split_f16(h[u], hi[u], lo[u]);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    if (lo_off > 0) *reinterpret_cast<uint4*>(dst + lo_off) = *reinterpret_cast<const uint4*>(lo);
}

// 32-byte stores (STG.256): one instruction writes a whole 32-byte L2 sector, where two 16-byte stores leave it partially
// written in between.  dst must be 32-byte aligned.
__device__ __forceinline__ void st_global_f32x8(float* dst, const float (&x)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(__float_as_uint(x[0])), "r"(__float_as_uint(x[1])),
                 "r"(__float_as_uint(x[2])), "r"(__float_as_uint(x[3])), "r"(__float_as_uint(x[4])), "r"(__float_as_uint(x[5])),
                 "r"(__float_as_uint(x[6])), "r"(__float_as_uint(x[7]))
                 : "memory");
}
__device__ __forceinline__ bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }
// 16 fp16 values (hi, and lo at +lo_off in the split mode) with one 32-byte store each when dst allows it
__device__ __forceinline__ void store_h16x16(__half* dst, int lo_off, const float (&h)[16]) {
    __align__(16) __half hi[16];
    __align__(16) __half lo[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) split_f16(h[u], hi[u], lo[u]);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(hi);
    if (aligned32(dst)) {
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
                     "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
    } else {
        reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(hi)[0];
        reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(hi)[1];
    }
    if (lo_off > 0) {
        reinterpret_cast<uint4*>(dst + lo_off)[0] = reinterpret_cast<const uint4*>(lo)[0];
        reinterpret_cast<uint4*>(dst + lo_off)[1] = reinterpret_cast<const uint4*>(lo)[1];
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// pre[8] holds the chunk's 32 bias values when the chunk is full (requested one chunk ahead by the caller, so that their
// L1 / L2 latency overlaps the previous chunk's arithmetic); the vocabulary's tail chunk loads its own.
__device__ __forceinline__ void request_bias32(float4 (&pre)[8], const float* __restrict__ bias, int n0, int N) {
    if (n0 + 32 <= N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pre[i] = __ldg(reinterpret_cast<const float4*>(bias + n0) + i);
    }
}

// ------------------------------------------------------------------------------------------------ epilogues
// Each epilogue thread owns ONE output row (TMEM lane) and walks its share of the tile's columns -- chunks
// [c0, c1) of 32 columns -- so row-wise statistics (softmax max / sum, top-k) need no cross-thread traffic.

template <int BLOCK_N>
__device__ __forceinline__ void epi_store(uint32_t taddr, int row, int n_base, int c0, int c1, const GemmParams& p) {
    const EpiParams& e = p.epi;
    const bool row_ok = row < p.M;
    const bool vec_ok = (p.N & 3) == 0;
    const bool keep = !(e.row_keep && row_ok && __ldg(e.row_keep + row) == 0.f);
    float4 nb[8];  // bias of the next chunk, requested one chunk ahead
    if (e.bias && vec_ok && row_ok) request_bias32(nb, e.bias, n_base + c0 * 32, p.N);
    // fp16 output rows that start at 16 (mod 32) bytes are written through a 16-byte carry (see below)
    const bool shifted = e.out16 && row_ok &&
                         (reinterpret_cast<uintptr_t>(e.out16 + static_cast<size_t>(row) * e.ld16 + n_base + c0 * 32) & 31) == 16;
    uint4 carry = make_uint4(0, 0, 0, 0);
    bool have_carry = false;
    __half* carry_at = nullptr;
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
        const int n0 = n_base + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        float4 cb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cb[i] = nb[i];
        if (e.bias && vec_ok && row_ok && c + 1 < c1) request_bias32(nb, e.bias, n0 + 32, p.N);
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        if (!row_ok) continue;
        if (e.bias) {
            if (vec_ok && n0 + 32 <= p.N) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[4 * i] += cb[i].x, v[4 * i + 1] += cb[i].y, v[4 * i + 2] += cb[i].z, v[4 * i + 3] += cb[i].w;
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (n0 + i < p.N) v[i] += __ldg(e.bias + n0 + i);
            }
        }
        if (e.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (!keep) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (e.out32) {
            float* o = e.out32 + static_cast<size_t>(row) * e.ld32 + n0;
            if (vec_ok && n0 + 32 <= p.N && aligned32(o)) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    const float x[8] = {v[i], v[i + 1], v[i + 2], v[i + 3], v[i + 4], v[i + 5], v[i + 6], v[i + 7]};
                    st_global_f32x8(o + i, x);
                }
            } else if (vec_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (n0 + i < p.N) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (n0 + i < p.N) o[i] = v[i];
            }
        }
        if (e.out16) {
            __half* o = e.out16 + static_cast<size_t>(row) * e.ld16 + n0;
            if (n0 + 32 <= p.N && e.lo16 == 0 && shifted) {
                // Row at 16 (mod 32) bytes -- every second row of the attention kernel's padded operands (pitch = cols + 8
                // halves): two 16-byte stores per sector would make L2 fetch each sector before merging (the projection GEMM
                // read 1.6x its algorithmic bytes).  The chunk's 64 bytes are written as [carry | first 16 B] + one aligned
                // 32-byte store + a new 16-byte carry, so only the two ends of the thread's 256-byte segment are partial.
                __align__(16) __half hh[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) hh[u] = __float2half_rn(v[u]);
                const uint4* q = reinterpret_cast<const uint4*>(hh);
                if (c == c0) {
                    *reinterpret_cast<uint4*>(o) = q[0];
                } else {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o - 8), "r"(carry.x), "r"(carry.y), "r"(carry.z),
                                 "r"(carry.w), "r"(q[0].x), "r"(q[0].y), "r"(q[0].z), "r"(q[0].w)
                                 : "memory");
                }
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 8), "r"(q[1].x), "r"(q[1].y), "r"(q[1].z), "r"(q[1].w),
                             "r"(q[2].x), "r"(q[2].y), "r"(q[2].z), "r"(q[2].w)
                             : "memory");
                carry = q[3];
                have_carry = true;
                carry_at = o + 24;
            } else if (n0 + 32 <= p.N) {
#pragma unroll
                for (int i = 0; i < 32; i += 16) {
                    float h[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) h[u] = v[i + u];
                    store_h16x16(o + i, e.lo16, h);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    if (n0 + i < p.N) {
                        float h[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) h[u] = v[i + u];
                        store_h16x8(o + i, e.lo16, h);
                    }
                }
            }
        }
        if (have_carry && (c + 1 == c1 || n0 + 64 > p.N)) {  // last full chunk of this thread's segment: flush the trailing 16 bytes
            *reinterpret_cast<uint4*>(carry_at) = carry;
            have_carry = false;
        }
    }
}

// Columns are packed gate-interleaved: n = 4*j + g, g in (i, f, g, o) -- torch.nn.LSTMCell gate order.
// The additive terms (bias or the hoisted per-image vector) and the previous cell state are fetched BEFORE the
// accumulator chunk is read from TMEM, so their L2 latency overlaps the TMEM load.
template <int BLOCK_N>
__device__ __forceinline__ void epi_lstm(uint32_t taddr, int row, int n_base, int c0, int c1, const GemmParams& p) {
    const EpiParams& e = p.epi;
    const bool row_ok = row < p.M;
    int prow = row;
    const float* add = e.bias;  // [N] shared by all rows ...
    if (row_ok) {
        if (e.parent) prow = __ldg(e.parent + row);
        if (e.rowadd) add = e.rowadd + static_cast<size_t>(row / e.rows_per_group) * e.rowadd_ld;  // ... or per image
    }
    const float* cin = e.c_in ? e.c_in + static_cast<size_t>(prow) * e.ldc : nullptr;
    const float* gat = (e.gather && row_ok) ? e.gather + static_cast<size_t>(__ldg(e.gather_idx + row)) * e.gather_ld : nullptr;
    // The additive terms of chunk c+1 (bias / hoisted per-image row, and the gathered embedding-gate row -- a random row
    // of a 155 MB table, i.e. an L2 or HBM access) are requested while chunk c is being computed: one chunk of load
    // latency is exposed per tile instead of four.
    float4 na[8], ng[8];
    auto request = [&](int c) {
        const int n0 = n_base + c * 32;
        if (row_ok && c < c1 && n0 < p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) na[i] = __ldg(reinterpret_cast<const float4*>(add + n0) + i);
            if (gat) {
#pragma unroll
                for (int i = 0; i < 8; ++i) ng[i] = __ldg(reinterpret_cast<const float4*>(gat + n0) + i);
            }
        }
    };
    request(c0);
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
        const int n0 = n_base + c * 32;
        if (n0 >= p.N) break;
        const int j0 = n0 >> 2;
        float4 ad[8];
        float4 cp0 = make_float4(0.f, 0.f, 0.f, 0.f), cp1 = cp0;
        if (row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                ad[i] = na[i];
                if (gat) ad[i].x += ng[i].x, ad[i].y += ng[i].y, ad[i].z += ng[i].z, ad[i].w += ng[i].w;
            }
            if (cin) {
                cp0 = *reinterpret_cast<const float4*>(cin + j0);
                cp1 = *reinterpret_cast<const float4*>(cin + j0 + 4);
            }
        }
        request(c + 1);
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        if (!row_ok) continue;
        const float cp[8] = {cp0.x, cp0.y, cp0.z, cp0.w, cp1.x, cp1.y, cp1.z, cp1.w};
        float cn[8], hn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float gi = v[4 * u] + ad[u].x, gf = v[4 * u + 1] + ad[u].y, gg = v[4 * u + 2] + ad[u].z, go = v[4 * u + 3] + ad[u].w;
            cn[u] = sigmoidf_acc(gf) * cp[u] + sigmoidf_acc(gi) * tanhf_acc(gg);
            hn[u] = sigmoidf_acc(go) * tanhf_acc(cn[u]);
        }
        float* co = e.c_out + static_cast<size_t>(row) * e.ldc + j0;
        if (aligned32(co)) {
            st_global_f32x8(co, cn);
        } else {
            *reinterpret_cast<float4*>(co) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(co + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
        }
        store_h16x8(e.out16 + static_cast<size_t>(row) * e.ld16 + j0, e.lo16, hn);
        if (e.h32) {
            float* ho = e.h32 + static_cast<size_t>(row) * e.ldh32 + j0;
            if (aligned32(ho)) {
                st_global_f32x8(ho, hn);
            } else {
                *reinterpret_cast<float4*>(ho) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                *reinterpret_cast<float4*>(ho + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
            }
        }
    }
}

// Issued by an epilogue thread as soon as it knows its next tile, i.e. while the tile's MMAs are still running:
// pulls the thread's additive rows (embedding-gate row of its word: a random 16 KB row of a 155 MB table, i.e. an
// HBM miss; per-image hoisted row; previous cell state) into L2 so that the epilogue proper sees L2 latency.
template <int BLOCK_N>
__device__ __forceinline__ void epi_lstm_prefetch(int row, int n_base, int c0, int c1, const GemmParams& p) {
    const EpiParams& e = p.epi;
    if (row >= p.M) return;
    const int n0 = n_base + c0 * 32;
    if (n0 >= p.N) return;
    const int lines = (c1 - c0);  // 32 fp32 columns = one 128-byte line per chunk
    if (e.gather) {
        const float* g = e.gather + static_cast<size_t>(__ldg(e.gather_idx + row)) * e.gather_ld + n0;
        for (int i = 0; i < lines; ++i) prefetch_l2(g + 32 * i);
    }
    if (e.rowadd) {
        const float* r = e.rowadd + static_cast<size_t>(row / e.rows_per_group) * e.rowadd_ld + n0;
        for (int i = 0; i < lines; ++i) prefetch_l2(r + 32 * i);
    }
    if (e.c_in) {
        const int prow = e.parent ? __ldg(e.parent + row) : row;
        prefetch_l2(e.c_in + static_cast<size_t>(prow) * e.ldc + (n0 >> 2));
    }
}

// Columns are packed (a_j, gate_j)-interleaved: n = 2*j + s.  nn.GLU: a * sigmoid(gate).
template <int BLOCK_N>
__device__ __forceinline__ void epi_glu(uint32_t taddr, int row, int n_base, int c0, int c1, const GemmParams& p) {
    const EpiParams& e = p.epi;
    const bool row_ok = row < p.M;
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
        const int n0 = n_base + c * 32;
        if (n0 >= p.N) break;
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        if (!row_ok) continue;
        const int j0 = n0 >> 1;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + i));
            v[i] += b.x, v[i + 1] += b.y, v[i + 2] += b.z, v[i + 3] += b.w;
        }
        float y[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) y[u] = v[2 * u] * sigmoidf_acc(v[2 * u + 1]);
        if (e.resid) {
            const float* r = e.resid + static_cast<size_t>(row) * e.ld_resid + j0;
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
                const float4 q = *reinterpret_cast<const float4*>(r + u);
                y[u] += q.x, y[u + 1] += q.y, y[u + 2] += q.z, y[u + 3] += q.w;
            }
        }
        if (e.out32) {
            float* o = e.out32 + static_cast<size_t>(row) * e.ld32 + j0;
            if (aligned32(o)) {
#pragma unroll
                for (int u = 0; u < 16; u += 8) {
                    const float x[8] = {y[u], y[u + 1], y[u + 2], y[u + 3], y[u + 4], y[u + 5], y[u + 6], y[u + 7]};
                    st_global_f32x8(o + u, x);
                }
            } else {
#pragma unroll
                for (int u = 0; u < 16; u += 4) *reinterpret_cast<float4*>(o + u) = make_float4(y[u], y[u + 1], y[u + 2], y[u + 3]);
            }
        }
        if (e.out16) store_h16x16(e.out16 + static_cast<size_t>(row) * e.ld16 + j0, e.lo16, y);
    }
}

// One 32-column chunk of logits: add the bias, mask the padded vocabulary tail, fold into the running
// (max, sum of exp(x - max)) pair.  exp via ex2.approx on log2e-prescaled arguments (one FFMA + one MUFU per element).
__device__ __forceinline__ float logits_chunk_stats(float (&v)[32], int n0, int N, const float* __restrict__ bias,
                                                    const float4 (&pre)[8], float& m, float& s) {
    float cmax = -INFINITY;
    if (n0 + 32 <= N) {
        float c4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[4 * i] += pre[i].x, v[4 * i + 1] += pre[i].y, v[4 * i + 2] += pre[i].z, v[4 * i + 3] += pre[i].w;
            c4[i & 3] = fmaxf(c4[i & 3], fmaxf(fmaxf(v[4 * i], v[4 * i + 1]), fmaxf(v[4 * i + 2], v[4 * i + 3])));
        }
        cmax = fmaxf(fmaxf(c4[0], c4[1]), fmaxf(c4[2], c4[3]));
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (n0 + i < N) {
                v[i] += __ldg(bias + n0 + i);
                cmax = fmaxf(cmax, v[i]);
            } else {
                v[i] = -INFINITY;
            }
        }
    }
    const float mn = fmaxf(m, cmax);
    const float mn2 = mn * LOG2E;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};  // four independent partial sums instead of one 32-long dependent chain
#pragma unroll
    for (int i = 0; i < 32; ++i) a4[i & 3] += ex2_ftz(fmaf(v[i], LOG2E, -mn2));  // exp2(-inf) = 0 for the padded columns
    s = s * ex2_ftz(fmaf(m, LOG2E, -mn2)) + ((a4[0] + a4[1]) + (a4[2] + a4[3]));
    m = mn;
    return cmax;
}

// v[i] for a run-time i out of registers: 31 selects.  (Indexing the array instead puts it into local memory -- 128 bytes per
// thread and chunk through the L1 / shared-memory arrays that the tensor core is reading its operands from: measured, the K = 1024
// logit GEMM fell from 1000 to 780 TFLOP/s.)
__device__ __forceinline__ float pick32(const float (&v)[32], int i) {
    float a[16], b[8], c[4];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = (i & 16) ? v[16 + j] : v[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (i & 8) ? a[8 + j] : a[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = (i & 4) ? b[4 + j] : b[j];
    const float d0 = (i & 2) ? c[2] : c[0], d1 = (i & 2) ? c[3] : c[1];
    return (i & 1) ? d1 : d0;
}

// Running per-row state of the fused log-softmax / top-k epilogue.  It is carried in registers across all the
// vocabulary tiles one CTA processes for a row block (a "run"), so the top-k warm-up is paid once per run and a
// row produces only runs * EPI_SPLIT partial records: (max, sum of exp(x - max), KTOP largest logits with their
// vocabulary indices, ties keep the lower index first).  The full logits row is never written to HBM.
template <int KTOP>
struct TopkState {
    float m, s;
    float tv[KTOP];
    int ti[KTOP];
    __device__ __forceinline__ void init() {
        m = -INFINITY, s = 0.f;
#pragma unroll
        for (int q = 0; q < KTOP; ++q) tv[q] = -INFINITY, ti[q] = 0x7FFFFFFF;
    }
    __device__ __forceinline__ void insert(float x, int idx) {  // x > tv[KTOP - 1]; ties keep the lower index first
        tv[KTOP - 1] = x;
        ti[KTOP - 1] = idx;
#pragma unroll
        for (int q = KTOP - 1; q > 0; --q) {
            if (tv[q] > tv[q - 1]) {
                const float fv = tv[q]; tv[q] = tv[q - 1]; tv[q - 1] = fv;
                const int iv = ti[q]; ti[q] = ti[q - 1]; ti[q - 1] = iv;
            }
        }
    }
    __device__ __forceinline__ void tile(uint32_t taddr, int n_base, int c0, int c1, const GemmParams& p) {
        // The chunk's 32 bias values (one 128-byte line, the same for every row) are loaded right before the accumulator, so
        // that the two latencies overlap; the line was pulled into L1 two chunks earlier.  (Holding the next chunk's values in
        // registers instead cost 32 register moves per chunk in this non-unrolled loop.)
        prefetch_l1(p.epi.bias + min(n_base + c0 * 32, p.N - 1));
        prefetch_l1(p.epi.bias + min(n_base + c0 * 32 + 32, p.N - 1));
#pragma unroll 1
        for (int c = c0; c < c1; ++c) {
            const int n0 = n_base + c * 32;
            if (n0 >= p.N) break;
            if (c + 2 < c1) prefetch_l1(p.epi.bias + min(n0 + 64, p.N - 1));
            float4 cb[8];
            request_bias32(cb, p.epi.bias, n0, p.N);
            float v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            const float cmax = logits_chunk_stats(v, n0, p.N, p.epi.bias, cb, m, s);
            if (cmax > tv[KTOP - 1]) {
                const float t0 = tv[KTOP - 1];
                if (t0 == -INFINITY) {  // the run's first chunk: every column enters
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (v[i] > tv[KTOP - 1]) insert(v[i], n0 + i);
                } else {
                    // Columns above the chunk-start threshold as a bit mask (branch-free, 32 independent compares), then only
                    // those, in column order: with 32 rows per warp "some lane has a new top-k entry" holds for most chunks
                    // of a run, and the unrolled 32-column insertion above is ~15 issue slots per column for all of them.
                    // v > t0  <=>  the sign of t0 - v (x - x is +0), shifted into the mask from column 31 down, two chains
                    uint32_t mhi = 0u, mlo = 0u;
#pragma unroll
                    for (int i = 15; i >= 0; --i) {
                        mhi = __funnelshift_l(__float_as_uint(t0 - v[16 + i]), mhi, 1);
                        mlo = __funnelshift_l(__float_as_uint(t0 - v[i]), mlo, 1);
                    }
                    uint32_t mask = (mhi << 16) | mlo;
                    while (mask != 0u) {
                        const int i = __ffs(static_cast<int>(mask)) - 1;
                        mask &= mask - 1u;
                        const float x = pick32(v, i);
                        if (x > tv[KTOP - 1]) insert(x, n0 + i);
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void flush(int row, int slot, const GemmParams& p) const {
        if (row >= p.M) return;
        constexpr int PS = topk_part_stride(KTOP);
        float* o = p.epi.part + (static_cast<size_t>(row) * p.epi.n_tiles + slot) * PS;
        o[0] = m;
        o[1] = s;
#pragma unroll
        for (int q = 0; q < KTOP; ++q) {
            o[2 + q] = tv[q];
            o[2 + KTOP + q] = __int_as_float(ti[q]);
        }
    }
};

// Greedy / multinomial draw: argmax over the vocabulary of (logit + Gumbel noise); with use_noise == 0 this is
// the plain argmax of ``sample``.  Also carries (max, sum-exp) for the log-prob of the drawn word.
struct DrawState {
    float m, s, best, best_raw;
    int best_i, forced;
    uint32_t rs;
    bool noisy;
    __device__ __forceinline__ void init(int row, const GemmParams& p) {
        m = -INFINITY, s = 0.f, best = -INFINITY, best_raw = 0.f, best_i = 0x7FFFFFFF;
        noisy = p.epi.use_noise != 0;
        int noise_row = row;
        if (p.epi.scst_n > 0) {
            const int img = row / (p.epi.scst_n + 1), j = row - img * (p.epi.scst_n + 1);
            noisy = noisy && j < p.epi.scst_n;  // the last row of an image's group is its greedy rollout
            noise_row = img * p.epi.scst_n + j;
        }
        const uint32_t seed = p.epi.seed_ptr ? __ldg(p.epi.seed_ptr) : p.epi.seed;
        rs = gumbel_row_step_hash(seed, static_cast<uint32_t>(noise_row), static_cast<uint32_t>(p.epi.step));
        forced = (p.epi.forced && row < p.M) ? __ldg(p.epi.forced + static_cast<size_t>(row) * p.epi.forced_ld + p.epi.step) : -1;
    }
    __device__ __forceinline__ void tile(uint32_t taddr, int n_base, int c0, int c1, const GemmParams& p) {
        // The chunk's 32 bias values (one 128-byte line, the same for every row) are loaded right before the accumulator, so
        // that the two latencies overlap; the line was pulled into L1 two chunks earlier.  (Holding the next chunk's values in
        // registers instead cost 32 register moves per chunk in this non-unrolled loop.)
        prefetch_l1(p.epi.bias + min(n_base + c0 * 32, p.N - 1));
        prefetch_l1(p.epi.bias + min(n_base + c0 * 32 + 32, p.N - 1));
#pragma unroll 1
        for (int c = c0; c < c1; ++c) {
            const int n0 = n_base + c * 32;
            if (n0 >= p.N) break;
            if (c + 2 < c1) prefetch_l1(p.epi.bias + min(n0 + 64, p.N - 1));
            float4 cb[8];
            request_bias32(cb, p.epi.bias, n0, p.N);
            float v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            const float cmax = logits_chunk_stats(v, n0, p.N, p.epi.bias, cb, m, s);
            if (p.epi.forced) {  // the given word "wins": its raw logit is what the log-prob needs
                const int f = forced - n0;
                if (f >= 0 && f < 32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i == f) best = 3.0e38f, best_i = forced, best_raw = v[i];
                }
            } else if (noisy) {
                // A column wins only if logit + noise > best, i.e. only with noise above best - (largest logit of the
                // chunk).  Whether a column's noise can exceed that bound is decided on the 24 hash bits alone (an integer
                // compare); the two logarithms are evaluated for the few columns that pass (~1e-3 of a uniform vocabulary
                // once the running best has settled, none of a trained model's low-probability words).  Exact: the columns
                // skipped could not have won, the others are evaluated as before and in the same order.
                // Branch-free filter, then the few survivors.  Per column: the hash up to its last multiply and ONE compare --
                // x = h ^ (h >> 16) keeps h's top 16 bits, so (x >> 8) >= thr needs h >= (thr << 8) & 0xFFFF0000 -- setting
                // a bit of a 32-bit mask; no per-column branch, 32 independent chains.  (ncu on the per-column-branch form:
                // 39 instructions per column at one issue per 5.6 cycles per warp -- with 32 rows per warp the rare "this
                // column might win" block ran for a fifth of all columns -- and the tensor pipe 36 % busy.)  The survivors of
                // all 32 lanes are then evaluated two at a time, in column order as before, so the draw is unchanged.
                // Padded columns of the vocabulary's tail chunk hold -inf and cannot win.
                const uint32_t thr = gumbel_pass_threshold(best - cmax - 1e-3f);
                if (thr == 0u) {  // nothing to compare against yet (the run's first chunk): every column, 32 independent chains
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float pv = v[i] + gumbel_from_bits(gumbel_bits(rs, static_cast<uint32_t>(n0 + i)));
                        if (pv > best) best = pv, best_i = n0 + i, best_raw = v[i];
                    }
                } else {
                    const uint32_t thr_hi = (thr << 8) & 0xFFFF0000u;
                    uint32_t mask = 0u;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        mask |= gumbel_bits_open(rs, static_cast<uint32_t>(n0 + i)) >= thr_hi ? (1u << i) : 0u;
                    while (mask != 0u) {  // two survivors per pass (their logits come out of registers by 31 selects each)
                        int ci[2];
                        float cp[2], cv[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            ci[q] = mask != 0u ? __ffs(static_cast<int>(mask)) - 1 : -1;
                            mask &= mask - 1u;  // 0 stays 0
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const int i = ci[q] < 0 ? 0 : ci[q];
                            cv[q] = pick32(v, i);
                            cp[q] = cv[q] + gumbel_from_bits(gumbel_bits(rs, static_cast<uint32_t>(n0 + i)));
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            if (ci[q] >= 0 && cp[q] > best) best = cp[q], best_i = n0 + ci[q], best_raw = cv[q];
                    }
                }
            } else if (cmax > best) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (v[i] > best) best = v[i], best_i = n0 + i, best_raw = v[i];
            }
        }
    }
    __device__ __forceinline__ void flush(int row, int slot, const GemmParams& p) const {
        if (row >= p.M) return;
        float* o = p.epi.part + (static_cast<size_t>(row) * p.epi.n_tiles + slot) * SAMPLE_PART_STRIDE;
        o[0] = m;
        o[1] = s;
        o[2] = best;
        o[3] = __int_as_float(best_i);
        o[4] = best_raw;
    }
};

// Work decomposition.  runs == 0: output tiles round-robin over the CTAs, M-fastest inside a GROUP of m_group row blocks
// and all column blocks of the group before the next group.  With m_group = number of resident CTAs (pairs) every CTA
// keeps its row block while it walks the columns, so the A operand is read from HBM once (a group's A rows stay in L2)
// and the co-resident CTAs share each weight tile; a plain M-fastest sweep re-reads A once per column block as soon as
// A outgrows L2 (the refiner's GEMMs: A = 113-226 MB).  runs > 0 (logit GEMM): a work item is (row block, run) = a
// contiguous range of vocabulary tiles of one row block, so the epilogue's per-row state persists across the item's tiles.
struct TileIter {
    int item, items, step, runs, num_m, num_n, m_group;
    int m_blk, n_blk, n_end, run;
    __device__ __forceinline__ TileIter(const GemmParams& p, int first = blockIdx.x, int stride = gridDim.x)
        : item(first), step(stride), runs(p.runs), num_m(p.num_m_blocks), num_n(p.num_n_blocks), m_group(p.m_group) {
        items = runs > 0 ? num_m * runs : num_m * num_n;
        n_blk = 0, n_end = 0, m_blk = 0, run = 0;
        if (m_group <= 0 || m_group > num_m) m_group = num_m;
    }
    // advance to the next tile of this CTA; returns false when done.  first_of_item / last_of_item delimit a run.
    __device__ __forceinline__ bool next(bool& first_of_item, bool& last_of_item) {
        if (runs > 0) {
            if (n_blk + 1 < n_end) {
                ++n_blk;
                first_of_item = false;
            } else {
                if (n_end != 0) item += step;
                if (item >= items) return false;
                m_blk = item / runs;
                run = item - m_blk * runs;
                n_blk = (run * num_n) / runs;
                n_end = ((run + 1) * num_n) / runs;
                first_of_item = true;
            }
            last_of_item = n_blk + 1 == n_end;
            return true;
        }
        if (n_end != 0) item += step;
        n_end = 1;
        if (item >= items) return false;
        const int per = m_group * num_n;
        const int grp = item / per, rem = item - grp * per;
        const int m0 = grp * m_group;
        const int gm = num_m - m0 < m_group ? num_m - m0 : m_group;  // the last group may be short
        n_blk = rem / gm;
        m_blk = m0 + rem - n_blk * gm;
        first_of_item = last_of_item = true;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ the kernel
template <int BLOCK_N, int EPI, int KTOP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    const uint32_t full_bar = smem_u32(bars);                     // [STAGES]
    const uint32_t empty_bar = smem_u32(bars + STAGES);           // [STAGES]
    const uint32_t tfull_bar = smem_u32(bars + 2 * STAGES);       // [2]
    const uint32_t tempty_bar = smem_u32(bars + 2 * STAGES + 2);  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_kb = p.k_blocks * p.passes;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, 32 * EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();  // set-up done: the next kernel of the stream may be scheduled behind this one ...
    griddep_wait();    // ... and this one touches its operands only once its predecessors have finished

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            TileIter it(p);
            bool f, l;
            while (it.next(f, l)) {
                const int m_blk = it.m_blk, n_blk = it.n_blk;
                for (int kk = 0; kk < total_kb; ++kk) {
                    const int pass = kk / p.k_blocks;
                    const int kb = kk - pass * p.k_blocks;
                    const int ka = kb * BLOCK_K + (pass == 2 ? p.a_lo_off : 0);
                    const int kbb = kb * BLOCK_K + (pass == 1 ? p.b_lo_off : 0);
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_BYTES;
                    mbar_arrive_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
                    tma_load_2d(sa, &tma_a, full_bar + 8 * stage, ka, m_blk * BLOCK_M);
                    tma_load_2d(sb, &tma_b, full_bar + 8 * stage, kbb, n_blk * BLOCK_N);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            TileIter it(p);
            bool f, l;
            while (it.next(f, l)) {
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kk = 0; kk < total_kb; ++kk) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t da = make_smem_desc_sw128(sa);
                    const uint64_t db = make_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 fp16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
                        umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kk | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar + 8 * stage);  // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
                umma_commit(tfull_bar + 8 * acc);  // accumulator complete -> epilogue
                if (++acc == 2) acc = 0, acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue warps (2 .. 2+EPI_WARPS) =====================
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int split = (warp - 2) >> 2;       // which share of the tile's columns
        constexpr int CHUNKS = BLOCK_N / 32;
        const int c0 = split * (CHUNKS / EPI_SPLIT), c1 = c0 + CHUNKS / EPI_SPLIT;
        int acc = 0;
        uint32_t acc_phase = 0;
        TopkState<KTOP> tk;
        DrawState sp;
        TileIter it(p);
        bool first, last;
        while (it.next(first, last)) {
            const int m_blk = it.m_blk, n_blk = it.n_blk;
            const int row = m_blk * BLOCK_M + quarter * 32 + lane;
            const int n_base = n_blk * BLOCK_N;
            if constexpr (EPI == EPI_LSTM) epi_lstm_prefetch<BLOCK_N>(row, n_base, c0, c1, p);
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
            if constexpr (EPI == EPI_STORE) epi_store<BLOCK_N>(taddr, row, n_base, c0, c1, p);
            else if constexpr (EPI == EPI_LSTM) epi_lstm<BLOCK_N>(taddr, row, n_base, c0, c1, p);
            else if constexpr (EPI == EPI_GLU) epi_glu<BLOCK_N>(taddr, row, n_base, c0, c1, p);
            else {
                // partial record slot: (run, column share) with runs, (vocabulary tile, column share) without
                const int slot = (p.runs > 0 ? it.run : n_blk) * EPI_SPLIT + split;
                if constexpr (EPI == EPI_TOPK) {
                    if (first) tk.init();
                    tk.tile(taddr, n_base, c0, c1, p);
                    if (last) tk.flush(row, slot, p);
                } else {
                    if (first) sp.init(row, p);
                    sp.tile(taddr, n_base, c0, c1, p);
                    if (last) sp.flush(row, slot, p);
                }
            }
            __syncwarp();
            tc_fence_before();
            mbar_arrive(tempty_bar + 8 * acc);
            if (++acc == 2) acc = 0, acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// Same GEMM with tcgen05.mma.cta_group::2: a cluster of two CTAs (one SM pair) computes a 256 x 256 output tile.
// Each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 weight rows); the pair MMA
// reads both halves, so per-SM shared-memory fill and L2->SM traffic drop from 48 KB to 32 KB per k-block and the
// operand reads per MMA from 12 KB to 8 KB -- the single-CTA kernel saturates the SM's shared-memory bandwidth
// at ~70 % tensor-pipe activity.  p.num_m_blocks counts 256-row PAIR blocks here.
//   full[s]   (leader only) : armed by the leader with the bytes of BOTH CTAs; both CTAs' TMA loads signal it
//   empty[s]  (each CTA)    : multicast tcgen05.commit from the leader's MMA thread
//   tfull[a]  (each CTA)    : multicast tcgen05.commit -> both epilogues
//   tempty[a] (leader only) : one arrive per epilogue warp of both CTAs (remote arrive from the peer)
struct GemmCfg2 {
    static constexpr int BLOCK_N = 256;
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = 6;
    static constexpr int TMEM_COLS = 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

// EW epilogue warps (a multiple of 4: EW / 4 warps per TMEM lane quarter share a tile's columns).  The sampling epilogue, the
// longest per column, runs with 12 (LOGIT_EPI_WARPS: 448 threads, 128 registers): 820 -> 864 TFLOP/s in the SCST step; the
// top-k epilogue measured the same with 8 and 12 (NIC 794 / 794, BUTD 1050 / 1028 TFLOP/s) and stays on 8.
template <int EPI, int KTOP, int EW = EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    using Cfg = GemmCfg2;
    constexpr int ES = EW / 4;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BLOCK_N = Cfg::BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    const uint32_t full_bar = smem_u32(bars);
    const uint32_t empty_bar = smem_u32(bars + STAGES);
    const uint32_t tfull_bar = smem_u32(bars + 2 * STAGES);
    const uint32_t tempty_bar = smem_u32(bars + 2 * STAGES + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int total_kb = p.k_blocks * p.passes;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, 2 * EW);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_cg2(smem_u32(tmem_slot), Cfg::TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits and TMEM allocation of BOTH CTAs are visible before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();  // set-up done: the next kernel of the stream may be scheduled behind this one ...
    griddep_wait();    // ... and this one touches its operands only once its predecessors have finished

    if (warp == 0) {
        // ===================== TMA producer (each CTA loads its A rows and its half of B) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            TileIter it(p, pair, npairs);
            bool f, l;
            while (it.next(f, l)) {
                const int m_blk = 2 * it.m_blk + static_cast<int>(rank), n_blk = it.n_blk;
                for (int kk = 0; kk < total_kb; ++kk) {
                    const int pass = kk / p.k_blocks;
                    const int kb = kk - pass * p.k_blocks;
                    const int ka = kb * BLOCK_K + (pass == 2 ? p.a_lo_off : 0);
                    const int kbb = kb * BLOCK_K + (pass == 1 ? p.b_lo_off : 0);
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_BYTES;
                    const uint32_t lead_full = mapa_shared(full_bar + 8 * stage, 0);
                    if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * Cfg::STAGE_BYTES);
                    tma_load_2d_cg2(sa, &tma_a, lead_full, ka, m_blk * BLOCK_M);
                    tma_load_2d_cg2(sb, &tma_b, lead_full, kbb, n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread of the LEADER CTA drives both tensor cores =====================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_f16(2 * BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            TileIter it(p, pair, npairs);
            bool f, l;
            while (it.next(f, l)) {
                mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kk = 0; kk < total_kb; ++kk) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t da = make_smem_desc_sw128(sa);
                    const uint64_t db = make_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kk | k) != 0 ? 1u : 0u);
                    umma_commit_cg2(empty_bar + 8 * stage, 3);  // frees the slot in BOTH CTAs
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
                umma_commit_cg2(tfull_bar + 8 * acc, 3);
                if (++acc == 2) acc = 0, acc_phase ^= 1;
            }
        }
    } else {
        // ===================== epilogue warps of both CTAs: each CTA owns its 128 accumulator rows =====================
        const int quarter = warp & 3;
        const int split = (warp - 2) >> 2;
        constexpr int CHUNKS = BLOCK_N / 32;
        const int c0 = split * CHUNKS / ES, c1 = (split + 1) * CHUNKS / ES;
        int acc = 0;
        uint32_t acc_phase = 0;
        TopkState<KTOP> tk;
        DrawState sp;
        TileIter it(p, pair, npairs);
        bool first, last;
        while (it.next(first, last)) {
            const int m_blk = 2 * it.m_blk + static_cast<int>(rank), n_blk = it.n_blk;
            const int row = m_blk * BLOCK_M + quarter * 32 + lane;
            const int n_base = n_blk * BLOCK_N;
            if constexpr (EPI == EPI_LSTM) epi_lstm_prefetch<BLOCK_N>(row, n_base, c0, c1, p);
            mbar_wait(tfull_bar + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
            if constexpr (EPI == EPI_STORE) epi_store<BLOCK_N>(taddr, row, n_base, c0, c1, p);
            else if constexpr (EPI == EPI_LSTM) epi_lstm<BLOCK_N>(taddr, row, n_base, c0, c1, p);
            else if constexpr (EPI == EPI_GLU) epi_glu<BLOCK_N>(taddr, row, n_base, c0, c1, p);
            else {
                const int slot = (p.runs > 0 ? it.run : n_blk) * ES + split;
                if constexpr (EPI == EPI_TOPK) {
                    if (first) tk.init();
                    tk.tile(taddr, n_base, c0, c1, p);
                    if (last) tk.flush(row, slot, p);
                } else {
                    if (first) sp.init(row, p);
                    sp.tile(taddr, n_base, c0, c1, p);
                    if (last) sp.flush(row, slot, p);
                }
            }
            __syncwarp();
            tc_fence_before();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar + 8 * acc, 0));
            if (++acc == 2) acc = 0, acc_phase ^= 1;
        }
    }

    tc_fence_before();
    cluster_sync_all();  // nobody frees TMEM / exits while the peer may still signal it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ chained pair kernel
// Two dependent GEMMs in ONE persistent launch of the CTA-pair kernel: phase 1 = a gate GEMM with the LSTM epilogue whose
// fp16 h' rows are phase 2's A operand (top-down LSTM -> dec_att(h1), BUTD_Model.py:265,58), phase 2 = bias + store.
// Phase-2 tiles are appended to every pair's work list, so they fill the last, partial wave of phase 1 and the second
// launch (ramp, pipeline fill, drain of a single-wave GEMM) disappears.  A phase-2 tile of row block m may read h' only
// once all column tiles of phase 1 have written it: every epilogue warp of phase 1 release-increments ready[m] after its
// stores; the TMA producers of phase 2 acquire-wait for the full count (all phase-1 tiles are in flight or done by then:
// they precede the phase-2 tiles in every pair's list).  The last producer to pass re-arms the counters.
// Measured on B200 at the benchmark batch (4608 rows): SLOWER than two launches (8.17 vs 7.63 ms per decode) -- with 288 + 72
// tiles on 74 pairs the phase-2 tiles queue behind four gate tiles on most pairs anyway, and every gate tile's epilogue now
// ends in a gpu-scope release; kept for CAPDEC_CHAIN=1 experiments, off by default.
struct ChainSync {
    int* ready;   // [num_m_blocks] arrivals of phase-1 epilogue warps (zero between launches)
    int* passed;  // [num_m_blocks] phase-2 producers that have passed the wait
    int ready_target;   // EPI_WARPS * 2 CTAs * num_n_blocks of phase 1
    int passed_target;  // 2 CTAs * num_n_blocks of phase 2
};

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_chain_kernel(const __grid_constant__ CUtensorMap tma_a1, const __grid_constant__ CUtensorMap tma_b1, const GemmParams p1,
                   const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2, const GemmParams p2,
                   const ChainSync cs) {
    using Cfg = GemmCfg2;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BLOCK_N = Cfg::BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    const uint32_t full_bar = smem_u32(bars);
    const uint32_t empty_bar = smem_u32(bars + STAGES);
    const uint32_t tfull_bar = smem_u32(bars + 2 * STAGES);
    const uint32_t tempty_bar = smem_u32(bars + 2 * STAGES + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a1);
        tma_prefetch_desc(&tma_b1);
        tma_prefetch_desc(&tma_a2);
        tma_prefetch_desc(&tma_b2);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, 2 * EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_cg2(smem_u32(tmem_slot), Cfg::TMEM_COLS);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();
    griddep_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int ph = 0; ph < 2; ++ph) {
                const GemmParams& p = ph == 0 ? p1 : p2;
                const CUtensorMap* ma = ph == 0 ? &tma_a1 : &tma_a2;
                const CUtensorMap* mb = ph == 0 ? &tma_b1 : &tma_b2;
                const int total_kb = p.k_blocks * p.passes;
                TileIter it(p, pair, npairs);
                bool f, l;
                while (it.next(f, l)) {
                    const int m_blk = 2 * it.m_blk + static_cast<int>(rank), n_blk = it.n_blk;
                    if (ph == 1) {  // phase 1 must have written every column of this row block's h'
                        const int* rdy = cs.ready + it.m_blk;
                        const long long t0 = clock64();
                        while (true) {
                            int v;
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(rdy) : "memory");
                            if (v >= cs.ready_target) break;
                            if (clock64() - t0 > 4000000000LL) {
                                printf("capdec: chained GEMM wait timed out (block %d row block %d: %d of %d)\n", blockIdx.x, it.m_blk, v,
                                       cs.ready_target);
                                __trap();
                            }
                        }
                        fence_proxy_async_global();  // the generic-proxy stores of h' are ordered before this thread's TMA reads
                        if (atomicAdd(cs.passed + it.m_blk, 1) == cs.passed_target - 1) cs.ready[it.m_blk] = 0, cs.passed[it.m_blk] = 0;
                    }
                    for (int kk = 0; kk < total_kb; ++kk) {
                        const int pass = kk / p.k_blocks;
                        const int kb = kk - pass * p.k_blocks;
                        const int ka = kb * BLOCK_K + (pass == 2 ? p.a_lo_off : 0);
                        const int kbb = kb * BLOCK_K + (pass == 1 ? p.b_lo_off : 0);
                        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint32_t sb = sa + Cfg::A_BYTES;
                        const uint32_t lead_full = mapa_shared(full_bar + 8 * stage, 0);
                        if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * Cfg::STAGE_BYTES);
                        tma_load_2d_cg2(sa, ma, lead_full, ka, m_blk * BLOCK_M);
                        tma_load_2d_cg2(sb, mb, lead_full, kbb, n_blk * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_f16(2 * BLOCK_M, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int ph = 0; ph < 2; ++ph) {
                const GemmParams& p = ph == 0 ? p1 : p2;
                const int total_kb = p.k_blocks * p.passes;
                TileIter it(p, pair, npairs);
                bool f, l;
                while (it.next(f, l)) {
                    mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                    for (int kk = 0; kk < total_kb; ++kk) {
                        mbar_wait(full_bar + 8 * stage, phase);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint64_t da = make_smem_desc_sw128(sa);
                        const uint64_t db = make_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_f16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc, (kk | k) != 0 ? 1u : 0u);
                        umma_commit_cg2(empty_bar + 8 * stage, 3);
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                    umma_commit_cg2(tfull_bar + 8 * acc, 3);
                    if (++acc == 2) acc = 0, acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;
        const int split = (warp - 2) >> 2;
        constexpr int CHUNKS = BLOCK_N / 32;
        const int c0 = split * (CHUNKS / EPI_SPLIT), c1 = c0 + CHUNKS / EPI_SPLIT;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int ph = 0; ph < 2; ++ph) {
            const GemmParams& p = ph == 0 ? p1 : p2;
            TileIter it(p, pair, npairs);
            bool first, last;
            while (it.next(first, last)) {
                const int m_blk = 2 * it.m_blk + static_cast<int>(rank), n_blk = it.n_blk;
                const int row = m_blk * BLOCK_M + quarter * 32 + lane;
                const int n_base = n_blk * BLOCK_N;
                if (ph == 0) epi_lstm_prefetch<BLOCK_N>(row, n_base, c0, c1, p);
                mbar_wait(tfull_bar + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
                if (ph == 0) epi_lstm<BLOCK_N>(taddr, row, n_base, c0, c1, p);
                else epi_store<BLOCK_N>(taddr, row, n_base, c0, c1, p);
                __syncwarp();
                tc_fence_before();
                if (lane == 0) {
                    mbar_arrive_cluster(mapa_shared(tempty_bar + 8 * acc, 0));
                    if (ph == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(cs.ready + it.m_blk) : "memory");
                }
                if (++acc == 2) acc = 0, acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace capdec
