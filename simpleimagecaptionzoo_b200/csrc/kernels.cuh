// Non-GEMM kernels of the caption decoder: weight packing, one-time feature preparation, the per-step
// attention kernels (HBM-bound), LayerNorm, and the beam / sampling bookkeeping kernels that also assemble
// the next step's GEMM operands (beam-state reorder by parent index + embedding gather).
#pragma once
#include "gemm.cuh"

namespace capdec {

constexpr int TOK_PAD = 0, TOK_STA = 1, TOK_END = 2;  // PreProcess/Build_caption_vocab.py:37-40
constexpr int MAX_ROWS = 8;                           // beam size / samples per image supported per CTA

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ================================================================================================ weights
// scale[n] = g[n] / ||v[n,:]||_2 in double (legacy torch weight_norm, dim=0).
__global__ void weightnorm_scale_kernel(const float* __restrict__ g, const float* __restrict__ v, int N, int K,
                                        double* __restrict__ scale) {
    const int n = blockIdx.x;
    if (n >= N) return;
    double acc = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const double x = v[static_cast<size_t>(n) * K + k];
        acc += x * x;
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
        scale[n] = static_cast<double>(g[n]) / sqrt(t);
    }
}

// Row permutations of the packed operand:  0 identity;  1 LSTM gate interleave (dst 4*j+g <- src g*Hh+j);
// 2 GLU interleave (dst 2*j+s <- src s*Hh+j).
__host__ __device__ __forceinline__ int packed_src_row(int n_dst, int mode, int Hh) {
    if (mode == 1) return (n_dst & 3) * Hh + (n_dst >> 2);
    if (mode == 2) return (n_dst & 1) * Hh + (n_dst >> 1);
    return n_dst;
}

// dst[n, dst_col + k] (hi) and dst[n, lo + dst_col + k] (lo) <- src[src_row(n), src_col + k] * scale[src_row]
__global__ void pack_weight_kernel(const float* __restrict__ src, int src_ld, int src_col, const double* __restrict__ scale,
                                   __half* __restrict__ dst, int dst_ld, int dst_lo, int dst_col, int N, int K, int mode,
                                   int Hh) {
    const size_t total = static_cast<size_t>(N) * K;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<size_t>(n) * K);
        const int sr = packed_src_row(n, mode, Hh);
        float w = src[static_cast<size_t>(sr) * src_ld + src_col + k];
        if (scale) w = static_cast<float>(static_cast<double>(w) * scale[sr]);
        __half hi, lo;
        split_f16(w, hi, lo);
        dst[static_cast<size_t>(n) * dst_ld + dst_col + k] = hi;
        if (dst_lo > 0) dst[static_cast<size_t>(n) * dst_ld + dst_lo + dst_col + k] = lo;
    }
}

// dst[n] = a[src_row(n)] (+ b[src_row(n)])
__global__ void pack_bias_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst, int N,
                                 int mode, int Hh) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int sr = packed_src_row(n, mode, Hh);
    dst[n] = b ? a[sr] + b[sr] : a[sr];
}

// scaled folded vector (the BUTD ``affine`` layer has one output row): dst[k] = v[k] * scale[0]
__global__ void fold_vector_kernel(const float* __restrict__ v, const double* __restrict__ scale, float* __restrict__ dst,
                                   int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) dst[k] = static_cast<float>(static_cast<double>(v[k]) * scale[0]);
}

// ================================================================================================ prepare
// fp32 [rows, cols] -> fp16 hi|lo operand
__global__ void cvt_f16_kernel(const float* __restrict__ src, size_t rows, int cols, __half* __restrict__ dst, int dst_ld,
                               int dst_lo, int dst_col, int relu = 0) {
    const size_t total4 = rows * static_cast<size_t>(cols) / 4;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t e = i * 4;
        const size_t r = e / cols;
        const int c = static_cast<int>(e - r * cols);
        float4 x = __ldg(reinterpret_cast<const float4*>(src + e));
        if (relu) x.x = fmaxf(x.x, 0.f), x.y = fmaxf(x.y, 0.f), x.z = fmaxf(x.z, 0.f), x.w = fmaxf(x.w, 0.f);
        __align__(8) __half hi[4];
        __align__(8) __half lo[4];
        split_f16(x.x, hi[0], lo[0]);
        split_f16(x.y, hi[1], lo[1]);
        split_f16(x.z, hi[2], lo[2]);
        split_f16(x.w, hi[3], lo[3]);
        __half* d = dst + r * dst_ld + dst_col + c;
        *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
        if (dst_lo > 0) *reinterpret_cast<uint2*>(d + dst_lo) = *reinterpret_cast<const uint2*>(lo);
    }
}

// mean over regions (optionally masked: sum(f*m)/sum(m), AoA_Model.py:422-425); writes fp32 and/or fp16 operand.
// T = float (the reference's feature format) or __half (packed fp16 feature shards, rows of ld elements).
template <typename T>
__global__ void region_mean_kernel(const T* __restrict__ feats, int ld, const float* __restrict__ mask, int R, int C,
                                   float* __restrict__ out32, __half* __restrict__ out16, int ld16, int lo16) {
    const int b = blockIdx.x;
    const T* f = feats + static_cast<size_t>(b) * R * ld;
    float msum = static_cast<float>(R);
    if (mask) {
        msum = 0.f;
        for (int r = 0; r < R; ++r) msum += mask[static_cast<size_t>(b) * R + r];
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) {
            const float x = static_cast<float>(f[static_cast<size_t>(r) * ld + c]);
            acc += mask ? x * mask[static_cast<size_t>(b) * R + r] : x;
        }
        const float m = acc / msum;
        if (out32) out32[static_cast<size_t>(b) * C + c] = m;
        if (out16) {
            __half hi, lo;
            split_f16(m, hi, lo);
            out16[static_cast<size_t>(b) * ld16 + c] = hi;
            if (lo16 > 0) out16[static_cast<size_t>(b) * ld16 + lo16 + c] = lo;
        }
    }
}

// BUTD ingest of one image per CTA in ONE pass over its features [R, D]: convert / copy into the row-padded fp16 operand
// layout (hi | lo in the split mode) and accumulate the mean over the regions (BUTD_Model.py:251) as the fp16 operand of
// the hoisted gate term.  T = float (reference format) or __half (packed feature shards).  D % 8 == 0.
template <typename T>
__global__ void __launch_bounds__(256) butd_ingest_kernel(const T* __restrict__ feats, int R, int D, __half* __restrict__ dst,
                                                          int dst_ld, int dst_lo, __half* __restrict__ mean16, int mean_ld,
                                                          int mean_lo) {
    const int b = blockIdx.x;
    const T* src = feats + static_cast<size_t>(b) * R * D;
    __half* out = dst + static_cast<size_t>(b) * R * dst_ld;
    const float rf = static_cast<float>(R);
    for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < R; ++r) {
            float x[8];
            if constexpr (sizeof(T) == 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * D + c));
                const float4 d = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * D + c) + 1);
                x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = d.x, x[5] = d.y, x[6] = d.z, x[7] = d.w;
                store_h16x8(out + static_cast<size_t>(r) * dst_ld + c, dst_lo, x);
            } else {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * D + c));
                *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * dst_ld + c) = raw;
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    x[2 * i] = f.x, x[2 * i + 1] = f.y;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += x[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = acc[i] / rf;
        store_h16x8(mean16 + static_cast<size_t>(b) * mean_ld + c, mean_lo, acc);
    }
}

__global__ void set_u32_kernel(uint32_t* p, uint32_t v) { *p = v; }

// fp16 operand rows (hi, + lo in the fp32-grade mode) -> fp32 rows at a caller-chosen row stride
__global__ void export_f32_kernel(const __half* __restrict__ src, int ld, int lo, int M, int n, float* __restrict__ dst, size_t dst_ld) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < static_cast<size_t>(M) * n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / n, c = i - r * n;
        float v = __half2float(src[r * ld + c]);
        if (lo > 0) v += __half2float(src[r * ld + lo + c]);
        dst[r * dst_ld + c] = v;
    }
}

__global__ void fill_f16_kernel(__half* p, size_t n, float v) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        p[i] = __float2half_rn(v);
}

// ================================================================================================ BUTD attention
// 8 consecutive elements as raw 16-byte vectors (issued as independent loads first, converted later, so that
// several loads per thread are in flight: the kernel is HBM-bound only if memory-level parallelism is high).
template <typename T> struct Raw8;
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void lds(const float* p) {
        a = *reinterpret_cast<const float4*>(p);
        b = *(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f), b = a; }
    __device__ __forceinline__ void get(float (&x)[8]) const {
        x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
    }
};
template <> struct Raw8<__half> {
    uint4 u;
    __device__ __forceinline__ void load(const __half* p) { u = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void lds(const __half* p) { u = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void zero() { u = make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void get(float (&x)[8]) const {
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            x[2 * i] = f.x, x[2 * i + 1] = f.y;
        }
    }
};

// Sum NP (power of two <= 32) per-lane values across the 32 lanes of a warp with NP-1 (+ log2(32/NP)) shuffles
// instead of 5*NP: at every step each lane keeps one half of its values and hands the other half to its partner.
// Returns the fully reduced value of output index `out_idx`; lanes with (lane & (32/NP - 1)) == 0 are the writers.
template <int NP>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[NP], int lane, int& out_idx) {
    int idx = 0;
#pragma unroll
    for (int st = 0; st < 5; ++st) {
        const int o = 16 >> st;
        constexpr int dummy = 0;
        (void)dummy;
        const int cnt = NP >> st;  // values still held per lane before this step (compile-time after unrolling)
        if (cnt > 1) {
            const int half = cnt >> 1;
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < NP / 2; ++i) {
                if (i < half) {
                    const float send = upper ? v[i] : v[i + half];
                    const float keep = upper ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            idx += upper ? half : 0;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        }
    }
    out_idx = idx;
    return v[0];
}

template <int KR> struct AttCfg {
    static constexpr int GR = KR <= 1 ? 8 : (KR <= 3 ? 6 : (KR <= 5 ? 4 : 2));  // regions per thread per group
    static constexpr int NV = KR * GR;                                          // partial sums per thread per group
    static constexpr int NP = NV <= 8 ? 8 : (NV <= 16 ? 16 : 32);
    static constexpr int LB = KR <= 3 ? 9 : (KR <= 5 ? 6 : 4);                  // feature loads in flight (phase 3)
};

// One CTA (256 threads) per image, all K rows (beams / samples) of the image together so that the image's
// projected features enc_ctx [R,A] and raw features [R,D] are read from HBM once per image-step, not per row.
//   e[k,r]   = w_aff . relu(enc_ctx[r,:] + dec_ctx[k,:]) + b_aff      (BUTD_Model.py:58-59, ReLU not tanh)
//   alpha    = softmax_r(e)                                            (:60)
//   ctx[k,:] = sum_r alpha[k,r] * feats[r,:]                           (:61)
// ctx is written straight into the language LSTM's fp16 operand buffer.  T = float (fp32-grade mode) or __half
// (fp16 mode: projected and raw features are read in the fp16 form the projection GEMM already uses -- half the
// HBM bytes).  Phase 1: a thread owns 8 columns of A (w and dec stay in registers) and GR regions of one parity
// per group; partial dot products are reduced with a transposing butterfly + one smem hop across the 4 warps of a
// parity.  A, D multiples of 8.
template <int KR, typename T>
__global__ void __launch_bounds__(256) butd_attention_kernel(const T* __restrict__ enc_ctx, int enc_ld, const T* __restrict__ feats,
                                                             int feats_ld, const float* __restrict__ dec_ctx,
                                                             const float* __restrict__ w_aff, float b_aff, int R, int A, int D,
                                                             int K, __half* __restrict__ ctx16, int ld16, int lo16,
                                                             float* __restrict__ alphas_out, size_t alpha_stride) {
    using C = AttCfg<KR>;
    constexpr int GR = C::GR, NV = C::NV, NP = C::NP, LB = C::LB;
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    extern __shared__ float sm[];
    float* s_e = sm;                    // [KR][R] scores, then alphas
    float* s_red = s_e + KR * R;        // [2][8 warps][NP]
    const int img = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const int parity = tid >> 7;        // which of the two interleaved region streams
    const int tcol = tid & 127;

    // ---------------- phase 1: attention scores
    const T* enc = enc_ctx + static_cast<size_t>(img) * R * enc_ld;
    const float* dec = dec_ctx + static_cast<size_t>(img) * K * A;
    const int n_groups = (R + 2 * GR - 1) / (2 * GR);
    for (int g = 0; g < n_groups; ++g) {
        float p[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) p[i] = 0.f;
        for (int a0 = tcol * 8; a0 < A; a0 += 1024) {
            Raw8<T> raw[GR];
#pragma unroll
            for (int i = 0; i < GR; ++i) {
                const int r = g * 2 * GR + parity + 2 * i;
                if (r < R) raw[i].load(enc + static_cast<size_t>(r) * enc_ld + a0);
                else raw[i].zero();
            }
            float w[8];
            Raw8<float> t;
            t.load(w_aff + a0);
            t.get(w);
            float d[KR][8];
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                if (k < K) {
                    t.load(dec + k * A + a0);
                    t.get(d[k]);
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) d[k][q] = 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < GR; ++i) {
                float x[8];
                raw[i].get(x);
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    float acc = p[k * GR + i];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc = fmaf(w[q], fmaxf(x[q] + d[k][q], 0.f), acc);
                    p[k * GR + i] = acc;
                }
            }
        }
        int oidx;
        const float tot = warp_transpose_reduce<NP>(p, lane, oidx);
        float* red = s_red + (g & 1) * 8 * NP;
        if ((lane & (32 / NP - 1)) == 0) red[warp * NP + oidx] = tot;
        __syncthreads();
        if (tid < 2 * NV) {
            const int par = tid / NV, v = tid - par * NV;
            const float sum = red[(par * 4 + 0) * NP + v] + red[(par * 4 + 1) * NP + v] + red[(par * 4 + 2) * NP + v] +
                              red[(par * 4 + 3) * NP + v];
            const int k = v / GR, i = v - k * GR;
            const int r = g * 2 * GR + par + 2 * i;
            if (r < R && k < K) s_e[k * R + r] = sum + b_aff;
        }
    }
    __syncthreads();

    // ---------------- phase 2: softmax over regions, one warp per row
    for (int k = warp; k < K; k += nwarp) {
        float m = -INFINITY;
        for (int r = lane; r < R; r += 32) m = fmaxf(m, s_e[k * R + r]);
        m = warp_max(m);
        float s = 0.f;
        for (int r = lane; r < R; r += 32) {
            const float ex = expf(s_e[k * R + r] - m);
            s_e[k * R + r] = ex;
            s += ex;
        }
        s = warp_sum(s);
        for (int r = lane; r < R; r += 32) {
            const float al = s_e[k * R + r] / s;
            s_e[k * R + r] = al;
            if (alphas_out) alphas_out[(static_cast<size_t>(img) * K + k) * alpha_stride + r] = al;
        }
    }
    __syncthreads();

    // ---------------- phase 3: attention-weighted feature sum
    const T* f = feats + static_cast<size_t>(img) * R * feats_ld;
    for (int d0 = tid * 8; d0 < D; d0 += blockDim.x * 8) {
        float acc[KR][8];
#pragma unroll
        for (int k = 0; k < KR; ++k)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[k][q] = 0.f;
        for (int r0 = 0; r0 < R; r0 += LB) {
            Raw8<T> raw[LB];
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                if (r0 + i < R) raw[i].load(f + static_cast<size_t>(r0 + i) * feats_ld + d0);
                else raw[i].zero();
            }
#pragma unroll
            for (int i = 0; i < LB; ++i) {
                if (r0 + i < R) {
                    float x[8];
                    raw[i].get(x);
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        if (k < K) {
                            const float al = s_e[k * R + r0 + i];
#pragma unroll
                            for (int q = 0; q < 8; ++q) acc[k][q] = fmaf(al, x[q], acc[k][q]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            if (k < K) store_h16x8(ctx16 + (static_cast<size_t>(img) * K + k) * ld16 + d0, lo16, acc[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ streaming form
// Persistent, copy-engine-fed variant of the kernel above (the one the decode loop uses when A <= 1024, D <= 2048 and
// rows are dense): CTAs loop over images; one producer thread streams each image's projected features (phase 1) and
// raw features (phase 3) into a shared-memory ring with cp.async.bulk (completion on mbarriers), running ahead of
// the 256 consumer threads across chunk AND image boundaries, so HBM reads never wait for the arithmetic.
template <int KR, typename T> struct AttStreamCfg {
    static constexpr int GR = KR <= 3 ? 6 : (KR <= 5 ? 4 : 2);
    static constexpr int NV = KR * GR;
    static constexpr int NP = NV <= 8 ? 8 : (NV <= 16 ? 16 : 32);
    static constexpr int RC3 = 6;                                            // regions per phase-3 chunk
    static constexpr int STAGE_BYTES = 12 * 1024 * static_cast<int>(sizeof(T));  // 12 regions x 1024 cols / 6 x 2048
    static constexpr int STAGES = sizeof(T) == 2 ? 4 : 3;
    static constexpr int THREADS = 288;                                      // warp 0 = producer, warps 1-8 = consumers
    static constexpr int CTAS_PER_SM = (sizeof(T) == 2 && KR <= 3) ? 2 : 1;  // register budget: no spills
};

template <int KR, typename T>
__global__ void __launch_bounds__(288, AttStreamCfg<KR, T>::CTAS_PER_SM)
butd_attention_stream_kernel(const T* __restrict__ enc_ctx, const T* __restrict__ feats, const float* __restrict__ dec_ctx,
                             const float* __restrict__ w_aff, float b_aff, int B, int R, int A, int D, int K,
                             __half* __restrict__ ctx16, int ld16, int lo16, float* __restrict__ alphas_out, size_t alpha_stride) {
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    using C = AttStreamCfg<KR, T>;
    constexpr int GR = C::GR, NV = C::NV, NP = C::NP, RC3 = C::RC3, STAGES = C::STAGES;
    extern __shared__ uint8_t att_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
    float* s_e = reinterpret_cast<float*>(bars + 2 * STAGES);  // [KR][R]
    float* s_red = s_e + KR * R;                               // [2][8][NP]
    const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = (R + 2 * GR - 1) / (2 * GR);
    const int n_chunks3 = (R + RC3 - 1) / RC3;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 8);  // one arrive per consumer warp
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ===================== producer: one thread issues every bulk copy, in consumer order =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int img = blockIdx.x; img < B; img += gridDim.x) {
                const T* enc = enc_ctx + static_cast<size_t>(img) * R * A;
                const T* f = feats + static_cast<size_t>(img) * R * D;
                for (int c = 0; c < n_groups + n_chunks3; ++c) {
                    const void* src;
                    uint32_t bytes;
                    if (c < n_groups) {
                        const int r0 = c * 2 * GR, nr = min(2 * GR, R - r0);
                        src = enc + static_cast<size_t>(r0) * A;
                        bytes = static_cast<uint32_t>(nr) * A * sizeof(T);
                    } else {
                        const int r0 = (c - n_groups) * RC3, nr = min(RC3, R - r0);
                        src = f + static_cast<size_t>(r0) * D;
                        bytes = static_cast<uint32_t>(nr) * D * sizeof(T);
                    }
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(full_bar + 8 * stage, bytes);
                    bulk_load_1d(smem_u32(smem + stage * C::STAGE_BYTES), src, bytes, full_bar + 8 * stage);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
        return;
    }

    // ===================== consumers (256 threads) =====================
    const int tid = threadIdx.x - 32, cw = warp - 1;
    const int parity = tid >> 7, tcol = tid & 127;
    const int a0 = tcol * 8;
    const bool a_ok = a0 < A;
    const int d0 = tid * 8;
    const bool d_ok = d0 < D;
    float w[8];
    {
        Raw8<float> t;
        if (a_ok) t.load(w_aff + a0); else t.zero();
        t.get(w);
    }
    float d[KR][8];
    auto load_dec = [&](int img) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            Raw8<float> t;
            if (a_ok && k < K && img < B) t.load(dec_ctx + (static_cast<size_t>(img) * K + k) * A + a0); else t.zero();
            t.get(d[k]);
        }
    };
    load_dec(blockIdx.x);
    int stage = 0;
    uint32_t phase = 0;
    int red_flip = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x) {
        // ---------------- phase 1: scores, one ring stage per group of 2*GR regions
        for (int g = 0; g < n_groups; ++g) {
            float p[NP];
#pragma unroll
            for (int i = 0; i < NP; ++i) p[i] = 0.f;
            mbar_wait(full_bar + 8 * stage, phase);
            const T* chunk = reinterpret_cast<const T*>(smem + stage * C::STAGE_BYTES);
            if (a_ok) {
#pragma unroll
                for (int i = 0; i < GR; ++i) {
                    const int rl = 2 * i + parity;
                    if (g * 2 * GR + rl < R) {
                        float x[8];
                        Raw8<T> raw;
                        raw.lds(chunk + static_cast<size_t>(rl) * A + a0);
                        raw.get(x);
#pragma unroll
                        for (int k = 0; k < KR; ++k) {
                            float acc = 0.f;
#pragma unroll
                            for (int q = 0; q < 8; ++q) acc = fmaf(w[q], fmaxf(x[q] + d[k][q], 0.f), acc);
                            p[k * GR + i] = acc;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
            if (++stage == STAGES) stage = 0, phase ^= 1;
            int oidx;
            const float tot = warp_transpose_reduce<NP>(p, lane, oidx);
            float* red = s_red + red_flip * 8 * NP;
            red_flip ^= 1;
            if ((lane & (32 / NP - 1)) == 0) red[cw * NP + oidx] = tot;
            named_bar_sync(1, 256);
            if (tid < 2 * NV) {
                const int par = tid / NV, v = tid - par * NV;
                const float sum = red[(par * 4 + 0) * NP + v] + red[(par * 4 + 1) * NP + v] + red[(par * 4 + 2) * NP + v] +
                                  red[(par * 4 + 3) * NP + v];
                const int k = v / GR, i = v - k * GR;
                const int r = g * 2 * GR + par + 2 * i;
                if (r < R && k < K) s_e[k * R + r] = sum + b_aff;
            }
        }
        load_dec(img + gridDim.x);  // next image's dec_att rows: latency hidden behind phases 2-3
        named_bar_sync(1, 256);
        // ---------------- phase 2: softmax over regions, one warp per row
        for (int k = cw; k < K; k += 8) {
            float m = -INFINITY;
            for (int r = lane; r < R; r += 32) m = fmaxf(m, s_e[k * R + r]);
            m = warp_max(m);
            float s = 0.f;
            for (int r = lane; r < R; r += 32) {
                const float ex = expf(s_e[k * R + r] - m);
                s_e[k * R + r] = ex;
                s += ex;
            }
            s = warp_sum(s);
            for (int r = lane; r < R; r += 32) {
                const float al = s_e[k * R + r] / s;
                s_e[k * R + r] = al;
                if (alphas_out) alphas_out[(static_cast<size_t>(img) * K + k) * alpha_stride + r] = al;
            }
        }
        named_bar_sync(1, 256);
        // ---------------- phase 3: weighted feature sum, one ring stage per RC3 regions
        float acc[KR][8];
#pragma unroll
        for (int k = 0; k < KR; ++k)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[k][q] = 0.f;
        for (int c = 0; c < n_chunks3; ++c) {
            mbar_wait(full_bar + 8 * stage, phase);
            const T* chunk = reinterpret_cast<const T*>(smem + stage * C::STAGE_BYTES);
            if (d_ok) {
#pragma unroll
                for (int i = 0; i < RC3; ++i) {
                    const int r = c * RC3 + i;
                    if (r < R) {
                        float x[8];
                        Raw8<T> raw;
                        raw.lds(chunk + static_cast<size_t>(i) * D + d0);
                        raw.get(x);
#pragma unroll
                        for (int k = 0; k < KR; ++k) {
                            const float al = s_e[k * R + r];
#pragma unroll
                            for (int q = 0; q < 8; ++q) acc[k][q] = fmaf(al, x[q], acc[k][q]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
            if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        if (d_ok) {
#pragma unroll
            for (int k = 0; k < KR; ++k)
                if (k < K) store_h16x8(ctx16 + (static_cast<size_t>(img) * K + k) * ld16 + d0, lo16, acc[k]);
        }
        named_bar_sync(1, 256);  // s_e is rewritten by the next image's phase 1
    }
}

// ------------------------------------------------------------------------------------------------ fp16 fragment form
// The fp16-mode attention kernel of the decode loop.  Same persistent bulk-copy ring as above, but the arithmetic
// is done on MMA fragments so that the kernel is bound by HBM, not by instruction issue (the FFMA form executes
// ~40 k warp instructions per image, this one ~10 k):
//   phase 1  e[k,r] = sum_a w[a]*relu(enc[r,a] + dec[k,a]): ldmatrix loads a 16-region x 16-column tile as an
//            m16n8k16 A fragment, the add + ReLU run as packed half2 ops on the fragment, and the dot product with
//            w is one MMA whose B fragment holds w in column 0 -- the reduction over a happens inside the MMA.
//   phase 3  ctx[k,d] = sum_r alpha[k,r]*feats[r,d]: feats^T tiles (ldmatrix.trans) are the A operand of m16n8k8,
//            alpha (beam = n) the B operand.
// The fp16 copies of the projected / raw features are stored in HBM with a 16-byte row pad (ld = cols + 8), so one
// contiguous bulk copy per chunk lands rows in shared memory at a pitch that makes ldmatrix bank-conflict free.
// Requires A % 16 == 0, A <= 1024, D % 32 == 0, D <= 2048, K <= 8.
struct AttMmaShape {
    int row1, row3;        // row pitch in bytes of a projected / raw feature row (global == smem: rows are padded by 16 B)
    int stage_bytes;       // max(16*row1, 8*row3)
    int rp;                // R rounded up to 16
};
__host__ __device__ inline AttMmaShape att_mma_shape(int R, int ld_enc, int ld_feats) {
    AttMmaShape s;
    s.row1 = ld_enc * 2;
    s.row3 = ld_feats * 2;
    s.stage_bytes = 16 * s.row1 > 8 * s.row3 ? 16 * s.row1 : 8 * s.row3;
    s.rp = (R + 15) / 16 * 16;
    return s;
}
constexpr int ATT_CG = 3;  // 16-row score chunks accumulated in registers between two cross-warp reductions (R <= 48: one)
template <int KR, int CTAS> struct AttMmaCfg {
    static constexpr int CTAS_PER_SM = CTAS;
    static constexpr int STAGES = CTAS == 2 ? 3 : (KR <= 5 ? 6 : 5);
    static constexpr int THREADS = 288;
};
__host__ __device__ inline size_t att_mma_smem_bytes(int KR, int stages, int R, int A, int ld_enc, int ld_feats) {
    const AttMmaShape s = att_mma_shape(R, ld_enc, ld_feats);
    return static_cast<size_t>(stages) * s.stage_bytes + 2 * stages * 8 + static_cast<size_t>(KR) * (A + 8) * 2 +
           static_cast<size_t>(KR) * s.rp * 4 + 8 * KR * ATT_CG * 16 * 4 + 8 * (s.rp + 8) * 2 + static_cast<size_t>(A) * 2 + 128;
}

// relu(x + d) on two fp16 lanes in ONE instruction: fma.rn.relu(x, 1, d) -- the product by 1 is exact, so the result is
// the correctly rounded sum, clamped at 0 (HFMA2.RELU instead of HADD2 + HMNMX2)
__device__ __forceinline__ uint32_t relu_add_h2(uint32_t x, uint32_t d) {
    uint32_t r;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0x3C003C00u), "r"(d));
    return r;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 r = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// plain spin on an mbarrier phase (try_wait suspends in hardware); traps instead of hanging after ~2^26 failed tries
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
    int tries = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++tries > (1 << 26)) __trap();
    }
}

// FULL: A == 1024 and D == 2048 (the reference's dims): every warp owns 8 score tiles and 16 context tiles, known at
// compile time (no per-tile predicates).
template <int KR, int CTAS, bool FULL>
__global__ void __launch_bounds__(288, CTAS)
butd_attention_mma_kernel(const __half* __restrict__ enc16, int ld_enc, const __half* __restrict__ feats16, int ld_feats,
                          size_t total_rows, const float* __restrict__ dec_ctx, const float* __restrict__ w_aff, float b_aff, int B,
                          int R, int A, int D, int K, __half* __restrict__ ctx16, int ld16, float* __restrict__ alphas_out,
                          size_t alpha_stride, int* __restrict__ done_ctr) {
    constexpr int STAGES = AttMmaCfg<KR, CTAS>::STAGES;
    const AttMmaShape sh = att_mma_shape(R, ld_enc, ld_feats);
    extern __shared__ __align__(128) uint8_t att_smem[];
    uint8_t* smem = att_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * sh.stage_bytes);
    __half* s_dec16 = reinterpret_cast<__half*>(bars + 2 * STAGES);          // [KR][A+8]
    float* s_e = reinterpret_cast<float*>(s_dec16 + KR * (A + 8));           // [KR][rp]
    float* s_part = s_e + KR * sh.rp;                                        // [8 warps][KR][ATT_CG*16]
    __half* s_alpha = reinterpret_cast<__half*>(s_part + 8 * KR * ATT_CG * 16);  // [8][rp+8]
    __half* s_w16 = s_alpha + 8 * (sh.rp + 8);                               // [A]
    const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + STAGES);
    const uint32_t ring = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks1 = (R + 15) / 16, n_chunks3 = (R + 7) / 8;

    // tail rows of a chunk (beyond the image's R regions) hold whatever followed in HBM / an earlier chunk: finite
    // fp16 data that is multiplied by alpha = 0 or ignored -- but never uninitialised bits
    for (int i = threadIdx.x; i < STAGES * sh.stage_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 8 * (sh.rp + 8) / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(s_alpha)[i] = 0u;
    for (int i = threadIdx.x; i < A; i += blockDim.x) s_w16[i] = __float2half_rn(w_aff[i]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 8);
        }
        fence_barrier_init();
    }
    fence_proxy_async_smem();  // the zero fill above (generic proxy) is ordered before the bulk copies (async proxy)
    __syncthreads();
    griddep_launch();
    griddep_wait();  // dec_ctx (and, on the first step, the projected features) come from earlier kernels of the stream

    if (warp == 0) {
        // ===================== producer: ONE contiguous bulk copy per chunk (16 / 8 padded rows) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t pol = l2_policy_evict_first();
            for (int img = blockIdx.x; img < B; img += gridDim.x) {
                for (int c = 0; c < n_chunks1 + n_chunks3; ++c) {
                    const bool p1 = c < n_chunks1;
                    const int r0 = p1 ? c * 16 : (c - n_chunks1) * 8;
                    const size_t row0 = static_cast<size_t>(img) * R + r0;
                    // only the image's own rows: the last chunk of a phase is short (R = 36: 16 + 16 + 4 and 4 x 8 + 4 rows);
                    // the rest of its slot keeps finite fp16 values of an earlier chunk, which get alpha = 0 / are ignored
                    size_t nr = p1 ? 16 : 8;
                    if (r0 + static_cast<int>(nr) > R) nr = R - r0;
                    const uint32_t bytes = static_cast<uint32_t>(nr) * (p1 ? sh.row1 : sh.row3);
                    const __half* src = p1 ? enc16 + row0 * ld_enc : feats16 + row0 * ld_feats;
                    mbar_wait_lean(empty_bar + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(full_bar + 8 * stage, bytes);
                    bulk_load_1d_hint(ring + stage * sh.stage_bytes, src, bytes, full_bar + 8 * stage, pol);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
        }
        return;
    }

    // ===================== consumers (256 threads, 8 warps) =====================
    const int tid = threadIdx.x - 32, cw = warp - 1;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t wmask = g == 0 ? 0xFFFFFFFFu : 0u;  // B fragment of the w-dot: w in output column 0 only
    // loop-invariant per-thread shared addresses (bytes).  Warp cw owns column tiles cw, cw+8, ... (phase 1) and the
    // 16 d-tiles [16cw, 16cw+16) (phase 3); tile j of a warp is a compile-time offset from these bases.
    const int nj1 = FULL ? 8 : max(0, min(8, ((A >> 4) - cw + 7) >> 3));   // valid phase-1 tiles of this warp
    const int nj3 = FULL ? 16 : max(0, min(16, (D >> 4) - cw * 16));      // valid phase-3 tiles of this warp
    const uint32_t w_addr = smem_u32(s_w16) + (cw * 16 + 2 * t) * 2;    // + 256*j (+16 for the upper 8 columns)
    const uint32_t dec_addr = smem_u32(s_dec16) + (cw * 16 + 2 * t) * 2;  // + k*(A+8)*2 + 256*j (+16)
    const uint32_t dec_pitch = (A + 8) * 2;
    const uint32_t ld1_off = ((lane & 7) + ((lane >> 3) & 1) * 8) * sh.row1 + (cw * 16 + (lane >> 4) * 8) * 2;  // + 256*j
    const uint32_t ld3_off = (lane & 7) * sh.row3 + ((cw * 16 + (lane >> 4)) * 16 + ((lane >> 3) & 1) * 8) * 2;  // + 64*jp
    const uint32_t alpha_addr = smem_u32(s_alpha) + (g * (sh.rp + 8) + 2 * t) * 2;                              // + 16*c
    float4 dn[KR];  // this thread's 4 columns of the next image's dec_att rows
    const bool dec_ok = tid * 4 < A;
    auto load_dec = [&](int img) {
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            dn[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dec_ok && k < K && img < B)
                dn[k] = __ldg(reinterpret_cast<const float4*>(dec_ctx + (static_cast<size_t>(img) * K + k) * A) + tid);
        }
    };
    load_dec(blockIdx.x);
    int stage = 0;
    uint32_t phase = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x) {
        if (dec_ok) {
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                uint2 v;
                v.x = pack_h2(dn[k].x, dn[k].y);
                v.y = pack_h2(dn[k].z, dn[k].w);
                *reinterpret_cast<uint2*>(s_dec16 + k * (A + 8) + tid * 4) = v;
            }
        }
        named_bar_sync(1, 256);
        // ---------------- phase 1: scores.  Each warp sums its 8 column tiles of up to ATT_CG row chunks in registers;
        // the 8 warps' partial sums meet in shared memory once per chunk group (once per image for R <= 48)
        const bool one_group = n_chunks1 <= ATT_CG;
        for (int c0 = 0; c0 < n_chunks1; c0 += ATT_CG) {
            float acc[ATT_CG][KR][4];
#pragma unroll
            for (int cc = 0; cc < ATT_CG; ++cc)
#pragma unroll
                for (int k = 0; k < KR; ++k) acc[cc][k][0] = acc[cc][k][1] = acc[cc][k][2] = acc[cc][k][3] = 0.f;
#pragma unroll
            for (int cc = 0; cc < ATT_CG; ++cc) {
                if (c0 + cc < n_chunks1) {
                    mbar_wait_lean(full_bar + 8 * stage, phase);
                    const uint32_t base1 = ring + stage * sh.stage_bytes + ld1_off;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (FULL || j < nj1) {
                            uint32_t x[4];
                            ldmatrix_x4(x, base1 + 256 * j);
                            const uint32_t wb0 = lds_b32(w_addr + 256 * j) & wmask;
                            const uint32_t wb1 = lds_b32(w_addr + 256 * j + 16) & wmask;
#pragma unroll
                            for (int k = 0; k < KR; ++k) {
                                const uint32_t dlo = lds_b32(dec_addr + k * dec_pitch + 256 * j);
                                const uint32_t dhi = lds_b32(dec_addr + k * dec_pitch + 256 * j + 16);
                                mma_m16n8k16_f16(acc[cc][k], relu_add_h2(x[0], dlo), relu_add_h2(x[1], dlo), relu_add_h2(x[2], dhi),
                                                 relu_add_h2(x[3], dhi), wb0, wb1);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
                    if (++stage == STAGES) stage = 0, phase ^= 1;
                }
            }
            if (t == 0) {
#pragma unroll
                for (int cc = 0; cc < ATT_CG; ++cc)
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        s_part[(cw * KR + k) * (ATT_CG * 16) + cc * 16 + g] = acc[cc][k][0];
                        s_part[(cw * KR + k) * (ATT_CG * 16) + cc * 16 + g + 8] = acc[cc][k][2];
                    }
            }
            if (c0 == 0) load_dec(img + gridDim.x);
            named_bar_sync(1, 256);
            if (!one_group) {  // long region lists: fold this group's partial sums into s_e, then reuse the buffer
                for (int i = tid; i < KR * ATT_CG * 16; i += 256) {
                    const int k = i / (ATT_CG * 16), q = i - k * (ATT_CG * 16);
                    float sum = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < 8; ++w8) sum += s_part[(w8 * KR + k) * (ATT_CG * 16) + q];
                    const int r = c0 * 16 + q;
                    if (r < R && k < K) s_e[k * sh.rp + r] = sum + b_aff;
                }
                named_bar_sync(1, 256);
            }
        }
        // ---------------- phase 2: softmax over regions -> fp16 alpha rows (B operand of phase 3)
        for (int k = cw; k < K; k += 8) {
            if (one_group) {  // the reduction over the 8 warps' partial sums is folded into the softmax warp's loads
                for (int r = lane; r < R; r += 32) {
                    float sum = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < 8; ++w8) sum += s_part[(w8 * KR + k) * (ATT_CG * 16) + r];
                    s_e[k * sh.rp + r] = sum + b_aff;
                }
                __syncwarp();
            }
            float m = -INFINITY;
            for (int r = lane; r < R; r += 32) m = fmaxf(m, s_e[k * sh.rp + r]);
            m = warp_max(m);
            float s = 0.f;
            for (int r = lane; r < R; r += 32) {
                const float ex = __expf(s_e[k * sh.rp + r] - m);
                s_e[k * sh.rp + r] = ex;
                s += ex;
            }
            s = warp_sum(s);
            const float inv = 1.0f / s;
            for (int r = lane; r < R; r += 32) {
                const float al = s_e[k * sh.rp + r] * inv;
                s_alpha[k * (sh.rp + 8) + r] = __float2half_rn(al);
                if (alphas_out) alphas_out[(static_cast<size_t>(img) * K + k) * alpha_stride + r] = al;
            }
        }
        named_bar_sync(1, 256);
        // ---------------- phase 3: ctx = alpha * feats
        float acc3[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc3[j][0] = acc3[j][1] = acc3[j][2] = acc3[j][3] = 0.f;
        for (int c = 0; c < n_chunks3; ++c) {
            mbar_wait_lean(full_bar + 8 * stage, phase);
            const uint32_t b0 = lds_b32(alpha_addr + 16 * c);
            const uint32_t base3 = ring + stage * sh.stage_bytes + ld3_off;
#pragma unroll
            for (int jp = 0; jp < 8; ++jp) {
                if (FULL || 2 * jp < nj3) {
                    uint32_t m4[4];
                    ldmatrix_x4_trans(m4, base3 + 64 * jp);
                    mma_m16n8k8_f16(acc3[2 * jp], m4[0], m4[1], b0);
                    mma_m16n8k8_f16(acc3[2 * jp + 1], m4[2], m4[3], b0);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
            if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        {
            // c0/c2: beam 2t, columns g / g+8 of the tile; c1/c3: beam 2t+1
            __half* o0 = ctx16 + (static_cast<size_t>(img) * K + 2 * t) * ld16 + cw * 256 + g;
            __half* o1 = o0 + ld16;
            const bool b0ok = 2 * t < K, b1ok = 2 * t + 1 < K;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (FULL || j < nj3) {
                    if (b0ok) {
                        o0[16 * j] = __float2half_rn(acc3[j][0]);
                        o0[16 * j + 8] = __float2half_rn(acc3[j][2]);
                    }
                    if (b1ok) {
                        o1[16 * j] = __float2half_rn(acc3[j][1]);
                        o1[16 * j + 8] = __float2half_rn(acc3[j][3]);
                    }
                }
            }
        }
        named_bar_sync(1, 256);  // s_dec16 / s_e / s_alpha are rewritten by the next image
    }
    // Small-batch path: a GEMM kernel that runs CONCURRENTLY with this one (smallm.cuh) waits for the context rows through
    // this counter -- images finished, cumulative over the decode.  The 256 consumers' stores happen-before the barrier above.
    if (done_ctr && tid == 0) {
        int n = 0;
        for (int img = blockIdx.x; img < B; img += gridDim.x) ++n;
        if (n) asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(done_ctr), "r"(n) : "memory");
    }
}

// ================================================================================================ AoA pieces
// LayerNorm of the reference (AoA_Model.py:14-25): gain*(x-mean)/(std_unbiased + eps) + bias.  One warp per row;
// writes the fp16 operand (query for linear_Q and for the AoA gate GEMM) and/or an fp32 copy (refined features).
__global__ void aoa_layernorm_kernel(const float* __restrict__ h, int M, int H, const float* __restrict__ gain,
                                     const float* __restrict__ bias, float eps, __half* __restrict__ q16, int ld16, int lo16,
                                     float* __restrict__ out32 = nullptr) {
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* x = h + static_cast<size_t>(row) * H;
    float s = 0.f;
    for (int i = lane; i < H; i += 32) s += x[i];
    const float mean = warp_sum(s) / static_cast<float>(H);
    float v = 0.f;
    for (int i = lane; i < H; i += 32) {
        const float d = x[i] - mean;
        v += d * d;
    }
    const float sd = sqrtf(warp_sum(v) / static_cast<float>(H - 1));
    const float inv = 1.0f / (sd + eps);
    for (int i = lane; i < H; i += 32) {
        const float y = gain[i] * (x[i] - mean) * inv + bias[i];
        if (out32) out32[static_cast<size_t>(row) * H + i] = y;
        if (q16) {
            __half hi, lo;
            split_f16(y, hi, lo);
            q16[static_cast<size_t>(row) * ld16 + i] = hi;
            if (lo16 > 0) q16[static_cast<size_t>(row) * ld16 + lo16 + i] = lo;
        }
    }
}

// Same LayerNorm for H = 128 * NV4: the row lives in registers (one global read), 16-byte loads, 8-byte fp16 stores.
template <int NV4>
__global__ void __launch_bounds__(256) aoa_layernorm_vec_kernel(const float* __restrict__ h, int M, const float* __restrict__ gain,
                                                                const float* __restrict__ bias, float eps, __half* __restrict__ q16,
                                                                int ld16, int lo16, float* __restrict__ out32) {
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    constexpr int H = 128 * NV4;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float4* x4 = reinterpret_cast<const float4*>(h + static_cast<size_t>(row) * H);
    float4 v[NV4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        v[i] = x4[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / static_cast<float>(H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        v[i].x -= mean, v[i].y -= mean, v[i].z -= mean, v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float inv = 1.0f / (sqrtf(warp_sum(q) / static_cast<float>(H - 1)) + eps);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int c = 4 * (lane + 32 * i);
        const float4 gn = __ldg(reinterpret_cast<const float4*>(gain + c)), bs = __ldg(reinterpret_cast<const float4*>(bias + c));
        const float y0 = gn.x * v[i].x * inv + bs.x, y1 = gn.y * v[i].y * inv + bs.y, y2 = gn.z * v[i].z * inv + bs.z,
                    y3 = gn.w * v[i].w * inv + bs.w;
        if (out32) *reinterpret_cast<float4*>(out32 + static_cast<size_t>(row) * H + c) = make_float4(y0, y1, y2, y3);
        if (q16) {
            __align__(8) __half hi[4];
            __align__(8) __half lo[4];
            split_f16(y0, hi[0], lo[0]);
            split_f16(y1, hi[1], lo[1]);
            split_f16(y2, hi[2], lo[2]);
            split_f16(y3, hi[3], lo[3]);
            __half* d = q16 + static_cast<size_t>(row) * ld16 + c;
            *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
            if (lo16 > 0) *reinterpret_cast<uint2*>(d + lo16) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// ------------------------------------------------------------------------------------------------ AoA refiner
// Multi-head SELF-attention of one AoA_Refine_Block (AoA_Model.py:41-69, 90-117, 136-138): queries = keys = values =
// the R regions of an image, per head  S = Q K^T / sqrt(d), masked_fill(mask == 0, -1e9), softmax over keys, X = P V.
// q/k/v are column blocks of the fused projection output qkv [B*R, 3H] (Q | K | V).  Output: the "att" half of the
// gate GEMM's operand [att | LN(x)].
//
// Generic form (fp32-grade mode, or head dims the fragment kernel does not take): one CTA per (image, head), K and V of
// the head in shared memory as fp32, one warp per query row: lanes <-> keys for the scores, lanes <-> columns for P V.
template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_as_float<__half>(const __half* p) { return __half2float(*p); }

template <typename T>
__global__ void __launch_bounds__(128) refine_attention_kernel(const T* __restrict__ qkv, int ld, const float* __restrict__ mask,
                                                               int R, int H, int nh, __half* __restrict__ x16, int ld16,
                                                               int lo16) {
    extern __shared__ float rsm[];
    const int d = H / nh;
    const int img = blockIdx.x / nh, hd = blockIdx.x - img * nh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    float* s_k = rsm;                     // [R][d+1]
    float* s_v = s_k + R * (d + 1);       // [R][d]
    float* s_q = s_v + R * d;             // [nwarp][d]
    float* s_p = s_q + nwarp * d;         // [nwarp][R]
    const T* base = qkv + static_cast<size_t>(img) * R * ld + hd * d;
    for (int i = threadIdx.x; i < R * d; i += blockDim.x) {
        const int r = i / d, c = i - r * d;
        s_k[r * (d + 1) + c] = ld_as_float<T>(base + static_cast<size_t>(r) * ld + H + c);
        s_v[r * d + c] = ld_as_float<T>(base + static_cast<size_t>(r) * ld + 2 * H + c);
    }
    __syncthreads();
    const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(d));
    float* q = s_q + warp * d;
    float* p = s_p + warp * R;
    for (int qr = warp; qr < R; qr += nwarp) {
        for (int c = lane; c < d; c += 32) q[c] = ld_as_float<T>(base + static_cast<size_t>(qr) * ld + c);
        __syncwarp();
        float m = -INFINITY;
        for (int r = lane; r < R; r += 32) {
            float acc = 0.f;
            const float* kr = s_k + r * (d + 1);
            for (int c = 0; c < d; ++c) acc = fmaf(q[c], kr[c], acc);
            acc *= inv_sqrt_d;
            if (mask && mask[static_cast<size_t>(img) * R + r] == 0.f) acc = -1e9f;
            p[r] = acc;
            m = fmaxf(m, acc);
        }
        m = warp_max(m);
        float sum = 0.f;
        for (int r = lane; r < R; r += 32) {
            const float ex = expf(p[r] - m);
            p[r] = ex;
            sum += ex;
        }
        sum = warp_sum(sum);
        __syncwarp();
        for (int c = lane; c < d; c += 32) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(p[r] / sum, s_v[r * d + c], acc);
            __half hi, lo;
            split_f16(acc, hi, lo);
            __half* o = x16 + (static_cast<size_t>(img) * R + qr) * ld16 + hd * d + c;
            *o = hi;
            if (lo16 > 0) o[lo16] = lo;
        }
        __syncwarp();
    }
}

// Fragment form (fp16 mode).  A CTA works on (image, group of G heads) items: the K and V slices of the group
// ([R, G*DH] fp16 each, 1 KB contiguous per row) are copied with cp.async into per-head shared-memory tiles (rows padded
// by 16 B -> conflict-free ldmatrix); one warp per (head, 16-query tile), so an SM holds 24 warps and the per-warp
// dependency chains overlap.  Per warp:  S = Q K^T  (m16n8k16: A = Q fragments read straight from global memory while
// the copies are in flight, B = K rows via ldmatrix), softmax over the keys on the accumulator fragments (row statistics
// via quad shuffles), P re-packed in registers as the A operand of  X = P V  (B = V via ldmatrix.trans).
// NKT = ceil(R / 16) key tiles = query tiles, DH = head dim.
template <int NKT, int DH, int G>
struct RefineMmaCfg {
    static constexpr int ROWS = 16 * NKT;
    static constexpr int LDS = DH + 8;                        // halves per shared-memory row
    static constexpr int TILE_BYTES = ROWS * LDS * 2;         // one head's K (or V) tile
    static constexpr int SMEM_BYTES = 2 * G * TILE_BYTES;     // K tiles then V tiles
    static constexpr int WARPS = NKT * G;
    // small key counts: many small CTAs per SM (up to 24 warps under the 64 K register file) so that the copy phase of
    // one CTA overlaps the MMA phase of its neighbours; larger ones need the registers for the score fragments
    static constexpr int SMEM_LIMIT = (227 * 1024) / (SMEM_BYTES + 1024);
    static constexpr int CTAS_PER_SM = NKT <= 3 ? (SMEM_LIMIT < 24 / WARPS ? SMEM_LIMIT : 24 / WARPS) : 1;
};

template <int NKT, int DH, int G>
__global__ void __launch_bounds__(32 * RefineMmaCfg<NKT, DH, G>::WARPS, RefineMmaCfg<NKT, DH, G>::CTAS_PER_SM)
refine_attention_mma_kernel(const __half* __restrict__ qkv, int ld, const float* __restrict__ mask, int B, int R, int H, int nh,
                            __half* __restrict__ x16, int ld16) {
    using C = RefineMmaCfg<NKT, DH, G>;
    extern __shared__ __align__(128) uint8_t rf_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int hl = warp / NKT, mt = warp - hl * NKT;  // head within the group, query tile
    const uint32_t s_base = smem_u32(rf_smem);
    const uint32_t sk = s_base + hl * C::TILE_BYTES, sv = s_base + (G + hl) * C::TILE_BYTES;
    // rows R .. ROWS-1 are never written by the copies: zero everything once (0 * garbage must not become NaN)
    for (int i = threadIdx.x; i < C::SMEM_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(rf_smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // ldmatrix lane offsets: "a" = 16x16 block in A-operand order (V^T via .trans); "b" = two 8-key groups as the B operand
    const uint32_t off_a = (((lane & 7) + ((lane >> 3) & 1) * 8) * C::LDS + (lane >> 4) * 8) * 2;
    const uint32_t off_b = (((lane & 7) + ((lane >> 4) & 1) * 8) * C::LDS + ((lane >> 3) & 1) * 8) * 2;
    const float scale = rsqrtf(static_cast<float>(DH)) * LOG2E;  // exp(x) = exp2(x * log2 e)
    constexpr int CH = DH / 8;                                     // 16-byte chunks per head row
    const int groups = nh / G;
    const int items = B * groups;
    const int q0 = mt * 16 + g, q1 = q0 + 8;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int img = item / groups, h0 = (item - img * groups) * G;
        const __half* src = qkv + static_cast<size_t>(img) * R * ld + h0 * DH;
        for (int i = threadIdx.x; i < 2 * R * G * CH; i += blockDim.x) {  // K then V rows of the head group
            const int part = i >= R * G * CH;
            const int rem = i - part * R * G * CH;
            const int r = rem / (G * CH), c = rem - r * (G * CH);
            const int hh = c / CH, cc = c - hh * CH;
            cp_async_16(s_base + ((part * G + hh) * C::TILE_BYTES) + (r * C::LDS + cc * 8) * 2,
                        src + static_cast<size_t>(r) * ld + (1 + part) * H + c * 8);
        }
        cp_async_commit();
        // this warp's Q fragments (rows q0 / q1 of head h0 + hl), straight from global memory
        uint32_t qa[DH / 16][4];
        {
            const __half* qr0 = src + static_cast<size_t>(q0) * ld + hl * DH + 2 * t;
            const __half* qr1 = src + static_cast<size_t>(q1) * ld + hl * DH + 2 * t;
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
                qa[ks][0] = q0 < R ? __ldg(reinterpret_cast<const uint32_t*>(qr0 + ks * 16)) : 0u;
                qa[ks][1] = q1 < R ? __ldg(reinterpret_cast<const uint32_t*>(qr1 + ks * 16)) : 0u;
                qa[ks][2] = q0 < R ? __ldg(reinterpret_cast<const uint32_t*>(qr0 + ks * 16 + 8)) : 0u;
                qa[ks][3] = q1 < R ? __ldg(reinterpret_cast<const uint32_t*>(qr1 + ks * 16 + 8)) : 0u;
            }
        }
        // key status bits of this lane's columns: bit (2*nt + e) <-> key nt*8 + 2t + e
        uint64_t in_range = 0, keep = 0;
#pragma unroll
        for (int nt = 0; nt < 2 * NKT; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = nt * 8 + 2 * t + e;
                if (key < R) {
                    in_range |= 1ull << (2 * nt + e);
                    if (!mask || __ldg(mask + static_cast<size_t>(img) * R + key) != 0.f) keep |= 1ull << (2 * nt + e);
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (mt * 16 < R) {
            float s[2 * NKT][4];
#pragma unroll
            for (int nt = 0; nt < 2 * NKT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
                for (int kt = 0; kt < NKT; ++kt) {
                    uint32_t b[4];
                    ldmatrix_x4(b, sk + (kt * 16 * C::LDS + ks * 16) * 2 + off_b);
                    mma_m16n8k16_f16(s[2 * kt], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[0], b[1]);
                    mma_m16n8k16_f16(s[2 * kt + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b[2], b[3]);
                }
            }
            // softmax over the keys for rows g (elements 0,1) and g+8 (elements 2,3), in the log2 domain
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 2 * NKT; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool in = (in_range >> (2 * nt + e)) & 1, kp = (keep >> (2 * nt + e)) & 1;
                    const float lo = in ? -1e9f * LOG2E : -INFINITY;  // masked_fill(mask == 0, -1e9) / key beyond R
                    s[nt][e] = (in && kp) ? s[nt][e] * scale : lo;
                    s[nt][2 + e] = (in && kp) ? s[nt][2 + e] * scale : lo;
                    m0 = fmaxf(m0, s[nt][e]);
                    m1 = fmaxf(m1, s[nt][2 + e]);
                }
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < 2 * NKT; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    s[nt][e] = exp2f(s[nt][e] - m0);
                    s[nt][2 + e] = exp2f(s[nt][2 + e] - m1);
                    sum0 += s[nt][e];
                    sum1 += s[nt][2 + e];
                }
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
            float o[DH / 8][4];
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < NKT; ++kt) {
                const uint32_t a0 = pack_h2(s[2 * kt][0] * inv0, s[2 * kt][1] * inv0);
                const uint32_t a1 = pack_h2(s[2 * kt][2] * inv1, s[2 * kt][3] * inv1);
                const uint32_t a2 = pack_h2(s[2 * kt + 1][0] * inv0, s[2 * kt + 1][1] * inv0);
                const uint32_t a3 = pack_h2(s[2 * kt + 1][2] * inv1, s[2 * kt + 1][3] * inv1);
#pragma unroll
                for (int dn = 0; dn < DH / 16; ++dn) {
                    uint32_t b[4];
                    ldmatrix_x4_trans(b, sv + (kt * 16 * C::LDS + dn * 16) * 2 + off_a);
                    mma_m16n8k16_f16(o[2 * dn], a0, a1, a2, a3, b[0], b[1]);
                    mma_m16n8k16_f16(o[2 * dn + 1], a0, a1, a2, a3, b[2], b[3]);
                }
            }
            __half* o0 = x16 + (static_cast<size_t>(img) * R + q0) * ld16 + (h0 + hl) * DH + 2 * t;
            __half* o1 = o0 + static_cast<size_t>(8) * ld16;
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) {
                if (q0 < R) *reinterpret_cast<uint32_t*>(o0 + 8 * j) = pack_h2(o[j][0], o[j][1]);
                if (q1 < R) *reinterpret_cast<uint32_t*>(o1 + 8 * j) = pack_h2(o[j][2], o[j][3]);
            }
        }
        __syncthreads();  // every warp is done with the tiles before the next item's copies land
    }
}

// Multi-head dot-product attention of the decoder AoA block for the K rows of one image (AoA_Model.py:41-69,
// 90-117) over the one-time K,V projections kv [B*R, 2H] (K in columns [0,H), V in [H,2H)):
//   s[k,h,r] = Q[k,h,:].K[r,h,:] / sqrt(d), masked_fill(mask==0, -1e9), softmax over r, x[k,h,:] = sum_r p V[r,h,:]
// Requires (H/heads)/4 to be a power of two.
template <int KR>
__global__ void __launch_bounds__(256) aoa_attention_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                            const float* __restrict__ mask, int R, int H, int nh, int K,
                                                            __half* __restrict__ x16, int ld16, int lo16,
                                                            float* __restrict__ alphas_out, size_t alpha_stride) {
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    extern __shared__ float sm[];
    float* s_q = sm;               // [KR][H]
    float* s_p = s_q + KR * H;     // [KR][nh][R]
    const int img = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const int d = H / nh;
    const int G = d / 4;  // float4 per head
    const float inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(d));

    for (int i = tid; i < K * H; i += blockDim.x) s_q[i] = q[static_cast<size_t>(img) * K * H + i];
    __syncthreads();

    const float* kvb = kv + static_cast<size_t>(img) * R * 2 * H;
    const int nf4 = H / 4;
    for (int r = warp; r < R; r += nwarp) {
        const float* kr = kvb + static_cast<size_t>(r) * 2 * H;
        const bool masked = mask && mask[static_cast<size_t>(img) * R + r] == 0.f;
        if (G >= 32) {
            for (int hd = 0; hd < nh; ++hd) {
                float acc[KR];
#pragma unroll
                for (int k = 0; k < KR; ++k) acc[k] = 0.f;
                for (int i = hd * G + lane; i < (hd + 1) * G; i += 32) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(kr) + i);
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        if (k < K) {
                            const float4 qq = *reinterpret_cast<const float4*>(s_q + k * H + 4 * i);
                            acc[k] += x.x * qq.x + x.y * qq.y + x.z * qq.z + x.w * qq.w;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    const float t = warp_sum(acc[k]);
                    if (lane == 0 && k < K) s_p[(k * nh + hd) * R + r] = masked ? -1e9f : t * inv_sqrt_d;
                }
            }
        } else {
            for (int i0 = 0; i0 < nf4; i0 += 32) {
                const int i = i0 + lane;
                float acc[KR];
#pragma unroll
                for (int k = 0; k < KR; ++k) acc[k] = 0.f;
                if (i < nf4) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(kr) + i);
#pragma unroll
                    for (int k = 0; k < KR; ++k) {
                        if (k < K) {
                            const float4 qq = *reinterpret_cast<const float4*>(s_q + k * H + 4 * i);
                            acc[k] = x.x * qq.x + x.y * qq.y + x.z * qq.z + x.w * qq.w;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    float t = acc[k];
                    for (int o = G >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (i < nf4 && (lane % G) == 0 && k < K) s_p[(k * nh + i / G) * R + r] = masked ? -1e9f : t * inv_sqrt_d;
                }
            }
        }
    }
    __syncthreads();

    for (int kh = warp; kh < K * nh; kh += nwarp) {  // softmax over regions per (row, head)
        float* p = s_p + static_cast<size_t>(kh) * R;
        float m = -INFINITY;
        for (int r = lane; r < R; r += 32) m = fmaxf(m, p[r]);
        m = warp_max(m);
        float s = 0.f;
        for (int r = lane; r < R; r += 32) {
            const float ex = expf(p[r] - m);
            p[r] = ex;
            s += ex;
        }
        s = warp_sum(s);
        for (int r = lane; r < R; r += 32) p[r] = p[r] / s;
    }
    __syncthreads();
    if (alphas_out) {  // mean over heads (AoA_Model.py:119)
        for (int i = tid; i < K * R; i += blockDim.x) {
            const int k = i / R, r = i - k * R;
            float a = 0.f;
            for (int hh = 0; hh < nh; ++hh) a += s_p[(k * nh + hh) * R + r];
            alphas_out[(static_cast<size_t>(img) * K + k) * alpha_stride + r] = a / static_cast<float>(nh);
        }
    }

    for (int i = tid; i < nf4; i += blockDim.x) {
        const int hd = i / G;
        float4 acc[KR];
#pragma unroll
        for (int k = 0; k < KR; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(kvb + static_cast<size_t>(r) * 2 * H + H) + i);
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                if (k < K) {
                    const float pr = s_p[(k * nh + hd) * R + r];
                    acc[k].x += pr * x.x, acc[k].y += pr * x.y, acc[k].z += pr * x.z, acc[k].w += pr * x.w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KR; ++k) {
            if (k < K) {
                __align__(8) __half hi[4];
                __align__(8) __half lo[4];
                split_f16(acc[k].x, hi[0], lo[0]);
                split_f16(acc[k].y, hi[1], lo[1]);
                split_f16(acc[k].z, hi[2], lo[2]);
                split_f16(acc[k].w, hi[3], lo[3]);
                __half* o = x16 + (static_cast<size_t>(img) * K + k) * ld16 + 4 * i;
                *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(hi);
                if (lo16 > 0) *reinterpret_cast<uint2*>(o + lo16) = *reinterpret_cast<const uint2*>(lo);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ AoA, fp16 fragment form
// Multi-head attention of the AoA decoder on MMA fragments behind the same persistent bulk-copy ring as the BUTD
// kernel (fp16 mode).  K and V projections are stored as separate fp16 matrices with padded rows (ld = H + 8).
// One warp owns one head (heads <= 8, head dim a multiple of 16):
//   phase 1  S[r, beam] = K[r, head, :] . Q[beam, head, :]   m16n8k16: A = K tile (ldmatrix), B = Q^T (beam = n)
//   phase 2  softmax over regions per (beam, head), mask -> -1e9 before it (AoA_Model.py:62-65)
//   phase 3  X[d, beam] = sum_r V[r, head, d] * P[beam, head, r]   A = V^T tile (ldmatrix.trans), B = P
__host__ __device__ inline size_t aoa_mma_fixed_smem(int KR, int R, int H, int nh) {
    const int rp = (R + 15) / 16 * 16;
    return 2 * 8 * 8 /*bars*/ + static_cast<size_t>(8) * (H + 8) * 2 /*q16*/ + static_cast<size_t>(KR) * nh * rp * 4 /*scores*/ +
           static_cast<size_t>(8) * nh * (rp + 8) * 2 /*p16*/ + static_cast<size_t>(rp) * 4 /*mask*/ + 128;
}

template <int KR>
__global__ void __launch_bounds__(288, 1)
aoa_attention_mma_kernel(const __half* __restrict__ k16, const __half* __restrict__ v16, int ld_kv, size_t total_rows,
                         const __half* __restrict__ q16, int ld_q, const float* __restrict__ mask, int B, int R, int H, int nh,
                         int K, int stages, __half* __restrict__ x16, int ld16, float* __restrict__ alphas_out,
                         size_t alpha_stride) {
    const int row_bytes = ld_kv * 2;
    const int stage_bytes = 16 * row_bytes;
    const int rp = (R + 15) / 16 * 16;
    const int d = H / nh;
    extern __shared__ __align__(128) uint8_t att_smem[];
    uint8_t* smem = att_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    __half* s_q = reinterpret_cast<__half*>(bars + 16);                 // [8][H+8], rows >= K stay zero
    float* s_s = reinterpret_cast<float*>(s_q + 8 * (H + 8));           // [KR][nh][rp]
    __half* s_p = reinterpret_cast<__half*>(s_s + KR * nh * rp);        // [8][nh][rp+8], rows >= K stay zero
    float* s_mask = reinterpret_cast<float*>(s_p + 8 * nh * (rp + 8));  // [rp]
    const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + 8);
    const uint32_t ring = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (R + 15) / 16;

    for (int i = threadIdx.x; i < stages * stage_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < 8 * (H + 8) / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(s_q)[i] = 0u;
    for (int i = threadIdx.x; i < 8 * nh * (rp + 8) / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(s_p)[i] = 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 8);
        }
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    __syncthreads();
    griddep_launch();
    griddep_wait();  // q16 comes from the preceding projection GEMM

    if (warp == 0) {
        if (lane == 0) {  // producer: K chunks then V chunks of every image, one contiguous copy each
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t pol = l2_policy_evict_first();
            for (int img = blockIdx.x; img < B; img += gridDim.x) {
                for (int c = 0; c < 2 * n_chunks; ++c) {
                    const bool kpart = c < n_chunks;
                    const int r0 = (kpart ? c : c - n_chunks) * 16;
                    const size_t row0 = static_cast<size_t>(img) * R + r0;
                    // only the image's own rows (R = 36: 16 + 16 + 4); the rest of a short chunk's slot keeps finite fp16
                    // values of an earlier chunk: their scores are never read and their probabilities are zero
                    const size_t nr = r0 + 16 > R ? R - r0 : 16;
                    const uint32_t bytes = static_cast<uint32_t>(nr) * row_bytes;
                    mbar_wait_lean(empty_bar + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx(full_bar + 8 * stage, bytes);
                    bulk_load_1d_hint(ring + stage * stage_bytes, (kpart ? k16 : v16) + row0 * ld_kv, bytes, full_bar + 8 * stage, pol);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        }
        return;
    }

    const int tid = threadIdx.x - 32, cw = warp - 1;
    const int g = lane >> 2, t = lane & 3;
    const bool has_head = cw < nh;
    const int hd = cw;                              // this warp's head
    const int nds = d >> 4;                         // 16-wide steps across the head dim (<= 16)
    const float inv_sqrt_d = rsqrtf(static_cast<float>(d));
    const uint32_t q_addr = smem_u32(s_q) + (g * (H + 8) + hd * d + 2 * t) * 2;            // + 32*ds (+16)
    const uint32_t ld1_off = ((lane & 7) + ((lane >> 3) & 1) * 8) * row_bytes + (hd * d + (lane >> 4) * 8) * 2;  // + 32*ds
    const uint32_t ld3_off = ((lane & 7) + ((lane >> 4) & 1) * 8) * row_bytes + (hd * d + ((lane >> 3) & 1) * 8) * 2;  // + 32*dt
    const uint32_t p_addr = smem_u32(s_p) + ((g * nh + hd) * (rp + 8) + 2 * t) * 2;         // + 32*c (+16)
    const int nq8 = K * H / 8;                      // uint4 vectors of this image's query rows
    int stage = 0;
    uint32_t phase = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x) {
        for (int i = tid; i < nq8; i += 256) {
            const int k = i / (H / 8), c8 = i - k * (H / 8);
            *reinterpret_cast<uint4*>(s_q + k * (H + 8) + c8 * 8) =
                __ldg(reinterpret_cast<const uint4*>(q16 + (static_cast<size_t>(img) * K + k) * ld_q) + c8);
        }
        for (int r = tid; r < rp; r += 256) s_mask[r] = (mask && r < R) ? mask[static_cast<size_t>(img) * R + r] : 1.f;
        named_bar_sync(1, 256);
        // ---------------- phase 1: scores
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait_lean(full_bar + 8 * stage, phase);
            if (has_head) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t base1 = ring + stage * stage_bytes + ld1_off;
#pragma unroll 4
                for (int ds = 0; ds < nds; ++ds) {
                    uint32_t x[4];
                    ldmatrix_x4(x, base1 + 32 * ds);
                    mma_m16n8k16_f16(acc, x[0], x[1], x[2], x[3], lds_b32(q_addr + 32 * ds), lds_b32(q_addr + 32 * ds + 16));
                }
                // c0/c1: region g, beams 2t/2t+1; c2/c3: region g+8
                const int r0 = c * 16 + g, r1 = r0 + 8;
                const float m0 = s_mask[r0], m1 = s_mask[r1];
                if (2 * t < K) {
                    s_s[((2 * t) * nh + hd) * rp + r0] = m0 == 0.f ? -1e9f : acc[0] * inv_sqrt_d;
                    s_s[((2 * t) * nh + hd) * rp + r1] = m1 == 0.f ? -1e9f : acc[2] * inv_sqrt_d;
                }
                if (2 * t + 1 < K) {
                    s_s[((2 * t + 1) * nh + hd) * rp + r0] = m0 == 0.f ? -1e9f : acc[1] * inv_sqrt_d;
                    s_s[((2 * t + 1) * nh + hd) * rp + r1] = m1 == 0.f ? -1e9f : acc[3] * inv_sqrt_d;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
            if (++stage == stages) stage = 0, phase ^= 1;
        }
        named_bar_sync(1, 256);
        // ---------------- phase 2: softmax over regions per (beam, head)
        for (int kh = cw; kh < K * nh; kh += 8) {
            float* sr = s_s + kh * rp;
            float m = -INFINITY;
            for (int r = lane; r < R; r += 32) m = fmaxf(m, sr[r]);
            m = warp_max(m);
            float sum = 0.f;
            for (int r = lane; r < R; r += 32) {
                const float ex = __expf(sr[r] - m);
                sr[r] = ex;
                sum += ex;
            }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int r = lane; r < R; r += 32) {
                const float pr = sr[r] * inv;
                sr[r] = pr;
                s_p[kh * (rp + 8) + r] = __float2half_rn(pr);
            }
        }
        named_bar_sync(1, 256);
        if (alphas_out) {  // attention map returned by the reference: mean over heads (AoA_Model.py:119)
            for (int i = tid; i < K * R; i += 256) {
                const int k = i / R, r = i - k * R;
                float a = 0.f;
                for (int hh = 0; hh < nh; ++hh) a += s_s[(k * nh + hh) * rp + r];
                alphas_out[(static_cast<size_t>(img) * K + k) * alpha_stride + r] = a / static_cast<float>(nh);
            }
        }
        // ---------------- phase 3: weighted value sum
        float acc3[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc3[j][0] = acc3[j][1] = acc3[j][2] = acc3[j][3] = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait_lean(full_bar + 8 * stage, phase);
            if (has_head) {
                const uint32_t b0 = lds_b32(p_addr + 32 * c), b1 = lds_b32(p_addr + 32 * c + 16);
                const uint32_t base3 = ring + stage * stage_bytes + ld3_off;
#pragma unroll
                for (int dt = 0; dt < 16; ++dt) {
                    if (dt < nds) {
                        uint32_t m4[4];
                        ldmatrix_x4_trans(m4, base3 + 32 * dt);
                        mma_m16n8k16_f16(acc3[dt], m4[0], m4[1], m4[2], m4[3], b0, b1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8 * stage);
            if (++stage == stages) stage = 0, phase ^= 1;
        }
        if (has_head) {
            __half* o0 = x16 + (static_cast<size_t>(img) * K + 2 * t) * ld16 + hd * d + g;
            __half* o1 = o0 + ld16;
            const bool b0ok = 2 * t < K, b1ok = 2 * t + 1 < K;
#pragma unroll
            for (int dt = 0; dt < 16; ++dt) {
                if (dt < nds) {
                    if (b0ok) {
                        o0[16 * dt] = __float2half_rn(acc3[dt][0]);
                        o0[16 * dt + 8] = __float2half_rn(acc3[dt][2]);
                    }
                    if (b1ok) {
                        o1[16 * dt] = __float2half_rn(acc3[dt][1]);
                        o1[16 * dt + 8] = __float2half_rn(acc3[dt][3]);
                    }
                }
            }
        }
        named_bar_sync(1, 256);
    }
}

// ================================================================================================ operand assembly
// After the bookkeeping of a step every row's next GEMM operands are rebuilt: recurrent states are gathered
// by parent row (the beam reorder of BUTD_Model.py:297-300), the chosen word is embedded (:264), AoA's
// mean + ctx input is formed (AoA_Model.py:441).
enum AdvKind { ADV_COPY16 = 0, ADV_EMBED = 1, ADV_MEAN_PLUS = 2, ADV_BCAST16 = 3 };
struct AdvOp {
    int kind;
    const void* src;   // COPY16/BCAST16: fp16 operand rows; EMBED: fp32 table [V,n]; MEAN_PLUS: fp32 [M,n] (may be null)
    int src_ld, src_lo;
    const float* aux;  // MEAN_PLUS: mean [B,n]
    __half* dst;
    int dst_ld, dst_lo;
    int n;             // elements per row
    int flag;          // EMBED: 1 = ReLU after the lookup (BUTD / AoA embed = Embedding+ReLU)
};
struct AdvOps {
    int n;
    AdvOp op[6];
};

// Rebuild the operand rows of all K slots of one image.  s_prow[slot] = absolute source (parent) row, s_tok[slot] =
// word fed to the next step (both in shared memory).  For every 16-byte vector the loads of all K slots are issued
// before the first store, so K (x2 with the lo halves) independent loads are in flight per thread.
template <int KR>
__device__ __forceinline__ void advance_image(const AdvOps& ops, int img, int K, const int* s_prow, const int* s_tok, int tid,
                                              int nthreads) {
    for (int q = 0; q < ops.n; ++q) {
        const AdvOp& o = ops.op[q];
        if (o.kind == ADV_COPY16 || o.kind == ADV_BCAST16) {
            const __half* src = static_cast<const __half*>(o.src);
            const bool lo = o.dst_lo > 0;
            for (int v = tid; v < (o.n >> 3); v += nthreads) {
                uint4 hi_v[KR], lo_v[KR];
#pragma unroll
                for (int s = 0; s < KR; ++s) {
                    if (s < K) {
                        const size_t srow = o.kind == ADV_COPY16 ? s_prow[s] : img;
                        hi_v[s] = reinterpret_cast<const uint4*>(src + srow * o.src_ld)[v];
                        if (lo) lo_v[s] = reinterpret_cast<const uint4*>(src + srow * o.src_ld + o.src_lo)[v];
                    }
                }
#pragma unroll
                for (int s = 0; s < KR; ++s) {
                    if (s < K) {
                        __half* dst = o.dst + static_cast<size_t>(img * K + s) * o.dst_ld;
                        reinterpret_cast<uint4*>(dst)[v] = hi_v[s];
                        if (lo) reinterpret_cast<uint4*>(dst + o.dst_lo)[v] = lo_v[s];
                    }
                }
            }
        } else {
            const bool embed = o.kind == ADV_EMBED;
            for (int v = tid; v < (o.n >> 2); v += nthreads) {
                float4 x[KR];
#pragma unroll
                for (int s = 0; s < KR; ++s) {
                    if (s < K) {
                        if (embed) {
                            x[s] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(o.src) +
                                                                         static_cast<size_t>(s_tok[s]) * o.src_ld) + v);
                        } else {  // ADV_MEAN_PLUS: mean[img] + ctx[parent row]
                            x[s] = __ldg(reinterpret_cast<const float4*>(o.aux + static_cast<size_t>(img) * o.n) + v);
                            if (o.src) {
                                const float4 c = reinterpret_cast<const float4*>(static_cast<const float*>(o.src) +
                                                                                 static_cast<size_t>(s_prow[s]) * o.src_ld)[v];
                                x[s].x += c.x, x[s].y += c.y, x[s].z += c.z, x[s].w += c.w;
                            }
                        }
                    }
                }
#pragma unroll
                for (int s = 0; s < KR; ++s) {
                    if (s < K) {
                        float4 y = x[s];
                        if (embed && o.flag) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
                        __align__(8) __half hi[4];
                        __align__(8) __half lw[4];
                        split_f16(y.x, hi[0], lw[0]);
                        split_f16(y.y, hi[1], lw[1]);
                        split_f16(y.z, hi[2], lw[2]);
                        split_f16(y.w, hi[3], lw[3]);
                        __half* dst = o.dst + static_cast<size_t>(img * K + s) * o.dst_ld + 4 * v;
                        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(hi);
                        if (o.dst_lo > 0) *reinterpret_cast<uint2*>(dst + o.dst_lo) = *reinterpret_cast<const uint2*>(lw);
                    }
                }
            }
        }
    }
}

// ================================================================================================ beam search
struct BeamState {
    int B, K, V, T;
    int* tok;          // [B*K] word fed to the next step
    float* cum;        // [B*K] cumulative log-prob of live slots (-inf = dead)
    int* parent;       // [B*K] absolute parent row of each slot (cell-state indirection)
    int* n_live;       // [B]
    int* seqs_in;      // [B*K, T+1]
    int* seqs_out;     // [B*K, T+1]
    float* best_score; // [B] best COMPLETED hypothesis so far (-inf = none)
    int* best_seq;     // [B, T+1]
    int* best_len;     // [B]
    int* hist_parent;  // [T, B*K] parent SLOT of every slot at every step (attention-map backtracking) or null
    int* best_pslot;   // [B] parent slot of the best completed hypothesis at its last step
};

// t = 0: initial state (all K slots <sta>, cum 0; BUTD_Model.py:247-250); no partials are read.
// parent_is_img: the first LSTM step reads a per-IMAGE cell state (NIC's primed c0), so parent[row] = image.
template <int KR>
__global__ void beam_init_kernel(BeamState s, AdvOps ops, int parent_is_img) {
    const int img = blockIdx.x;
    const int L = s.T + 1;
    for (int i = threadIdx.x; i < s.K * L; i += blockDim.x) {
        const int slot = i / L, pos = i - slot * L;
        const int v = pos == 0 ? TOK_STA : TOK_PAD;
        s.seqs_in[(static_cast<size_t>(img) * s.K + slot) * L + pos] = v;
        s.seqs_out[(static_cast<size_t>(img) * s.K + slot) * L + pos] = v;
    }
    for (int i = threadIdx.x; i < L; i += blockDim.x) s.best_seq[static_cast<size_t>(img) * L + i] = TOK_PAD;
    if (threadIdx.x < s.K) {
        const int row = img * s.K + threadIdx.x;
        s.tok[row] = TOK_STA;
        s.cum[row] = 0.f;
        s.parent[row] = parent_is_img ? img : row;
    }
    if (threadIdx.x == 0) {
        s.n_live[img] = s.K;
        s.best_score[img] = -INFINITY;
        s.best_len[img] = 0;
    }
    __shared__ int i_prow[MAX_ROWS], i_tok[MAX_ROWS];
    if (threadIdx.x < MAX_ROWS) i_prow[threadIdx.x] = img * s.K + threadIdx.x, i_tok[threadIdx.x] = TOK_STA;
    __syncthreads();
    advance_image<KR>(ops, img, s.K, i_prow, i_tok, threadIdx.x, blockDim.x);
}

// One CTA (128 threads) per image.  Merges the per-(row, N-tile) partials of the logit GEMM into log-softmax
// scores, takes the top-k over (live rows x vocabulary) and applies the reference's bookkeeping
// (BUTD_Model.py:271-302 == NIC_Model.py:175-202 == AoA_Model.py:456-488):
//   step 1 looks at row 0 only; selected <end> candidates complete (running best, strict '>' = first max) and
//   shrink the beam; survivors keep their sorted order; states follow their parent.
template <int KTOP, int KR>
__global__ void __launch_bounds__(128, KR <= 3 ? 12 : 8) beam_step_kernel(const float* __restrict__ part, int n_tiles, BeamState s, int t,
                                                        AdvOps ops, int* __restrict__ done_ctr) {
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    constexpr int PS = topk_part_stride(KTOP);
    __shared__ float c_val[MAX_ROWS][MAX_ROWS];
    __shared__ int c_idx[MAX_ROWS][MAX_ROWS];
    __shared__ int s_parent[MAX_ROWS], s_tok[MAX_ROWS];
    __shared__ int s_n_new, s_best_src;
    const int img = blockIdx.x;
    const int K = s.K, L = s.T + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nl = s.n_live[img];
    const int rows_considered = (t == 1) ? min(nl, 1) : nl;

    for (int r = warp; r < K; r += 4) {
        if (r >= rows_considered) {
            if (lane < K) c_val[r][lane] = -INFINITY, c_idx[r][lane] = 0x7FFFFFFF;
            continue;
        }
        const int row = img * K + r;
        const float* pr = part + static_cast<size_t>(row) * n_tiles * PS;
        float m = -INFINITY;
        for (int j = lane; j < n_tiles; j += 32) m = fmaxf(m, pr[j * PS]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < n_tiles; j += 32) sum += pr[j * PS + 1] * expf(pr[j * PS] - m);
        sum = warp_sum(sum);
        const float lse = m + logf(sum);
        const float base = s.cum[row];
        // K best (value desc, index asc) of the n_tiles*KTOP candidates, extracted in order
        float pv = INFINITY;
        int pi = -1;
        const int ncand = n_tiles * KTOP;
        for (int j = 0; j < K; ++j) {
            float bv = -INFINITY;
            int bi = 0x7FFFFFFF;
            for (int c = lane; c < ncand; c += 32) {
                const int tile = c / KTOP, q = c - tile * KTOP;
                const float v = pr[tile * PS + 2 + q];
                const int i = __float_as_int(pr[tile * PS + 2 + KTOP + q]);
                const bool after_prev = (v < pv) || (v == pv && i > pi);
                const bool better = (v > bv) || (v == bv && i < bi);
                if (after_prev && better) bv = v, bi = i;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
            }
            pv = bv, pi = bi;
            if (lane == 0) {
                c_val[r][j] = (bi == 0x7FFFFFFF) ? -INFINITY : base + (bv - lse);
                c_idx[r][j] = bi;
            }
        }
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        // global top-nl over rows x K candidates, ordered by (score desc, flat index asc).  This thread only DECIDES
        // (survivors, their parents, the best completed hypothesis); the token histories are copied by the whole CTA below.
        int taken[MAX_ROWS];  // how many candidates of each row are consumed (rows are sorted already)
        for (int r = 0; r < K; ++r) taken[r] = 0;
        int n_new = 0;
        int best_src = -1;  // slot whose history + <end> became the best completed hypothesis in this step
        float best = s.best_score[img];
        for (int j = 0; j < nl; ++j) {
            float bv = -INFINITY;
            int br = -1;
            for (int r = 0; r < rows_considered; ++r) {
                if (taken[r] >= K) continue;
                const float v = c_val[r][taken[r]];
                if (c_idx[r][taken[r]] == 0x7FFFFFFF) continue;
                if (br < 0 || v > bv) bv = v, br = r;  // ties: lower row first == lower flat index
            }
            if (br < 0) break;
            const int word = c_idx[br][taken[br]];
            taken[br]++;
            if (word == TOK_END) {
                if (bv > best) {  // strict: first maximum wins (BUTD_Model.py:307)
                    best = bv;
                    best_src = br;
                }
            } else {
                s_parent[n_new] = img * K + br;  // absolute parent row
                s_tok[n_new] = word;
                s.cum[img * K + n_new] = bv;
                ++n_new;
            }
        }
        s.best_score[img] = best;
        s.n_live[img] = n_new;
        if (best_src >= 0) {
            s.best_len[img] = t + 1;
            if (s.best_pslot) s.best_pslot[img] = best_src;
        }
        s_n_new = n_new;
        s_best_src = best_src;
        for (int q = n_new; q < K; ++q) {
            s_parent[q] = img * K;
            s_tok[q] = TOK_PAD;
            s.cum[img * K + q] = -INFINITY;
        }
    }
    __syncthreads();
    {  // token histories: survivor q = history of its parent slot + its word; best completed = history + <end> + <pad>...
        const int* sin = s.seqs_in + static_cast<size_t>(img) * K * L;
        int* sout = s.seqs_out + static_cast<size_t>(img) * K * L;
        const int n_new = s_n_new;
        for (int i = threadIdx.x; i < n_new * L; i += blockDim.x) {
            const int q = i / L, p = i - q * L;
            if (p < t) sout[i] = sin[(s_parent[q] - img * K) * L + p];
            else if (p == t) sout[i] = s_tok[q];
        }
        if (s_best_src >= 0) {
            int* bs = s.best_seq + static_cast<size_t>(img) * L;
            for (int p = threadIdx.x; p < L; p += blockDim.x)
                bs[p] = p < t ? sin[s_best_src * L + p] : (p == t ? TOK_END : TOK_PAD);
        }
    }
    if (threadIdx.x < K) {
        s.parent[img * K + threadIdx.x] = s_parent[threadIdx.x];
        s.tok[img * K + threadIdx.x] = s_tok[threadIdx.x];
        if (s.hist_parent) s.hist_parent[static_cast<size_t>(t - 1) * s.B * K + img * K + threadIdx.x] = s_parent[threadIdx.x] - img * K;
    }
    advance_image<KR>(ops, img, K, s_parent, s_tok, threadIdx.x, blockDim.x);
    // Small-batch path: the next step's gate GEMM launch runs CONCURRENTLY with this kernel (it streams its weights meanwhile) and
    // loads the operand rows assembled above only once every image has passed here -- images done, cumulative over the decode.
    if (done_ctr) {
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(done_ctr) : "memory");
    }
}

// Result selection (BUTD_Model.py:306-315): the best COMPLETED hypothesis if any, else live slot 0 (slots stay
// sorted by score).  seqs = the buffer written by the last step.
// alphas_out [B, T, R]: attention map of the returned hypothesis at every step it took (zero rows after its end),
// reconstructed by walking the per-step parent slots back from its last step (the reference reorders a per-beam
// alpha history instead, BUTD_Model.py:280-281,303).  One warp per image.
__global__ void beam_finalize_kernel(BeamState s, const int* __restrict__ seqs, int* __restrict__ tokens,
                                     float* __restrict__ seq_logprob, int* __restrict__ lengths,
                                     const float* __restrict__ alpha_step, float* __restrict__ alphas_out, int R) {
    const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (img >= s.B) return;
    const int L = s.T + 1;
    const bool done = s.best_score[img] > -INFINITY;
    const int* src = done ? s.best_seq + static_cast<size_t>(img) * L : seqs + static_cast<size_t>(img) * s.K * L;
    for (int i = lane; i < L; i += 32) tokens[static_cast<size_t>(img) * L + i] = src[i];
    if (lane == 0) {
        if (seq_logprob) seq_logprob[img] = done ? s.best_score[img] : s.cum[img * s.K];
        if (lengths) lengths[img] = done ? s.best_len[img] : L;
    }
    if (alphas_out) {
        const size_t MK = static_cast<size_t>(s.B) * s.K;
        int t_end = done ? s.best_len[img] - 1 : s.T;        // number of generated words
        int cur = done ? s.best_pslot[img] : s.hist_parent[static_cast<size_t>(s.T - 1) * MK + img * s.K];  // live slot 0
        for (int t = s.T; t > t_end; --t)
            for (int r = lane; r < R; r += 32) alphas_out[(static_cast<size_t>(img) * s.T + t - 1) * R + r] = 0.f;
        for (int t = t_end; t >= 1; --t) {
            const float* a = alpha_step + (static_cast<size_t>(t - 1) * MK + img * s.K + cur) * R;
            for (int r = lane; r < R; r += 32) alphas_out[(static_cast<size_t>(img) * s.T + t - 1) * R + r] = a[r];
            if (t > 1) cur = s.hist_parent[static_cast<size_t>(t - 2) * MK + img * s.K + cur];
        }
    }
}

// ================================================================================================ sampling
struct SampleState {
    int B, n, V, T;
    int* tok;          // [B*n]
    int* unfinished;   // [B*n]
    int* parent;       // [B*n] cell-state row indirection (identity after the first step)
    int* live_count;   // [T+1] rows still unfinished after step t (the reference's batch-wide early break)
    int* tokens;       // [B*n, T] output
    float* logprobs;   // [B*n, T] output or null
    int multinomial;
    int scst;          // 1: n = n_samples + 1 rows per image, the last one a greedy rollout (no <end> handling)
    int* greedy_tokens;  // scst: [B, T] output of the greedy rows; tokens / logprobs then hold [B*(n-1), T] sample rows
};

template <int KR>
__global__ void sample_init_kernel(SampleState s, AdvOps ops, int parent_is_img) {
    const int img = blockIdx.x;
    if (threadIdx.x < s.n) {
        const int row = img * s.n + threadIdx.x;
        s.tok[row] = TOK_STA;
        s.unfinished[row] = 1;
        s.parent[row] = parent_is_img ? img : row;
    }
    if (img == 0)
        for (int i = threadIdx.x; i <= s.T; i += blockDim.x) s.live_count[i] = 0;
    __shared__ int i_prow[MAX_ROWS], i_tok[MAX_ROWS];
    if (threadIdx.x < MAX_ROWS) i_prow[threadIdx.x] = img * s.n + threadIdx.x, i_tok[threadIdx.x] = TOK_STA;
    __syncthreads();
    advance_image<KR>(ops, img, s.n, i_prow, i_tok, threadIdx.x, blockDim.x);
}

// One CTA per image, one warp per row.  sample (BUTD_Model.py:183-188): word = argmax; sample_rl (:221-233):
// word ~ multinomial via Gumbel-max, logprob gathered, <end> and everything after it stored as 0, 0 fed back.
template <int KR>
__global__ void __launch_bounds__(128) sample_step_kernel(const float* __restrict__ part, int n_tiles, SampleState s, int t,
                                                          AdvOps ops) {
    griddep_launch();
    griddep_wait();  // the inputs come from earlier kernels of the stream
    __shared__ int s_tok[MAX_ROWS], s_prow[MAX_ROWS];
    const int img = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PS = SAMPLE_PART_STRIDE;
    if (threadIdx.x < MAX_ROWS) s_prow[threadIdx.x] = img * s.n + threadIdx.x;
    const bool stopped = s.multinomial && t > 0 && s.live_count[t - 1] == 0;  // reference broke out of the loop
    for (int r = warp; r < s.n; r += 4) {
        const int row = img * s.n + r;
        const float* pr = part + static_cast<size_t>(row) * n_tiles * PS;
        float m = -INFINITY;
        for (int j = lane; j < n_tiles; j += 32) m = fmaxf(m, pr[j * PS]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < n_tiles; j += 32) sum += pr[j * PS + 1] * expf(pr[j * PS] - m);
        sum = warp_sum(sum);
        const float lse = m + logf(sum);
        float bv = -INFINITY, braw = 0.f;
        int bi = 0x7FFFFFFF;
        for (int j = lane; j < n_tiles; j += 32) {
            const float v = pr[j * PS + 2];
            const int i = __float_as_int(pr[j * PS + 3]);
            if (v > bv || (v == bv && i < bi)) bv = v, bi = i, braw = pr[j * PS + 4];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const float orw = __shfl_xor_sync(0xffffffffu, braw, o);
            if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi, braw = orw;
        }
        if (lane == 0) {
            int word = bi;
            const bool greedy_row = s.scst && r == s.n - 1;
            const int out_row = s.scst ? img * (s.n - 1) + r : row;  // sample rows are stored densely without the greedy ones
            if (s.multinomial && !greedy_row) {
                int unf = s.unfinished[row];
                unf = unf && (word != TOK_END);
                word = unf ? word : 0;
                s.unfinished[row] = unf;
                if (!stopped) {
                    s.tokens[static_cast<size_t>(out_row) * s.T + t] = word;
                    if (s.logprobs) s.logprobs[static_cast<size_t>(out_row) * s.T + t] = braw - lse;
                    if (unf) atomicAdd(s.live_count + t, 1);
                }
            } else if (greedy_row) {
                s.greedy_tokens[static_cast<size_t>(img) * s.T + t] = word;
            } else {
                if (s.tokens) s.tokens[static_cast<size_t>(row) * s.T + t] = word;
                if (s.logprobs) s.logprobs[static_cast<size_t>(row) * s.T + t] = braw - lse;
            }
            s.tok[row] = word;
            s.parent[row] = row;
            s_tok[r] = word;
        }
    }
    __syncthreads();
    advance_image<KR>(ops, img, s.n, s_prow, s_tok, threadIdx.x, blockDim.x);
}

}  // namespace capdec
