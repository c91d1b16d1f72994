// Small-batch decode GEMMs for sm_100a: "swap-AB" tcgen05 tiles + split-K, one persistent launch for up to three dependent
// GEMM phases of a decode step.
//
// At the reference's own evaluation shapes (BASELINE configs[0]: 16 images x beam 3 = 48 rows; Utils.py:72-73: ONE image per
// beam-search call) the 256 x 256 pair tiles of gemm.cuh leave > 80 % of every MMA's rows empty and only a handful of SM
// pairs busy.  A decode step there is bound by how fast the ~72 MB of fp16 weights stream from L2 through the SMs, so this
// kernel turns the product around:
//
//   D^T[n, m] = W[n, :] . X[m, :]      weights are the 128-row UMMA "A" operand (M = 128), the <= 128 activation rows are
//                                      the "N" operand (N = N_ACT = 16 / 64 / 128): no MMA row is wasted on padding
//
// and cuts the K loop of every 128-row weight tile into `ksplit` pieces so that all 148 SMs pull weights at once (one
// (tile, split) item per CTA).  Partial tiles go to an L2-resident scratch slab, transposed to [activation row][weight
// row] so that both the partial store and the reduction are coalesced.  The `ksplit` CTAs of a tile then wait for each
// other at the tile's counter and SHARE the epilogue: each takes a contiguous share of the activation rows, sums the slabs
// in split order (deterministic) into shared memory and runs the fused epilogue on the sums -- the same epilogues as
// gemm.cuh: bias / store, LSTMCell pointwise, GLU, log-softmax partials + top-k, Gumbel-max draw.  A phase whose tiles
// already fill the SMs (the vocabulary GEMM: 75 tiles) is not split: its accumulators go from TMEM to shared memory directly.
//
// Phases: a launch carries up to SM_MAX_PHASES GEMMs that depend on each other (top-down LSTM gates -> dec_att;
// language LSTM gates -> vocabulary logits); between two phases all CTAs meet at a grid barrier (the grid is one CTA per
// SM, all co-resident) instead of paying a kernel boundary.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue (thread == TMEM
// lane == weight row of the tile).
#pragma once
#include "gemm.cuh"

namespace capdec {

constexpr int SM_THREADS = 192;
constexpr int SM_EPI_THREADS = 128;
constexpr int SM_MAX_PHASES = 3;
constexpr int SM_TILE_N = 128;  // weight rows per tile (UMMA M)
constexpr int SM_TILE_LD = 132;  // row pitch (floats) of the summed tile in shared memory: 16-byte aligned, float4 accesses conflict-free

struct SmallPhase {
    CUtensorMap map_w;  // weights [N_w, K] fp16, box 64 x 128 rows, 128B swizzle
    CUtensorMap map_x;  // activations [rows, K] fp16, box 64 x N_ACT rows
    int N_w, M;         // output features (weight rows), activation rows (M <= N_ACT)
    int k_blocks, passes, w_lo_off, x_lo_off;
    int tiles, ksplit;
    int epi, ktop;
    int w_hint;         // L2 priority of the weight lines: 0 = default, 1 = evict-first, 2 = evict-last
    const int* wait_ctr;  // non-null: activation columns below wait_kb k-blocks are written by a kernel running CONCURRENTLY (the
    int wait_target;      //   attention); their loads wait until *wait_ctr >= wait_target, the weight tiles are requested before
    int wait_kb;
    EpiParams e;
};

struct SmallParams {
    int n_phases;
    float* slabs;        // [item][N_ACT][128] fp32 partial tiles (L2-resident scratch)
    int* counters;       // [tile][2] arrivals at / departures from the tile's reduction (zero between uses)
    unsigned* bar;       // grid-barrier counters: [SM_MAX_PHASES][2] arrivals / departures, zero between uses
    unsigned long long* trace;  // CAPDEC_TRACE=1: [0] = launch counter, then 16 %globaltimer stamps of CTA 0 per launch; else null
    SmallPhase ph[SM_MAX_PHASES];
};

template <int N_ACT>
struct SmallCfg {
    static constexpr int A_BYTES = SM_TILE_N * BLOCK_K * 2;
    static constexpr int B_BYTES = N_ACT * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = N_ACT <= 16 ? 8 : (N_ACT <= 64 ? 7 : 4);
    static constexpr int ACC_COLS = N_ACT < 32 ? 32 : N_ACT;  // TMEM columns per accumulator buffer (tcgen05.ld reads 32 at a time)
    static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
    static constexpr int TILE_BYTES = N_ACT * SM_TILE_LD * 4;  // the summed tile [activation row][weight row] the epilogues read
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + TILE_BYTES + 256 + 1024;
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- epilogues.  They read the SUMMED tile from shared memory: `tile[lr][n]`, lr = local row of this CTA's share of the
// activation rows (global row m = row0 + lr, nrows of them), n = weight row of the 128-row tile.  With ksplit == 1 the share
// is every row and the tile comes straight from TMEM; with ksplit > 1 the `ksplit` CTAs of the tile meet at its counter and
// each sums its contiguous share of the rows from the L2-resident partial slabs (small_sum_share).
constexpr int SM_U = 4;

// tile[lr][4j..4j+3] = sum over the splits (in split order: deterministic) for this CTA's rows; thread (w, j) of the 128
// epilogue threads takes rows w, w+4, ...; the loads of 4 rows x 4 splits are in flight together (one L2 round trip).
// `colbias` (logit phases): per-column additive term folded into the tile, -inf for the padded vocabulary tail.
template <int N_ACT>
__device__ __forceinline__ void small_sum_share(const float* slab0, int ksplit, int row0, int nrows, float* tile, int t, bool logits,
                                                const float* bias, int n_base, int N_w) {
    const int w = t >> 5, j = t & 31;
    float cb[4] = {0.f, 0.f, 0.f, 0.f};
    if (logits) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cb[i] = (n_base + 4 * j + i < N_w) ? __ldg(bias + n_base + 4 * j + i) : -INFINITY;
    }
    constexpr size_t SS = static_cast<size_t>(N_ACT) * SM_TILE_N;
    for (int lb = w; lb < nrows; lb += 4 * SM_U) {
        float4 a[SM_U];
#pragma unroll
        for (int u = 0; u < SM_U; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < ksplit; s0 += 4) {
            float4 v[SM_U][4];
#pragma unroll
            for (int u = 0; u < SM_U; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (lb + 4 * u < nrows && s0 + i < ksplit)
                        v[u][i] = ldcg4(slab0 + (s0 + i) * SS + static_cast<size_t>(row0 + lb + 4 * u) * SM_TILE_N + 4 * j);
#pragma unroll
            for (int u = 0; u < SM_U; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (lb + 4 * u < nrows && s0 + i < ksplit)
                        a[u].x += v[u][i].x, a[u].y += v[u][i].y, a[u].z += v[u][i].z, a[u].w += v[u][i].w;
        }
#pragma unroll
        for (int u = 0; u < SM_U; ++u)
            if (lb + 4 * u < nrows)
                *reinterpret_cast<float4*>(tile + (lb + 4 * u) * SM_TILE_LD + 4 * j) =
                    make_float4(a[u].x + cb[0], a[u].y + cb[1], a[u].z + cb[2], a[u].w + cb[3]);
    }
}

__device__ __forceinline__ float4 tile4(const float* tile, int lr, int j) {
    return *reinterpret_cast<const float4*>(tile + lr * SM_TILE_LD + 4 * j);
}

// thread (w, j): rows lr = w, w+4, ... ; weight rows n_base + 4j .. 4j+3
__device__ __forceinline__ void small_epi_store(const SmallPhase& P, const float* tile, int tile_idx, int row0, int nrows, int t) {
    const EpiParams& e = P.e;
    const int w = t >> 5, j = t & 31;
    const int n0 = tile_idx * SM_TILE_N + 4 * j;
    if (n0 >= P.N_w) return;
    const bool full = n0 + 4 <= P.N_w;
    float b[4] = {0.f, 0.f, 0.f, 0.f};
    if (e.bias) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (n0 + i < P.N_w) b[i] = __ldg(e.bias + n0 + i);
    }
    for (int lr = w; lr < nrows; lr += 4) {
        const int m = row0 + lr;
        const float4 a = tile4(tile, lr, j);
        float v[4] = {a.x + b[0], a.y + b[1], a.z + b[2], a.w + b[3]};
        if (e.relu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (e.row_keep && __ldg(e.row_keep + m) == 0.f) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = 0.f;
        }
        if (e.out32) {
            float* o = e.out32 + static_cast<size_t>(m) * e.ld32 + n0;
            if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (n0 + i < P.N_w) o[i] = v[i];
            }
        }
        if (e.out16) {
            __half* o = e.out16 + static_cast<size_t>(m) * e.ld16 + n0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (n0 + i < P.N_w) {
                    __half hi, lo;
                    split_f16(v[i], hi, lo);
                    o[i] = hi;
                    if (e.lo16 > 0) o[e.lo16 + i] = lo;
                }
            }
        }
    }
}

// gate-interleaved weight rows n = 4*unit + gate (i, f, g, o): the four sums of a float4 are one hidden unit's gates.
// The additive terms of SM_U rows are requested together (they were prefetched into L1 while the weights streamed).
__device__ __forceinline__ void small_epi_lstm(const SmallPhase& P, const float* tile, int tile_idx, int row0, int nrows, int t) {
    const EpiParams& e = P.e;
    const int w = t >> 5, j = t & 31;
    const int n0 = tile_idx * SM_TILE_N + 4 * j;
    if (n0 >= P.N_w) return;
    const int unit = n0 >> 2;
    for (int lb = w; lb < nrows; lb += 4 * SM_U) {
        int gidx[SM_U], prow[SM_U];
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
            const int m = row0 + lb + 4 * u;
            const bool ok = lb + 4 * u < nrows;
            gidx[u] = (e.gather && ok) ? __ldcg(e.gather_idx + m) : 0;  // (L2 loads: a phase that waited for the bookkeeping kernel
            prow[u] = (e.parent && ok) ? __ldcg(e.parent + m) : m;      //  running beside it reads what that kernel just wrote)
        }
        float4 ad[SM_U], g[SM_U];
        float cp[SM_U];
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
            const int m = row0 + lb + 4 * u;
            if (lb + 4 * u < nrows) {
                const float* add = e.rowadd ? e.rowadd + static_cast<size_t>(m / e.rows_per_group) * e.rowadd_ld : e.bias;
                ad[u] = __ldg(reinterpret_cast<const float4*>(add + n0));
                g[u] = e.gather ? __ldg(reinterpret_cast<const float4*>(e.gather + static_cast<size_t>(gidx[u]) * e.gather_ld + n0))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                cp[u] = e.c_in ? e.c_in[static_cast<size_t>(prow[u]) * e.ldc + unit] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
            const int lr = lb + 4 * u, m = row0 + lr;
            if (lr >= nrows) break;
            const float4 a = tile4(tile, lr, j);
            const float gi = a.x + (ad[u].x + g[u].x), gf = a.y + (ad[u].y + g[u].y), gg = a.z + (ad[u].z + g[u].z),
                        go = a.w + (ad[u].w + g[u].w);
            const float cn = sigmoidf_acc(gf) * cp[u] + sigmoidf_acc(gi) * tanhf_acc(gg);
            const float hn = sigmoidf_acc(go) * tanhf_acc(cn);
            e.c_out[static_cast<size_t>(m) * e.ldc + unit] = cn;
            __half hi, lo;
            split_f16(hn, hi, lo);
            __half* o = e.out16 + static_cast<size_t>(m) * e.ld16 + unit;
            o[0] = hi;
            if (e.lo16 > 0) o[e.lo16] = lo;
            if (e.h32) e.h32[static_cast<size_t>(m) * e.ldh32 + unit] = hn;
        }
    }
}

// (a, gate)-interleaved weight rows n = 2*unit + s: a float4 holds two units
__device__ __forceinline__ void small_epi_glu(const SmallPhase& P, const float* tile, int tile_idx, int row0, int nrows, int t) {
    const EpiParams& e = P.e;
    const int w = t >> 5, j = t & 31;
    const int n0 = tile_idx * SM_TILE_N + 4 * j;
    if (n0 >= P.N_w) return;
    const int unit = n0 >> 1;
    const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0));
    for (int lr = w; lr < nrows; lr += 4) {
        const int m = row0 + lr;
        const float4 a = tile4(tile, lr, j);
        float y0 = (a.x + b.x) * sigmoidf_acc(a.y + b.y);
        float y1 = (a.z + b.z) * sigmoidf_acc(a.w + b.w);
        if (e.resid) {
            const float2 r = *reinterpret_cast<const float2*>(e.resid + static_cast<size_t>(m) * e.ld_resid + unit);
            y0 += r.x, y1 += r.y;
        }
        if (e.out32) {
            float* o = e.out32 + static_cast<size_t>(m) * e.ld32 + unit;
            o[0] = y0, o[1] = y1;
        }
        if (e.out16) {
            __half* o = e.out16 + static_cast<size_t>(m) * e.ld16 + unit;
            __half h0, l0, h1, l1;
            split_f16(y0, h0, l0);
            split_f16(y1, h1, l1);
            o[0] = h0, o[1] = h1;
            if (e.lo16 > 0) o[e.lo16] = l0, o[e.lo16 + 1] = l1;
        }
    }
}

__device__ __forceinline__ bool cand_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// Vocabulary tile of the logit GEMM.  TPR = 1 / 2 / 4 / 8 adjacent threads share an activation row (as many as the 128
// epilogue threads allow), each walks its 128 / TPR logits in shared memory -- no cross-lane traffic until the short merge
// at the end.  Writes the (row, tile) partial record gemm.cuh's epilogues write per (row, run, column share): max, sum of
// exp(x - max), then the KTOP largest logits + indices (top-k) or the arg-max of logit (+ Gumbel noise) with its raw logit.
__device__ __forceinline__ int small_tpr(int nrows) {
    int tpr = 1;
    while (tpr < 8 && 2 * tpr * nrows <= SM_EPI_THREADS) tpr *= 2;
    return tpr;
}

template <int KTOP>
__device__ __forceinline__ void small_epi_topk(const SmallPhase& P, const float* tile, int tile_idx, int row0, int nrows, int t) {
    const EpiParams& e = P.e;
    const int tpr = small_tpr(nrows);
    const int lr = t / tpr, part = t - lr * tpr;
    const bool active = lr < nrows;
    const int seg = SM_TILE_N / tpr;                  // logits per thread
    const int c0 = part * seg;                        // first column of this thread's segment
    const int n_base = tile_idx * SM_TILE_N;
    const float* row = tile + (active ? lr : 0) * SM_TILE_LD;  // logits + bias, padded vocabulary tail = -inf
    // pass 1: maximum of the segment
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int c = c0; c < c0 + seg; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(row + c);
        m4[0] = fmaxf(m4[0], a.x), m4[1] = fmaxf(m4[1], a.y), m4[2] = fmaxf(m4[2], a.z), m4[3] = fmaxf(m4[3], a.w);
    }
    float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    for (int o = 1; o < tpr; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float mx2 = mx * LOG2E;
    // pass 2: sum of exp(x - max) and the segment's KTOP largest (value desc, index asc)
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    float tv[KTOP];
    int ti[KTOP];
#pragma unroll
    for (int q = 0; q < KTOP; ++q) tv[q] = -INFINITY, ti[q] = 0x7FFFFFFF;
    for (int c = c0; c < c0 + seg; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(row + c);
        const float x[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] += ex2_ftz(fmaf(x[i], LOG2E, -mx2));  // exp2(-inf) = 0 for the padded columns
        if (fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])) > tv[KTOP - 1]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (x[i] > tv[KTOP - 1]) {
                    tv[KTOP - 1] = x[i], ti[KTOP - 1] = n_base + c + i;
#pragma unroll
                    for (int q = KTOP - 1; q > 0; --q) {
                        if (tv[q] > tv[q - 1]) {
                            const float fv = tv[q]; tv[q] = tv[q - 1]; tv[q - 1] = fv;
                            const int iv = ti[q]; ti[q] = ti[q - 1]; ti[q - 1] = iv;
                        }
                    }
                }
            }
        }
    }
    float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    for (int o = 1; o < tpr; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    constexpr int PS = topk_part_stride(KTOP);
    float* rec = e.part + (static_cast<size_t>(row0 + (active ? lr : 0)) * e.n_tiles + tile_idx) * PS;
    if (active && part == 0) rec[0] = mx, rec[1] = sum;
    // merge the tpr sorted lists: KTOP rounds of "best head wins"; the winner pops its head
#pragma unroll
    for (int q = 0; q < KTOP; ++q) {
        float wv = tv[0];
        int wi = ti[0];
        for (int o = 1; o < tpr; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (cand_better(ov, oi, wv, wi)) wv = ov, wi = oi;
        }
        if (wi == ti[0] && wi != 0x7FFFFFFF) {
#pragma unroll
            for (int r = 0; r + 1 < KTOP; ++r) tv[r] = tv[r + 1], ti[r] = ti[r + 1];
            tv[KTOP - 1] = -INFINITY, ti[KTOP - 1] = 0x7FFFFFFF;
        }
        if (active && part == 0) rec[2 + q] = wv, rec[2 + KTOP + q] = __int_as_float(wi);
    }
}

__device__ __forceinline__ void small_epi_sample(const SmallPhase& P, const float* tile, int tile_idx, int row0, int nrows, int t) {
    const EpiParams& e = P.e;
    const int tpr = small_tpr(nrows);
    const int lr = t / tpr, part = t - lr * tpr;
    const bool active = lr < nrows;
    const int seg = SM_TILE_N / tpr, c0 = part * seg, n_base = tile_idx * SM_TILE_N;
    const float* row = tile + (active ? lr : 0) * SM_TILE_LD;  // logits + bias, padded vocabulary tail = -inf
    const int m = row0 + (active ? lr : 0);
    bool noisy = e.use_noise != 0;
    int noise_row = m;
    if (e.scst_n > 0) {  // groups of scst_n sampled rows + one greedy row per image (DrawState::init)
        const int img = m / (e.scst_n + 1), r = m - img * (e.scst_n + 1);
        noisy = noisy && r < e.scst_n;
        noise_row = img * e.scst_n + r;
    }
    const uint32_t seed = e.seed_ptr ? __ldg(e.seed_ptr) : e.seed;
    const uint32_t rs = gumbel_row_step_hash(seed, static_cast<uint32_t>(noise_row), static_cast<uint32_t>(e.step));
    const int forced = (e.forced && active) ? __ldg(e.forced + static_cast<size_t>(m) * e.forced_ld + e.step) : -1;
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int c = c0; c < c0 + seg; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(row + c);
        m4[0] = fmaxf(m4[0], a.x), m4[1] = fmaxf(m4[1], a.y), m4[2] = fmaxf(m4[2], a.z), m4[3] = fmaxf(m4[3], a.w);
    }
    const float smax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));  // largest logit of this thread's segment
    float mx = smax;
    for (int o = 1; o < tpr; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float mx2 = mx * LOG2E;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    float bv = -INFINITY, braw = 0.f;
    int bi = 0x7FFFFFFF;
    for (int c = c0; c < c0 + seg; c += 4) {
        const float4 a = *reinterpret_cast<const float4*>(row + c);
        const float x[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] += ex2_ftz(fmaf(x[i], LOG2E, -mx2));
        if (e.forced) {
            const int f = forced - (n_base + c);
            if (f >= 0 && f < 4) bv = 3.0e38f, bi = forced, braw = f == 0 ? x[0] : (f == 1 ? x[1] : (f == 2 ? x[2] : x[3]));
        } else if (noisy) {
            // a column can win only with noise above best - (largest logit of the segment): decided on the hash bits alone,
            // the two logarithms run for the few columns that pass (DrawState::tile, gemm.cuh)
            const uint32_t thr = gumbel_pass_threshold(bv - smax - 1e-3f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t hb = gumbel_bits(rs, static_cast<uint32_t>(n_base + c + i));
                if ((hb >> 8) >= thr && x[i] > -INFINITY) {
                    const float pv = x[i] + gumbel_from_bits(hb);
                    if (pv > bv) bv = pv, bi = n_base + c + i, braw = x[i];  // ascending index: the first maximum wins
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (x[i] > bv) bv = x[i], bi = n_base + c + i, braw = x[i];
        }
    }
    float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    for (int o = 1; o < tpr; o <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const float orw = __shfl_xor_sync(0xffffffffu, braw, o);
        if (cand_better(ov, oi, bv, bi)) bv = ov, bi = oi, braw = orw;
    }
    if (active && part == 0) {
        float* rec = e.part + (static_cast<size_t>(m) * e.n_tiles + tile_idx) * SAMPLE_PART_STRIDE;
        rec[0] = mx, rec[1] = sum, rec[2] = bv, rec[3] = __int_as_float(bi), rec[4] = braw;
    }
}

// All CTAs of the (co-resident, one per SM) grid meet; the phase's global writes -- made with generic stores -- are
// ordered before the next phase's TMA (async proxy) reads of them.  ctr[0] counts arrivals, ctr[1] departures: the last
// CTA to leave zeroes both, so a barrier's counters are always zero between two uses whatever launches (captured graphs,
// eager launches, other phase counts) follow each other.
__device__ __forceinline__ void small_grid_barrier(unsigned* ctr) {
    fence_proxy_async_all();
    __syncthreads();  // every thread's writes of the phase happen-before thread 0's release below
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        const long long t0 = clock64();
        while (true) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= gridDim.x) break;
            if (clock64() - t0 > 4000000000LL) {
                printf("capdec: grid barrier timed out (block %d, %u of %u arrived)\n", blockIdx.x, v, gridDim.x);
                __trap();
            }
        }
    }
    __syncthreads();
    fence_proxy_async_all();
    if (threadIdx.x == 33) {  // an idle lane of the MMA warp: the producer thread is not delayed by the round trip
        if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
            ctr[0] = 0;
            ctr[1] = 0;
        }
    }
}

__device__ __forceinline__ void small_stamp(unsigned long long* trace, int slot, int k) {
    if (trace && slot >= 0 && slot < 4000) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        trace[1 + 16 * slot + k] = t;
    }
}


// Issued by an epilogue thread BEFORE it waits for its tile's accumulator: pulls what its first batch of rows will read
// besides the partial sums -- bias / hoisted per-image row, the gathered embedding-gate row (a random row of a 155 MB
// table: an HBM access), the parent's cell state -- into L1 while the weights are still streaming.
__device__ __forceinline__ void small_epi_prefetch(const SmallPhase& P, int tile_idx, int row0, int nrows, int t) {
    const EpiParams& e = P.e;
    const int w = t >> 5, j = t & 31;
    const int n0 = tile_idx * SM_TILE_N + 4 * j;
    if (n0 >= P.N_w) return;
    if (P.epi != EPI_LSTM) {
        if (e.bias && (j & 7) == 0) prefetch_l1(e.bias + n0);  // 8 threads share a 128-byte line
        return;
    }
#pragma unroll
    for (int u = 0; u < SM_U; ++u) {
        const int lr = w + 4 * u, m = row0 + lr;
        if (lr >= nrows) break;
        if ((j & 7) == 0) {
            const float* add = e.rowadd ? e.rowadd + static_cast<size_t>(m / e.rows_per_group) * e.rowadd_ld : e.bias;
            prefetch_l1(add + n0);
            if (e.gather) prefetch_l1(e.gather + static_cast<size_t>(__ldg(e.gather_idx + m)) * e.gather_ld + n0);
        }
        if (e.c_in && j == 0) prefetch_l1(e.c_in + static_cast<size_t>(e.parent ? __ldg(e.parent + m) : m) * e.ldc + (n0 >> 2));
    }
}

template <int N_ACT>
__global__ void __launch_bounds__(SM_THREADS, 1) smallm_kernel(const __grid_constant__ SmallParams p) {
    using Cfg = SmallCfg<N_ACT>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);  // 1024-byte aligned, still a shared-space pointer (LDS / STS)
    float* s_tile = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::TILE_BYTES);
    const uint32_t full_bar = smem_u32(bars);
    const uint32_t empty_bar = smem_u32(bars + STAGES);
    const uint32_t tfull_bar = smem_u32(bars + 2 * STAGES);
    const uint32_t tempty_bar = smem_u32(bars + 2 * STAGES + 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    volatile int* s_slot = reinterpret_cast<volatile int*>(tmem_slot + 1);
    if (threadIdx.x == 0) {
        *s_slot = (p.trace && blockIdx.x == 0) ? static_cast<int>(atomicAdd(p.trace, 1ull)) : -1;
        small_stamp(p.trace, *s_slot, 0);
    }

    if (warp == 0 && lane == 0) {
        for (int q = 0; q < p.n_phases; ++q) {
            tma_prefetch_desc(&p.ph[q].map_w);
            tma_prefetch_desc(&p.ph[q].map_x);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar + 8 * s, 1);
            mbar_init(tempty_bar + 8 * s, SM_EPI_THREADS);
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
    const int slot = *s_slot;
    griddep_launch();
    griddep_wait();
    if (threadIdx.x == 0) small_stamp(p.trace, slot, 1);

    // pipeline positions persist across the phases
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int pre = 0, pre_stage = 0;  // producer: k-blocks of the next phase whose weight tiles are already requested

    for (int q = 0; q < p.n_phases; ++q) {
        const SmallPhase& P = p.ph[q];
        const int items = P.tiles * P.ksplit;
        const int total_kb = P.k_blocks * P.passes;
        if (warp == 0) {
            // ===================== TMA producer =====================
            // The first `pre` k-blocks of this CTA's first item had their WEIGHT tiles requested before the grid barrier
            // that precedes the phase (weights do not depend on the previous phase); only the activation tiles follow here.
            if (lane == 0) {
                const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();
                // K order of an item.  Normally the contiguous k-blocks [kk0, kk1) of its split.  When the activation columns below
                // wait_kb k-blocks come from a kernel running concurrently (wait_ctr), every split takes its share of the
                // INDEPENDENT k-blocks first and its share of the dependent ones last: all CTAs stream weights and multiply while the
                // other kernel runs, request the dependent weight tiles ahead, and after the counter passes only those k-blocks'
                // activation tiles are missing.  (The MMA warp just consumes n_kb stages per item: the order is the producer's.)
                const bool mixed = P.wait_ctr != nullptr && P.passes == 1 && P.wait_kb % P.ksplit == 0 && (total_kb - P.wait_kb) % P.ksplit == 0;
                for (int item = blockIdx.x; item < items; item += gridDim.x) {
                    const int tile = item / P.ksplit, split = item - tile * P.ksplit;
                    const int kk0 = (split * total_kb) / P.ksplit, kk1 = ((split + 1) * total_kb) / P.ksplit;
                    int n_ind = kk1 - kk0, ind0 = kk0, n_dep = 0, dep0 = 0;
                    if (mixed) {
                        n_dep = P.wait_kb / P.ksplit, dep0 = split * n_dep;
                        n_ind = (total_kb - P.wait_kb) / P.ksplit, ind0 = P.wait_kb + split * n_ind;
                    } else if (P.wait_ctr != nullptr && kk0 < P.wait_kb) {
                        n_dep = n_ind, dep0 = ind0, n_ind = 0;  // contiguous split that touches the dependent columns: all of it waits
                    }
                    const int n_kb = n_ind + n_dep;
                    for (int i = 0; i < n_kb; ++i) {
                        const int kk = i < n_ind ? ind0 + i : dep0 + (i - n_ind);
                        if (i == n_ind && n_dep > 0) {
                            // entering the dependent k-blocks: request their weight tiles (as many as the ring holds), THEN wait
                            // for the counter; the loop body below adds the activation tiles
                            if (pre == 0) {
                                pre_stage = stage;
                                for (int u = i; u < n_kb && pre < STAGES; ++u, ++pre) {
                                    const int ku = dep0 + (u - n_ind);
                                    const int pass_u = ku / P.k_blocks;
                                    const int kw = (ku - pass_u * P.k_blocks) * BLOCK_K + (pass_u == 1 ? P.w_lo_off : 0);
                                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                                    mbar_arrive_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
                                    if (P.w_hint == 0) tma_load_2d(smem_base + stage * Cfg::STAGE_BYTES, &P.map_w, full_bar + 8 * stage, kw, tile * SM_TILE_N);
                                    else tma_load_2d_hint(smem_base + stage * Cfg::STAGE_BYTES, &P.map_w, full_bar + 8 * stage, kw, tile * SM_TILE_N,
                                                          P.w_hint == 1 ? pol_first : pol_last);
                                    if (++stage == STAGES) stage = 0, phase ^= 1;
                                }
                            }
                            const long long t0 = clock64();
                            while (true) {
                                int v;
                                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(P.wait_ctr) : "memory");
                                if (v >= P.wait_target) break;
                                if (clock64() - t0 > 4000000000LL) {
                                    printf("capdec: wait for the concurrent attention kernel timed out (block %d: %d of %d)\n", blockIdx.x, v, P.wait_target);
                                    __trap();
                                }
                            }
                            fence_proxy_async_all();  // its generic-proxy stores are ordered before this thread's TMA reads
                        }
                        const int pass = kk / P.k_blocks;
                        const int kb = kk - pass * P.k_blocks;
                        const int kx = kb * BLOCK_K + (pass == 2 ? P.x_lo_off : 0);  // hi*hi, hi(x)*lo(w), lo(x)*hi(w)
                        const int kw = kb * BLOCK_K + (pass == 1 ? P.w_lo_off : 0);
                        if (pre > 0) {  // weight tile already in flight into pre_stage
                            tma_load_2d(smem_base + pre_stage * Cfg::STAGE_BYTES + Cfg::A_BYTES, &P.map_x, full_bar + 8 * pre_stage, kx, 0);
                            if (++pre_stage == STAGES) pre_stage = 0;
                            --pre;
                            continue;
                        }
                        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        mbar_arrive_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
                        if (P.w_hint == 0) tma_load_2d(sa, &P.map_w, full_bar + 8 * stage, kw, tile * SM_TILE_N);
                        else tma_load_2d_hint(sa, &P.map_w, full_bar + 8 * stage, kw, tile * SM_TILE_N, P.w_hint == 1 ? pol_first : pol_last);
                        tma_load_2d(sa + Cfg::A_BYTES, &P.map_x, full_bar + 8 * stage, kx, 0);
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                }
                if (q + 1 < p.n_phases) {  // request the next phase's first weight tiles now: they land during the grid barrier
                    const SmallPhase& Pn = p.ph[q + 1];
                    const int item = blockIdx.x;
                    if (item < Pn.tiles * Pn.ksplit && Pn.wait_ctr == nullptr) {  // (a waiting phase orders its k-blocks itself)
                        const int n_kb = Pn.k_blocks * Pn.passes;
                        const int tile = item / Pn.ksplit, split = item - tile * Pn.ksplit;
                        const int kk0 = (split * n_kb) / Pn.ksplit, kk1 = ((split + 1) * n_kb) / Pn.ksplit;
                        pre_stage = stage;
                        for (int kk = kk0; kk < kk1 && pre < STAGES; ++kk, ++pre) {
                            const int pass = kk / Pn.k_blocks;
                            const int kw = (kk - pass * Pn.k_blocks) * BLOCK_K + (pass == 1 ? Pn.w_lo_off : 0);
                            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                            mbar_arrive_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
                            if (Pn.w_hint == 0) tma_load_2d(smem_base + stage * Cfg::STAGE_BYTES, &Pn.map_w, full_bar + 8 * stage, kw, tile * SM_TILE_N);
                            else tma_load_2d_hint(smem_base + stage * Cfg::STAGE_BYTES, &Pn.map_w, full_bar + 8 * stage, kw, tile * SM_TILE_N,
                                                  Pn.w_hint == 1 ? pol_first : pol_last);
                            if (++stage == STAGES) stage = 0, phase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                constexpr uint32_t idesc = make_idesc_f16(SM_TILE_N, N_ACT);
                for (int item = blockIdx.x; item < items; item += gridDim.x) {
                    const int tile = item / P.ksplit, split = item - tile * P.ksplit;
                    const int kk0 = (split * total_kb) / P.ksplit, kk1 = ((split + 1) * total_kb) / P.ksplit;
                    mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
                    for (int kk = kk0; kk < kk1; ++kk) {
                        mbar_wait(full_bar + 8 * stage, phase);
                        if (kk == kk0 && item == blockIdx.x && q < 2) small_stamp(p.trace, slot, 2 + 6 * q);
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint64_t da = make_smem_desc_sw128(sa);
                        const uint64_t db = make_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kk > kk0 || k > 0) ? 1u : 0u);
                        umma_commit(empty_bar + 8 * stage);
                        if (++stage == STAGES) stage = 0, phase ^= 1;
                    }
                    umma_commit(tfull_bar + 8 * acc);
                    if (item == blockIdx.x && q < 2) small_stamp(p.trace, slot, 3 + 6 * q);
                    if (++acc == 2) acc = 0, acc_phase ^= 1;
                }
            }
        } else {
            // ===================== epilogue warps =====================
            const int quarter = warp & 3;
            const int n_local = quarter * 32 + lane;
            const int t = threadIdx.x - 64;
            const int rpc = (P.M + P.ksplit - 1) / P.ksplit;  // activation rows per CTA of a tile
            const bool is_logits = P.epi == EPI_TOPK || P.epi == EPI_SAMPLE;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int tile = item / P.ksplit;
                const int split = item - tile * P.ksplit;
                const int row0 = split * rpc;
                const int nrows = P.M - row0 < rpc ? (P.M - row0 > 0 ? P.M - row0 : 0) : rpc;
                if (P.wait_ctr == nullptr) small_epi_prefetch(P, tile, row0, nrows, t);  // (a waiting phase's indices are not final yet)
                mbar_wait(tfull_bar + 8 * acc, acc_phase);
                if (t == 0 && item == blockIdx.x && q < 2) small_stamp(p.trace, slot, 4 + 6 * q);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(quarter * 32) << 16);
                // accumulator -> [activation row][weight row]: straight into the shared tile when this CTA holds the whole K
                // range, else into this item's L2-resident partial slab
                float* dst = P.ksplit == 1 ? s_tile : p.slabs + static_cast<size_t>(item) * N_ACT * SM_TILE_N;
                const int dst_ld = P.ksplit == 1 ? SM_TILE_LD : SM_TILE_N;
                // logit phases: bias added (and the padded vocabulary tail set to -inf) as the tile is written
                const int n_col = tile * SM_TILE_N + n_local;
                const float colb = (is_logits && P.ksplit == 1) ? (n_col < P.N_w ? __ldg(P.e.bias + n_col) : -INFINITY) : 0.f;
#pragma unroll 1
                for (int c = 0; c < Cfg::ACC_COLS / 32; ++c) {
                    if (c * 32 >= P.M) break;  // warp-uniform
                    float v[32];
                    tmem_ld_32x32(taddr + c * 32, v);
                    if (P.ksplit == 1) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < P.M) dst[(c * 32 + i) * dst_ld + n_local] = v[i] + colb;
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < P.M) __stcg(dst + (c * 32 + i) * dst_ld + n_local, v[i]);
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar + 8 * acc);
                if (++acc == 2) acc = 0, acc_phase ^= 1;
                int* ctr = p.counters + 2 * tile;
                if (P.ksplit > 1) {
                    // every CTA of the tile waits until all ksplit partials are in L2, then sums its share of the rows
                    named_bar_sync(1, SM_EPI_THREADS);  // the 128 threads' slab stores happen-before thread 0's release
                    if (t == 0) {
                        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
                        const long long t0 = clock64();
                        while (true) {
                            int v;
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
                            if (v >= P.ksplit) break;
                            if (clock64() - t0 > 4000000000LL) {
                                printf("capdec: split-K wait timed out (block %d tile %d: %d of %d)\n", blockIdx.x, tile, v, P.ksplit);
                                __trap();
                            }
                        }
                    }
                    named_bar_sync(1, SM_EPI_THREADS);
                    if (t == 0 && item == blockIdx.x && q < 2) small_stamp(p.trace, slot, 5 + 6 * q);
                    small_sum_share<N_ACT>(p.slabs + static_cast<size_t>(tile) * P.ksplit * N_ACT * SM_TILE_N, P.ksplit, row0, nrows, s_tile, t,
                                           is_logits, P.e.bias, tile * SM_TILE_N, P.N_w);
                }
                named_bar_sync(1, SM_EPI_THREADS);  // the summed tile is complete
                switch (P.epi) {
                    case EPI_STORE: small_epi_store(P, s_tile, tile, row0, nrows, t); break;
                    case EPI_LSTM: small_epi_lstm(P, s_tile, tile, row0, nrows, t); break;
                    case EPI_GLU: small_epi_glu(P, s_tile, tile, row0, nrows, t); break;
                    case EPI_TOPK:
                        if (P.ktop <= 4) small_epi_topk<4>(P, s_tile, tile, row0, nrows, t);
                        else small_epi_topk<8>(P, s_tile, tile, row0, nrows, t);
                        break;
                    default: small_epi_sample(P, s_tile, tile, row0, nrows, t); break;
                }
                named_bar_sync(1, SM_EPI_THREADS);  // the tile is rewritten by the next item
                if (t == 0) {
                    if (item == blockIdx.x && q < 2) small_stamp(p.trace, slot, 6 + 6 * q);
                    // the last CTA to leave re-arms the tile's counters for their next use (a later phase or launch)
                    if (P.ksplit > 1 && atomicAdd(ctr + 1, 1) == P.ksplit - 1) ctr[0] = 0, ctr[1] = 0;
                }
            }
        }
        if (q + 1 < p.n_phases) {
            small_grid_barrier(p.bar + 2 * q);
            if (threadIdx.x == 0 && q < 2) small_stamp(p.trace, slot, 7 + 6 * q);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) small_stamp(p.trace, slot, 14);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace capdec
