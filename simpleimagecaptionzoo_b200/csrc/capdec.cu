// libcapdec.so -- host side of the B200-native caption decoder behind the C ABI of include/capdec.h.
// Owns packed weights + workspace, builds the TMA tensor maps, and enqueues the per-step kernel sequence
// (no host synchronisation inside the decode loop).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/capdec.h"
#include "kernels.cuh"
#include "smallm.cuh"
#include "cider.cuh"

using namespace capdec;

namespace {

thread_local std::string g_create_error;

struct Act16 {  // fp16 GEMM operand, row-major, hi | lo halves
    __half* p = nullptr;
    int rows = 0, cols = 0, ld = 0, lo = 0;
};

struct Raw {
    float* d = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;
};

// A GEMM operand as the launchers see it: the tensor map the large-tile kernels load through, plus where it came from (the
// small-batch kernel builds maps with its own box shapes from the same rows).
struct Operand {
    CUtensorMap map;
    const __half* base = nullptr;  // first element (column offset applied)
    int rows = 0, ld = 0, cols = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

constexpr int BN = 256;  // BLOCK_N of every GEMM instantiation

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: set it once per (kernel, device),
// whichever handle on whichever GPU of the process launches the kernel first.
cudaError_t smem_attr(const void* kern, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({kern, dev})) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.insert({kern, dev});
    return e;
}

}  // namespace

struct capdec_handle {
    capdec_config cfg{};
    std::string err;
    int num_sms = 148;
    bool split = false;
    bool weights_ready = false;
    bool prepared = false;
    int B = 0, R = 0;  // prepared batch
    const float* feats = nullptr;
    const float* mask = nullptr;
    int64_t launches = 0;
    bool pair_gemm = true;             // CAPDEC_GEMM_1CTA=1 selects the single-CTA GEMM kernel instead of the CTA-pair one
    int logit_ew = LOGIT_EPI_WARPS;    // CAPDEC_LOGIT_EW=8: the sampling epilogue on 8 warps like the other epilogues (A/B)
    bool mgroup_split = true;          // CAPDEC_MGROUP_SPLIT=0: one row block per pair in the large-M tile walk (round 1's walk)
    bool no_stream_attention = false;  // CAPDEC_NO_STREAM_ATTENTION=1: use the non-persistent attention kernel
    int att_variant = 0;               // CAPDEC_ATT_VARIANT=1: FFMA streaming kernel instead of the MMA-fragment kernel
    // small-batch path (smallm.cuh): swap-AB split-K GEMMs for <= small_rows activation rows (CAPDEC_NO_SMALLM=1 disables,
    // CAPDEC_SMALLM_ROWS overrides the row limit, CAPDEC_NO_FUSE=1 keeps every GEMM of a step in its own launch)
    int small_rows = 128;
    bool small_fuse = true;
    float* small_slabs = nullptr;
    int* small_counters = nullptr;
    unsigned* small_bar = nullptr;  // SM_MAX_PHASES x (arrivals, departures) grid-barrier counters, zero between uses
    // attention kernel CONCURRENT with the [language gates -> logits] launch on the small-batch path (CAPDEC_NO_OVERLAP=1: in series)
    bool small_overlap = true;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int* att_done = nullptr;  // images whose context rows are written, cumulative over a decode
    int* beam_done = nullptr;  // images whose bookkeeping + operand assembly is done, cumulative over a decode
    cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
    bool small_hint = true;      // CAPDEC_SMALL_HINT=0: no L2 eviction-priority hints on the small-batch kernel's weight loads
    bool chain = false;          // CAPDEC_CHAIN=1: top-down gates and dec_att in ONE launch of the chained pair kernel (measured slower)
    int* chain_sync = nullptr;   // [2][row blocks] ready / passed counters of the chained pair kernel (zero between launches)
    int chain_blocks = 0;
    unsigned long long* small_trace = nullptr;  // CAPDEC_TRACE=1: device-side timeline of the small-batch kernel (capdec_debug_trace)
    bool prof = false;  // bracket every launch with CUDA events (capdec_profile)
    struct ProfRec {
        int cat;
        double flops;
        cudaEvent_t a, b;
    };
    std::vector<ProfRec> recs;
    std::vector<void*> allocs;
    std::map<std::string, Raw> raw;

    // dims
    int H = 0, E = 0, A = 0, D = 0, V = 0, NH = 0, Bmax = 0, Rmax = 0, Kmax = 0, Tmax = 0, Mmax = 0;
    int n_tiles_v = 0;

    // packed weights
    Act16 W_pred, W_l1, W_l2, W_aux1, W_aux2, W_aux3, W_emb, emb16;
    float* emb_gates = nullptr;  // [V, 4H] gate pre-activations contributed by each vocabulary word (embedding x W_ih slice)
    float *b_pred = nullptr, *b_l1 = nullptr, *b_l2 = nullptr, *b_aux1 = nullptr, *b_aux2 = nullptr, *b_aux3 = nullptr;
    float* w_aff = nullptr;
    float b_aff = 0.f;
    const float *ln_gain = nullptr, *ln_bias = nullptr;
    double* scale_tmp = nullptr;

    // activations / workspace
    Act16 feats16, enc16, mean16, XA, XB, Hb, Hb2, Xp, H0, k16, v16, q16;
    float *enc_ctx = nullptr, *G0 = nullptr, *dec_ctx = nullptr, *kv32 = nullptr, *mean32 = nullptr, *q32 = nullptr,
          *ctx32 = nullptr, *h32 = nullptr, *c0 = nullptr;
    float* c1[2] = {nullptr, nullptr};
    float* c2[2] = {nullptr, nullptr};
    float* part = nullptr;
    int *tok = nullptr, *parent = nullptr, *n_live = nullptr, *best_seq = nullptr, *best_len = nullptr, *unfinished = nullptr,
        *live_count = nullptr;
    int* seqs[2] = {nullptr, nullptr};
    float *cum = nullptr, *best_score = nullptr;
    // CUDA graph of one whole beam-search decode (all steps), replayed while the shapes / buffers stay the same
    bool use_pdl = false;    // CAPDEC_PDL=1: programmatic dependent launch inside the decode loop (measured neutral)
    bool use_graphs = true;  // CAPDEC_NO_GRAPH=1 disables
    // cache of captured decodes, least recently used entry replaced (a ragged last batch or alternating beam sizes /
    // decode kinds do not re-capture every call)
    enum { GK_BEAM = 0, GK_SAMPLE = 1, GK_SCST = 2, GK_SCORE = 3 };
    struct GraphKey {
        int kind, B, R, K, T, mode, outs;  // K = beam or rows per image; outs = which optional outputs are produced
        const void* feats;  // fp32-grade mode reads the caller's features inside the loop
        bool masked;
        bool operator==(const GraphKey& o) const {
            return kind == o.kind && B == o.B && R == o.R && K == o.K && T == o.T && mode == o.mode && outs == o.outs &&
                   feats == o.feats && masked == o.masked;
        }
    };
    struct GraphEntry {
        GraphKey key;
        cudaGraphExec_t exec;
        int64_t launches;
        uint64_t used;
    };
    static constexpr size_t GRAPH_CACHE = 8;
    std::vector<GraphEntry> graphs;
    uint64_t graph_clock = 0;
    int64_t graph_captures = 0;
    uint32_t* seed_dev = nullptr;     // sampling seed of the (replayed) rollout, written before every launch
    int* out_sample_tokens = nullptr;  // library-owned outputs of a captured rollout (copied to the caller's buffers)
    float* out_sample_logprobs = nullptr;
    int* out_greedy = nullptr;
    int* forced_buf = nullptr;         // teacher-forced words of capdec_score at a stable address
    float* states_out = nullptr;       // capdec_score_states: where the per-step predict inputs go ([M, T, H] fp32) or null
    float* mask_buf = nullptr;  // library-owned copy of the region mask (stable address for the captured decode)
    int* out_tokens = nullptr;
    float* out_scores = nullptr;
    int* out_lengths = nullptr;
    float* alpha_step = nullptr;  // [Tmax, Mmax, Rmax] per-step attention maps (allocated on first request)
    int *hist_parent = nullptr, *best_pslot = nullptr;

    // AoA encoder side in front of the decoder (img_feats_porjection + aoa_refine, AoA_Model.py:122-162,661-665):
    // present when the checkpoint carries those entries; workspace allocated when they are finalized
    struct RefineLayer {
        Act16 W_qkv, W_glu;  // [3H, H] = linear_Q ; linear_K ; linear_V stacked, [2H, 2H] (a, gate)-interleaved
        float *b_qkv = nullptr, *b_glu = nullptr;
        const float *ln_gain = nullptr, *ln_bias = nullptr;
    };
    std::vector<RefineLayer> refine;
    bool refiner_ready = false, refiner_allocated = false;
    Act16 W_proj, bu16, XR, qkv16;  // projection weight [H, D]; fp16 bottom-up features; [att | LN(x)] operand; Q|K|V (fp16 mode)
    float *b_proj = nullptr, *qkv32 = nullptr, *xres = nullptr, *refined = nullptr;
    const float *rfinal_gain = nullptr, *rfinal_bias = nullptr;
};

namespace {

#define CK(h, expr)                                                                                         \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess) {                                                                            \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                                  \
            return CAPDEC_ERR_CUDA;                                                                         \
        }                                                                                                   \
    } while (0)

#define CKS(h, expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s != CAPDEC_OK) return _s; \
    } while (0)

int fail(capdec_handle* h, int code, const std::string& msg) {
    h->err = msg;
    return code;
}

// Optional per-launch timing (CUDA events on the launching stream), by kernel category.
void prof_begin(capdec_handle* h, int cat, double flops, cudaStream_t st) {
    if (!h->prof) return;
    capdec_handle::ProfRec r{cat, flops, nullptr, nullptr};
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    h->recs.push_back(r);
}
void prof_end(capdec_handle* h, cudaStream_t st) {
    if (!h->prof || h->recs.empty()) return;
    cudaEventRecord(h->recs.back().b, st);
}

// Launch a decode-loop kernel with programmatic dependent launch: its blocks may be scheduled, and run their set-up, while
// the previous kernel of the stream drains (every such kernel calls griddep_wait() before it reads what a predecessor
// wrote).  Not while per-launch timing is on (the event records between launches would hide the overlap anyway).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(const capdec_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (h->use_pdl && !h->prof) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename T>
int dalloc(capdec_handle* h, T** out, size_t n, bool zero = true) {
    void* p = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        return CAPDEC_ERR_NOMEM;
    }
    if (zero) CK(h, cudaMemset(p, 0, bytes));
    h->allocs.push_back(p);
    *out = static_cast<T*>(p);
    return CAPDEC_OK;
}

// pad > 0 (fp16 mode only): extra halves per row, so that rows bulk-copied to shared memory are bank-conflict free
int alloc_act(capdec_handle* h, Act16* a, int rows, int cols, int pad = 0) {
    a->rows = rows;
    a->cols = cols;
    a->ld = cols * (h->split ? 2 : 1) + (h->split ? 0 : pad);
    a->lo = h->split ? cols : 0;
    return dalloc(h, &a->p, static_cast<size_t>(rows) * a->ld);
}

// Tensor map over rows x (ld - col_off) fp16 starting at column col_off; box = 64 (K) x box_rows, 128B swizzle.
int make_map(capdec_handle* h, CUtensorMap* m, const __half* base, int rows, int ld, int col_off, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(h, CAPDEC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ld - col_off), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(__half)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base + col_off), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, CAPDEC_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(static_cast<int>(r)));
    return CAPDEC_OK;
}
int make_operand(capdec_handle* h, Operand* o, const __half* base, int rows, int ld, int col_off, int box_rows) {
    o->base = base + col_off;
    o->rows = rows;
    o->ld = ld;
    o->cols = ld - col_off;
    return make_map(h, &o->map, base, rows, ld, col_off, box_rows);
}
int map_a(capdec_handle* h, Operand* m, const Act16& a, int col_off = 0) {
    return make_operand(h, m, a.p, a.rows, a.ld, col_off, BLOCK_M);
}
// weight operand: the single-CTA kernel loads 256-row boxes, the CTA-pair kernel 128-row halves
int map_b(capdec_handle* h, Operand* m, const Act16& a) {
    return make_operand(h, m, a.p, a.rows, a.ld, 0, h->pair_gemm ? BN / 2 : BN);
}

template <int EPI, int KTOP>
int launch_gemm_t(capdec_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    auto kern = gemm_kernel<BN, EPI, KTOP>;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES));
    const int tiles = p.runs > 0 ? p.num_m_blocks * p.runs : p.num_m_blocks * p.num_n_blocks;
    const int grid = tiles < h->num_sms ? tiles : h->num_sms;
    GemmParams pg = p;
    pg.m_group = grid;
    const int cat = EPI == EPI_LSTM ? CAPDEC_CAT_GEMM_LSTM : EPI == EPI_STORE ? CAPDEC_CAT_GEMM_STORE
                  : EPI == EPI_GLU ? CAPDEC_CAT_GEMM_GLU : CAPDEC_CAT_GEMM_LOGITS;
    prof_begin(h, cat, 2.0 * p.M * p.N * (static_cast<double>(p.k_blocks) * BLOCK_K), st);
    CK(h, launch_pdl(h, kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, ma, mb, pg));
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// D[M,N] = A[M,Kdim] * B[N,Kdim]^T with the chosen epilogue.
template <int EPI, int KTOP, int EW = EPI_WARPS>
int launch_gemm2_t(capdec_handle* h, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, cudaStream_t st) {
    auto kern = gemm2_kernel<EPI, KTOP, EW>;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), GemmCfg2::SMEM_BYTES));
    const int items = p.runs > 0 ? p.num_m_blocks * p.runs : p.num_m_blocks * p.num_n_blocks;
    const int pairs = items < h->num_sms / 2 ? items : h->num_sms / 2;
    GemmParams pg = p;
    // Tile walk: groups of m_group row blocks, all column blocks of a group before the next group.  With one row block per
    // pair (m_group = pairs) a pair re-reads its 256 x K operand rows once per column block, with the other pairs' rows and
    // the group's output in between -- ~110 MB at K = 2048, more than L2 keeps: the projection GEMM read 1.66x its algorithmic
    // bytes.  Groups of pairs / column-blocks row blocks put the column blocks of a row block on DIFFERENT pairs at the same
    // time, so the operand rows are fetched once and shared in L2 while they are hot.
    pg.m_group = pairs;
    if (h->mgroup_split && p.runs == 0 && p.num_m_blocks >= pairs && p.num_n_blocks >= 2) {
        pg.m_group = pairs / p.num_n_blocks;
        if (pg.m_group < 1) pg.m_group = 1;
    }
    const int cat = EPI == EPI_LSTM ? CAPDEC_CAT_GEMM_LSTM : EPI == EPI_STORE ? CAPDEC_CAT_GEMM_STORE
                  : EPI == EPI_GLU ? CAPDEC_CAT_GEMM_GLU : CAPDEC_CAT_GEMM_LOGITS;
    prof_begin(h, cat, 2.0 * p.M * p.N * (static_cast<double>(p.k_blocks) * BLOCK_K), st);
    CK(h, launch_pdl(h, kern, dim3(2 * pairs), dim3(64 + 32 * EW), GemmCfg2::SMEM_BYTES, st, ma, mb, pg));
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// number of vocabulary-tile runs per row block of the logit GEMM: fill the SMs with (row block, run) items
int logit_runs(const capdec_handle* h, int M, int N) {
    const int rows = h->pair_gemm ? 2 * BLOCK_M : BLOCK_M;  // row-block height of a work item
    const int workers = h->pair_gemm ? h->num_sms / 2 : h->num_sms;
    const int num_m = (M + rows - 1) / rows, num_n = (N + BN - 1) / BN;
    int runs = workers / num_m;
    if (runs < 1) runs = 1;
    if (runs > num_n) runs = num_n;
    return runs;
}

// ------------------------------------------------------------------------------------------------ small-batch path
struct SmallDesc {
    int epi, ktop;
    const Operand* x;  // activations [M, K]
    int x_lo;
    const Operand* w;  // weights [N, K]
    int w_lo;
    int M, N, Kdim;
    EpiParams e;
    const int* wait_ctr = nullptr;  // see SmallPhase::wait_ctr
    int wait_target = 0, wait_kb = 0;
};

bool small_ok(const capdec_handle* h, int M) { return h->small_slabs != nullptr && M <= h->small_rows; }

constexpr int SMALL_MAX_ITEMS = 2 * 148;

// K split of a phase.  An item costs its k-blocks (~0.3 us each: the weight stream is L2-bound) plus a fixed ~6 us chain of
// L2 round trips (partial store, arrive / wait at the tile's counter, reduction loads, epilogue), so a second round over
// the SMs is never worth it: the largest split that keeps every item in ONE round wins -- top-down gates (32 tiles x 32
// k-blocks) -> 4, dec_att (8 x 16) -> 8, language gates (32 x 64) -> 4, logits (75 x 16) -> 1
int small_ksplit(int tiles, int total_kb, int num_sms) {
    int best = 1;
    double best_cost = 1e30;
    for (int ks = 1; ks <= 8 && ks <= total_kb; ++ks) {
        const int items = tiles * ks;
        if (items > SMALL_MAX_ITEMS) break;
        const int rounds = (items + num_sms - 1) / num_sms;
        const int per = (total_kb + ks - 1) / ks;
        const double cost = rounds * (per + 20.0) + 0.5 * ((ks + 3) / 4);  // + one reduction round trip per four splits
        if (cost < best_cost) best_cost = cost, best = ks;
    }
    return best;
}

template <int N_ACT>
int launch_small_t(capdec_handle* h, SmallParams& p, const SmallDesc* d, int n, cudaStream_t st, int sms) {
    using C = SmallCfg<N_ACT>;
    auto kern = smallm_kernel<N_ACT>;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
    int max_items = 1;
    for (int q = 0; q < n; ++q) {
        SmallPhase& P = p.ph[q];
        CKS(h, make_map(h, &P.map_w, d[q].w->base, d[q].w->rows, d[q].w->ld, 0, SM_TILE_N));
        CKS(h, make_map(h, &P.map_x, d[q].x->base, d[q].x->rows, d[q].x->ld, 0, N_ACT));
        if (P.tiles * P.ksplit > max_items) max_items = P.tiles * P.ksplit;
    }
    const int grid = max_items < sms ? max_items : sms;
    CK(h, launch_pdl(h, kern, dim3(n > 1 ? sms : grid), dim3(SM_THREADS), C::SMEM_BYTES, st, p));
    return CAPDEC_OK;
}

// One launch for up to SM_MAX_PHASES dependent GEMMs on the same <= small_rows activation rows.
int launch_small(capdec_handle* h, const SmallDesc* d, int n, cudaStream_t st, int grid_limit = 0) {
    if (n < 1 || n > SM_MAX_PHASES) return fail(h, CAPDEC_ERR_INVALID, "small-batch launch: bad phase count");
    const int sms = (grid_limit > 0 && grid_limit < h->num_sms) ? grid_limit : h->num_sms;  // SMs this launch may count on
    SmallParams p{};
    p.n_phases = n;
    p.slabs = h->small_slabs;
    p.counters = h->small_counters;
    p.bar = h->small_bar;
    p.trace = h->small_trace;
    int max_m = 0;
    double flops = 0.0;
    for (int q = 0; q < n; ++q) {
        const SmallDesc& D = d[q];
        if (D.Kdim % BLOCK_K || D.M <= 0 || D.N <= 0) return fail(h, CAPDEC_ERR_INVALID, "small-batch GEMM: K must be a multiple of 64");
        SmallPhase& P = p.ph[q];
        P.N_w = D.N, P.M = D.M;
        P.k_blocks = D.Kdim / BLOCK_K;
        P.passes = h->split ? 3 : 1;
        P.w_lo_off = D.w_lo, P.x_lo_off = D.x_lo;
        P.tiles = (D.N + SM_TILE_N - 1) / SM_TILE_N;
        P.ksplit = small_ksplit(P.tiles, P.k_blocks * P.passes, sms);
        P.epi = D.epi, P.ktop = D.ktop;
        P.wait_ctr = D.wait_ctr, P.wait_target = D.wait_target, P.wait_kb = D.wait_kb;
        // the step's weights (72 MB) cycle through an L2 that keeps ~60 MB of them: the vocabulary matrix is streamed
        // evict-first so that the gate matrices (evict-last) survive from one step to the next
        P.w_hint = h->small_hint ? ((D.epi == EPI_TOPK || D.epi == EPI_SAMPLE) ? 1 : 2) : 0;
        P.e = D.e;
        if (D.epi == EPI_TOPK || D.epi == EPI_SAMPLE) P.e.n_tiles = P.tiles;  // one partial record per (row, 128-word tile)
        if (D.M > max_m) max_m = D.M;
        flops += 2.0 * D.M * D.N * D.Kdim;
    }
    const int epi0 = d[0].epi;
    const int cat = epi0 == EPI_LSTM ? CAPDEC_CAT_GEMM_LSTM : epi0 == EPI_STORE ? CAPDEC_CAT_GEMM_STORE
                  : epi0 == EPI_GLU ? CAPDEC_CAT_GEMM_GLU : CAPDEC_CAT_GEMM_LOGITS;
    prof_begin(h, cat, flops, st);
    int status;
    if (max_m <= 16) status = launch_small_t<16>(h, p, d, n, st, sms);
    else if (max_m <= 64) status = launch_small_t<64>(h, p, d, n, st, sms);
    else status = launch_small_t<128>(h, p, d, n, st, sms);
    prof_end(h, st);
    CKS(h, status);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// number of partial records per row the logit GEMM writes for M rows (what the bookkeeping kernels merge)
// (the pair kernel's sampling epilogue runs on logit_ew warps, its top-k epilogue on EPI_WARPS: see dispatch)
int logit_slots(const capdec_handle* h, int M, int N, int epi) {
    if (small_ok(h, M)) return (N + SM_TILE_N - 1) / SM_TILE_N;
    return logit_runs(h, M, N) * (h->pair_gemm && epi == EPI_SAMPLE ? h->logit_ew / 4 : EPI_SPLIT);
}

int alloc_small(capdec_handle* h) {
    const char* off = getenv("CAPDEC_NO_SMALLM");
    if (off && off[0] == '1') return CAPDEC_OK;
    const char* rows = getenv("CAPDEC_SMALLM_ROWS");
    if (rows) h->small_rows = atoi(rows) < 128 ? atoi(rows) : 128;
    const char* hint = getenv("CAPDEC_SMALL_HINT");
    h->small_hint = !(hint && hint[0] == '0');
    const char* nf = getenv("CAPDEC_NO_FUSE");
    h->small_fuse = !(nf && nf[0] == '1');
    if (h->small_rows <= 0) return CAPDEC_OK;
    CKS(h, dalloc(h, &h->small_counters, 4096));
    {
        const char* tr = getenv("CAPDEC_TRACE");
        if (tr && tr[0] == '1') CKS(h, dalloc(h, &h->small_trace, 1 + 16 * 4000));
    }
    CKS(h, dalloc(h, &h->small_bar, 2 * SM_MAX_PHASES));
    return dalloc(h, &h->small_slabs, static_cast<size_t>(SMALL_MAX_ITEMS) * 128 * SM_TILE_N, false);
}

int launch_gemm(capdec_handle* h, int epi, int ktop, const Operand& oa, int a_lo, const Operand& ob, int b_lo, int M,
                int N, int Kdim, const EpiParams& e, cudaStream_t st) {
    if (Kdim % BLOCK_K) return fail(h, CAPDEC_ERR_INVALID, "GEMM K must be a multiple of 64");
    if (M <= 0 || N <= 0) return fail(h, CAPDEC_ERR_INVALID, "empty GEMM");
    if (small_ok(h, M)) {
        SmallDesc d{epi, ktop, &oa, a_lo, &ob, b_lo, M, N, Kdim, e};
        return launch_small(h, &d, 1, st);
    }
    const CUtensorMap& ma = oa.map;
    const CUtensorMap& mb = ob.map;
    GemmParams p{};
    p.M = M;
    p.N = N;
    p.k_blocks = Kdim / BLOCK_K;
    p.passes = h->split ? 3 : 1;
    p.a_lo_off = a_lo;
    p.b_lo_off = b_lo;
    p.num_m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    p.num_n_blocks = (N + BN - 1) / BN;
    p.runs = (epi == EPI_TOPK || epi == EPI_SAMPLE) ? logit_runs(h, M, N) : 0;
    p.epi = e;
    if (h->pair_gemm) {  // CTA-pair kernel: 256-row blocks; `mb` must be a 128-row-box map (map_b)
        p.num_m_blocks = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
        switch (epi) {
            case EPI_STORE: return launch_gemm2_t<EPI_STORE, 1>(h, ma, mb, p, st);
            case EPI_LSTM: return launch_gemm2_t<EPI_LSTM, 1>(h, ma, mb, p, st);
            case EPI_GLU: return launch_gemm2_t<EPI_GLU, 1>(h, ma, mb, p, st);
            case EPI_SAMPLE:
                if (h->logit_ew == LOGIT_EPI_WARPS) return launch_gemm2_t<EPI_SAMPLE, 1, LOGIT_EPI_WARPS>(h, ma, mb, p, st);
                return launch_gemm2_t<EPI_SAMPLE, 1>(h, ma, mb, p, st);
            case EPI_TOPK:
                if (ktop <= 4) return launch_gemm2_t<EPI_TOPK, 4>(h, ma, mb, p, st);
                return launch_gemm2_t<EPI_TOPK, 8>(h, ma, mb, p, st);
        }
        return fail(h, CAPDEC_ERR_INVALID, "unknown epilogue");
    }
    switch (epi) {
        case EPI_STORE: return launch_gemm_t<EPI_STORE, 1>(h, ma, mb, p, st);
        case EPI_LSTM: return launch_gemm_t<EPI_LSTM, 1>(h, ma, mb, p, st);
        case EPI_GLU: return launch_gemm_t<EPI_GLU, 1>(h, ma, mb, p, st);
        case EPI_SAMPLE: return launch_gemm_t<EPI_SAMPLE, 1>(h, ma, mb, p, st);
        case EPI_TOPK:
            if (ktop <= 4) return launch_gemm_t<EPI_TOPK, 4>(h, ma, mb, p, st);
            return launch_gemm_t<EPI_TOPK, 8>(h, ma, mb, p, st);
    }
    return fail(h, CAPDEC_ERR_INVALID, "unknown epilogue");
}

int ktop_for(int beam) { return beam <= 4 ? 4 : 8; }

// ------------------------------------------------------------------------------------------------ weights
const Raw* find_raw(capdec_handle* h, const std::string& name) {
    auto it = h->raw.find(name);
    return it == h->raw.end() ? nullptr : &it->second;
}

int need(capdec_handle* h, const std::string& name, std::vector<int64_t> shape, const Raw** out) {
    const Raw* r = find_raw(h, name);
    if (!r) return fail(h, CAPDEC_ERR_WEIGHT, "missing state_dict entry: " + name);
    if (r->shape != shape) {
        std::string s = "shape mismatch for " + name + ": got [";
        for (auto d : r->shape) s += std::to_string(d) + ",";
        s += "] expected [";
        for (auto d : shape) s += std::to_string(d) + ",";
        return fail(h, CAPDEC_ERR_WEIGHT, s + "]");
    }
    *out = r;
    return CAPDEC_OK;
}

int grid_for(size_t n, int block = 256) {
    size_t g = (n + block - 1) / block;
    return static_cast<int>(g > 4096 ? 4096 : (g ? g : 1));
}

// pack src[:, src_col : src_col+K] (optionally weight-norm scaled by g) into dst[:, dst_col : dst_col+K]
int pack_segment(capdec_handle* h, const Raw* src, int src_col, int K, const Raw* g, Act16& dst, int dst_col, int mode, int Hh,
                 cudaStream_t st) {
    const int N = dst.rows;
    const int src_ld = static_cast<int>(src->shape[1]);
    const double* scale = nullptr;
    if (g) {
        weightnorm_scale_kernel<<<N, 256, 0, st>>>(g->d, src->d, N, src_ld, h->scale_tmp);
        CK(h, cudaGetLastError());
        scale = h->scale_tmp;
    }
    pack_weight_kernel<<<grid_for(static_cast<size_t>(N) * K), 256, 0, st>>>(src->d, src_ld, src_col, scale, dst.p, dst.ld, dst.lo,
                                                                             dst_col, N, K, mode, Hh);
    CK(h, cudaGetLastError());
    return CAPDEC_OK;
}

int pack_bias(capdec_handle* h, const Raw* a, const Raw* b, float* dst, int N, int mode, int Hh, cudaStream_t st) {
    pack_bias_kernel<<<(N + 255) / 256, 256, 0, st>>>(a->d, b ? b->d : nullptr, dst, N, mode, Hh);
    CK(h, cudaGetLastError());
    return CAPDEC_OK;
}

// emb_gates[v, :] = act(embed[v, :]) * W_emb^T : the fed-back word's contribution to the LSTM gates becomes a row
// gather in the gate GEMM's epilogue instead of E columns of its K loop (embed + Linear folded once at load;
// BUTD_Model.py:264-265, NIC_Model.py:172-173, AoA_Model.py:439-441).
int build_embedding_gates(capdec_handle* h, const Raw* emb, int relu, cudaStream_t st) {
    const int V = h->V, E = h->E, H = h->H;
    cvt_f16_kernel<<<grid_for(static_cast<size_t>(V) * E / 4), 256, 0, st>>>(emb->d, V, E, h->emb16.p, h->emb16.ld, h->emb16.lo, 0, relu);
    CK(h, cudaGetLastError());
    Operand ma, mb;
    CKS(h, map_a(h, &ma, h->emb16));
    CKS(h, map_b(h, &mb, h->W_emb));
    EpiParams e{};
    e.out32 = h->emb_gates;
    e.ld32 = 4 * H;
    return launch_gemm(h, EPI_STORE, 1, ma, h->emb16.lo, mb, h->W_emb.lo, V, 4 * H, E, e, st);
}

int finalize_predict(capdec_handle* h, cudaStream_t st) {
    const Raw *g, *v, *b;
    CKS(h, need(h, "predict.weight_g", {h->V, 1}, &g));
    CKS(h, need(h, "predict.weight_v", {h->V, h->H}, &v));
    CKS(h, need(h, "predict.bias", {h->V}, &b));
    CKS(h, pack_segment(h, v, 0, h->H, g, h->W_pred, 0, 0, 0, st));
    CKS(h, pack_bias(h, b, nullptr, h->b_pred, h->V, 0, 0, st));
    return CAPDEC_OK;
}

int finalize_butd(capdec_handle* h, cudaStream_t st) {
    const int H = h->H, E = h->E, A = h->A, D = h->D;
    const Raw *g, *v, *b, *wih, *whh, *bih, *bhh;
    // attention projections (weight-normed Linears, BUTD_Model.py:43-45)
    CKS(h, need(h, "atten.enc_att.weight_g", {A, 1}, &g));
    CKS(h, need(h, "atten.enc_att.weight_v", {A, D}, &v));
    CKS(h, need(h, "atten.enc_att.bias", {A}, &b));
    CKS(h, pack_segment(h, v, 0, D, g, h->W_aux1, 0, 0, 0, st));  // W_aux1 = enc_att [A, D]
    CKS(h, pack_bias(h, b, nullptr, h->b_aux1, A, 0, 0, st));
    CKS(h, need(h, "atten.dec_att.weight_g", {A, 1}, &g));
    CKS(h, need(h, "atten.dec_att.weight_v", {A, H}, &v));
    CKS(h, need(h, "atten.dec_att.bias", {A}, &b));
    CKS(h, pack_segment(h, v, 0, H, g, h->W_aux2, 0, 0, 0, st));  // W_aux2 = dec_att [A, H]
    CKS(h, pack_bias(h, b, nullptr, h->b_aux2, A, 0, 0, st));
    CKS(h, need(h, "atten.affine.weight_g", {1, 1}, &g));
    CKS(h, need(h, "atten.affine.weight_v", {1, A}, &v));
    CKS(h, need(h, "atten.affine.bias", {1}, &b));
    weightnorm_scale_kernel<<<1, 256, 0, st>>>(g->d, v->d, 1, A, h->scale_tmp);
    CK(h, cudaGetLastError());
    fold_vector_kernel<<<(A + 255) / 256, 256, 0, st>>>(v->d, h->scale_tmp, h->w_aff, A);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(&h->b_aff, b->d, sizeof(float), cudaMemcpyDeviceToHost, st));
    // top-down attention LSTM: input cat[h2, mean, emb] (BUTD_Model.py:265) -> operand [h2 | h1]; the mean slice is
    // hoisted per image (W_aux3), the emb slice per vocabulary word (W_emb -> emb_gates)
    CKS(h, need(h, "TD_atten.weight_ih", {4 * H, H + D + E}, &wih));
    CKS(h, need(h, "TD_atten.weight_hh", {4 * H, H}, &whh));
    CKS(h, need(h, "TD_atten.bias_ih", {4 * H}, &bih));
    CKS(h, need(h, "TD_atten.bias_hh", {4 * H}, &bhh));
    CKS(h, pack_segment(h, wih, 0, H, nullptr, h->W_l1, 0, 1, H, st));
    CKS(h, pack_segment(h, whh, 0, H, nullptr, h->W_l1, H, 1, H, st));
    CKS(h, pack_segment(h, wih, H + D, E, nullptr, h->W_emb, 0, 1, H, st));
    CKS(h, pack_segment(h, wih, H, D, nullptr, h->W_aux3, 0, 1, H, st));  // W_aux3 = mean-feature slice [4H, D]
    CKS(h, pack_bias(h, bih, bhh, h->b_l1, 4 * H, 1, H, st));
    // language LSTM: input cat[ctx, h1] (BUTD_Model.py:268) -> operand [ctx | h1 | h2]
    CKS(h, need(h, "language_model.weight_ih", {4 * H, D + H}, &wih));
    CKS(h, need(h, "language_model.weight_hh", {4 * H, H}, &whh));
    CKS(h, need(h, "language_model.bias_ih", {4 * H}, &bih));
    CKS(h, need(h, "language_model.bias_hh", {4 * H}, &bhh));
    CKS(h, pack_segment(h, wih, 0, D + H, nullptr, h->W_l2, 0, 1, H, st));
    CKS(h, pack_segment(h, whh, 0, H, nullptr, h->W_l2, D + H, 1, H, st));
    CKS(h, pack_bias(h, bih, bhh, h->b_l2, 4 * H, 1, H, st));
    const Raw* emb;
    CKS(h, need(h, "embed.0.weight", {h->V, E}, &emb));
    return build_embedding_gates(h, emb, 1, st);  // embed = Embedding + ReLU (BUTD_Model.py:77-81)
}

int finalize_nic(capdec_handle* h, cudaStream_t st) {
    const int H = h->H, E = h->E;
    const Raw *wih, *whh, *bih, *bhh, *emb;
    CKS(h, need(h, "lstm.weight_ih", {4 * H, E}, &wih));
    CKS(h, need(h, "lstm.weight_hh", {4 * H, H}, &whh));
    CKS(h, need(h, "lstm.bias_ih", {4 * H}, &bih));
    CKS(h, need(h, "lstm.bias_hh", {4 * H}, &bhh));
    CKS(h, pack_segment(h, wih, 0, E, nullptr, h->W_emb, 0, 1, H, st));  // also the operand of the priming step
    CKS(h, pack_segment(h, whh, 0, H, nullptr, h->W_l1, 0, 1, H, st));
    CKS(h, pack_bias(h, bih, bhh, h->b_l1, 4 * H, 1, H, st));
    CKS(h, need(h, "embed.weight", {h->V, E}, &emb));
    return build_embedding_gates(h, emb, 0, st);  // plain nn.Embedding (NIC_Model.py:47)
}

int finalize_aoa(capdec_handle* h, cudaStream_t st) {
    const int H = h->H, E = h->E;
    const Raw *wih, *whh, *bih, *bhh, *w, *b, *emb, *gn, *bs;
    // lstm input cat[emb, mean+ctx] (AoA_Model.py:441) -> operand [mean+ctx | h]; emb slice -> emb_gates
    CKS(h, need(h, "lstm.weight_ih", {4 * H, E + H}, &wih));
    CKS(h, need(h, "lstm.weight_hh", {4 * H, H}, &whh));
    CKS(h, need(h, "lstm.bias_ih", {4 * H}, &bih));
    CKS(h, need(h, "lstm.bias_hh", {4 * H}, &bhh));
    CKS(h, pack_segment(h, wih, E, H, nullptr, h->W_l1, 0, 1, H, st));
    CKS(h, pack_segment(h, whh, 0, H, nullptr, h->W_l1, H, 1, H, st));
    CKS(h, pack_segment(h, wih, 0, E, nullptr, h->W_emb, 0, 1, H, st));
    CKS(h, pack_bias(h, bih, bhh, h->b_l1, 4 * H, 1, H, st));
    CKS(h, need(h, "aoa_block.linear_Q.weight", {H, H}, &w));
    CKS(h, need(h, "aoa_block.linear_Q.bias", {H}, &b));
    CKS(h, pack_segment(h, w, 0, H, nullptr, h->W_aux1, 0, 0, 0, st));  // W_aux1 = linear_Q
    CKS(h, pack_bias(h, b, nullptr, h->b_aux1, H, 0, 0, st));
    // W_aux2 = [linear_K ; linear_V] stacked along N -> one projection GEMM, kv [B*R, 2H]
    CKS(h, need(h, "aoa_block.linear_K.weight", {H, H}, &w));
    CKS(h, need(h, "aoa_block.linear_K.bias", {H}, &b));
    {
        Act16 top = h->W_aux2;
        top.rows = H;
        CKS(h, pack_segment(h, w, 0, H, nullptr, top, 0, 0, 0, st));
        CKS(h, pack_bias(h, b, nullptr, h->b_aux2, H, 0, 0, st));
    }
    CKS(h, need(h, "aoa_block.linear_V.weight", {H, H}, &w));
    CKS(h, need(h, "aoa_block.linear_V.bias", {H}, &b));
    {
        Act16 bot = h->W_aux2;
        bot.rows = H;
        bot.p = h->W_aux2.p + static_cast<size_t>(H) * h->W_aux2.ld;
        CKS(h, pack_segment(h, w, 0, H, nullptr, bot, 0, 0, 0, st));
        CKS(h, pack_bias(h, b, nullptr, h->b_aux2 + H, H, 0, 0, st));
    }
    // AoA information/gate Linear(2H -> 2H) + GLU on cat[att, query] (AoA_Model.py:85-88,118)
    CKS(h, need(h, "aoa_block.aoa_module.0.weight", {2 * H, 2 * H}, &w));
    CKS(h, need(h, "aoa_block.aoa_module.0.bias", {2 * H}, &b));
    CKS(h, pack_segment(h, w, 0, 2 * H, nullptr, h->W_aux3, 0, 2, H, st));
    CKS(h, pack_bias(h, b, nullptr, h->b_aux3, 2 * H, 2, H, st));
    CKS(h, need(h, "embed.0.weight", {h->V, E}, &emb));
    CKS(h, need(h, "h_norm.gain", {H}, &gn));
    CKS(h, need(h, "h_norm.bias", {H}, &bs));
    h->ln_gain = gn->d;
    h->ln_bias = bs->d;
    return build_embedding_gates(h, emb, 1, st);  // embed = Embedding + ReLU (AoA_Model.py:206-210)
}

// reference LayerNorm over the rows of x [M, H] -> fp16 operand and/or fp32 copy
int launch_layernorm(capdec_handle* h, const float* x, int M, const float* gain, const float* bias, __half* q16, int ld16, int lo16,
                     float* out32, cudaStream_t st) {
    prof_begin(h, CAPDEC_CAT_OTHER, 0.0, st);
    const dim3 grid((M + 7) / 8), block(256);
    if (h->H == 1024) CK(h, launch_pdl(h, aoa_layernorm_vec_kernel<8>, grid, block, 0, st, x, M, gain, bias, 1e-6f, q16, ld16, lo16, out32));
    else if (h->H == 512) CK(h, launch_pdl(h, aoa_layernorm_vec_kernel<4>, grid, block, 0, st, x, M, gain, bias, 1e-6f, q16, ld16, lo16, out32));
    else CK(h, launch_pdl(h, aoa_layernorm_kernel, grid, block, 0, st, x, M, h->H, gain, bias, 1e-6f, q16, ld16, lo16, out32));
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// ------------------------------------------------------------------------------------------------ AoA encoder side
bool has_refiner_weights(capdec_handle* h) { return find_raw(h, "img_feats_porjection.0.weight") != nullptr; }

// Pack img_feats_porjection + every aoa_refine layer found in the checkpoint; allocate the refiner workspace (once).
int finalize_refiner(capdec_handle* h, cudaStream_t st) {
    const int H = h->H, D = h->D;
    if (D <= 0 || D % 64) return fail(h, CAPDEC_ERR_INVALID, "AoA refiner needs enc_dim (bottom-up feature width) as a multiple of 64");
    const Raw *w, *b, *gn, *bs;
    const bool first = !h->refiner_allocated;  // a failed first attempt (missing entry) frees nothing and is not repeated
    if (first && !h->refine.empty()) return fail(h, CAPDEC_ERR_STATE, "an earlier finalize of the AoA refiner failed; create a new handle");
    if (first) {
        CKS(h, alloc_act(h, &h->W_proj, H, D));
        CKS(h, dalloc(h, &h->b_proj, H));
    }
    CKS(h, need(h, "img_feats_porjection.0.weight", {H, D}, &w));
    CKS(h, need(h, "img_feats_porjection.0.bias", {H}, &b));
    CKS(h, pack_segment(h, w, 0, D, nullptr, h->W_proj, 0, 0, 0, st));
    CKS(h, pack_bias(h, b, nullptr, h->b_proj, H, 0, 0, st));
    int n_layers = 0;
    while (find_raw(h, "aoa_refine.aoa_layers." + std::to_string(n_layers) + ".aoa_block.linear_Q.weight")) ++n_layers;
    if (n_layers == 0) return fail(h, CAPDEC_ERR_WEIGHT, "missing state_dict entry: aoa_refine.aoa_layers.0.aoa_block.linear_Q.weight");
    if (!first && static_cast<int>(h->refine.size()) != n_layers) return fail(h, CAPDEC_ERR_WEIGHT, "number of aoa_refine layers changed");
    if (first) h->refine.resize(n_layers);
    for (int l = 0; l < n_layers; ++l) {
        capdec_handle::RefineLayer& L = h->refine[l];
        const std::string p = "aoa_refine.aoa_layers." + std::to_string(l) + ".";
        if (first) {
            CKS(h, alloc_act(h, &L.W_qkv, 3 * H, H));
            CKS(h, alloc_act(h, &L.W_glu, 2 * H, 2 * H));
            CKS(h, dalloc(h, &L.b_qkv, 3 * H));
            CKS(h, dalloc(h, &L.b_glu, 2 * H));
        }
        const char* names[3] = {"linear_Q", "linear_K", "linear_V"};
        for (int part = 0; part < 3; ++part) {  // one fused projection GEMM per layer: rows [Q ; K ; V]
            CKS(h, need(h, p + "aoa_block." + names[part] + ".weight", {H, H}, &w));
            CKS(h, need(h, p + "aoa_block." + names[part] + ".bias", {H}, &b));
            Act16 blk = L.W_qkv;
            blk.rows = H;
            blk.p = L.W_qkv.p + static_cast<size_t>(part) * H * L.W_qkv.ld;
            CKS(h, pack_segment(h, w, 0, H, nullptr, blk, 0, 0, 0, st));
            CKS(h, pack_bias(h, b, nullptr, L.b_qkv + part * H, H, 0, 0, st));
        }
        CKS(h, need(h, p + "aoa_block.aoa_module.0.weight", {2 * H, 2 * H}, &w));
        CKS(h, need(h, p + "aoa_block.aoa_module.0.bias", {2 * H}, &b));
        CKS(h, pack_segment(h, w, 0, 2 * H, nullptr, L.W_glu, 0, 2, H, st));
        CKS(h, pack_bias(h, b, nullptr, L.b_glu, 2 * H, 2, H, st));
        CKS(h, need(h, p + "sublayer.norm.gain", {H}, &gn));
        CKS(h, need(h, p + "sublayer.norm.bias", {H}, &bs));
        L.ln_gain = gn->d;
        L.ln_bias = bs->d;
    }
    CKS(h, need(h, "aoa_refine.norm.gain", {H}, &gn));
    CKS(h, need(h, "aoa_refine.norm.bias", {H}, &bs));
    h->rfinal_gain = gn->d;
    h->rfinal_bias = bs->d;
    if (first) {
        const size_t BR = static_cast<size_t>(h->Bmax) * h->Rmax;
        CKS(h, alloc_act(h, &h->bu16, static_cast<int>(BR), D));
        CKS(h, alloc_act(h, &h->XR, static_cast<int>(BR), 2 * H));
        if (h->split) CKS(h, dalloc(h, &h->qkv32, BR * 3 * H));
        else CKS(h, alloc_act(h, &h->qkv16, static_cast<int>(BR), 3 * H));
        CKS(h, dalloc(h, &h->xres, BR * H));
        CKS(h, dalloc(h, &h->refined, BR * H));
        h->refiner_allocated = true;
    }
    h->refiner_ready = true;
    return CAPDEC_OK;
}

template <int NKT, int DH, int G>
int launch_refine_att_mma_t(capdec_handle* h, int B, int R, const float* mask, cudaStream_t st) {
    using C = RefineMmaCfg<NKT, DH, G>;
    auto kern = refine_attention_mma_kernel<NKT, DH, G>;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), C::SMEM_BYTES));
    const int items = B * (h->NH / G);
    const int cap = h->num_sms * C::CTAS_PER_SM;
    kern<<<items < cap ? items : cap, 32 * C::WARPS, C::SMEM_BYTES, st>>>(h->qkv16.p, h->qkv16.ld, mask, B, R, h->H, h->NH, h->XR.p,
                                                                           h->XR.ld);
    return CAPDEC_OK;
}
// (query/key tiles, heads per CTA)
template <int DH>
int launch_refine_att_mma(capdec_handle* h, int B, int R, const float* mask, cudaStream_t st) {
    const int nh = h->NH;
    if (R <= 48) return launch_refine_att_mma_t<3, DH, 1>(h, B, R, mask, st);
    if (R <= 64 && nh % 2 == 0) return launch_refine_att_mma_t<4, DH, 2>(h, B, R, mask, st);
    if (R <= 112 && nh % 2 == 0) return launch_refine_att_mma_t<7, DH, 2>(h, B, R, mask, st);
    return launch_refine_att_mma_t<13, DH, 1>(h, B, R, mask, st);
}

// self-attention of one refiner layer: qkv -> XR[:, 0:H]
int launch_refine_att(capdec_handle* h, int B, int R, const float* mask, cudaStream_t st) {
    const int H = h->H, nh = h->NH, d = H / nh;
    prof_begin(h, CAPDEC_CAT_ATTENTION, 0.0, st);
    int status = CAPDEC_OK;
    if (!h->split && R <= 208 && (d == 128 || d == 64) && h->att_variant != 1) {
        status = d == 128 ? launch_refine_att_mma<128>(h, B, R, mask, st) : launch_refine_att_mma<64>(h, B, R, mask, st);
    } else {
        const size_t smem = (static_cast<size_t>(R) * (2 * d + 1) + 4 * (d + R)) * sizeof(float);
        if (smem > 227 * 1024) return fail(h, CAPDEC_ERR_INVALID, "refiner attention tile does not fit shared memory");
        if (h->split) {
            CK(h, smem_attr(reinterpret_cast<const void*>(refine_attention_kernel<float>), 227 * 1024));
            refine_attention_kernel<float><<<B * nh, 128, smem, st>>>(h->qkv32, 3 * H, mask, R, H, nh, h->XR.p, h->XR.ld, h->XR.lo);
        } else {
            CK(h, smem_attr(reinterpret_cast<const void*>(refine_attention_kernel<__half>), 227 * 1024));
            refine_attention_kernel<__half><<<B * nh, 128, smem, st>>>(h->qkv16.p, h->qkv16.ld, mask, R, H, nh, h->XR.p, h->XR.ld, h->XR.lo);
        }
    }
    prof_end(h, st);
    CKS(h, status);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// bu_feats [B,R,D] fp32 (+ prefix mask [B,R]) -> h->refined [B*R, H] fp32
int run_refiner(capdec_handle* h, const float* bu, const float* mask, int B, int R, cudaStream_t st, const __half* bu_f16 = nullptr) {
    const int H = h->H, D = h->D;
    const size_t BR = static_cast<size_t>(B) * R;
    const int N = static_cast<int>(BR);
    Operand ma, mb;
    if (bu_f16) {  // packed fp16 shard rows are the GEMM operand as they are
        CK(h, cudaMemcpyAsync(h->bu16.p, bu_f16, BR * D * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    } else {
        cvt_f16_kernel<<<grid_for(BR * D / 4), 256, 0, st>>>(bu, BR, D, h->bu16.p, h->bu16.ld, h->bu16.lo, 0);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    {  // x = relu(W_p bu + b_p), zeros in the padded rows (pack_wrapper, AoA_Model.py:650-655)
        CKS(h, map_a(h, &ma, h->bu16));
        CKS(h, map_b(h, &mb, h->W_proj));
        EpiParams e{};
        e.bias = h->b_proj;
        e.out32 = h->xres;
        e.ld32 = H;
        e.relu = 1;
        e.row_keep = mask;
        CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->bu16.lo, mb, h->W_proj.lo, N, H, D, e, st));
    }
    for (auto& L : h->refine) {
        // n = LN(x) -> XR[:, H:2H]  (SublayerConnection: norm first, AoA_Model.py:37)
        CKS(h, launch_layernorm(h, h->xres, N, L.ln_gain, L.ln_bias, h->XR.p + H, h->XR.ld, h->XR.lo, nullptr, st));
        {  // Q | K | V = n W^T + b  (AoA_Model.py:113-115), one GEMM
            CKS(h, map_a(h, &ma, h->XR, H));
            CKS(h, map_b(h, &mb, L.W_qkv));
            EpiParams e{};
            e.bias = L.b_qkv;
            if (h->split) {
                e.out32 = h->qkv32;
                e.ld32 = 3 * H;
            } else {
                e.out16 = h->qkv16.p;
                e.ld16 = h->qkv16.ld;
            }
            CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->XR.lo, mb, L.W_qkv.lo, N, 3 * H, H, e, st));
        }
        CKS(h, launch_refine_att(h, B, R, mask, st));
        {  // x += GLU(W [att, n] + b)  (AoA_Model.py:118 + residual :38)
            CKS(h, map_a(h, &ma, h->XR));
            CKS(h, map_b(h, &mb, L.W_glu));
            EpiParams e{};
            e.bias = L.b_glu;
            e.out32 = h->xres;
            e.ld32 = H;
            e.resid = h->xres;
            e.ld_resid = H;
            CKS(h, launch_gemm(h, EPI_GLU, 1, ma, h->XR.lo, mb, L.W_glu.lo, N, 2 * H, 2 * H, e, st));
        }
    }
    return launch_layernorm(h, h->xres, N, h->rfinal_gain, h->rfinal_bias, nullptr, 0, 0, h->refined, st);
}

// ------------------------------------------------------------------------------------------------ per-arch steps
struct StepCtx {
    int M = 0;         // rows = B * rows_per_image
    int K = 0;         // rows per image
    int t = 0;         // 1-based step
    int cur = 0;       // cell-state ping-pong index (read cur, write cur^1)
    int logits_epi = EPI_TOPK;
    int ktop = 4;
    uint32_t seed = 0;
    const uint32_t* seed_ptr = nullptr;  // captured rollouts read the seed from device memory
    int use_noise = 0;
    int scst_n = 0;  // > 0: rows = scst_n sampled + 1 greedy rollout per image
    bool first_from_c0 = false;
    const int* forced = nullptr;  // teacher-forced words [row * forced_ld + step] (capdec_score) or null
    int forced_ld = 0;
    float* states = nullptr;  // capdec_score_states: this step's slice of the [M, T, H] predict-input export (row stride states_ld) or null
    size_t states_ld = 0;
    int* att_done = nullptr;  // small-batch overlap: counter the attention kernel adds its finished images to, or null
    const int* beam_ctr = nullptr;  // small-batch overlap: the previous step's bookkeeping kernel runs beside this step's first launch,
    int beam_target = 0;            //   whose operand loads wait until *beam_ctr >= beam_target
    float* alphas = nullptr;  // where this step's attention maps go ([row * alpha_stride + region]) or null
    size_t alpha_stride = 0;
};

EpiParams logits_epi(capdec_handle* h, const StepCtx& c) {
    EpiParams e{};
    e.bias = h->b_pred;
    e.part = h->part;
    e.n_tiles = logit_slots(h, c.M, h->V, c.logits_epi);
    e.seed = c.seed;
    e.seed_ptr = c.seed_ptr;
    e.step = c.t - 1;
    e.use_noise = c.use_noise;
    e.scst_n = c.scst_n;
    e.forced = c.forced;
    e.forced_ld = c.forced_ld;
    return e;
}

// Gate GEMM (LSTM epilogue) -> bias/store GEMM on the h' it wrote, as ONE launch of the chained pair kernel (gemm.cuh).
int launch_chain(capdec_handle* h, const SmallDesc& d1, const SmallDesc& d2, cudaStream_t st) {
    auto fill = [&](const SmallDesc& d, GemmParams& p) {
        p = GemmParams{};
        p.M = d.M, p.N = d.N;
        p.k_blocks = d.Kdim / BLOCK_K;
        p.passes = h->split ? 3 : 1;
        p.a_lo_off = d.x_lo, p.b_lo_off = d.w_lo;
        p.num_m_blocks = (d.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
        p.num_n_blocks = (d.N + BN - 1) / BN;
        p.runs = 0;
        p.epi = d.e;
    };
    GemmParams p1, p2;
    fill(d1, p1);
    fill(d2, p2);
    const int items1 = p1.num_m_blocks * p1.num_n_blocks, items2 = p2.num_m_blocks * p2.num_n_blocks;
    const int most = items1 > items2 ? items1 : items2;
    const int pairs = most < h->num_sms / 2 ? most : h->num_sms / 2;
    p1.m_group = pairs, p2.m_group = pairs;
    ChainSync cs{h->chain_sync, h->chain_sync + h->chain_blocks, EPI_WARPS * 2 * p1.num_n_blocks, 2 * p2.num_n_blocks};
    auto kern = gemm2_chain_kernel;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), GemmCfg2::SMEM_BYTES));
    CK(h, launch_pdl(h, kern, dim3(2 * pairs), dim3(GEMM_THREADS), GemmCfg2::SMEM_BYTES, st, d1.x->map, d1.w->map, p1, d2.x->map, d2.w->map, p2, cs));
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// Run dependent GEMMs on the same activation rows: ONE persistent launch with grid barriers between them on the
// small-batch path (smallm.cuh), one launch each otherwise (or while per-launch timing is on).
int run_gemms(capdec_handle* h, const SmallDesc* d, int n, int M, cudaStream_t st, int grid_limit = 0) {
    if (n > 1 && small_ok(h, M) && h->small_fuse && !h->prof) return launch_small(h, d, n, st, grid_limit);
    if (n == 2 && !small_ok(h, M) && h->chain && h->pair_gemm && !h->prof && d[0].epi == EPI_LSTM && d[1].epi == EPI_STORE &&
        d[0].M == d[1].M && (d[0].M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) <= h->chain_blocks)
        return launch_chain(h, d[0], d[1], st);
    for (int q = 0; q < n; ++q)
        CKS(h, launch_gemm(h, d[q].epi, d[q].ktop, *d[q].x, d[q].x_lo, *d[q].w, d[q].w_lo, d[q].M, d[q].N, d[q].Kdim, d[q].e, st));
    return CAPDEC_OK;
}

template <int KR, typename T>
int launch_butd_att_t(capdec_handle* h, const StepCtx& c, const T* enc, int enc_ld, const T* feats, int feats_ld, cudaStream_t st) {
    auto kern = butd_attention_kernel<KR, T>;
    const size_t smem = (static_cast<size_t>(KR) * h->R + 2 * 8 * AttCfg<KR>::NP) * sizeof(float);
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), 160 * 1024));
    if (smem > 160 * 1024) return fail(h, CAPDEC_ERR_INVALID, "attention tile does not fit shared memory");
    prof_begin(h, CAPDEC_CAT_ATTENTION, 0.0, st);
    kern<<<h->B, 256, smem, st>>>(enc, enc_ld, feats, feats_ld, h->dec_ctx, h->w_aff, h->b_aff, h->R, h->A, h->D, c.K, h->XB.p, h->XB.ld,
                                  h->XB.lo, c.alphas, c.alpha_stride);
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

template <int KR, typename T>
int launch_butd_att_stream_t(capdec_handle* h, const StepCtx& c, const T* enc, const T* feats, cudaStream_t st) {
    using C = AttStreamCfg<KR, T>;
    auto kern = butd_attention_stream_kernel<KR, T>;
    const size_t smem = static_cast<size_t>(C::STAGES) * C::STAGE_BYTES + 2 * C::STAGES * 8 +
                        (static_cast<size_t>(KR) * h->R + 2 * 8 * C::NP) * sizeof(float) + 128;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), 200 * 1024));
    if (smem > 200 * 1024) return fail(h, CAPDEC_ERR_INVALID, "attention ring does not fit shared memory");
    const int per_sm = C::CTAS_PER_SM;
    const int grid = h->B < h->num_sms * per_sm ? h->B : h->num_sms * per_sm;
    prof_begin(h, CAPDEC_CAT_ATTENTION, 0.0, st);
    kern<<<grid, C::THREADS, smem, st>>>(enc, feats, h->dec_ctx, h->w_aff, h->b_aff, h->B, h->R, h->A, h->D, c.K, h->XB.p, h->XB.ld,
                                         h->XB.lo, c.alphas, c.alpha_stride);
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

template <int KR, int CTAS, bool FULL>
int launch_butd_att_mma_t(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    using C = AttMmaCfg<KR, CTAS>;
    auto kern = butd_attention_mma_kernel<KR, CTAS, FULL>;
    const size_t smem = att_mma_smem_bytes(KR, C::STAGES, h->R, h->A, h->enc16.ld, h->feats16.ld);
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), 227 * 1024));
    if (smem > 227 * 1024) return fail(h, CAPDEC_ERR_INVALID, "attention ring does not fit shared memory");
    const int cap = h->num_sms * C::CTAS_PER_SM;
    const int grid = h->B < cap ? h->B : cap;
    prof_begin(h, CAPDEC_CAT_ATTENTION, 0.0, st);
    CK(h, launch_pdl(h, kern, dim3(grid), dim3(C::THREADS), smem, st, h->enc16.p, h->enc16.ld, h->feats16.p, h->feats16.ld,
                     static_cast<size_t>(h->B) * h->R, h->dec_ctx, h->w_aff, h->b_aff, h->B, h->R, h->A, h->D, c.K, h->XB.p, h->XB.ld,
                     c.alphas, c.alpha_stride, c.att_done));
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}
template <int KR, int CTAS>
int launch_butd_att_mma_d(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    if (h->A == 1024 && h->D == 2048) return launch_butd_att_mma_t<KR, CTAS, true>(h, c, st);
    return launch_butd_att_mma_t<KR, CTAS, false>(h, c, st);
}
template <int KR>
int launch_butd_att_mma(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    // two CTAs per SM (3-stage rings) when the per-CTA state is small enough, else one CTA with a deep ring
    if (KR <= 3 && h->att_variant != 2 &&
        2 * (att_mma_smem_bytes(KR, AttMmaCfg<KR, 2>::STAGES, h->R, h->A, h->enc16.ld, h->feats16.ld) + 1024) <= 228 * 1024)
        return launch_butd_att_mma_d<KR, 2>(h, c, st);
    return launch_butd_att_mma_d<KR, 1>(h, c, st);
}

// fp16 mode reads the fp16 copies (projected features written by the projection GEMM, raw features converted for
// it); the fp32-grade mode reads the caller's fp32 features and the fp32 projection.
template <int KR>
int launch_butd_att(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    // streaming (persistent, bulk-copy fed) kernel when one region row of A / D columns fits the ring layout
    const bool stream_ok = h->A <= 1024 && h->D <= 2048 && !h->no_stream_attention;
    if (h->split) {
        if (stream_ok) return launch_butd_att_stream_t<KR, float>(h, c, h->enc_ctx, h->feats, st);
        return launch_butd_att_t<KR, float>(h, c, h->enc_ctx, h->A, h->feats, h->D, st);
    }
    if (stream_ok && h->A % 16 == 0 && h->D % 32 == 0 && h->att_variant != 1) return launch_butd_att_mma<KR>(h, c, st);
    return launch_butd_att_t<KR, __half>(h, c, h->enc16.p, h->enc16.ld, h->feats16.p, h->feats16.ld, st);
}

bool aoa_mma_ok(const capdec_handle* h) {
    const int d = h->NH > 0 ? h->H / h->NH : 0;
    return !h->split && h->NH <= 8 && d % 16 == 0 && d <= 256 && h->att_variant != 1;
}

template <int KR>
int launch_aoa_att_mma(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    auto kern = aoa_attention_mma_kernel<KR>;
    const size_t fixed = aoa_mma_fixed_smem(KR, h->R, h->H, h->NH);
    const size_t stage_bytes = static_cast<size_t>(16) * h->k16.ld * 2;
    int stages = static_cast<int>((226 * 1024 - fixed) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return fail(h, CAPDEC_ERR_INVALID, "AoA attention ring does not fit shared memory");
    const size_t smem = fixed + stages * stage_bytes;
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), 227 * 1024));
    const int grid = h->B < h->num_sms ? h->B : h->num_sms;
    prof_begin(h, CAPDEC_CAT_ATTENTION, 0.0, st);
    CK(h, launch_pdl(h, kern, dim3(grid), dim3(288), smem, st, h->k16.p, h->v16.p, h->k16.ld, static_cast<size_t>(h->B) * h->R,
                     h->q16.p, h->q16.ld, h->mask, h->B, h->R, h->H, h->NH, c.K, stages, h->XB.p, h->XB.ld, c.alphas, c.alpha_stride));
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

template <int KR>
int launch_aoa_att(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    if (aoa_mma_ok(h)) return launch_aoa_att_mma<KR>(h, c, st);
    auto kern = aoa_attention_kernel<KR>;
    const size_t smem = (static_cast<size_t>(KR) * h->H + static_cast<size_t>(KR) * h->NH * h->R) * sizeof(float);
    CK(h, smem_attr(reinterpret_cast<const void*>(kern), 160 * 1024));
    if (smem > 160 * 1024) return fail(h, CAPDEC_ERR_INVALID, "attention tile does not fit shared memory");
    prof_begin(h, CAPDEC_CAT_ATTENTION, 0.0, st);
    kern<<<h->B, 256, smem, st>>>(h->q32, h->kv32, h->mask, h->R, h->H, h->NH, c.K, h->XB.p, h->XB.ld, h->XB.lo, c.alphas,
                                  c.alpha_stride);
    prof_end(h, st);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

// BUTD: XA = [h2 | h1] (top-down LSTM operand), XB = [ctx | h1 | h2] (language LSTM operand), Hb2 = new h2.
int step_butd(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    const int H = h->H, A = h->A, D = h->D;
    Operand x_td, w_td, x_da, w_da, x_lm, w_lm, x_pr, w_pr;
    SmallDesc g[2];
    {  // top-down attention LSTM (BUTD_Model.py:265)
        CKS(h, map_a(h, &x_td, h->XA));
        CKS(h, map_b(h, &w_td, h->W_l1));
        EpiParams e{};
        e.rowadd = h->G0;
        e.rowadd_ld = 4 * H;
        e.rows_per_group = c.K;
        e.gather = h->emb_gates;
        e.gather_idx = h->tok;
        e.gather_ld = 4 * H;
        e.c_in = h->c1[c.cur];
        e.c_out = h->c1[c.cur ^ 1];
        e.ldc = H;
        e.parent = h->parent;
        e.out16 = h->XB.p + D;
        e.ld16 = h->XB.ld;
        e.lo16 = h->XB.lo;
        g[0] = SmallDesc{EPI_LSTM, 1, &x_td, h->XA.lo, &w_td, h->W_l1.lo, c.M, 4 * H, H + H, e};
        if (c.beam_ctr) {  // every operand row comes from the bookkeeping kernel running beside this launch
            g[0].wait_ctr = c.beam_ctr;
            g[0].wait_target = c.beam_target;
            g[0].wait_kb = (H + H) / BLOCK_K;
        }
    }
    {  // dec_att(h1) (BUTD_Model.py:58)
        CKS(h, map_a(h, &x_da, h->XB, D));
        CKS(h, map_b(h, &w_da, h->W_aux2));
        EpiParams e{};
        e.bias = h->b_aux2;
        e.out32 = h->dec_ctx;
        e.ld32 = A;
        g[1] = SmallDesc{EPI_STORE, 1, &x_da, h->XB.lo, &w_da, h->W_aux2.lo, c.M, A, H, e};
    }
    CKS(h, run_gemms(h, g, 2, c.M, st));
    // Small-batch path: the attention kernel runs on a side stream CONCURRENTLY with the [language gates -> logits] launch.
    // The language LSTM's K range is [ctx | h1 | h2]: its weight tiles and the h1 / h2 splits do not depend on the attention,
    // so that launch (shrunk by the SMs the attention's one-CTA-per-image grid may occupy) starts right away and only the
    // activation loads of its ctx k-blocks wait for the attention's counter.
    const bool overlap = h->small_overlap && h->side && small_ok(h, c.M) && h->small_fuse && !h->prof && !h->split && !c.alphas &&
                         !h->no_stream_attention && h->att_variant != 1 && h->A <= 1024 && h->D <= 2048 && h->A % 16 == 0 && h->D % 32 == 0 &&
                         h->num_sms - h->B >= 4 * ((4 * h->H + SM_TILE_N - 1) / SM_TILE_N);  // the shrunk launch still holds the gate GEMM's 4-way split in one round (B <= 20)
    StepCtx ca = c;
    cudaStream_t att_st = st;
    if (overlap) {
        ca.att_done = h->att_done;
        att_st = h->side;
        CK(h, cudaEventRecord(h->ev_fork, st));
        CK(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    }
    if (c.K <= 1) CKS(h, launch_butd_att<1>(h, ca, att_st));
    else if (c.K <= 3) CKS(h, launch_butd_att<3>(h, ca, att_st));
    else if (c.K <= 5) CKS(h, launch_butd_att<5>(h, ca, att_st));
    else CKS(h, launch_butd_att<8>(h, ca, att_st));
    if (overlap) CK(h, cudaEventRecord(h->ev_join, h->side));
    {  // language LSTM (BUTD_Model.py:268)
        CKS(h, map_a(h, &x_lm, h->XB));
        CKS(h, map_b(h, &w_lm, h->W_l2));
        EpiParams e{};
        e.bias = h->b_l2;
        e.c_in = h->c2[c.cur];
        e.c_out = h->c2[c.cur ^ 1];
        e.ldc = H;
        e.parent = h->parent;
        e.out16 = h->Hb2.p;
        e.ld16 = h->Hb2.ld;
        e.lo16 = h->Hb2.lo;
        g[0] = SmallDesc{EPI_LSTM, 1, &x_lm, h->XB.lo, &w_lm, h->W_l2.lo, c.M, 4 * H, D + H + H, e};
        if (overlap) {  // the ctx columns (k-blocks below D / 64) arrive from the concurrent attention kernel
            g[0].wait_ctr = h->att_done;
            g[0].wait_target = h->B * c.t;  // cumulative over the decode: reset_state zeroes the counter
            g[0].wait_kb = D / BLOCK_K;
        }
    }
    CKS(h, map_a(h, &x_pr, h->Hb2));
    CKS(h, map_b(h, &w_pr, h->W_pred));
    g[1] = SmallDesc{c.logits_epi, c.ktop, &x_pr, h->Hb2.lo, &w_pr, h->W_pred.lo, c.M, h->V, H, logits_epi(h, c)};
    CKS(h, run_gemms(h, g, 2, c.M, st, overlap ? h->num_sms - h->B : 0));
    if (overlap) CK(h, cudaStreamWaitEvent(st, h->ev_join, 0));  // the side stream joins: later work is ordered after the attention too
    return CAPDEC_OK;
}

AdvOp op_copy(const Act16& src, int src_col, const Act16& dst, int dst_col, int n) {
    AdvOp o{};
    o.kind = ADV_COPY16;
    o.src = src.p + src_col;
    o.src_ld = src.ld;
    o.src_lo = src.lo;
    o.dst = dst.p + dst_col;
    o.dst_ld = dst.ld;
    o.dst_lo = dst.lo;
    o.n = n;
    return o;
}
AdvOps adv_butd(capdec_handle* h, bool init) {
    const int H = h->H, D = h->D;
    AdvOps a{};
    if (!init) {
        a.op[a.n++] = op_copy(h->Hb2, 0, h->XA, 0, H);      // h2 -> top-down operand
        a.op[a.n++] = op_copy(h->Hb2, 0, h->XB, D + H, H);  // h2 -> language operand (recurrent part)
        a.op[a.n++] = op_copy(h->XB, D, h->XA, H, H);       // h1 -> top-down operand (recurrent part)
    }
    return a;
}

// NIC: XA = [h], Hb = new h.
int step_nic(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    const int H = h->H;
    Operand x_l, w_l, x_pr, w_pr;
    SmallDesc g[2];
    CKS(h, map_a(h, &x_l, h->XA));
    CKS(h, map_b(h, &w_l, h->W_l1));
    EpiParams e{};
    e.bias = h->b_l1;
    e.gather = h->emb_gates;
    e.gather_idx = h->tok;
    e.gather_ld = 4 * H;
    e.c_in = c.first_from_c0 ? h->c0 : h->c1[c.cur];
    e.c_out = h->c1[c.cur ^ 1];
    e.ldc = H;
    e.parent = h->parent;
    e.out16 = h->Hb.p;
    e.ld16 = h->Hb.ld;
    e.lo16 = h->Hb.lo;
    g[0] = SmallDesc{EPI_LSTM, 1, &x_l, h->XA.lo, &w_l, h->W_l1.lo, c.M, 4 * H, H, e};
    CKS(h, map_a(h, &x_pr, h->Hb));
    CKS(h, map_b(h, &w_pr, h->W_pred));
    g[1] = SmallDesc{c.logits_epi, c.ktop, &x_pr, h->Hb.lo, &w_pr, h->W_pred.lo, c.M, h->V, H, logits_epi(h, c)};
    return run_gemms(h, g, 2, c.M, st);
}

AdvOps adv_nic(capdec_handle* h, bool init) {
    const int H = h->H;
    AdvOps a{};
    if (init) {
        AdvOp o = op_copy(h->H0, 0, h->XA, 0, H);  // primed hidden state of the image (NIC_Model.py:164,170)
        o.kind = ADV_BCAST16;
        a.op[a.n++] = o;
    } else {
        a.op[a.n++] = op_copy(h->Hb, 0, h->XA, 0, H);
    }
    return a;
}

// AoA: XA = [mean+ctx | h] (LSTM operand), XB = [att | query] (AoA gate operand), Hb = new h, Hb2 = ctx fp16.
int step_aoa(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    const int H = h->H;
    Operand ma, mb;
    SmallDesc g[2];
    {
        CKS(h, map_a(h, &ma, h->XA));
        CKS(h, map_b(h, &mb, h->W_l1));
        EpiParams e{};
        e.bias = h->b_l1;
        e.gather = h->emb_gates;
        e.gather_idx = h->tok;
        e.gather_ld = 4 * H;
        e.c_in = h->c1[c.cur];
        e.c_out = h->c1[c.cur ^ 1];
        e.ldc = H;
        e.parent = h->parent;
        e.out16 = h->Hb.p;
        e.ld16 = h->Hb.ld;
        e.lo16 = h->Hb.lo;
        e.h32 = h->h32;
        e.ldh32 = H;
        CKS(h, launch_gemm(h, EPI_LSTM, 1, ma, h->XA.lo, mb, h->W_l1.lo, c.M, 4 * H, H + H, e, st));
    }
    CKS(h, launch_layernorm(h, h->h32, c.M, h->ln_gain, h->ln_bias, h->XB.p + H, h->XB.ld, h->XB.lo, nullptr, st));
    {  // linear_Q (AoA_Model.py:113)
        CKS(h, map_a(h, &ma, h->XB, H));
        CKS(h, map_b(h, &mb, h->W_aux1));
        EpiParams e{};
        e.bias = h->b_aux1;
        if (aoa_mma_ok(h)) {
            e.out16 = h->q16.p;
            e.ld16 = h->q16.ld;
        } else {
            e.out32 = h->q32;
            e.ld32 = H;
        }
        CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->XB.lo, mb, h->W_aux1.lo, c.M, H, H, e, st));
    }
    if (c.K <= 1) CKS(h, launch_aoa_att<1>(h, c, st));
    else if (c.K <= 3) CKS(h, launch_aoa_att<3>(h, c, st));
    else if (c.K <= 5) CKS(h, launch_aoa_att<5>(h, c, st));
    else CKS(h, launch_aoa_att<8>(h, c, st));
    {  // AoA gate: GLU(Linear(cat[att, query])) (AoA_Model.py:118)
        CKS(h, map_a(h, &ma, h->XB));
        CKS(h, map_b(h, &mb, h->W_aux3));
        EpiParams e{};
        e.bias = h->b_aux3;
        e.out32 = h->ctx32;
        e.ld32 = H;
        e.out16 = h->Hb2.p;
        e.ld16 = h->Hb2.ld;
        e.lo16 = h->Hb2.lo;
        g[0] = SmallDesc{EPI_GLU, 1, &ma, h->XB.lo, &mb, h->W_aux3.lo, c.M, 2 * H, 2 * H, e};
    }
    Operand x_pr, w_pr;
    CKS(h, map_a(h, &x_pr, h->Hb2));
    CKS(h, map_b(h, &w_pr, h->W_pred));
    g[1] = SmallDesc{c.logits_epi, c.ktop, &x_pr, h->Hb2.lo, &w_pr, h->W_pred.lo, c.M, h->V, H, logits_epi(h, c)};
    return run_gemms(h, g, 2, c.M, st);
}

AdvOps adv_aoa(capdec_handle* h, bool init) {
    const int H = h->H;
    AdvOps a{};
    AdvOp m{};
    m.kind = ADV_MEAN_PLUS;
    m.src = init ? nullptr : h->ctx32;
    m.src_ld = H;
    m.aux = h->mean32;
    m.dst = h->XA.p;
    m.dst_ld = h->XA.ld;
    m.dst_lo = h->XA.lo;
    m.n = H;
    a.op[a.n++] = m;
    if (!init) a.op[a.n++] = op_copy(h->Hb, 0, h->XA, H, H);
    return a;
}

int run_step(capdec_handle* h, const StepCtx& c, cudaStream_t st) {
    int status;
    switch (h->cfg.arch) {
        case CAPDEC_ARCH_BUTD: status = step_butd(h, c, st); break;
        case CAPDEC_ARCH_NIC: status = step_nic(h, c, st); break;
        default: status = step_aoa(h, c, st); break;
    }
    CKS(h, status);
    if (c.states) {  // the rows `predict` was applied to in this step (BUTD_Model.py:270 h2, NIC_Model.py:174 h, AoA_Model.py:455 ctx)
        const Act16& x = h->cfg.arch == CAPDEC_ARCH_NIC ? h->Hb : h->Hb2;
        export_f32_kernel<<<grid_for(static_cast<size_t>(c.M) * h->H), 256, 0, st>>>(x.p, x.ld, x.lo, c.M, h->H, c.states, c.states_ld);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    return CAPDEC_OK;
}
AdvOps adv_ops(capdec_handle* h, bool init) {
    switch (h->cfg.arch) {
        case CAPDEC_ARCH_BUTD: return adv_butd(h, init);
        case CAPDEC_ARCH_NIC: return adv_nic(h, init);
        default: return adv_aoa(h, init);
    }
}

// zero the recurrent operand slots and the cell state before a decode (states start at 0:
// BUTD_Model.py:92-95,261-262; AoA_Model.py:223-227)
int reset_state(capdec_handle* h, int M, cudaStream_t st) {
    CK(h, cudaMemsetAsync(h->XA.p, 0, static_cast<size_t>(M) * h->XA.ld * sizeof(__half), st));
    if (h->XB.p) CK(h, cudaMemsetAsync(h->XB.p, 0, static_cast<size_t>(M) * h->XB.ld * sizeof(__half), st));
    CK(h, cudaMemsetAsync(h->c1[0], 0, static_cast<size_t>(M) * h->H * sizeof(float), st));
    if (h->c2[0]) CK(h, cudaMemsetAsync(h->c2[0], 0, static_cast<size_t>(M) * h->H * sizeof(float), st));
    if (h->att_done) CK(h, cudaMemsetAsync(h->att_done, 0, sizeof(int), st));
    if (h->beam_done) CK(h, cudaMemsetAsync(h->beam_done, 0, sizeof(int), st));
    return CAPDEC_OK;
}

// Launch the captured form of a decode (capturing it first when the cache has no entry for `key`): `enqueue` issues the
// whole kernel sequence on `st` with library-owned output buffers.
template <typename F>
int replay_graph(capdec_handle* h, const capdec_handle::GraphKey& key, cudaStream_t st, F&& enqueue) {
    capdec_handle::GraphEntry* hit = nullptr;
    for (auto& g : h->graphs)
        if (g.key == key) hit = &g;
    if (!hit) {
        cudaGraph_t graph = nullptr;
        CK(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int64_t l0 = h->launches;
        const int status = enqueue();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        const int64_t n = h->launches - l0;
        h->launches = l0;
        if (status != CAPDEC_OK) {
            if (graph) cudaGraphDestroy(graph);
            return status;
        }
        CK(h, ce);
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        CK(h, ie);
        if (h->graphs.size() >= capdec_handle::GRAPH_CACHE) {  // replace the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < h->graphs.size(); ++i)
                if (h->graphs[i].used < h->graphs[lru].used) lru = i;
            cudaGraphExecDestroy(h->graphs[lru].exec);
            h->graphs.erase(h->graphs.begin() + static_cast<long>(lru));
        }
        h->graphs.push_back({key, exec, n, 0});
        hit = &h->graphs.back();
        h->graph_captures++;
    }
    hit->used = ++h->graph_clock;
    CK(h, cudaGraphLaunch(hit->exec, st));
    h->launches += hit->launches;
    return CAPDEC_OK;
}

}  // namespace

struct capdec_cider {
    int device = 0;
    std::string err;
    uint64_t* keys = nullptr;  // device open-addressing table
    float* df = nullptr;
    uint64_t slots = 0;
    double log_ref_len = 0.0;
};

// ================================================================================================ C ABI
extern "C" {

// ------------------------------------------------------------------------------------------------ CIDEr-D reward
uint64_t capdec_cider_ngram_key(const int32_t* ids, int32_t k) { return (ids && k > 0) ? cider_ngram_key(ids, k) : 0; }

const char* capdec_cider_last_error(const capdec_cider* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int capdec_cider_create(int32_t device, capdec_cider** out) {
    if (!out) return CAPDEC_ERR_INVALID;
    *out = nullptr;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        g_create_error = "capdec_cider needs an sm_100a (B200) device; there is no fallback path";
        cudaGetLastError();
        return CAPDEC_ERR_CUDA;
    }
    capdec_cider* c = new capdec_cider();
    c->device = device;
    *out = c;
    return CAPDEC_OK;
}

void capdec_cider_destroy(capdec_cider* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->keys);
    cudaFree(c->df);
    delete c;
}

int capdec_cider_set_df(capdec_cider* c, const uint64_t* keys, const float* df, int64_t n, double log_ref_len) {
    if (!c || n < 0 || (n > 0 && (!keys || !df))) return CAPDEC_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    cudaFree(c->keys);
    cudaFree(c->df);
    c->keys = nullptr, c->df = nullptr, c->slots = 0;
    c->log_ref_len = log_ref_len;
    if (n == 0) return CAPDEC_OK;
    uint64_t slots = 16;
    while (slots < static_cast<uint64_t>(n) * 2) slots <<= 1;
    std::vector<uint64_t> hk(slots, 0);
    std::vector<float> hv(slots, 0.f);
    for (int64_t i = 0; i < n; ++i) {
        if (keys[i] == 0) {
            c->err = "cider_set_df: key 0 is reserved";
            return CAPDEC_ERR_INVALID;
        }
        uint64_t s = cider_mix64(keys[i]) & (slots - 1);
        while (hk[s] != 0 && hk[s] != keys[i]) s = (s + 1) & (slots - 1);
        hk[s] = keys[i];
        hv[s] = df[i];
    }
    CK(c, cudaMalloc(reinterpret_cast<void**>(&c->keys), slots * sizeof(uint64_t)));
    CK(c, cudaMalloc(reinterpret_cast<void**>(&c->df), slots * sizeof(float)));
    CK(c, cudaMemcpy(c->keys, hk.data(), slots * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->df, hv.data(), slots * sizeof(float), cudaMemcpyHostToDevice));
    c->slots = slots;
    return CAPDEC_OK;
}

int capdec_cider_reward(capdec_cider* c, const int32_t* gen, int32_t n_per_image, const int32_t* greedy, int32_t batch,
                        int32_t max_seq, const int32_t* ref_tokens, const int32_t* ref_lens, const int32_t* ref_offsets,
                        int32_t ref_ld, double sigma, double weight, float* rewards, float* scores, void* stream) {
    if (!c) return CAPDEC_ERR_INVALID;
    if (!gen || !greedy || !ref_tokens || !ref_lens || !ref_offsets || !rewards || batch <= 0 || max_seq <= 0 || ref_ld <= 0 ||
        n_per_image <= 0 || n_per_image > MAX_ROWS || sigma <= 0.0) {
        c->err = "cider_reward: null pointer or size out of range (1 <= n_per_image <= 8)";
        return CAPDEC_ERR_INVALID;
    }
    CK(c, cudaSetDevice(c->device));
    const size_t smem = static_cast<size_t>(n_per_image + 1 + CIDER_REF_SLOTS) * sizeof(CiderVec);
    CK(c, smem_attr(reinterpret_cast<const void*>(cider_reward_kernel), static_cast<int>((CIDER_MAX_HYPS + CIDER_REF_SLOTS) * sizeof(CiderVec))));
    CiderTable tab{c->keys, c->df, c->slots ? c->slots - 1 : 0, c->log_ref_len};
    cider_reward_kernel<<<batch, 32 * (n_per_image + 1), smem, static_cast<cudaStream_t>(stream)>>>(
        tab, gen, n_per_image, greedy, max_seq, ref_tokens, ref_lens, ref_offsets, ref_ld, sigma, weight, rewards, scores);
    CK(c, cudaGetLastError());
    return CAPDEC_OK;
}

int capdec_abi_version(void) { return CAPDEC_ABI_VERSION; }

const char* capdec_last_error(const capdec_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t capdec_launch_count(const capdec_handle* h) { return h ? h->launches : 0; }

int64_t capdec_graph_captures(const capdec_handle* h) { return h ? h->graph_captures : 0; }

int capdec_debug_trace(capdec_handle* h, uint64_t* dst, int64_t n) {
    if (!h || !dst || n <= 0) return CAPDEC_ERR_INVALID;
    if (!h->small_trace) return fail(h, CAPDEC_ERR_STATE, "debug trace is off (CAPDEC_TRACE=1 before capdec_create)");
    CK(h, cudaSetDevice(h->cfg.device));
    CK(h, cudaDeviceSynchronize());
    const int64_t m = n < 1 + 16 * 4000 ? n : 1 + 16 * 4000;
    CK(h, cudaMemcpy(dst, h->small_trace, static_cast<size_t>(m) * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CK(h, cudaMemset(h->small_trace, 0, sizeof(uint64_t)));
    return CAPDEC_OK;
}

void capdec_destroy(capdec_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_fork2) cudaEventDestroy(h->ev_fork2);
    if (h->ev_join2) cudaEventDestroy(h->ev_join2);
    if (h->side) cudaStreamDestroy(h->side);
    for (void* p : h->allocs) cudaFree(p);
    for (auto& kv : h->raw) cudaFree(kv.second.d);
    delete h;
}

static int create_impl(capdec_handle* h) {
    const capdec_config& c = h->cfg;
    if (c.arch < 0 || c.arch > 2) return fail(h, CAPDEC_ERR_INVALID, "unknown arch");
    h->H = c.hidden_dim, h->E = c.embed_dim, h->A = c.atten_dim, h->D = c.enc_dim, h->V = c.vocab_size, h->NH = c.num_heads;
    h->Bmax = c.max_batch, h->Rmax = c.max_regions, h->Kmax = c.max_rows, h->Tmax = c.max_seq;
    h->split = c.math_mode == CAPDEC_MATH_F16X3;
    if (c.math_mode != CAPDEC_MATH_F16 && c.math_mode != CAPDEC_MATH_F16X3) return fail(h, CAPDEC_ERR_INVALID, "unknown math_mode");
    const int H = h->H, E = h->E, V = h->V;
    if (H <= 0 || H % 64 || E <= 0 || E % 64) return fail(h, CAPDEC_ERR_INVALID, "hidden_dim and embed_dim must be multiples of 64");
    if (V < 4) return fail(h, CAPDEC_ERR_INVALID, "vocab_size must be >= 4");
    if (h->Bmax <= 0 || h->Tmax <= 0 || h->Kmax <= 0 || h->Kmax > MAX_ROWS)
        return fail(h, CAPDEC_ERR_INVALID, "max_batch/max_seq must be > 0 and 1 <= max_rows <= 8");
    if (c.arch == CAPDEC_ARCH_BUTD && (h->A <= 0 || h->A % 64 || h->D <= 0 || h->D % 64 || h->Rmax <= 0))
        return fail(h, CAPDEC_ERR_INVALID, "BUTD needs atten_dim, enc_dim multiples of 64 and max_regions > 0");
    if (c.arch == CAPDEC_ARCH_AOA) {
        if (h->NH <= 0 || H % h->NH || h->Rmax <= 0) return fail(h, CAPDEC_ERR_INVALID, "AoA needs hidden_dim % num_heads == 0");
        const int d = H / h->NH;
        if (d % 4 || ((d / 4) & (d / 4 - 1))) return fail(h, CAPDEC_ERR_INVALID, "AoA head dim / 4 must be a power of two");
    }
    CK(h, cudaSetDevice(c.device));
    cudaDeviceProp prop;
    CK(h, cudaGetDeviceProperties(&prop, c.device));
    if (prop.major != 10) return fail(h, CAPDEC_ERR_CUDA, "libcapdec needs an sm_100a (B200) device; there is no fallback path");
    h->num_sms = prop.multiProcessorCount;
    {
        const char* e = getenv("CAPDEC_NO_STREAM_ATTENTION");
        h->no_stream_attention = e && e[0] == '1';
        const char* g1 = getenv("CAPDEC_GEMM_1CTA");
        h->pair_gemm = !(g1 && g1[0] == '1');
        const char* mg = getenv("CAPDEC_MGROUP_SPLIT");
        h->mgroup_split = !(mg && mg[0] == '0');
        const char* ew = getenv("CAPDEC_LOGIT_EW");
        if (ew && atoi(ew) == EPI_WARPS) h->logit_ew = EPI_WARPS;
        const char* ng = getenv("CAPDEC_NO_GRAPH");
        h->use_graphs = !(ng && ng[0] == '1');
        const char* np = getenv("CAPDEC_PDL");
        h->use_pdl = np && np[0] == '1';
        const char* v = getenv("CAPDEC_ATT_VARIANT");
        h->att_variant = v ? atoi(v) : 0;
    }
    h->Mmax = h->Bmax * h->Kmax;
    h->n_tiles_v = ((V + BN - 1) / BN) * LOGIT_EPI_SPLIT_MAX;  // partial slots per row: (N tile, column share)
    const int M = h->Mmax;

    CKS(h, dalloc(h, &h->scale_tmp, static_cast<size_t>(V > 4 * H ? V : 4 * H)));
    CKS(h, alloc_act(h, &h->W_pred, V, H));
    CKS(h, alloc_act(h, &h->W_emb, 4 * H, E));
    CKS(h, alloc_act(h, &h->emb16, V, E));
    CKS(h, dalloc(h, &h->emb_gates, static_cast<size_t>(V) * 4 * H));
    CKS(h, dalloc(h, &h->b_pred, V));
    CKS(h, dalloc(h, &h->b_l1, 4 * H));
    const size_t part_stride = topk_part_stride(8) > SAMPLE_PART_STRIDE ? topk_part_stride(8) : SAMPLE_PART_STRIDE;
    CKS(h, dalloc(h, &h->part, static_cast<size_t>(M) * h->n_tiles_v * part_stride));
    CKS(h, dalloc(h, &h->c1[0], static_cast<size_t>(M) * H));
    CKS(h, dalloc(h, &h->c1[1], static_cast<size_t>(M) * H));
    CKS(h, dalloc(h, &h->tok, M));
    CKS(h, dalloc(h, &h->parent, M));
    CKS(h, dalloc(h, &h->cum, M));
    CKS(h, dalloc(h, &h->unfinished, M));
    CKS(h, dalloc(h, &h->live_count, h->Tmax + 1));
    CKS(h, dalloc(h, &h->n_live, h->Bmax));
    CKS(h, dalloc(h, &h->best_score, h->Bmax));
    CKS(h, dalloc(h, &h->best_len, h->Bmax));
    CKS(h, dalloc(h, &h->best_seq, static_cast<size_t>(h->Bmax) * (h->Tmax + 1)));
    CKS(h, dalloc(h, &h->seqs[0], static_cast<size_t>(M) * (h->Tmax + 1)));
    CKS(h, dalloc(h, &h->seqs[1], static_cast<size_t>(M) * (h->Tmax + 1)));
    CKS(h, dalloc(h, &h->out_tokens, static_cast<size_t>(h->Bmax) * (h->Tmax + 1)));
    if (c.arch == CAPDEC_ARCH_AOA) CKS(h, dalloc(h, &h->mask_buf, static_cast<size_t>(h->Bmax) * h->Rmax));
    CKS(h, dalloc(h, &h->out_scores, h->Bmax));
    CKS(h, dalloc(h, &h->out_lengths, h->Bmax));
    CKS(h, dalloc(h, &h->seed_dev, 1));
    CKS(h, alloc_small(h));
    if (h->small_slabs && c.arch == CAPDEC_ARCH_BUTD) {
        const char* no = getenv("CAPDEC_NO_OVERLAP");
        h->small_overlap = !(no && no[0] == '1');
        CKS(h, dalloc(h, &h->att_done, 1));
        CKS(h, dalloc(h, &h->beam_done, 1));
        CK(h, cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
        CK(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        CK(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    {
        const char* nc = getenv("CAPDEC_CHAIN");
        h->chain = nc && nc[0] == '1';
        h->chain_blocks = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) + 1;
        CKS(h, dalloc(h, &h->chain_sync, 2 * static_cast<size_t>(h->chain_blocks)));
    }
    CKS(h, dalloc(h, &h->out_sample_tokens, static_cast<size_t>(M) * h->Tmax));
    CKS(h, dalloc(h, &h->out_sample_logprobs, static_cast<size_t>(M) * h->Tmax));
    CKS(h, dalloc(h, &h->out_greedy, static_cast<size_t>(h->Bmax) * h->Tmax));
    CKS(h, dalloc(h, &h->forced_buf, static_cast<size_t>(M) * h->Tmax));

    if (c.arch == CAPDEC_ARCH_BUTD) {
        const int A = h->A, D = h->D;
        const size_t BR = static_cast<size_t>(h->Bmax) * h->Rmax;
        CKS(h, alloc_act(h, &h->W_l1, 4 * H, H + H));
        CKS(h, alloc_act(h, &h->W_l2, 4 * H, D + H + H));
        CKS(h, alloc_act(h, &h->W_aux1, A, D));
        CKS(h, alloc_act(h, &h->W_aux2, A, H));
        CKS(h, alloc_act(h, &h->W_aux3, 4 * H, D));
        CKS(h, dalloc(h, &h->b_l2, 4 * H));
        CKS(h, dalloc(h, &h->b_aux1, A));
        CKS(h, dalloc(h, &h->b_aux2, A));
        CKS(h, dalloc(h, &h->w_aff, A));
        CKS(h, alloc_act(h, &h->feats16, static_cast<int>(BR), D, 8));
        CKS(h, alloc_act(h, &h->mean16, h->Bmax, D));
        if (h->split) CKS(h, dalloc(h, &h->enc_ctx, BR * A));
        else CKS(h, alloc_act(h, &h->enc16, static_cast<int>(BR), A, 8));
        CKS(h, dalloc(h, &h->G0, static_cast<size_t>(h->Bmax) * 4 * H));
        CKS(h, alloc_act(h, &h->XA, M, H + H));
        CKS(h, alloc_act(h, &h->XB, M, D + H + H));
        CKS(h, alloc_act(h, &h->Hb2, M, H));
        CKS(h, dalloc(h, &h->dec_ctx, static_cast<size_t>(M) * A));
        CKS(h, dalloc(h, &h->c2[0], static_cast<size_t>(M) * H));
        CKS(h, dalloc(h, &h->c2[1], static_cast<size_t>(M) * H));
    } else if (c.arch == CAPDEC_ARCH_NIC) {
        CKS(h, alloc_act(h, &h->W_l1, 4 * H, H));
        CKS(h, alloc_act(h, &h->XA, M, H));
        CKS(h, alloc_act(h, &h->Hb, M, H));
        CKS(h, alloc_act(h, &h->Xp, h->Bmax, E));
        CKS(h, alloc_act(h, &h->H0, h->Bmax, H));
        CKS(h, dalloc(h, &h->c0, static_cast<size_t>(h->Bmax) * H));
    } else {
        const size_t BR = static_cast<size_t>(h->Bmax) * h->Rmax;
        CKS(h, alloc_act(h, &h->W_l1, 4 * H, H + H));
        CKS(h, alloc_act(h, &h->W_aux1, H, H));
        CKS(h, alloc_act(h, &h->W_aux2, 2 * H, H));
        CKS(h, alloc_act(h, &h->W_aux3, 2 * H, 2 * H));
        CKS(h, dalloc(h, &h->b_aux1, H));
        CKS(h, dalloc(h, &h->b_aux2, 2 * H));
        CKS(h, dalloc(h, &h->b_aux3, 2 * H));
        CKS(h, alloc_act(h, &h->feats16, static_cast<int>(BR), H));
        if (aoa_mma_ok(h)) {
            CKS(h, alloc_act(h, &h->k16, static_cast<int>(BR), H, 8));
            CKS(h, alloc_act(h, &h->v16, static_cast<int>(BR), H, 8));
            CKS(h, alloc_act(h, &h->q16, M, H));
        } else {
            CKS(h, dalloc(h, &h->kv32, BR * 2 * H));
        }
        CKS(h, dalloc(h, &h->mean32, static_cast<size_t>(h->Bmax) * H));
        CKS(h, alloc_act(h, &h->XA, M, H + H));
        CKS(h, alloc_act(h, &h->XB, M, 2 * H));
        CKS(h, alloc_act(h, &h->Hb, M, H));
        CKS(h, alloc_act(h, &h->Hb2, M, H));
        CKS(h, dalloc(h, &h->q32, static_cast<size_t>(M) * H));
        CKS(h, dalloc(h, &h->ctx32, static_cast<size_t>(M) * H));
        CKS(h, dalloc(h, &h->h32, static_cast<size_t>(M) * H));
    }
    return CAPDEC_OK;
}

int capdec_create(const capdec_config* cfg, capdec_handle** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return CAPDEC_ERR_INVALID;
    }
    capdec_handle* h = new capdec_handle();
    h->cfg = *cfg;
    const int s = create_impl(h);
    if (s != CAPDEC_OK) {
        g_create_error = h->err;
        capdec_destroy(h);
        *out = nullptr;
        return s;
    }
    *out = h;
    return CAPDEC_OK;
}

int capdec_load_weight(capdec_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim, void* stream) {
    if (!h || !name || !data || !shape || ndim < 1 || ndim > 4) return h ? fail(h, CAPDEC_ERR_INVALID, "bad load_weight argument") : CAPDEC_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->cfg.device));
    Raw r;
    r.numel = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] <= 0) return fail(h, CAPDEC_ERR_INVALID, std::string("bad shape for ") + name);
        r.shape.push_back(shape[i]);
        r.numel *= static_cast<size_t>(shape[i]);
    }
    auto it = h->raw.find(name);
    if (it != h->raw.end()) {
        cudaFree(it->second.d);
        h->raw.erase(it);
    }
    CK(h, cudaMalloc(reinterpret_cast<void**>(&r.d), r.numel * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(r.d, data, r.numel * sizeof(float), cudaMemcpyDefault, st);
    if (e == cudaSuccess) {
        cudaPointerAttributes pa;
        // pageable / pinned host memory: make the copy complete before the caller may reuse the buffer
        if (cudaPointerGetAttributes(&pa, data) != cudaSuccess || pa.type != cudaMemoryTypeDevice) e = cudaStreamSynchronize(st);
        cudaGetLastError();
    }
    if (e != cudaSuccess) {
        cudaFree(r.d);
        return fail(h, CAPDEC_ERR_CUDA, std::string("copy of ") + name + ": " + cudaGetErrorString(e));
    }
    h->raw[name] = r;
    h->weights_ready = false;
    return CAPDEC_OK;
}

int capdec_finalize_weights(capdec_handle* h, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->cfg.device));
    CKS(h, finalize_predict(h, st));
    if (h->cfg.arch == CAPDEC_ARCH_BUTD) CKS(h, finalize_butd(h, st));
    else if (h->cfg.arch == CAPDEC_ARCH_NIC) CKS(h, finalize_nic(h, st));
    else {
        CKS(h, finalize_aoa(h, st));
        if (has_refiner_weights(h)) CKS(h, finalize_refiner(h, st));
    }
    CK(h, cudaStreamSynchronize(st));  // b_aff is read back; packing is a one-time load cost
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);  // kernel parameters baked into a captured decode (e.g. b_aff) may have changed
    h->graphs.clear();
    h->weights_ready = true;
    return CAPDEC_OK;
}

// feats16 != null: the features arrive as dense fp16 rows (packed feature shards) instead of fp32
static int prepare_impl(capdec_handle* h, const float* feats, const float* mask, int32_t batch, int32_t regions, cudaStream_t st,
                        const __half* feats16 = nullptr);

int capdec_prepare_f16(capdec_handle* h, const uint16_t* feats16, const float* mask, int32_t batch, int32_t regions, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->weights_ready) return fail(h, CAPDEC_ERR_STATE, "capdec_prepare_f16 before capdec_finalize_weights");
    if (!feats16 || batch <= 0 || batch > h->Bmax) return fail(h, CAPDEC_ERR_INVALID, "prepare_f16: batch out of range or null feats");
    if (h->split) return fail(h, CAPDEC_ERR_INVALID, "prepare_f16: the fp32-grade math mode (f16x3) needs fp32 features");
    if (h->cfg.arch == CAPDEC_ARCH_NIC) return fail(h, CAPDEC_ERR_INVALID, "prepare_f16: NIC takes its (B, E) image embedding in fp32");
    if (regions <= 0 || regions > h->Rmax) return fail(h, CAPDEC_ERR_INVALID, "prepare_f16: regions out of range");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->cfg.device));
    const __half* f16 = reinterpret_cast<const __half*>(feats16);
    if (h->cfg.arch == CAPDEC_ARCH_AOA && h->refiner_ready) {  // fp16 bottom-up features -> projection + refiner
        h->prepared = false;
        const float* m = nullptr;
        if (mask) {
            CK(h, cudaMemcpyAsync(h->mask_buf, mask, static_cast<size_t>(batch) * regions * sizeof(float), cudaMemcpyDeviceToDevice, st));
            m = h->mask_buf;
        }
        CKS(h, run_refiner(h, nullptr, m, batch, regions, st, f16));
        return prepare_impl(h, h->refined, m, batch, regions, st);
    }
    return prepare_impl(h, nullptr, mask, batch, regions, st, f16);
}

int capdec_prepare(capdec_handle* h, const float* feats, const float* mask, int32_t batch, int32_t regions, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->weights_ready) return fail(h, CAPDEC_ERR_STATE, "capdec_prepare before capdec_finalize_weights");
    if (!feats || batch <= 0 || batch > h->Bmax) return fail(h, CAPDEC_ERR_INVALID, "prepare: batch out of range or null feats");
    CK(h, cudaSetDevice(h->cfg.device));
    return prepare_impl(h, feats, mask, batch, regions, static_cast<cudaStream_t>(stream));
}

int capdec_prepare_bottom_up(capdec_handle* h, const float* bu_feats, const float* mask, int32_t batch, int32_t regions,
                             void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->weights_ready) return fail(h, CAPDEC_ERR_STATE, "capdec_prepare_bottom_up before capdec_finalize_weights");
    if (h->cfg.arch != CAPDEC_ARCH_AOA) return fail(h, CAPDEC_ERR_INVALID, "prepare_bottom_up: only the AoA captioners have an encoder-side refiner");
    if (!h->refiner_ready)
        return fail(h, CAPDEC_ERR_STATE, "prepare_bottom_up: the checkpoint carried no img_feats_porjection / aoa_refine entries");
    if (!bu_feats || batch <= 0 || batch > h->Bmax) return fail(h, CAPDEC_ERR_INVALID, "prepare_bottom_up: batch out of range or null feats");
    if (regions <= 0 || regions > h->Rmax) return fail(h, CAPDEC_ERR_INVALID, "prepare_bottom_up: regions out of range");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->cfg.device));
    h->prepared = false;
    const float* m = nullptr;
    if (mask) {  // library-owned copy first: the refiner and the (possibly replayed) decode read it at a fixed address
        CK(h, cudaMemcpyAsync(h->mask_buf, mask, static_cast<size_t>(batch) * regions * sizeof(float), cudaMemcpyDeviceToDevice, st));
        m = h->mask_buf;
    }
    CKS(h, run_refiner(h, bu_feats, m, batch, regions, st));
    return prepare_impl(h, h->refined, m, batch, regions, st);
}

int capdec_get_refined(capdec_handle* h, float* dst, void* stream) {
    if (!h || !dst) return h ? fail(h, CAPDEC_ERR_INVALID, "get_refined: null destination") : CAPDEC_ERR_INVALID;
    if (!h->refiner_ready || !h->prepared || h->feats != h->refined)
        return fail(h, CAPDEC_ERR_STATE, "get_refined: no batch prepared with capdec_prepare_bottom_up");
    CK(h, cudaSetDevice(h->cfg.device));
    CK(h, cudaMemcpyAsync(dst, h->refined, static_cast<size_t>(h->B) * h->R * h->H * sizeof(float), cudaMemcpyDeviceToDevice,
                          static_cast<cudaStream_t>(stream)));
    return CAPDEC_OK;
}

static int prepare_impl(capdec_handle* h, const float* feats, const float* mask, int32_t batch, int32_t regions, cudaStream_t st,
                        const __half* feats16) {
    const int H = h->H, E = h->E;
    h->prepared = false;
    Operand ma, mb;
    if (h->cfg.arch == CAPDEC_ARCH_NIC) {
        // priming step: (h,c) = lstm(image_embedding, (0,0))  (NIC_Model.py:52-56)
        cvt_f16_kernel<<<grid_for(static_cast<size_t>(batch) * E / 4), 256, 0, st>>>(feats, batch, E, h->Xp.p, h->Xp.ld, h->Xp.lo, 0);
        CK(h, cudaGetLastError());
        h->launches++;
        CKS(h, map_a(h, &ma, h->Xp));
        CKS(h, map_b(h, &mb, h->W_emb));
        EpiParams e{};
        e.bias = h->b_l1;
        e.c_out = h->c0;
        e.ldc = H;
        e.out16 = h->H0.p;
        e.ld16 = h->H0.ld;
        e.lo16 = h->H0.lo;
        CKS(h, launch_gemm(h, EPI_LSTM, 1, ma, h->Xp.lo, mb, h->W_emb.lo, batch, 4 * H, E, e, st));
        h->R = 0;
    } else {
        if (regions <= 0 || regions > h->Rmax) return fail(h, CAPDEC_ERR_INVALID, "prepare: regions out of range");
        const size_t BR = static_cast<size_t>(batch) * regions;
        if (h->cfg.arch == CAPDEC_ARCH_BUTD) {
            const int A = h->A, D = h->D;
            if (mask) return fail(h, CAPDEC_ERR_INVALID, "BUTD takes no region mask");
            // one pass: fp16 operand rows (padded for the attention kernel's bulk copies) + the mean over the regions
            if (feats16)
                butd_ingest_kernel<__half><<<batch, 256, 0, st>>>(feats16, regions, D, h->feats16.p, h->feats16.ld, h->feats16.lo,
                                                                  h->mean16.p, h->mean16.ld, h->mean16.lo);
            else
                butd_ingest_kernel<float><<<batch, 256, 0, st>>>(feats, regions, D, h->feats16.p, h->feats16.ld, h->feats16.lo,
                                                                 h->mean16.p, h->mean16.ld, h->mean16.lo);
            CK(h, cudaGetLastError());
            h->launches += 1;
            {  // enc_att(feats): once per image instead of once per step and beam (BUTD_Model.py:57)
                CKS(h, map_a(h, &ma, h->feats16));
                CKS(h, map_b(h, &mb, h->W_aux1));
                EpiParams e{};
                e.bias = h->b_aux1;
                if (h->split) {
                    e.out32 = h->enc_ctx;
                    e.ld32 = A;
                } else {
                    e.out16 = h->enc16.p;
                    e.ld16 = h->enc16.ld;
                }
                CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->feats16.lo, mb, h->W_aux1.lo, static_cast<int>(BR), A, D, e, st));
            }
            {  // hoisted, step-invariant part of the top-down LSTM gates: W_ih[:, mean] * mean + b_ih + b_hh
                CKS(h, map_a(h, &ma, h->mean16));
                CKS(h, map_b(h, &mb, h->W_aux3));
                EpiParams e{};
                e.bias = h->b_l1;
                e.out32 = h->G0;
                e.ld32 = 4 * H;
                CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->mean16.lo, mb, h->W_aux3.lo, batch, 4 * H, D, e, st));
            }
        } else {
            if (feats16 || !feats)
                return fail(h, CAPDEC_ERR_INVALID, "prepare_f16: AoA refined features are taken in fp32 (fp16 applies to bu_feats)");
            cvt_f16_kernel<<<grid_for(BR * H / 4), 256, 0, st>>>(feats, BR, H, h->feats16.p, h->feats16.ld, h->feats16.lo, 0);
            CK(h, cudaGetLastError());
            region_mean_kernel<float><<<batch, 256, 0, st>>>(feats, H, mask, regions, H, h->mean32, nullptr, 0, 0);
            CK(h, cudaGetLastError());
            h->launches += 2;
            CKS(h, map_a(h, &ma, h->feats16));
            if (aoa_mma_ok(h)) {  // fp16 K and V as separate row-padded matrices for the fragment kernel
                for (int part = 0; part < 2; ++part) {
                    CKS(h, make_operand(h, &mb, h->W_aux2.p + static_cast<size_t>(part) * H * h->W_aux2.ld, H, h->W_aux2.ld, 0,
                                        h->pair_gemm ? BN / 2 : BN));
                    EpiParams e{};
                    e.bias = h->b_aux2 + part * H;
                    Act16& dst = part == 0 ? h->k16 : h->v16;
                    e.out16 = dst.p;
                    e.ld16 = dst.ld;
                    CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->feats16.lo, mb, h->W_aux2.lo, static_cast<int>(BR), H, H, e, st));
                }
            } else {
                CKS(h, map_b(h, &mb, h->W_aux2));
                EpiParams e{};
                e.bias = h->b_aux2;
                e.out32 = h->kv32;
                e.ld32 = 2 * H;
                CKS(h, launch_gemm(h, EPI_STORE, 1, ma, h->feats16.lo, mb, h->W_aux2.lo, static_cast<int>(BR), 2 * H, H, e, st));
            }
        }
        h->R = regions;
    }
    h->B = batch;
    h->feats = feats;
    h->mask = nullptr;
    if (mask) {  // keep a copy: the decode loop (possibly a replayed graph) reads it at a fixed address
        if (mask != h->mask_buf)
            CK(h, cudaMemcpyAsync(h->mask_buf, mask, static_cast<size_t>(batch) * regions * sizeof(float), cudaMemcpyDeviceToDevice, st));
        h->mask = h->mask_buf;
    }
    h->prepared = true;
    return CAPDEC_OK;
}

static int enqueue_beam_search(capdec_handle* h, int32_t beam, int32_t max_seq, int32_t* tokens, float* seq_logprob,
                               int32_t* lengths, float* alphas, cudaStream_t st);

int capdec_beam_search(capdec_handle* h, int32_t beam, int32_t max_seq, int32_t* tokens, float* seq_logprob, int32_t* lengths,
                       float* alphas, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->prepared) return fail(h, CAPDEC_ERR_STATE, "capdec_beam_search before capdec_prepare");
    if (beam <= 0 || beam > h->Kmax || max_seq <= 0 || max_seq > h->Tmax || !tokens)
        return fail(h, CAPDEC_ERR_INVALID, "beam_search: beam / max_seq out of range or null tokens");
    if (alphas && h->cfg.arch == CAPDEC_ARCH_NIC) return fail(h, CAPDEC_ERR_INVALID, "NIC has no attention maps; pass alphas = NULL");
    if (static_cast<int64_t>(beam) * h->V < beam) return fail(h, CAPDEC_ERR_INVALID, "beam larger than vocabulary");
    if (alphas && !h->alpha_step) {  // first request: history buffers (not part of the steady-state workspace)
        CKS(h, dalloc(h, &h->alpha_step, static_cast<size_t>(h->Tmax) * h->Mmax * h->Rmax));
        CKS(h, dalloc(h, &h->hist_parent, static_cast<size_t>(h->Tmax) * h->Mmax));
        CKS(h, dalloc(h, &h->best_pslot, h->Bmax));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(h, cudaSetDevice(h->cfg.device));
    // Replay the whole decode (6 launches x max_seq steps) as one CUDA graph while shapes and buffers are unchanged.
    // Not on the legacy default stream (cannot be captured), not while profiling, not with the attention-map output.
    if (!h->use_graphs || h->prof || alphas || st == nullptr)
        return enqueue_beam_search(h, beam, max_seq, tokens, seq_logprob, lengths, alphas, st);
    const capdec_handle::GraphKey key{capdec_handle::GK_BEAM, h->B, h->R, beam, max_seq, 0, 0,
                                      h->split ? static_cast<const void*>(h->feats) : nullptr, h->mask != nullptr};
    CKS(h, replay_graph(h, key, st, [&]() {
        return enqueue_beam_search(h, beam, max_seq, h->out_tokens, h->out_scores, h->out_lengths, nullptr, st);
    }));
    const int B = h->B;
    CK(h, cudaMemcpyAsync(tokens, h->out_tokens, static_cast<size_t>(B) * (max_seq + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (seq_logprob) CK(h, cudaMemcpyAsync(seq_logprob, h->out_scores, static_cast<size_t>(B) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (lengths) CK(h, cudaMemcpyAsync(lengths, h->out_lengths, static_cast<size_t>(B) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return CAPDEC_OK;
}

static int enqueue_beam_search(capdec_handle* h, int32_t beam, int32_t max_seq, int32_t* tokens, float* seq_logprob,
                               int32_t* lengths, float* alphas, cudaStream_t st) {
    const int B = h->B, K = beam, M = B * K;
    CKS(h, reset_state(h, M, st));
    BeamState s{};
    s.B = B, s.K = K, s.V = h->V, s.T = max_seq;
    s.tok = h->tok, s.cum = h->cum, s.parent = h->parent, s.n_live = h->n_live;
    s.best_score = h->best_score, s.best_seq = h->best_seq, s.best_len = h->best_len;
    s.seqs_in = h->seqs[0], s.seqs_out = h->seqs[1];
    s.hist_parent = alphas ? h->hist_parent : nullptr;
    s.best_pslot = alphas ? h->best_pslot : nullptr;
    const bool nic = h->cfg.arch == CAPDEC_ARCH_NIC;
    // NIC: step 1 reads the primed cell state of the image, so parent[row] starts as the image index into c0
    if (K <= 1) beam_init_kernel<1><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    else if (K <= 3) beam_init_kernel<3><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    else if (K <= 5) beam_init_kernel<5><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    else beam_init_kernel<8><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    CK(h, cudaGetLastError());
    h->launches++;
    const int n_slots = logit_slots(h, M, h->V, EPI_TOPK);  // partial records per row written by the logit GEMM
    StepCtx c{};
    c.M = M, c.K = K, c.logits_epi = EPI_TOPK, c.ktop = ktop_for(K);
    const AdvOps ops = adv_ops(h, false);
    // Small-batch BUTD path: step t's bookkeeping kernel and step t+1's [top-down gates -> dec_att] launch run concurrently (the
    // launch streams its weight tiles and waits for the bookkeeping's counter before it loads the operand rows)
    const bool beam_overlap = h->cfg.arch == CAPDEC_ARCH_BUTD && h->small_overlap && h->side && h->beam_done && small_ok(h, M) && h->small_fuse &&
                              !h->prof && !h->split && !alphas && B <= 20;  // (measured: +2 % up to 20 images, -1.6 % at 42)
    for (int t = 1; t <= max_seq; ++t) {
        c.t = t;
        c.cur = (t - 1) & 1;
        c.first_from_c0 = nic && t == 1;
        c.alphas = alphas ? h->alpha_step + static_cast<size_t>(t - 1) * M * h->R : nullptr;
        c.alpha_stride = h->R;
        if (beam_overlap) {
            c.beam_ctr = h->beam_done;
            c.beam_target = B * (t - 1);  // cumulative: reset_state zeroed the counter
        }
        CKS(h, run_step(h, c, st));
        s.seqs_in = h->seqs[(t + 1) & 1];
        s.seqs_out = h->seqs[t & 1];
        cudaStream_t bst = st;
        int* done = nullptr;
        if (beam_overlap) {  // the bookkeeping goes to the side stream: the next step's first launch starts beside it
            CK(h, cudaEventRecord(h->ev_fork2, st));
            CK(h, cudaStreamWaitEvent(h->side, h->ev_fork2, 0));
            bst = h->side;
            done = h->beam_done;
        }
        prof_begin(h, CAPDEC_CAT_BOOKKEEPING, 0.0, bst);
        if (K <= 1) CK(h, launch_pdl(h, beam_step_kernel<4, 1>, dim3(B), dim3(128), 0, bst, h->part, n_slots, s, t, ops, done));
        else if (K <= 3) CK(h, launch_pdl(h, beam_step_kernel<4, 3>, dim3(B), dim3(128), 0, bst, h->part, n_slots, s, t, ops, done));
        else if (K <= 4) CK(h, launch_pdl(h, beam_step_kernel<4, 5>, dim3(B), dim3(128), 0, bst, h->part, n_slots, s, t, ops, done));
        else if (K <= 5) CK(h, launch_pdl(h, beam_step_kernel<8, 5>, dim3(B), dim3(128), 0, bst, h->part, n_slots, s, t, ops, done));
        else CK(h, launch_pdl(h, beam_step_kernel<8, 8>, dim3(B), dim3(128), 0, bst, h->part, n_slots, s, t, ops, done));
        prof_end(h, bst);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    if (beam_overlap) {  // the side stream joins before the result is selected
        CK(h, cudaEventRecord(h->ev_join2, h->side));
        CK(h, cudaStreamWaitEvent(st, h->ev_join2, 0));
    }
    beam_finalize_kernel<<<(B + 3) / 4, 128, 0, st>>>(s, h->seqs[max_seq & 1], tokens, seq_logprob, lengths, h->alpha_step, alphas, h->R);
    CK(h, cudaGetLastError());
    h->launches++;
    return CAPDEC_OK;
}

static int sample_impl(capdec_handle* h, int32_t mode, int32_t n_per_image, uint64_t seed, int32_t max_seq, int32_t* tokens,
                       float* logprobs, float* alphas, const int32_t* forced, cudaStream_t st, int32_t* greedy_tokens = nullptr,
                       const uint32_t* seed_ptr = nullptr);

// Rollouts replay a captured graph like beam search does: the seed travels through device memory, the teacher-forced
// words and the outputs through library-owned buffers (stable addresses), copied from / to the caller's around the launch.
static int sample_entry(capdec_handle* h, int kind, int32_t mode, int32_t n, uint64_t seed, int32_t max_seq, int32_t* tokens,
                        float* logprobs, float* alphas, const int32_t* forced, cudaStream_t st, int32_t* greedy_tokens = nullptr) {
    if (!h->use_graphs || h->prof || alphas || st == nullptr)
        return sample_impl(h, mode, n, seed, max_seq, tokens, logprobs, alphas, forced, st, greedy_tokens);
    const int B = h->B, M = B * n, M_out = greedy_tokens ? B * (n - 1) : M;
    set_u32_kernel<<<1, 1, 0, st>>>(h->seed_dev, static_cast<uint32_t>(seed & 0xFFFFFFFFu));
    CK(h, cudaGetLastError());
    if (forced) CK(h, cudaMemcpyAsync(h->forced_buf, forced, static_cast<size_t>(M) * max_seq * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    const capdec_handle::GraphKey key{kind, B, h->R, n, max_seq, mode, (tokens ? 1 : 0) | (logprobs ? 2 : 0),
                                      h->split ? static_cast<const void*>(h->feats) : nullptr, h->mask != nullptr};
    CKS(h, replay_graph(h, key, st, [&]() {
        return sample_impl(h, mode, n, 0, max_seq, tokens ? h->out_sample_tokens : nullptr, logprobs ? h->out_sample_logprobs : nullptr,
                           nullptr, forced ? h->forced_buf : nullptr, st, greedy_tokens ? h->out_greedy : nullptr, h->seed_dev);
    }));
    if (tokens) CK(h, cudaMemcpyAsync(tokens, h->out_sample_tokens, static_cast<size_t>(M_out) * max_seq * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (logprobs) CK(h, cudaMemcpyAsync(logprobs, h->out_sample_logprobs, static_cast<size_t>(M_out) * max_seq * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (greedy_tokens) CK(h, cudaMemcpyAsync(greedy_tokens, h->out_greedy, static_cast<size_t>(B) * max_seq * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return CAPDEC_OK;
}

int capdec_sample(capdec_handle* h, int32_t mode, int32_t n_per_image, uint64_t seed, int32_t max_seq, int32_t* tokens,
                  float* logprobs, float* alphas, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->prepared) return fail(h, CAPDEC_ERR_STATE, "capdec_sample before capdec_prepare");
    if (n_per_image <= 0 || n_per_image > h->Kmax || max_seq <= 0 || max_seq > h->Tmax || !tokens)
        return fail(h, CAPDEC_ERR_INVALID, "sample: n_per_image / max_seq out of range or null tokens");
    if (mode != CAPDEC_SAMPLE_GREEDY && mode != CAPDEC_SAMPLE_MULTINOMIAL) return fail(h, CAPDEC_ERR_INVALID, "unknown sample mode");
    if (alphas && h->cfg.arch == CAPDEC_ARCH_NIC) return fail(h, CAPDEC_ERR_INVALID, "NIC has no attention maps; pass alphas = NULL");
    CK(h, cudaSetDevice(h->cfg.device));
    return sample_entry(h, capdec_handle::GK_SAMPLE, mode, n_per_image, seed, max_seq, tokens, logprobs, alphas, nullptr,
                        static_cast<cudaStream_t>(stream));
}

int capdec_scst_rollout(capdec_handle* h, int32_t n_per_image, uint64_t seed, int32_t max_seq, int32_t* sample_tokens,
                        float* sample_logprobs, int32_t* greedy_tokens, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->prepared) return fail(h, CAPDEC_ERR_STATE, "capdec_scst_rollout before capdec_prepare");
    if (n_per_image <= 0 || n_per_image + 1 > h->Kmax || max_seq <= 0 || max_seq > h->Tmax || !sample_tokens || !greedy_tokens)
        return fail(h, CAPDEC_ERR_INVALID, "scst_rollout: n_per_image + 1 must be <= max_rows; null outputs or max_seq out of range");
    CK(h, cudaSetDevice(h->cfg.device));
    return sample_entry(h, capdec_handle::GK_SCST, CAPDEC_SAMPLE_MULTINOMIAL, n_per_image + 1, seed, max_seq, sample_tokens,
                        sample_logprobs, nullptr, nullptr, static_cast<cudaStream_t>(stream), greedy_tokens);
}

int capdec_score(capdec_handle* h, const int32_t* tokens, int32_t n_per_image, int32_t max_seq, float* logprobs, void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!h->prepared) return fail(h, CAPDEC_ERR_STATE, "capdec_score before capdec_prepare");
    if (n_per_image <= 0 || n_per_image > h->Kmax || max_seq <= 0 || max_seq > h->Tmax || !tokens || !logprobs)
        return fail(h, CAPDEC_ERR_INVALID, "score: n_per_image / max_seq out of range or null tokens / logprobs");
    CK(h, cudaSetDevice(h->cfg.device));
    return sample_entry(h, capdec_handle::GK_SCORE, CAPDEC_SAMPLE_GREEDY, n_per_image, 0, max_seq, nullptr, logprobs, nullptr, tokens,
                        static_cast<cudaStream_t>(stream));
}

int capdec_score_states(capdec_handle* h, const int32_t* tokens, int32_t n_per_image, int32_t max_seq, float* logprobs, float* states,
                        void* stream) {
    if (!h) return CAPDEC_ERR_INVALID;
    if (!states) return capdec_score(h, tokens, n_per_image, max_seq, logprobs, stream);
    if (!h->prepared) return fail(h, CAPDEC_ERR_STATE, "capdec_score_states before capdec_prepare");
    if (n_per_image <= 0 || n_per_image > h->Kmax || max_seq <= 0 || max_seq > h->Tmax || !tokens || !logprobs)
        return fail(h, CAPDEC_ERR_INVALID, "score_states: n_per_image / max_seq out of range or null tokens / logprobs");
    CK(h, cudaSetDevice(h->cfg.device));
    h->states_out = states;  // caller's buffer: this pass is enqueued directly (not replayed from the graph cache)
    const int status = sample_impl(h, CAPDEC_SAMPLE_GREEDY, n_per_image, 0, max_seq, nullptr, logprobs, nullptr, tokens,
                                   static_cast<cudaStream_t>(stream));
    h->states_out = nullptr;
    return status;
}

// greedy_tokens != null: the SCST pair of rollouts in one pass -- n_per_image rows per image of which the LAST is the greedy
// rollout (written to greedy_tokens [B, T]); tokens / logprobs then hold the n_per_image - 1 sampled rows per image.
static int sample_impl(capdec_handle* h, int32_t mode, int32_t n_per_image, uint64_t seed, int32_t max_seq, int32_t* tokens,
                       float* logprobs, float* alphas, const int32_t* forced, cudaStream_t st, int32_t* greedy_tokens,
                       const uint32_t* seed_ptr) {
    const int B = h->B, n = n_per_image, M = B * n;
    const int M_out = greedy_tokens ? B * (n - 1) : M;
    CKS(h, reset_state(h, M, st));
    if (tokens) CK(h, cudaMemsetAsync(tokens, 0, static_cast<size_t>(M_out) * max_seq * sizeof(int32_t), st));
    if (logprobs) CK(h, cudaMemsetAsync(logprobs, 0, static_cast<size_t>(M_out) * max_seq * sizeof(float), st));
    SampleState s{};
    s.B = B, s.n = n, s.V = h->V, s.T = max_seq;
    s.tok = h->tok, s.unfinished = h->unfinished, s.live_count = h->live_count, s.parent = h->parent;
    s.tokens = tokens, s.logprobs = logprobs;
    s.multinomial = mode == CAPDEC_SAMPLE_MULTINOMIAL;
    s.scst = greedy_tokens != nullptr;
    s.greedy_tokens = greedy_tokens;
    const bool nic = h->cfg.arch == CAPDEC_ARCH_NIC;
    if (n <= 1) sample_init_kernel<1><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    else if (n <= 3) sample_init_kernel<3><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    else if (n <= 5) sample_init_kernel<5><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    else sample_init_kernel<8><<<B, 128, 0, st>>>(s, adv_ops(h, true), nic ? 1 : 0);
    CK(h, cudaGetLastError());
    h->launches++;
    const int n_slots = logit_slots(h, M, h->V, EPI_SAMPLE);
    StepCtx c{};
    c.M = M, c.K = n, c.logits_epi = EPI_SAMPLE, c.ktop = 1;
    c.seed = static_cast<uint32_t>(seed & 0xFFFFFFFFu);
    c.seed_ptr = seed_ptr;
    c.use_noise = s.multinomial;
    c.scst_n = greedy_tokens ? n - 1 : 0;
    c.forced = forced;
    c.forced_ld = max_seq;
    const AdvOps ops = adv_ops(h, false);
    for (int t = 1; t <= max_seq; ++t) {
        c.t = t;
        c.cur = (t - 1) & 1;
        c.first_from_c0 = nic && t == 1;
        c.states = (forced && h->states_out) ? h->states_out + static_cast<size_t>(t - 1) * h->H : nullptr;  // [row, t, unit]
        c.states_ld = static_cast<size_t>(max_seq) * h->H;
        c.alphas = alphas ? alphas + static_cast<size_t>(t - 1) * h->R : nullptr;  // [row, t, region]
        c.alpha_stride = static_cast<size_t>(max_seq) * h->R;
        CKS(h, run_step(h, c, st));
        prof_begin(h, CAPDEC_CAT_BOOKKEEPING, 0.0, st);
        if (n <= 1) CK(h, launch_pdl(h, sample_step_kernel<1>, dim3(B), dim3(128), 0, st, h->part, n_slots, s, t - 1, ops));
        else if (n <= 3) CK(h, launch_pdl(h, sample_step_kernel<3>, dim3(B), dim3(128), 0, st, h->part, n_slots, s, t - 1, ops));
        else if (n <= 5) CK(h, launch_pdl(h, sample_step_kernel<5>, dim3(B), dim3(128), 0, st, h->part, n_slots, s, t - 1, ops));
        else CK(h, launch_pdl(h, sample_step_kernel<8>, dim3(B), dim3(128), 0, st, h->part, n_slots, s, t - 1, ops));
        prof_end(h, st);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    return CAPDEC_OK;
}

int capdec_profile(capdec_handle* h, int32_t enable) {
    if (!h) return CAPDEC_ERR_INVALID;
    for (auto& r : h->recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->recs.clear();
    h->prof = enable != 0;
    return CAPDEC_OK;
}

int capdec_profile_read(capdec_handle* h, double* ms, double* flops, int64_t* launches) {
    if (!h || !ms || !flops || !launches) return CAPDEC_ERR_INVALID;
    for (int i = 0; i < CAPDEC_NUM_CATEGORIES; ++i) ms[i] = 0.0, flops[i] = 0.0, launches[i] = 0;
    for (auto& r : h->recs) {
        if (!r.b) continue;
        CK(h, cudaEventSynchronize(r.b));
        float t = 0.f;
        CK(h, cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cat] += t;
        flops[r.cat] += r.flops;
        launches[r.cat] += 1;
    }
    return capdec_profile(h, h->prof ? 1 : 0);
}

int capdec_test_gemm(const float* a, const float* b, const float* bias, float* d, int32_t m, int32_t n, int32_t k, int32_t math_mode,
                     void* stream) {
    capdec_handle tmp;
    capdec_handle* h = &tmp;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "capdec_test_gemm needs an sm_100a device";
        return CAPDEC_ERR_CUDA;
    }
    h->cfg.device = dev;
    h->num_sms = prop.multiProcessorCount;
    h->split = math_mode == CAPDEC_MATH_F16X3;
    {
        const char* g1 = getenv("CAPDEC_GEMM_1CTA");
        h->pair_gemm = !(g1 && g1[0] == '1');
    }
    int status = CAPDEC_OK;
    Act16 A16, B16;
    do {
        if ((status = alloc_small(h)) != CAPDEC_OK) break;
        if ((status = alloc_act(h, &A16, m, k)) != CAPDEC_OK) break;
        if ((status = alloc_act(h, &B16, n, k)) != CAPDEC_OK) break;
        cvt_f16_kernel<<<grid_for(static_cast<size_t>(m) * k / 4), 256, 0, st>>>(a, m, k, A16.p, A16.ld, A16.lo, 0);
        cvt_f16_kernel<<<grid_for(static_cast<size_t>(n) * k / 4), 256, 0, st>>>(b, n, k, B16.p, B16.ld, B16.lo, 0);
        Operand ma, mb;
        if ((status = map_a(h, &ma, A16)) != CAPDEC_OK) break;
        if ((status = map_b(h, &mb, B16)) != CAPDEC_OK) break;
        EpiParams e{};
        e.bias = bias;
        e.out32 = d;
        e.ld32 = n;
        if ((status = launch_gemm(h, EPI_STORE, 1, ma, A16.lo, mb, B16.lo, m, n, k, e, st)) != CAPDEC_OK) break;
        cudaError_t ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {
            h->err = std::string("test gemm: ") + cudaGetErrorString(ce);
            status = CAPDEC_ERR_CUDA;
        }
    } while (0);
    for (void* p : h->allocs) cudaFree(p);
    if (status != CAPDEC_OK) g_create_error = h->err;
    return status;
}

// Test hook: time `iters` back-to-back launches of one GEMM shape with the given epilogue on synthetic operands.
int capdec_test_gemm_time(int32_t m, int32_t n, int32_t k, int32_t epi, int32_t math_mode, int32_t iters, float* us_per_launch) {
    capdec_handle tmp;
    capdec_handle* h = &tmp;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
        g_create_error = "capdec_test_gemm_time needs an sm_100a device";
        return CAPDEC_ERR_CUDA;
    }
    h->cfg.device = dev;
    h->num_sms = prop.multiProcessorCount;
    h->split = math_mode == CAPDEC_MATH_F16X3;
    {
        const char* g1 = getenv("CAPDEC_GEMM_1CTA");
        h->pair_gemm = !(g1 && g1[0] == '1');
    }
    int status = CAPDEC_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        Act16 A16, B16, H16;
        float *bias = nullptr, *c0 = nullptr, *c1 = nullptr, *out = nullptr, *part = nullptr;
        if ((status = alloc_small(h)) != CAPDEC_OK) break;
        if ((status = alloc_act(h, &A16, m, k)) != CAPDEC_OK) break;
        if ((status = alloc_act(h, &B16, n, k)) != CAPDEC_OK) break;
        fill_f16_kernel<<<1024, 256>>>(A16.p, static_cast<size_t>(m) * A16.ld, 0.01f);
        fill_f16_kernel<<<1024, 256>>>(B16.p, static_cast<size_t>(n) * B16.ld, 0.01f);
        if ((status = dalloc(h, &bias, n)) != CAPDEC_OK) break;
        Operand ma, mb;
        if ((status = map_a(h, &ma, A16)) != CAPDEC_OK) break;
        if ((status = map_b(h, &mb, B16)) != CAPDEC_OK) break;
        EpiParams e{};
        e.bias = bias;
        if (epi == EPI_STORE) {
            if ((status = dalloc(h, &out, static_cast<size_t>(m) * n)) != CAPDEC_OK) break;
            e.out32 = out;
            e.ld32 = n;
        } else if (epi == EPI_LSTM) {
            if ((status = dalloc(h, &c0, static_cast<size_t>(m) * n / 4)) != CAPDEC_OK) break;
            if ((status = dalloc(h, &c1, static_cast<size_t>(m) * n / 4)) != CAPDEC_OK) break;
            if ((status = alloc_act(h, &H16, m, n / 4)) != CAPDEC_OK) break;
            e.c_in = c0, e.c_out = c1, e.ldc = n / 4;
            e.out16 = H16.p, e.ld16 = H16.ld, e.lo16 = H16.lo;
        } else if (epi == EPI_TOPK) {
            const int nt = logit_slots(h, m, n, EPI_TOPK);
            if ((status = dalloc(h, &part, static_cast<size_t>(m) * nt * topk_part_stride(4))) != CAPDEC_OK) break;
            e.part = part;
            e.n_tiles = nt;
        } else {
            status = fail(h, CAPDEC_ERR_INVALID, "unsupported epilogue for the timing hook");
            break;
        }
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 3 && status == CAPDEC_OK; ++i) status = launch_gemm(h, epi, 4, ma, A16.lo, mb, B16.lo, m, n, k, e, nullptr);
        if (status != CAPDEC_OK) break;
        cudaEventRecord(e0, nullptr);
        for (int i = 0; i < iters && status == CAPDEC_OK; ++i) status = launch_gemm(h, epi, 4, ma, A16.lo, mb, B16.lo, m, n, k, e, nullptr);
        cudaEventRecord(e1, nullptr);
        cudaError_t ce = cudaEventSynchronize(e1);
        if (ce != cudaSuccess) {
            status = fail(h, CAPDEC_ERR_CUDA, std::string("gemm timing: ") + cudaGetErrorString(ce));
            break;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *us_per_launch = 1e3f * ms / iters;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    for (void* p : h->allocs) cudaFree(p);
    if (status != CAPDEC_OK) g_create_error = h->err;
    return status;
}

}  // extern "C"
