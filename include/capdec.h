/* capdec -- C ABI of the B200-native batched caption decoder (libcapdec.so).
 *
 * Drop-in boundary for the decode loop of zyj0021200/simpleImageCaptionZoo: the library replaces what the
 * reference runs inside
 *     DecoderRNN.beam_search_sample / sample / sample_rl      Models/BUTD_Model.py:153-318
 *     DecoderRNN.beam_search_sample / sample / sample_rl      Models/NIC_Model.py:100-212
 *     AoA_Decoder.beam_search_sample / sample / sample_rl     Models/AoA_Model.py:295-502
 * and is called from the Python Engine mirror (simpleimagecaptionzoo_b200/engine.py) the way
 * Engine.eval_captions_json_generation (Engine.py:274-300) and Engine.SCST_training_epoch (Engine.py:251-272)
 * call model.beam_search_sampler / model.sampler / model.sampler_rl.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative capdec_status,
 * with a message available from capdec_last_error().  No C++ exception crosses the boundary.  All GPU work is
 * enqueued on the caller's stream (a cudaStream_t passed as void*); nothing synchronises the device unless
 * stated.  One handle per (process, device); a handle is not re-entrant.  The caller owns every input and
 * output buffer and keeps it alive until the stream work has completed; the library owns its packed weights
 * and workspace, allocated in capdec_create (no allocation inside the decode loop).
 *
 * Concurrency on one device: decodes of <= 128 rows (batch x beam) run persistent kernels whose CTAs wait for each other
 * (split-K exchange, grid barriers) and assume one CTA per SM is resident -- issue the decodes of one device one after
 * the other (what the binding does: every call is ordered against the caller's current stream), as the reference does.
 * Two such decodes racing for the SMs from different streams can starve each other; every device-side wait is bounded
 * (~2 s) and then fails the launch with an error instead of hanging.
 */
#ifndef CAPDEC_H_
#define CAPDEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPDEC_ABI_VERSION 1

typedef enum capdec_status {
    CAPDEC_OK = 0,
    CAPDEC_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    CAPDEC_ERR_CUDA = -2,      /* a CUDA call failed (message has the CUDA error string) */
    CAPDEC_ERR_STATE = -3,     /* call order violated (e.g. decode before prepare) */
    CAPDEC_ERR_NOMEM = -4,
    CAPDEC_ERR_WEIGHT = -5     /* missing / mis-shaped state_dict entry */
} capdec_status;

typedef enum capdec_arch {
    CAPDEC_ARCH_NIC = 0,       /* Models/NIC_Model.py  DecoderRNN  (one LSTMCell primed by the image)      */
    CAPDEC_ARCH_BUTD = 1,      /* Models/BUTD_Model.py DecoderRNN  (top-down attention, two LSTMCells)      */
    CAPDEC_ARCH_AOA = 2        /* Models/AoA_Model.py  AoA_Decoder (LSTMCell + multi-head AoA)              */
} capdec_arch;

typedef enum capdec_math {
    CAPDEC_MATH_F16 = 0,       /* GEMM operands rounded to fp16 (10-bit mantissa, tf32-grade), fp32 accumulate */
    CAPDEC_MATH_F16X3 = 1      /* operands split hi+lo fp16, three tensor-core passes: fp32-grade products   */
} capdec_math;

typedef enum capdec_sample_mode {
    CAPDEC_SAMPLE_GREEDY = 0,      /* DecoderRNN.sample: argmax, max_seq fixed steps, no <end> handling        */
    CAPDEC_SAMPLE_MULTINOMIAL = 1  /* DecoderRNN.sample_rl: draw ~ softmax, <end> stored as 0, early break     */
} capdec_sample_mode;

/* Mirrors the model-settings keys read by Utils.model_construction (Utils.py:161-203):
 * embed_dim, hidden_dim, atten_dim (+ enc_dim = 2048, vocab size = len(caption_vocab), 8 AoA heads). */
typedef struct capdec_config {
    int32_t arch;          /* capdec_arch */
    int32_t hidden_dim;    /* H */
    int32_t embed_dim;     /* E */
    int32_t atten_dim;     /* A   (BUTD only) */
    int32_t enc_dim;       /* D   (BUTD: region feature width, 2048; AoA: width of the bottom-up features fed to
                              capdec_prepare_bottom_up, 2048 -- may be 0 when only refined features are decoded) */
    int32_t vocab_size;    /* V */
    int32_t num_heads;     /* AoA only */
    int32_t max_batch;     /* largest number of images per prepare() */
    int32_t max_regions;   /* largest R (36 bottom-up boxes, 49/196 grid cells); ignored for NIC */
    int32_t max_rows;      /* largest beam size / samples per image */
    int32_t max_seq;       /* largest number of decode steps */
    int32_t math_mode;     /* capdec_math */
    int32_t device;        /* CUDA device ordinal */
} capdec_config;

typedef struct capdec_handle capdec_handle;

int capdec_abi_version(void);

/* Allocate packed-weight storage and the decode workspace on cfg->device. */
int capdec_create(const capdec_config* cfg, capdec_handle** out);
void capdec_destroy(capdec_handle* h);
const char* capdec_last_error(const capdec_handle* h);  /* h may be NULL: error of the last failed create */

/* Hand one tensor of the reference's decoder state_dict to the library, under its reference name without the
 * "decoder." prefix (e.g. "predict.weight_g", "TD_atten.weight_ih", "embed.0.weight"; the checkpoint layout is
 * the one Engine.load_from_checkpoint reads, Engine.py:43-70).  data is fp32, row-major, host or device memory;
 * it is copied before the call returns control of the buffer (stream-ordered for device memory). */
int capdec_load_weight(capdec_handle* h, const char* name, const float* data, const int64_t* shape, int32_t ndim,
                       void* stream);
/* Fold weight-norm (w = g*v/||v||), fuse b_ih+b_hh, split W_ih by input segment, interleave LSTM gates / GLU
 * halves and convert to the fp16 (hi|lo) operand layout.  Fails with CAPDEC_ERR_WEIGHT if a tensor is missing. */
int capdec_finalize_weights(capdec_handle* h, void* stream);

/* One-time, step-invariant work for a batch of images (the part the reference recomputes every step):
 *   BUTD: feats [B,R,D] -> enc_att projection [B,R,A] (BUTD_Model.py:57), mean feature (:251), hoisted
 *         W_ih[:,mean]*mean + b_ih + b_hh of the top-down LSTM (:265).
 *   AoA : feats = refined features [B,R,H]; mask [B,R] float {0,1} or NULL -> masked mean (AoA_Model.py:422-425),
 *         K,V projections (:114-115).
 *   NIC : feats = image embedding [B,E] (R ignored) -> priming LSTM step (NIC_Model.py:52-56).
 * feats (and mask) are DEVICE pointers and must stay valid until the following decode call has completed. */
int capdec_prepare(capdec_handle* h, const float* feats, const float* mask, int32_t batch, int32_t regions, void* stream);

/* capdec_prepare for features that arrive as dense fp16 rows [B,R,D] (uint16_t = IEEE binary16 bits) -- the packed feature
 * shards of simpleimagecaptionzoo_b200/feature_store.py, which replace the reference's per-image zlib .npz files read
 * one by one (Datasets.py:138-145; half the host->device bytes, no conversion pass).  BUTD: D = enc_dim.  AoA with the
 * refiner loaded: D = enc_dim, runs what capdec_prepare_bottom_up runs.  fp16 math mode only (the library rounds fp32
 * features to fp16 for its operands anyway: same captions as capdec_prepare on the same rounded values). */
int capdec_prepare_f16(capdec_handle* h, const uint16_t* feats16, const float* mask, int32_t batch, int32_t regions, void* stream);

/* AoADetection_Captioner / AoASpatial_Captioner from the bottom-up (or CNN grid) features: the encoder-side half the
 * reference runs in front of AoA_Decoder on every sampler call (AoA_Model.py:748-751, 598-601) --
 *   img_feats_porjection = Linear(enc_dim -> H) + ReLU under pack_wrapper (:650-655, 661-665: rows where mask == 0
 *   come out as zeros), then AoA_Refine_Core (:140-162): N x [x += GLU(Linear([MHA(LN x), LN x]))] + final LayerNorm --
 * followed by what capdec_prepare does on the refined features.  Needs the checkpoint's "img_feats_porjection.*" and
 * "aoa_refine.*" entries (loaded under those names; the number of layers is taken from the checkpoint) and
 * cfg.enc_dim = width of bu_feats.  bu_feats [B,R,enc_dim] fp32 and mask [B,R] float {0,1} (prefix mask, or NULL) are
 * DEVICE pointers; both may be released once the stream work of this call has completed.  AoA only. */
int capdec_prepare_bottom_up(capdec_handle* h, const float* bu_feats, const float* mask, int32_t batch, int32_t regions,
                             void* stream);
/* Copy the refined features [B,R,H] fp32 of the batch prepared by capdec_prepare_bottom_up to dst (device memory):
 * what aoa_refine returns in the reference (AoA_Model.py:751).  Test / inspection hook. */
int capdec_get_refined(capdec_handle* h, float* dst, void* stream);

/* Batched beam search with the reference's exact bookkeeping (shrinking beam, best COMPLETED hypothesis wins,
 * no length normalisation; BUTD_Model.py:236-318) for the prepared batch.  Device outputs:
 *   tokens      [B, 1+max_seq] int32: <sta>=1 first, words, <end>=2 when completed, then <pad>=0
 *   seq_logprob [B] fp32 sum of log-probs of the emitted words        (may be NULL)
 *   lengths     [B] int32 valid entries of tokens incl. <sta>/<end>   (may be NULL)
 *   alphas      [B, max_seq, R] fp32 attention map of the returned hypothesis at every step it took, zero rows after
 *               its end (BUTD: softmax over regions, BUTD_Model.py:60; AoA: mean over heads, AoA_Model.py:119);
 *               may be NULL; must be NULL for NIC */
int capdec_beam_search(capdec_handle* h, int32_t beam, int32_t max_seq, int32_t* tokens, float* seq_logprob,
                       int32_t* lengths, float* alphas, void* stream);

/* Greedy (``sample``) or multinomial (``sample_rl``) rollout, n_per_image rows per prepared image.
 *   tokens   [B*n_per_image, max_seq] int32 (no <sta>)
 *   logprobs [B*n_per_image, max_seq] fp32 log-prob of each stored word (multinomial; may be NULL for greedy)
 *   alphas   [B*n_per_image, max_seq, R] fp32 attention maps (may be NULL; must be NULL for NIC)
 * The multinomial draw is the Gumbel-max form of torch.multinomial(exp(logprobs),1) with counter-based noise
 * keyed by (seed, row, step, word). */
int capdec_sample(capdec_handle* h, int32_t mode, int32_t n_per_image, uint64_t seed, int32_t max_seq, int32_t* tokens,
                  float* logprobs, float* alphas, void* stream);

/* The two rollouts of an SCST step (Engine.SCST_training_epoch, Engine.py:258-262: model.sampler then model.sampler_rl on
 * the same batch) in ONE pass over the prepared batch: n_per_image multinomial rows + one greedy row per image share every
 * GEMM / attention launch (the image's features are read once per step for all of them).  Row for row the outputs equal
 * capdec_sample(MULTINOMIAL, n_per_image, seed) and capdec_sample(GREEDY, 1): the sampled rows keep their noise streams.
 *   sample_tokens [B*n_per_image, max_seq] int32, sample_logprobs [B*n_per_image, max_seq] fp32 (may be NULL),
 *   greedy_tokens [B, max_seq] int32.  Needs n_per_image + 1 <= cfg.max_rows. */
int capdec_scst_rollout(capdec_handle* h, int32_t n_per_image, uint64_t seed, int32_t max_seq, int32_t* sample_tokens,
                        float* sample_logprobs, int32_t* greedy_tokens, void* stream);

/* Teacher-forced scoring of given word sequences for the prepared batch: what the reference's decoder ``forward`` computes
 * for given captions (BUTD_Model.py:97-151, NIC_Model.py:58-98, AoA_Model.py:229-293) followed by log_softmax + gather --
 * the ``seqLogprobs`` of an arbitrary rollout (Utils.py:290-317 RewardCriterion's input), forward values only.
 *   tokens   [B*n_per_image, max_seq] int32 words WITHOUT <sta> (the layout capdec_sample returns); word t is scored
 *            given <sta> and words 0..t-1 of the same row, and fed back as the next input whatever it is (0 = <pad> too)
 *   logprobs [B*n_per_image, max_seq] fp32 log p(word t | image, previous words) */
int capdec_score(capdec_handle* h, const int32_t* tokens, int32_t n_per_image, int32_t max_seq, float* logprobs, void* stream);

/* capdec_score that also exports what `predict` was applied to at every step -- h2 (BUTD_Model.py:270), h (NIC_Model.py:174),
 * the AoA context (AoA_Model.py:455) -- as states [B*n_per_image, max_seq, hidden_dim] fp32: with it the host rebuilds
 * predict + log_softmax + gather under autograd (scst.differentiable_logprobs), i.e. `seqLogprobs` WITH a graph through the
 * vocabulary layer for RewardCriterion (Utils.py:290-317, Engine.py:261-272).  Gradients of the recurrent weights (BPTT
 * through the LSTMs / attention) are the training path and stay with the reference.  states = NULL: same as capdec_score. */
int capdec_score_states(capdec_handle* h, const int32_t* tokens, int32_t n_per_image, int32_t max_seq, float* logprobs, float* states,
                        void* stream);

/* ---- CIDEr-D self-critical reward (the SCST step after the two rollouts) --------------------------------------------
 * Replaces Utils.get_self_critical_reward (Utils.py:319-367) -> CiderD.compute_score (cider/pyciderevalcap/ciderD/
 * ciderD.py:32-55) -> CiderScorer.compute_cider (ciderD_scorer.py:127-206, df mode "<dataset>-train") on word ids. */
typedef struct capdec_cider capdec_cider;
int capdec_cider_create(int32_t device, capdec_cider** out);
void capdec_cider_destroy(capdec_cider* c);
const char* capdec_cider_last_error(const capdec_cider* c);
/* 64-bit key of an n-gram of k word ids (the host side hashes the document-frequency table with the same function). */
uint64_t capdec_cider_ngram_key(const int32_t* ids, int32_t k);
/* Document frequencies of the training corpus (the reference's cider/data/<dataset>-train.p: 'document_frequency',
 * 'ref_len'): n entries, keys[i] = capdec_cider_ngram_key of the n-gram, df[i] = number of training images whose
 * references contain it; log_ref_len = log(number of training images).  HOST arrays, copied. */
int capdec_cider_set_df(capdec_cider* c, const uint64_t* keys, const float* df, int64_t n, double log_ref_len);
/* rewards[b*n + i] = weight * (CIDEr-D(sample i of image b) - CIDEr-D(greedy rollout of image b)), DEVICE pointers:
 *   gen        [B*n_per_image, max_seq] int32 as capdec_sample(MULTINOMIAL) stores them (<end> and after = 0)
 *   greedy     [B, max_seq] int32 as capdec_sample(GREEDY) stores them (the caption ends before the first <end>)
 *   ref_tokens [n_refs_total, ref_ld] int32 word ids of the reference captions (out-of-vocabulary words get ids >=
 *              vocabulary size on the host side), ref_lens [n_refs_total] (<= 65), ref_offsets [B+1]: image b owns the
 *              references [ref_offsets[b], ref_offsets[b+1])
 *   sigma = 6 (ciderD.py:25); scores [B, n_per_image+1] (samples then greedy) or NULL. */
int capdec_cider_reward(capdec_cider* c, const int32_t* gen, int32_t n_per_image, const int32_t* greedy, int32_t batch,
                        int32_t max_seq, const int32_t* ref_tokens, const int32_t* ref_lens, const int32_t* ref_offsets,
                        int32_t ref_ld, double sigma, double weight, float* rewards, float* scores, void* stream);

/* Number of kernels the library launched on behalf of this handle since create (bench.py's gpu_launches). */
int64_t capdec_launch_count(const capdec_handle* h);

/* Number of decode graphs captured so far.  Beam search and the rollouts replay a CUDA graph of their whole kernel
 * sequence, cached per (kind, batch, regions, rows per image, max_seq, ...) -- up to 8 entries, least recently used
 * replaced -- so a ragged last batch or alternating beam sizes capture once each. */
int64_t capdec_graph_captures(const capdec_handle* h);

/* Tuning hook (CAPDEC_TRACE=1 in the environment at create): %globaltimer stamps CTA 0 of the small-batch GEMM kernel took
 * along its phases -- dst[0] = launches recorded, then 16 stamps (ns) per launch; resets the launch counter.  HOST buffer. */
int capdec_debug_trace(capdec_handle* h, uint64_t* dst, int64_t n);

/* Per-launch timing for bench.py's roofline line: while enabled, every kernel the handle launches is bracketed by
 * CUDA events on the launching stream.  capdec_profile_read waits for the recorded events and returns, per
 * category, the summed device time (ms), the algorithmic FLOPs of the GEMM launches (2*M*N*K, one pass) and the
 * launch count; it then clears the records.  Arrays have CAPDEC_NUM_CATEGORIES entries. */
typedef enum capdec_category {
    CAPDEC_CAT_GEMM_LSTM = 0,    /* gate GEMM + fused LSTMCell pointwise */
    CAPDEC_CAT_GEMM_STORE = 1,   /* projection GEMMs (enc_att, dec_att, Q, K/V, hoisted mean term) */
    CAPDEC_CAT_GEMM_GLU = 2,     /* AoA gate GEMM + fused GLU */
    CAPDEC_CAT_GEMM_LOGITS = 3,  /* vocabulary GEMM + fused log-softmax partials / top-k / Gumbel-max */
    CAPDEC_CAT_ATTENTION = 4,    /* additive (BUTD) / multi-head (AoA) attention over the regions */
    CAPDEC_CAT_BOOKKEEPING = 5,  /* beam / sampling bookkeeping + state reorder + embedding gather */
    CAPDEC_CAT_OTHER = 6,
    CAPDEC_NUM_CATEGORIES = 7
} capdec_category;
int capdec_profile(capdec_handle* h, int32_t enable);
int capdec_profile_read(capdec_handle* h, double* ms, double* flops, int64_t* launches);

/* Test hook: D[M,N] = A[M,K] * B[N,K]^T (+bias[N]) through the library's tcgen05 GEMM, fp32 device buffers in/out.
 * K must be a multiple of 64. */
int capdec_test_gemm(const float* a, const float* b, const float* bias, float* d, int32_t m, int32_t n, int32_t k,
                     int32_t math_mode, void* stream);

/* Test hook: mean device time (us) of `iters` back-to-back launches of one GEMM shape with epilogue `epi`
 * (0 = store, 1 = LSTM with N = 4*H gate-interleaved columns, 3 = logits top-k) on synthetic operands. */
int capdec_test_gemm_time(int32_t m, int32_t n, int32_t k, int32_t epi, int32_t math_mode, int32_t iters, float* us_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* CAPDEC_H_ */
