/*
 * das_b200.h - C ABI of the B200-native active-selection scoring library (libdas_b200.so).
 *
 * This is the drop-in boundary for the scoring hot path of
 * nihalsid/deep-active-semantic-segmentation.  The reference has no FFI layer of its own: its
 * boundary is the Python method surface of active_selection (all .py files).  The mirror of that
 * surface lives in deep_active_semantic_segmentation_b200/active_selection/ and calls ONLY
 * the functions declared here (via ctypes; see INTEGRATION.md for the binding stub).  Each entry
 * point cites the reference code whose body it replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  All data pointers are DEVICE pointers
 *    (sm_100a, HBM) unless a comment says "host".  Tensors are contiguous, row-major, NCHW.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  No entry point
 *    synchronises the host; all work is enqueued on `stream`.
 *  - return value: DAS_OK (0) or a negative das_status; das_strerror() names it;
 *    das_last_cuda_error() returns the cudaError_t behind DAS_ERR_CUDA.
 *  - the caller owns every buffer; state / workspace sizes come from the *_bytes() queries.
 *  - every entry point that enqueues work takes a `das_handle*` first: one opaque handle per device
 *    (das_handle_create).  The handle holds what used to be process state: the device ordinal and its
 *    SM count (grid sizing), the tuning options (read from the environment ONCE, at creation), the
 *    k-center step scratch and a cache of encoded TMA descriptors.  A call made while another device
 *    is current switches to the handle's device for its duration.
 *  - thread-compatible, not thread-safe: one host thread per handle / state / workspace at a time;
 *    different handles (devices) may be used from different threads concurrently.
 *  - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef DAS_B200_H
#define DAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAS_ABI_VERSION 4

typedef enum das_status {
    DAS_OK = 0,
    DAS_ERR_INVALID_ARG = -1,   /* NULL pointer, non-positive size, pass index out of range ... */
    DAS_ERR_UNSUPPORTED = -2,   /* e.g. more than DAS_MAX_CLASSES classes, k > DAS_TOPK_MAX_K   */
    DAS_ERR_CUDA = -3,          /* a CUDA runtime call failed; see das_last_cuda_error()         */
    DAS_ERR_MISALIGNED = -4     /* a pointer violates the documented alignment                   */
} das_status;

const char* das_strerror(int status);
int das_abi_version(void);
/* cudaError_t behind the last DAS_ERR_CUDA returned to the calling thread */
int das_last_cuda_error(void);
/* number of kernels this library has launched since load, all handles (bench.py's "gpu_launches") */
uint64_t das_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Per-device handle ("one opaque per-device handle")
 * ------------------------------------------------------------------------------------------ */
typedef struct das_handle das_handle;

/* tuning options; defaults come from the environment variable named in the comment, read once in
 * das_handle_create, and can be changed per handle with das_handle_set_option */
enum {
    DAS_OPT_MC_TMA = 0,        /* DAS_MC_TMA       1: TMA-ring single-shot kernel when eligible (default), 0: LDG kernel   */
    DAS_OPT_MC_TMA_CTAS = 1,   /* DAS_MC_TMA_CTAS  CTAs per SM of the TMA kernel, 0 = per class count (default)           */
    DAS_OPT_MC_UP_WARPS = 2,   /* DAS_MC_UP_WARPS  kernel variant of the fused upsample: 0 = by class count and width (default),
                                  4 | 15 = pixel pairs with that many consumer warps, 220 | 216 = one pixel per lane with
                                  20 | 16 consumer + 4 producer warps                                                    */
    DAS_OPT_GEMM_2CTA = 3,     /* DAS_GEMM_2CTA    1: CTA-pair tcgen05 distance GEMM (default), 0: one CTA per tile       */
    DAS_OPT_KC_CLUSTER = 4,    /* DAS_KC_CLUSTER   1: k-center loop inside one thread-block cluster (default), 0: chain   */
    DAS_OPT_MC_L2_PERSIST = 5, /* DAS_MC_L2_PERSIST 1: streaming accumulators pinned in L2 when they fit (default), 0: off.
                                * Streaming launches (das_mc_accumulate, das_mc_finalize, a fused last group) then carry an
                                * access-policy window over the head of the state, and the first of them raises the
                                * device's persisting-L2 limit to its maximum; lines stay "persisting" until other
                                * persisting traffic replaces them or the application calls
                                * cudaCtxResetPersistingL2Cache().  The single-shot path never touches any of this. */
    DAS_OPT_COUNT = 6
};

/* device: CUDA ordinal, or -1 for the calling thread's current device */
int das_handle_create(int device, das_handle** out);
int das_handle_destroy(das_handle* h);
int das_handle_device(const das_handle* h);       /* ordinal, or a negative das_status */
int das_handle_sm_count(const das_handle* h);     /* multiprocessors of the device     */
/* L2 geometry the handle works with (bytes): total L2, the largest persisting carve-out, the largest access-policy window */
int das_handle_l2_info(const das_handle* h, size_t* l2_bytes, size_t* persisting_max, size_t* window_max);
int das_handle_set_option(das_handle* h, int option, int value);
int das_handle_get_option(const das_handle* h, int option, int* value);

/* ------------------------------------------------------------------------------------------
 * Monte-Carlo uncertainty reduction  (K1 accumulate + K2 finalize)
 *
 * Replaces the bodies of
 *   ActiveSelectionMCDropout._get_vote_entropy_for_batch      active_selection/mc_dropout.py:30-80
 *   ActiveSelectionMCNoise._get_vote_entropy_for_batch_with_* active_selection/mc_noise.py:21-44,62-84,86-114
 *   ActiveSelectionCEAL._get_entropies / get_least_*          active_selection/ceal.py:19-126   (T = 1)
 * and the per-image `torch.mean(...).cpu().item()` pooling (mc_dropout.py:189, mc_noise.py:56,
 * ceal.py:59,95,123).
 * ------------------------------------------------------------------------------------------ */

#define DAS_MAX_CLASSES 32      /* register-resident class vector; votes are uint8            */
#define DAS_MAX_PASSES 255      /* vote histogram counters are 8 bit                          */
#define DAS_MAX_PASS_GROUP 32   /* passes consumed by one das_mc_accumulate launch            */

enum {
    DAS_MC_VOTES = 1,  /* keep the per-pass argmax votes   -> vote entropy (reference-pinned)   */
    DAS_MC_PROBS = 2,  /* keep running sum of softmax probabilities and of per-pass entropies   */
                       /* -> predictive entropy, BALD, confidence, margin, expected entropy     */
    DAS_MC_SINGLE_SHOT = 4 /* all T (<= DAS_MAX_PASS_GROUP) passes arrive in ONE                */
                       /* das_mc_accumulate_finalize call: the state holds only block partials  */
};

/* order of the per-image scores written by das_mc_finalize */
enum {
    DAS_SCORE_VOTE_ENTROPY = 0,
    DAS_SCORE_PRED_ENTROPY = 1,
    DAS_SCORE_BALD = 2,
    DAS_SCORE_CONFIDENCE = 3,
    DAS_SCORE_MARGIN = 4,
    DAS_SCORE_EXPECTED_ENTROPY = 5,
    DAS_N_SCORES = 6
};

typedef struct das_mc_desc {
    int32_t B;      /* images in the batch                                   */
    int32_t C;      /* classes, 2..DAS_MAX_CLASSES                           */
    int32_t H, W;   /* pixels                                                */
    int32_t T_cap;  /* passes the state can hold, 1..DAS_MAX_PASSES          */
    int32_t flags;  /* DAS_MC_VOTES | DAS_MC_PROBS [| DAS_MC_SINGLE_SHOT]    */
} das_mc_desc;

/* Size of the caller-allocated state for `desc` (256-byte aligned device buffer). Layout:
 * sum_p f32 [B,C,H,W] | sum_entropy f32 [B,H,W] | votes u8 [B,T_cap,H,W] | block partials. */
int das_mc_state_bytes(const das_mc_desc* desc, size_t* bytes);

/* Optional: zero the state. das_mc_accumulate(pass_begin = 0) initialises it anyway. */
int das_mc_reset(das_handle* h, const das_mc_desc* desc, void* state, void* stream);

/* K1.  Consume `n_passes` (1..DAS_MAX_PASS_GROUP) Monte-Carlo passes in ONE launch.
 * pass_logits: HOST array of n_passes DEVICE pointers, each f32 [B,C,H,W] logits of one stochastic
 * forward (the value of `model(image_batch)`, mc_dropout.py:40).  Every logit is read exactly once;
 * per pixel the kernel takes the first-max argmax (the vote), the max-subtracted softmax and its
 * entropy, and adds them to the running state.  Passes pass_begin .. pass_begin+n_passes-1 are
 * recorded; pass_begin == 0 (re)initialises the state.  n_passes == 1 is the pure streaming form;
 * a larger group trades resident logits for fewer state round trips. */
int das_mc_accumulate(das_handle* h, const das_mc_desc* desc, void* state, const float* const* pass_logits,
                      int n_passes, int pass_begin, void* stream);

/* K2.  Turn the state after T passes into per-pixel maps and per-image scores.
 * labels: f32 [B,H,W] or NULL; a pixel is valid iff 0 <= label < C (mc_dropout.py:45).  Invalid
 *   pixels get 0 in the entropy-type maps (mc_dropout.py:49, ceal.py:119) and 1 in confidence /
 *   margin (ceal.py:39,91).
 * maps (each f32 [B,H,W] or NULL = not wanted):
 *   vote_entropy   -sum_c p_c log2(p_c + 1e-12), p_c = votes_c / T         mc_dropout.py:46-48
 *   pred_entropy   same formula on p_bar = mean_t softmax(x_t)               ceal.py:116-118 when T = 1
 *   bald           pred_entropy - mean_t entropy(softmax(x_t))               (composed)
 *   confidence     max_c p_bar                                               ceal.py:36
 *   margin         largest - second largest p_bar                            ceal.py:84-90
 * weak_labels: u8 [B,H,W] or NULL: vote of pass 0, 255 where invalid         ceal.py:157-163
 * image_scores: f32 [B, DAS_N_SCORES] or NULL: mean over ALL H*W pixels of each map (a score whose
 *   accumulator is not in desc->flags is written as NaN).
 * Scores are reduced in a fixed order (deterministic, independent of how a pool is sharded). */
int das_mc_finalize(das_handle* h, const das_mc_desc* desc, void* state, const float* labels, int T,
                    float* vote_entropy, float* pred_entropy, float* bald, float* confidence,
                    float* margin, uint8_t* weak_labels, float* image_scores, void* stream);

/* K1+K2 fused for the LAST pass group of a batch: consumes passes pass_begin .. pass_begin+n_passes-1
 * like das_mc_accumulate, then finalises with T = pass_begin + n_passes like das_mc_finalize - but the
 * accumulators stay in registers, the state is not written back.  With pass_begin == 0 (the whole
 * Monte-Carlo stack of a batch in one group; required when DAS_MC_SINGLE_SHOT is set) no state is read
 * or written at all: HBM traffic is the logits, once, plus the requested outputs.  Same outputs,
 * same arithmetic and same fixed reduction order as the two-call form. */
int das_mc_accumulate_finalize(das_handle* h, const das_mc_desc* desc, void* state, const float* const* pass_logits,
                               int n_passes, int pass_begin, const float* labels, float* vote_entropy,
                               float* pred_entropy, float* bald, float* confidence, float* margin,
                               uint8_t* weak_labels, float* image_scores, void* stream);

/* K1+K2 fused WITH the network's final bilinear upsample.  The reference model ends with
 *     x = F.interpolate(low_res_x, size=input.size()[2:], mode='bilinear', align_corners=True)   models/deeplab.py:59
 * (also models/unet.py:58, models/fastscnn.py:22); the selectors then reduce x (mc_dropout.py:40, ceal.py:111).
 * This entry point takes `low_res_x` of every pass instead - f32 [B,C,h,w], HOST array of n_passes DEVICE
 * pointers - interpolates to desc->H x desc->W inside the kernel with ATen's align_corners indices / weights
 * (scale = float(h-1)/float(H-1), src = scale*dst, lerps as fma(l0, a, l1*b): bit-identical to ATen's vectorised
 * CPU kernel at DeepLab's shapes) and runs the same reduction as das_mc_accumulate_finalize(pass_begin = 0):
 * the full-resolution logits never exist in HBM (16x less traffic at stride 4).  All T = n_passes
 * (<= DAS_MAX_PASS_GROUP) passes arrive in this one call; `state` only provides the block partials (any desc
 * flags combination, DAS_MC_SINGLE_SHOT recommended).  Pointers need 4-byte alignment only.
 * DAS_ERR_UNSUPPORTED when a 16-pixel output tile would read more than 6 source rows / columns (upsampling
 * factors below ~3.75: query das_mc_upsample_supported first and use F.interpolate + das_mc_accumulate_finalize). */
int das_mc_upsample_accumulate_finalize(das_handle* hd, const das_mc_desc* desc, void* state,
                                        const float* const* pass_lowres_logits, int n_passes, int h, int w,
                                        const float* labels, float* vote_entropy, float* pred_entropy,
                                        float* bald, float* confidence, float* margin, uint8_t* weak_labels,
                                        float* image_scores, void* stream);
/* 1 if das_mc_upsample_accumulate_finalize handles the h x w -> H x W interpolation, else 0 (host only; hd may be NULL:
 * the default options are assumed). */
int das_mc_upsample_supported(const das_handle* hd, int h, int w, int H, int W);
/* Which kernel das_mc_upsample_accumulate_finalize would launch for `desc` (B, C, H, W, flags) and an h x w source
 * (host only; hd may be NULL: the default options): 0 = unsupported shape, 4 | 15 = pixel pairs per lane with that many
 * consumer warps, 220 | 216 = one pixel per lane with 20 | 16 consumer + 4 producer warps (DAS_OPT_MC_UP_WARPS). */
int das_mc_upsample_variant(const das_handle* hd, const das_mc_desc* desc, int h, int w);

/* Device pointer to the recorded votes, u8 [B,T_cap,H,W] (test / debugging aid). */
int das_mc_votes_ptr(const das_mc_desc* desc, void* state, uint8_t** votes);

/* ------------------------------------------------------------------------------------------
 * Region scoring
 * ------------------------------------------------------------------------------------------ */

/* ActiveSelectionMCDropout.suppress_labeled_entropy (mc_dropout.py:110-121):
 * zero maps[i, r:r+h, c:c+w] for each of the n records (i, r, c, h, w) in `rects` (device, int32). */
int das_suppress_rects(das_handle* h, float* maps, int B, int H, int W, const int32_t* rects, int n, void* stream);
/* The same with the records in HOST memory (the caller's Python list of labelled regions, region_cityscapes.py:54-61):
 * they travel in the kernel parameters, 128 per launch - no host-to-device copy, nothing to keep alive after the call. */
int das_suppress_rects_host(das_handle* h, float* maps, int B, int H, int W, const int32_t* host_rects, int n,
                            void* stream);

/* Input perturbation of the MC-noise selectors on the device:
 *     noise = np.random.normal(0, 0.125, image_batch.shape); model(image_batch + noise)           mc_noise.py:26-27
 * out[i] = x[i] + sigma * z[i], z ~ N(0,1) from Philox4x32-10 + Box-Muller, a pure function of (seed, stream_id, i):
 * the same triple reproduces the same noise, different stream ids (one per pass) are independent streams.  One pass
 * over the data (HBM bound); out may alias x. */
int das_add_gaussian_noise(das_handle* h, const float* x, size_t n, float sigma, uint64_t seed, uint64_t stream_id,
                           float* out, void* stream);

/* a += b elementwise (combined noise + dropout vote entropy, mc_noise.py:141,165) */
int das_add_maps(das_handle* h, float* a, const float* b, size_t n, void* stream);

/* Stride-1 'valid' RxR box sum, the conv2d-with-ones of mc_dropout.py:148-149:
 * out[b,r,c] = sum_{i<R,j<R} maps[b,r+i,c+j], out is f32 [B,H-R+1,W-R+1].  Sums are formed in
 * fp64 sliding windows and rounded once.  minmax: device f32[2] = {min,max}; it is COMBINED with the
 * values already there (initialise with das_minmax_init), so a pool can be processed in batches
 * (mc_dropout.py:152-153).  workspace: das_box_sum_workspace_bytes(). */
int das_box_sum_workspace_bytes(int B, int H, int W, int R, size_t* bytes);
int das_minmax_init(das_handle* h, float* minmax, void* stream);
int das_box_sum(das_handle* h, const float* maps, int B, int H, int W, int R, float* out, float* minmax,
                void* workspace, void* stream);

/* x = (x + (-min)) * (1 / (max - min)) in float32, exactly as mc_dropout.py:154-155. */
int das_minmax_normalise(das_handle* h, float* score_maps, size_t n, const float* minmax, void* stream);

/* Per-image greedy NMS pick sequences - the image-local part of
 * ActiveSelectionMCDropout.square_nms (mc_dropout.py:82-108).  For each of the N images (one CTA
 * each): repeat { first flat argmax (r,c); record; zero [r-R,r+R) x [c-R,c+R) } while picks < kmax
 * and (it is the first pick or the map max >= stop).  MUTATES score_maps like the reference.
 * cand_score f32 [N,kmax] (unused slots = -inf), cand_rc int32 [N,kmax,2], cand_count int32 [N],
 * cand_flat int64 [N,kmax] or NULL: ((image_offset + i) * H2 + r) * W2 + c, the flat index of the pick in the
 *   un-sharded pool (-1 in unused slots).
 * Within an image the picks are sorted by (score desc, flat index asc), so the reference's global greedy loop is
 * exactly: das_topk over the flattened cand_score table (ties keep table order = flat order) and the stop rule
 * of mc_dropout.py:87,105 on that prefix - see active_selection/base.py:region_tail in the Python mirror. */
int das_nms_sequences(das_handle* h, float* score_maps, int N, int H2, int W2, int R, int kmax, float stop,
                      float* cand_score, int32_t* cand_rc, int32_t* cand_count, long long image_offset,
                      int64_t* cand_flat, void* stream);

/* ------------------------------------------------------------------------------------------
 * Accuracy-predictor selectors (active_selection/accuracy.py)
 * One pass over logits f32 [B,C,H,W] (the segmentation output, or the C = 2 error-predictor head) and the
 * float32 labels.  A pixel is valid iff 0 <= label < num_classes (accuracy.py:31,56,116); labels == NULL:
 * every pixel valid and wrong_count = 0.  image_scores f32 [B, DAS_ACC_N]:
 *   DAS_ACC_WRONG_COUNT     sum_valid [label != argmax_c logits]                accuracy.py:30-33
 *   DAS_ACC_P0_SUM          sum_valid softmax(logits)[0]                        accuracy.py:55-58
 *   DAS_ACC_NOT_ARGMAX_SUM  sum_valid (1 - argmax_c logits)                     accuracy.py:60-64
 *   DAS_ACC_UNSURE_MEAN     mean_valid (4 p1 - 4 p1^2), p1 = softmax[1]         accuracy.py:117-118 (NaN if none valid)
 *   DAS_ACC_VALID_COUNT     number of valid pixels
 * p0_map f32 [B,H,W] or NULL: softmax[0] with invalid pixels = 0 (accuracy.py:159-162), the input of the region
 * tail (das_suppress_rects -> das_box_sum -> das_minmax_normalise -> das_nms_sequences).
 * ------------------------------------------------------------------------------------------ */
enum {
    DAS_ACC_WRONG_COUNT = 0,
    DAS_ACC_P0_SUM = 1,
    DAS_ACC_NOT_ARGMAX_SUM = 2,
    DAS_ACC_UNSURE_MEAN = 3,
    DAS_ACC_VALID_COUNT = 4,
    DAS_ACC_N = 5
};
int das_accuracy_workspace_bytes(int B, int H, int W, size_t* bytes);
int das_accuracy_scores(das_handle* h, const float* logits, int B, int C, int H, int W, const float* labels, int num_classes,
                        float* p0_map, float* image_scores, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Max-subset representativeness (active_selection/max_subset.py:17-39)
 * Greedy facility location: D[n,i] = ||x_n - y_i|| (float64 arithmetic; stored as float32 for float32 features,
 * as sklearn.pairwise_distances returns them), then k times: the FIRST unselected candidate i minimising
 * sum_n min(min_d[n], D[n,i]) (the reference maximises the negated sum with a strict '>'), min_d = min(min_d, D[:,i]).
 * X [N,D] pool features, Y [M,D] candidate features, both float32 (is_f64 = 0) or float64 (1), device, row major.
 * picks: device int32 [k] candidate indices in pick order (-1 once every candidate is taken).
 * ------------------------------------------------------------------------------------------ */
int das_maxsubset_workspace_bytes(int N, int M, int D, int is_f64, size_t* bytes);
int das_maxsubset_greedy(das_handle* h, const void* X, const void* Y, int N, int M, int D, int is_f64, int k, int32_t* picks,
                         void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ranking  (K3)
 * Replaces `sorted(zip(scores, images), key=score, reverse=...)[:k]`
 * (mc_dropout.py:195, ceal.py:69,97,130, mc_noise.py:59,128,147): a STABLE sort - equal scores keep
 * input order - realised as a block radix select + sort on the composite key (score, position).
 * ids: int64 [n] payload or NULL (then the position itself).  k is clamped to n.
 * ------------------------------------------------------------------------------------------ */
#define DAS_TOPK_MAX_K 4096
int das_topk_workspace_bytes(int n, int k, size_t* bytes);
int das_topk(das_handle* h, const float* scores, const int64_t* ids, int n, int k, int descending,
             float* out_scores, int64_t* out_ids, void* workspace, void* stream);

/* Multi-GPU ranking: each rank's best candidates travel as RECORDS - int64 pairs
 * {float32 score bits in the low 32 bits, id} - so that ONE all-gather of a [k_slots, 2] int64 block per rank carries
 * them, and the merge runs on the device of every rank.
 * das_topk_records: like das_topk, but writes exactly k_slots records {score, ids[pos] + id_offset} (ids == NULL: the
 *   position + id_offset = the global image index of a contiguous shard); slots beyond min(k_slots, n) are padding
 *   records {-inf (descending) / +inf (ascending), -1}.  n == 0 (empty shard) writes padding only.
 * das_topk_merge: the first k_slots of the stable ranking of a gathered record table (n_records = ranks * k_slots
 *   records in rank order).  Ties keep table order = (rank, local order) = global index order for contiguous shards,
 *   i.e. exactly the stable sort of the un-sharded pool; padding records (id < 0) rank below every candidate. */
int das_topk_records(das_handle* h, const float* scores, const int64_t* ids, int n, int k_slots, int descending,
                     long long id_offset, int64_t* records, void* stream);
int das_topk_merge(das_handle* h, const int64_t* records, int n_records, int k_slots, int descending,
                   int64_t* out_records, void* stream);

/* ------------------------------------------------------------------------------------------
 * Core-set k-center greedy  (K4)
 * Replaces ActiveSelectionCoreSet._updated_distances / _select_batch (core_set.py:17-38):
 * sklearn float64 euclidean distances + numpy argmax / minimum.
 * feats: f32 [N,D] (the reference stores float32 network outputs widened to float64,
 * core_set.py:50,63, so float32 holds the same values exactly).  Distances are accumulated in
 * fp64 from exact float32 differences.  Rows [row_begin,row_end) are this rank's shard.
 * ------------------------------------------------------------------------------------------ */

/* Tensor-core distance filter (optional, K4 tcgen05 path).
 * das_kcenter_filter_build makes, for the row shard [row_begin,row_end) of `feats`, a device blob holding
 *   - a bf16 copy of the features and their exact float64 squared norms,
 *   - dt[c, i] ~ ||f_i - f_c||^2 for EVERY candidate centre c in [0,N) and every shard row i, computed by one
 *     bf16 tcgen05 GEMM (TMA-fed, FP32 accumulation in TMEM) fused with the |a|^2 + |b|^2 - 2ab epilogue.
 * |dt - exact| <= 2^-7 (|f_i|^2 + |f_c|^2) is guaranteed (bf16 rounding + fp32 accumulation, doubled), so
 * das_kcenter_init / _step / _greedy use dt only to SKIP rows whose float64 minimum provably cannot change
 * and re-evaluate the others exactly: min_d2 and the picks are bit-identical with and without a filter.
 * The blob (1024-byte aligned, das_kcenter_filter_bytes) is N*rows*4 bytes + O(N*D): 0.44 GB for N = 10 000.
 * Pass filter = NULL to the functions below for the plain float64 path (any N). */
int das_kcenter_filter_bytes(int N, int D, int rows, size_t* bytes);
int das_kcenter_filter_build(das_handle* h, const float* feats, int N, int D, int row_begin, int row_end, void* filter,
                             void* stream);
/* HOST out: stats2[0] = exact float64 row evaluations so far, stats2[1] = rows screened (synchronises `stream`) */
int das_kcenter_filter_stats(das_handle* h, const void* filter, int N, int D, int rows, unsigned long long* stats2, void* stream);

/* min_d2[i - row_begin] = min_l ||f_i - f_centers[l]||^2 (fp64) for the L initial centres
 * (core_set.py:19,32-36); also writes the packed argmax key of the shard to key2:
 * key2[0] = float64 bits of max min_d2 (order preserving since d2 >= 0), key2[1] = its row index, ties ->
 * lowest i (np.argmax, core_set.py:22). */
int das_kcenter_init(das_handle* h, const float* feats, int N, int D, int row_begin, int row_end,
                     const int32_t* centers, int L, double* min_d2, unsigned long long* key2,
                     const void* filter, void* stream);

/* One greedy step: centre = the row index held in *centre_idx (device int32); for every shard row
 * min_d2 = min(min_d2, ||f_i - f_centre||^2) (core_set.py:26,37-38) and the new shard argmax is
 * written to key2[0] (bits of the max), key2[1] (row index). */
int das_kcenter_step(das_handle* h, const float* feats, int N, int D, int row_begin, int row_end,
                     const int32_t* centre_idx, double* min_d2, unsigned long long* key2,
                     const void* filter, void* stream);

/* Whole single-GPU greedy loop, no host round trip per step: picks int32 [K], min_d f64 [N]
 * (final euclidean min-distances, i.e. sqrt).  workspace: das_kcenter_workspace_bytes().
 * filter: blob built with row_begin = 0, row_end = N, or NULL. */
int das_kcenter_workspace_bytes(const das_handle* h, int N, int D, size_t* bytes);
int das_kcenter_greedy(das_handle* h, const float* feats, int N, int D, const int32_t* centers, int L, int K,
                       int32_t* picks, double* min_d, void* workspace, const void* filter, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAS_B200_H */
