/*
 * moc_b200.h - C ABI of the B200 (sm_100a) MOC per-slide hot path.
 *
 * The reference (xmed-lab/MOC) is pure Python/PyTorch and has no FFI of its
 * own; its "operator interface" for this path is a handful of Python functions
 * in main_moc.py and utils/patch_selection_classifier*.py.  Each entry point
 * below names the reference code it replaces (file:line relative to the
 * reference checkout).  The Python mirror of those functions lives in the
 * moc_b200 package and calls this library through ctypes; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", no exceptions, no torch types.  Every function returns
 *    MOC_OK (0) or a negative MOC_E_* code; moc_last_error() gives the text.
 *  - All pointers are DEVICE pointers unless the name ends in _h.
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*;
 *    from PyTorch: torch.cuda.current_stream().cuda_stream).  No hidden
 *    synchronisation, no hidden allocation: the caller owns every buffer,
 *    including workspaces sized by the *_workspace_bytes() queries.
 *  - Bags ("slides") live in one ragged row-major fp32 buffer feat[total_rows][512];
 *    slide i owns rows offsets[i] .. offsets[i+1]-1.
 *  - Keys are stored as planes (structure of arrays): plane p, row r is
 *    keys[p * key_stride + r].  For C classes the planes are
 *       [0,C)    L[:,c]            patch-prompt similarity        (main_moc.py:336)
 *       [C,2C)   softmax(L)[:,c]   row softmax over classes       (_index.py:34)
 *       2C       |top1 - top2| of the row of L                    (_index.py:46-48)
 *       2C+1     sum of background similarities  Le[:, C:]        (_index.py:71-75)
 *       2C+2     max of background similarities  Le[:, C:]        (main_moc.py:365)
 *    From MOC_KEYS_COMPACT_MIN_CLASSES classes on (EBRAINS-30) the C softmax planes are not stored: the scoring
 *    kernels write C+4 planes
 *       [0,C)    L[:,c]
 *       C        lse = log sum_c exp(L[:,c])      (= max + log sum exp(L - max))
 *       C+1      |top1 - top2|                    C+2   sum of background      C+3   max of background
 *    and every consumer (gate / combination, backward, ablation, gather) forms softmax(L)[:,c] = expf(L[:,c] - lse)
 *    on the fly (within 3e-7 relative of the stored form), while the softmax SELECTION ranks a slide's patches by
 *    L[:,c] - lse, the logarithm of the same quantity: 136 instead of 252 bytes written per patch for 30 classes,
 *    and two plane reads without an exponential per key in the selection.  moc_num_key_planes() and moc_key_plane()
 *    give the counts and indices for a class count; moc_expand_keys() materialises the full 2C+3-plane layout from
 *    either (what the stand-alone helpers and tests index directly).  moc_row_keys always writes the full layout.
 */
#ifndef MOC_B200_H
#define MOC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOC_FEAT_DIM 512   /* CONCH embedding width                          */
#define MOC_HIDDEN 64      /* senet hidden width            main_moc.py:302  */
#define MOC_GATES 4        /* one gate per classifier       main_moc.py:315  */
#define MOC_MAX_COLS 64    /* C + number of background prompts, this build   */
#define MOC_BANK_MAX_CLASSES 8   /* un-collapsed prompt bank: classes ...        */
#define MOC_BANK_MAX_COLS 256    /* ... and bank prompts + background prompts    */
#define MOC_KEYS_COMPACT_MIN_CLASSES 9   /* from here on keys are C+4 planes (above) */

#define MOC_OK 0
#define MOC_E_ARG (-1)        /* null pointer / negative size / bad flag      */
#define MOC_E_SHAPE (-2)      /* shape outside what this build supports       */
#define MOC_E_WORKSPACE (-3)  /* workspace too small                          */
#define MOC_E_CUDA (-4)       /* a CUDA runtime call failed                   */

/* discard / active bit masks: bit m set = classifier m is discarded / active.
 * Order follows --discard_classifiers (main_moc.py:39). */
#define MOC_CLS_TOPK 1u
#define MOC_CLS_DELTA_SOFTMAX 2u
#define MOC_CLS_DELTA_DIFF 4u
#define MOC_CLS_BOTTOMK 8u
#define MOC_CLS_ALL 15u
/* OR-ed into active_mask (moc_head_forward) / flags (moc_senet_forward): run the gate MLP on the 3xTF32 kernel, which
 * has no |x| limit, whatever the class count.  Callers set it to redo a pass whose domain flag came back non-zero. */
#define MOC_HEAD_WIDE_DOMAIN 0x100u

const char* moc_last_error(void);
int moc_version(void);
/* number of key planes the scoring kernels write for C classes: 2C+3, or C+4 from MOC_KEYS_COMPACT_MIN_CLASSES on */
int moc_num_key_planes(int n_classes);
/* index of a named plane in that layout: which = MOC_PLANE_*; -1 when the layout does not store it (SOFTMAX0 in the
 * compact layout, LSE in the full one).  SOFTMAX0 is the softmax plane of class 0 (class c: + c). */
#define MOC_PLANE_TOP0 0
#define MOC_PLANE_SOFTMAX0 1
#define MOC_PLANE_DIFF 2
#define MOC_PLANE_BG_SUM 3
#define MOC_PLANE_BG_MAX 4
#define MOC_PLANE_LSE 5
int moc_key_plane(int n_classes, int which);
/* keys (either layout, n_rows rows) -> the full 2C+3-plane layout in `full` (stride full_stride); a plain copy when
 * the class count already uses it */
int moc_expand_keys(const float* keys, int64_t key_stride, int n_classes, int64_t n_rows, float* full,
                    int64_t full_stride, void* stream);

/* ---- prompt matrices ------------------------------------------------------
 * Packs W [512,C] and the background columns of W_ext [512,C_ext] (both
 * row-major as torch.stack(..., dim=1) leaves them, utils/zeroshot_utils.py:50)
 * into the K-major layout the scoring kernel keeps on chip:
 * packed[col][512], col < C from W, then the C_ext-C background columns of
 * W_ext, zero-padded to moc_packed_cols() columns.  The first C columns of
 * W_ext never influence slide_process (they only re-order rows inside the
 * bottom-k set, _index.py:82-86) and are not packed. */
int moc_packed_cols(int n_classes, int n_ext);
int moc_pack_prompts(const float* w, int n_classes, const float* w_ext, int n_ext,
                     float* packed, void* stream);

/* A prompt BANK (several prompt embeddings per class: classnames x templates)
 * collapses to one unit column per class exactly as the reference does it
 * offline (utils/zeroshot_utils.py:29-50): every prompt row is L2-normalised,
 * the rows of a class are averaged, the mean is divided by its norm.
 * bank [n_prompts][512] row-major; class c owns rows class_offsets[c] ..
 * class_offsets[c+1]-1 (device int32 [n_classes+1]); w_out [512][n_classes]
 * row-major, i.e. directly usable as zeroshot_weights(_ext).  Scoring against
 * the collapsed column equals the mean of the per-prompt scores up to that
 * one scale, so a bank of any size costs nothing on the streaming path. */
int moc_collapse_prompt_bank(const float* bank, const int32_t* class_offsets, int n_classes, float* w_out,
                             void* stream);

/* The same bank kept UN-COLLAPSED on the scoring path (BASELINE.json configs[2], SURVEY.md section 8d: the dense
 * contraction stress case): every prompt is a column of a tensor-core contraction (tcgen05 FP16x3, fp32-accurate),
 * and the kernel's epilogue takes the per-class sum and rescales it by 1 / (P_c * ||mean_p normalise(w_p)||), which is
 * exactly `feat @ w_out[:, c]` for the w_out moc_collapse_prompt_bank builds (utils/zeroshot_utils.py:38-44), so the
 * keys equal moc_score_keys' on the collapsed matrix within the usual tolerance.
 * bank / class_offsets as above (n_prompts = class_offsets[n_classes]); bg [n_bg][512] row-major are the background
 * prompts (the columns C.. of zeroshot_weights_ext, transposed), used as given.  2 <= n_classes <=
 * MOC_BANK_MAX_CLASSES, n_prompts + n_bg <= MOC_BANK_MAX_COLS.  image: moc_prompt_bank_tc_bytes() bytes, 16-byte
 * aligned, built once per bank; an int32 flag at moc_prompt_bank_tc_flag_offset() is set when a score came out
 * non-finite (|x| >= 65504 or non-finite features), like moc_prompts_tc_flag_offset.  keys as moc_score_keys. */
size_t moc_prompt_bank_tc_bytes(int n_prompts, int n_bg);
size_t moc_prompt_bank_tc_flag_offset(int n_prompts, int n_bg);
int moc_prepare_prompt_bank_tc(const float* bank, const int32_t* class_offsets, int n_classes, int n_prompts,
                               const float* bg, int n_bg, void* image, size_t image_bytes, void* stream);
int moc_score_keys_bank_tc(const float* feat, int64_t n_rows, const void* image, int n_classes, int n_prompts,
                           int n_bg, int normalize, float* keys, int64_t key_stride, void* stream);

/* ---- a2 + selection keys: the streaming kernel ----------------------------
 * Replaces `feat @ zeroshot_weights`, `feat @ zeroshot_weights_ext`
 * (main_moc.py:336-337) and the per-row arithmetic of the four selectors
 * (_index.py:34, :46-48, :71-75) and of the score planes (main_moc.py:359-366).
 * Reads every row of feat exactly once (2048 B / patch).  normalize != 0
 * L2-normalises each row first (models/model_adapters.py:188; off in MOC). */
int moc_score_keys(const float* feat, int64_t n_rows, const float* packed, int n_classes, int n_ext,
                   int normalize, float* keys, int64_t key_stride, void* stream);
/* Same, on at most max_ctas persistent CTAs (0 = the default: 132 on a 148-SM B200, where the streaming kernel reads
 * HBM fastest: 6.85 TB/s against 6.44 TB/s with one CTA on every SM).  A smaller number leaves SMs free for kernels
 * running on another stream; the best count differs slightly between individual GPUs (132 on most, 136 on some), which
 * MocEngine calibrates once per process.  Ignored by the wide-prompt-set kernels. */
int moc_score_keys_ex(const float* feat, int64_t n_rows, const float* packed, int n_classes, int n_ext,
                      int normalize, float* keys, int64_t key_stride, int max_ctas, void* stream);

/* ---- a2 for wide prompt sets: the same contract on the tensor cores ---------
 * With more than ~10 prompt columns (EBRAINS-30: 30 classes + 4 background)
 * main_moc.py:336-337 is a dense contraction that CUDA-core FMAs cannot keep
 * HBM-bound.  moc_score_keys_tc streams the patches through tcgen05.mma
 * (FP16 x 3-product split with fp32 accumulation in TMEM: fp32-level accuracy,
 * |x| < 65504 required) against a resident image of the prompts that
 * moc_prepare_prompts_tc builds once from the packed matrix (no host
 * synchronisation: the prompt scale is chosen on the device).  Works for any
 * C_ext <= MOC_MAX_COLS; moc_score_keys (CUDA cores, prompts in registers) is
 * the faster one for C_ext <= 8.  If a score comes out non-finite (non-finite
 * or out-of-range input) the int at moc_prompts_tc_flag_offset() inside the
 * image is set to 1; the caller may read it back whenever it synchronises. */
size_t moc_prompts_tc_bytes(int n_classes, int n_ext);
size_t moc_prompts_tc_flag_offset(int n_classes, int n_ext);
int moc_prepare_prompts_tc(const float* packed, int n_classes, int n_ext, void* prompts_tc,
                           size_t prompts_tc_bytes, void* stream);
int moc_score_keys_tc(const float* feat, int64_t n_rows, const void* prompts_tc, int n_classes, int n_ext,
                      int normalize, float* keys, int64_t key_stride, void* stream);

/* ---- a3..a7: four top-J selections, union, ascending compaction -----------
 * Replaces index_{topj,delta_softmax,delta_diff,bottomk_irrel}_classifier
 * (_index.py:17-87) plus the set union / sort of main_moc.py:341-354, for
 * n_slides bags at once.  row_mask (nullable, one byte per row, 0 = dropped)
 * plays the role of the random half mask (main_moc.py:329-331): dropped rows do
 * not exist for the selection, maxj = min(topj, kept rows), and sel_local holds
 * indices into the *masked* bag exactly like the reference's selected_index.
 * Output region of slide i is [sel_base[i], sel_base[i+1]) (caller-computed
 * prefix sums of moc_select_capacity(); n_slides+1 entries); sel_rows are
 * absolute rows of feat, ascending, and the unused tail of a region is -1. */
int64_t moc_select_capacity(int64_t n_rows_of_slide, int n_classes, int topj);
size_t moc_select_workspace_bytes(int64_t total_rows, int n_slides);
int moc_select_union(const float* keys, int64_t key_stride, const int64_t* offsets, int n_slides,
                     int64_t total_rows, int n_classes, int topj, unsigned discard_mask, const uint8_t* row_mask,
                     const int64_t* sel_base, int32_t* sel_rows, int32_t* sel_local, int32_t* sel_count,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- stand-alone sorted top-J (the selectors' public return value) --------
 * values[i * ld] for i < n; writes the indices of the min(j,n) largest (or
 * smallest) values in descending (ascending) value order, ties by lower index,
 * as int64 to idx_out[r * out_ld].  n_cols independent columns per call:
 * column c reads values + c * col_stride and writes idx_out + c.
 * Replaces Tensor.topk(maxj, 0, largest, True) in _index.py:25,35,50,81,85. */
int moc_topj_sorted(const float* values, int64_t n, int64_t ld, int n_cols, int64_t col_stride,
                    int j, int largest, int64_t* idx_out, int64_t out_ld, float* val_out, void* stream);

/* ---- helpers behind the stand-alone selector / pooling functions ------------
 * These take the [N,Ct] row-major logits the reference's functions are called
 * with (columns < n_fg are classes, the rest background) instead of features. */
int moc_row_keys(const float* logits, int64_t n, int64_t ld, int n_fg, int n_total, float* keys,
                 int64_t key_stride, void* stream);
/* dst[r][c] = src[idx[r]][c] for c < n_cols  (logits[indices] in the pooling functions) */
int moc_take_rows(const float* src, int64_t ld, const int64_t* idx, int64_t n_idx, int n_cols, float* dst,
                  void* stream);
/* out[c] = mean of vals[0..j-1][c]  (values[:min(j,maxj)].mean(dim=0), patch_selection_classifier.py:27) */
int moc_col_prefix_mean(const float* vals, int64_t ld, int n_cols, int j, float* out, void* stream);

/* ---- a8..a11: head forward -------------------------------------------------
 * For every selected row: gather the 512-vector, senet gate
 * sigmoid(W2 relu(W1 x + b1) + b2) (main_moc.py:299-312,:390), the gated sum of
 * the four score planes (main_moc.py:391-403 / :482-492, planes chosen by
 * active_mask), then per (slide, class) the mean of the min(topk, S) largest
 * (topj_pooling, utils/patch_selection_classifier.py:18-32).
 * The first layer runs on the tensor cores, fp32-accurate through operand
 * splitting: FP16x3 (tcgen05 kind::f16, W1 resident in shared memory; features
 * must satisfy |x| < 4094, beyond that the gates come out non-finite) for up to
 * 8 classes, 3xTF32 (kind::tf32, no range limit) for wider class sets; the
 * workspace (moc_head_forward_workspace_bytes(), 16-byte aligned) holds W1 split
 * and swizzled for them.  MOC_HEAD_IMPL=f16|tf32|simt in the environment forces
 * one kernel (simt = the CUDA-core cross-check).
 * gate [S_total,4], final [S_total,C], bag_logits [n_slides,C],
 * pool_pos [n_slides,C,topk] (positions inside the slide's selected list,
 * -1 padded), all indexed through sel_base/sel_count. gate may be null. */
size_t moc_head_forward_workspace_bytes(void);
/* Byte offset, inside that workspace, of an int32 "domain flag": the FP16x3 gate kernel ORs 1 into it when a gate
 * pre-activation came out non-finite (a feature with |x| >= 4094, or a non-finite one).  It is never cleared by the
 * library: the caller zeroes it before a pass and, at a point where it synchronises anyway, reads it and repeats the
 * pass with MOC_HEAD_WIDE_DOMAIN when it is set.  The reference (fp32 matmul, main_moc.py:303) is finite for any
 * finite feature, so the repeat restores parity for inputs outside the fast kernel's range. */
size_t moc_head_domain_flag_offset(void);
int moc_head_forward(const float* feat, const float* keys, int64_t key_stride, int n_classes,
                     const int64_t* sel_base, const int32_t* sel_rows, const int32_t* sel_count,
                     int n_slides, int64_t sel_capacity_total,
                     const float* w1, const float* b1, const float* w2, const float* b2,
                     unsigned active_mask, int topk,
                     float* gate, float* final_scores, float* bag_logits, int32_t* pool_pos,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ablation_evaluation (main_moc.py:523-582): un-gated avg (mode 0) / sum (1) / max (2) of the four planes
 * of every selected row, then the same top-K pooling. */
int moc_ablation_forward(const float* keys, int64_t key_stride, int n_classes, const int64_t* sel_base,
                         const int32_t* sel_rows, const int32_t* sel_count, int n_slides,
                         int64_t sel_capacity_total, int mode, int topk, float* final_scores,
                         float* bag_logits, void* stream);

/* per-(slide,class) top-K mean of one plane selected by another plane:
 * zs_evaluation's pooling (main_moc.py:427-432) with
 * select plane == value plane (topj_pooling) or softmax / delta planes. */
int moc_pool_topk(const float* keys, int64_t key_stride, const int64_t* offsets, int n_slides,
                  int n_classes, int topk, int select_plane0, int select_plane_step, int select_smallest,
                  int value_plane0, int value_plane_step, float* bag_logits, void* stream);

/* selected_feat = feat[selected_index] and the four [S,C] score planes as dense
 * row-major tensors (main_moc.py:355-366), for callers that want the
 * reference's slide_process() return value.  out_feat or the planes may be null. */
int moc_gather_selected(const float* feat, const float* keys, int64_t key_stride, int n_classes,
                        const int32_t* sel_rows, int64_t n_sel, float* out_feat, float* plane_top,
                        float* plane_softmax, float* plane_diff, float* plane_bg, void* stream);

/* senet as a stand-alone module on a dense [n_rows,512] input (main_moc.py:299-312):
 * gate = sigmoid(W2 relu(W1 x + b1) + b2) and, for autograd, the parameter
 * gradient given d(loss)/d(gate) [n_rows,4] (rows whose gradient is all-zero are skipped). */
int moc_senet_forward(const float* x, int64_t n_rows, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* gate, unsigned flags, void* workspace, size_t workspace_bytes,
                      void* stream);
size_t moc_senet_backward_workspace_bytes(int64_t n_rows);
int moc_senet_backward(const float* x, int64_t n_rows, const float* dgate, const float* w1, const float* b1,
                       const float* w2, const float* b2, float* grads, void* workspace, size_t workspace_bytes,
                       void* stream);

/* ---- dense per-patch layers of the secondary MIL heads (tensor cores) -------
 * y[n][m] = act( sum_k x[n][k] * w[m][k] + bias[m] ) for every patch n: one
 * nn.Linear (+ activation) applied to a whole bag, the building block of
 * Conch_CLIP_Ada.adapter (models/model_adapters.py:152-157), CLAM_SB / ABMIL
 * (models/model_clam.py:83-91, :41-64) and MIL_fc (models/model_mil.py:17-23).
 * w is nn.Linear.weight, [n_out][k] row-major; k a multiple of 32; bias may be
 * null.  Columns < split use act0, the others act1 (two layers that share
 * their input, e.g. the tanh and sigmoid branches of the gated attention,
 * run as one call on stacked weights).  tcgen05 3xTF32: ~1e-5 relative. */
#define MOC_ACT_NONE 0
#define MOC_ACT_RELU 1
#define MOC_ACT_TANH 2
#define MOC_ACT_SIGMOID 3
size_t moc_linear_workspace_bytes(int n_out, int k);
int moc_linear_forward(const float* x, int64_t ldx, int64_t n_rows, int k, const float* w, const float* bias,
                       int n_out, int act0, int split, int act1, float* y, int64_t ldy, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Weight gradient of such a layer (autograd of y = x W^T): dw[m][j] (+)= sum_n g[n][m] * x[n][j]
 * with g = d(loss)/dy [n_rows][n_out] - the contraction runs over the patches of the bag.
 * tcgen05 3xTF32, split over the bag with a fixed-order (deterministic) reduction.
 * n_out and k multiples of 4; accumulate != 0 adds to dw instead of overwriting it. */
size_t moc_linear_wgrad_workspace_bytes(int64_t n_rows, int n_out, int k);
int moc_linear_wgrad(const float* g, int64_t ldg, int n_out, const float* x, int64_t ldx, int k, int64_t n_rows,
                     float* dw, int64_t lddw, int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Conch_CLIP_Ada after its adapter MLP (models/model_adapters.py:186-190): f =
 * adapted * clip_ratio + x * (1 - clip_ratio), f /= |f|, logits = f @ classifier
 * ([512][C] row-major).  adapted == NULL gives forward_disable_ada (:211-214):
 * normalise, score.  logits are written as planes logits[c * ld + n] so that
 * moc_pool_topk finishes the per-class top-j mean (:173-183). */
int moc_adapter_scores(const float* x, const float* adapted, float clip_ratio, const float* classifier,
                       int n_classes, int64_t n_rows, float* logits, int64_t ld, void* stream);

/* Attn_Net_Gated (models/model_clam.py:59-62) after the two branch layers:
 * a_raw[n] = sum_d ab[n][d] * ab[n][hidden + d] * wc[d] + bc, ab = [tanh branch |
 * sigmoid branch] as one moc_linear_forward on stacked weights writes it. */
int moc_gated_attention_scores(const float* ab, int64_t ld, int hidden, const float* wc, float bc, int64_t n_rows,
                               float* a_raw, void* stream);

/* CLAM_SB.forward_single (models/model_clam.py:181, :209-212): softmax of a_raw
 * over the bag, pooled = softmax(a_raw) h, logits = w_cls pooled + b_cls
 * (w_cls [C][width] row-major), probs = softmax(logits), y_hat = argmax.
 * pooled / probs / y_hat may be null.  Deterministic two-stage reduction. */
size_t moc_attention_pool_workspace_bytes(int64_t n_rows, int width);
int moc_attention_pool(const float* a_raw, const float* h, int64_t ldh, int width, int64_t n_rows,
                       const float* w_cls, const float* b_cls, int n_classes, float* pooled, float* logits,
                       float* probs, int32_t* y_hat, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of ABMIL = CLAM_SB(instance_loss_fn=None).forward_single for one bag, from
 * dlogits = d(loss)/d(logits) [C]  (what loss.backward() computes through autograd in
 * utils/core_utils.py:398-414 for models/model_clam.py:175-212).  Inputs are the forward's
 * activations: h1 = relu(fc x) [N][width], ab = [tanh branch | sigmoid branch] [N][2*hidden],
 * a_raw [N], pooled [width]; w_ab = [attention_a.weight; attention_b.weight] stacked
 * [2*hidden][width], wc = attention_c.weight [hidden], w_cls [C][width].
 * Outputs (written, not accumulated): d_wfc [width][k_in], d_bfc [width], d_wab [2*hidden][width],
 * d_bab [2*hidden], d_wc [hidden], d_bc [1], d_wcls [C][width], d_bcls [C].
 * width a multiple of 128, hidden a multiple of 16 (<= 512); workspace 256-byte aligned. */
size_t moc_abmil_backward_workspace_bytes(int64_t n_rows, int k_in, int width, int hidden);
int moc_abmil_backward(const float* x, int64_t ldx, int k_in, int64_t n_rows, const float* h1, int64_t ldh,
                       int width, const float* ab, int64_t ldab, int hidden, const float* a_raw,
                       const float* pooled, const float* w_ab, const float* wc, const float* w_cls,
                       int n_classes, const float* dlogits, float* d_wfc, float* d_bfc, float* d_wab,
                       float* d_bab, float* d_wc, float* d_bc, float* d_wcls, float* d_bcls, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Backward pieces of the two heads whose gradient lives on a few rows only.
 * Conch_CLIP_Ada.forward (models/model_adapters.py:185-193): for R "virtual rows" (a row pooled into the top-j mean
 * of class cls_of_row[v], carrying g_of_row[v] = d(loss)/d(its logit)), x_rows / a2_rows [R][512] the gathered
 * features and adapter outputs: da2 [R][512] = gradient at the adapter output (through the blend, the
 * normalisation and the adapter's last ReLU).  The adapter's weight gradients then are moc_linear_wgrad of da2 and
 * of moc_mask_positive(da2 W2, a1).  moc_transpose: out[c][r] = in[r][c].
 * MIL_fc.forward (models/model_mil.py:30-51): gradients of the two Linear layers from dtop = d(loss)/d(top_instance)
 * [C]; only the selected instance (x_row [k_in], hid_row [width] after ReLU) carries gradient. */
int moc_adapter_backward_rows(const float* x_rows, const float* a2_rows, float clip_ratio, const float* classifier,
                              int n_classes, const int32_t* cls_of_row, const float* g_of_row, int64_t n_rows,
                              float* da2, void* stream);
int moc_mask_positive(float* g, const float* ref, int64_t n, void* stream);
int moc_transpose(const float* in, int rows, int cols, float* out, void* stream);
int moc_mil_fc_backward(const float* x_row, int k_in, const float* hid_row, int width, const float* w_last,
                        int n_classes, const float* dtop, float* d_w0, float* d_b0, float* d_wl, float* d_bl,
                        void* stream);

/* per-row softmax of [n_rows][n_cols] logits (MIL_fc, models/model_mil.py:38) */
int moc_row_softmax(const float* logits, int64_t ld, int n_cols, int64_t n_rows, float* probs, int64_t ldp,
                    void* stream);

/* ---- a12: loss, backward, optimiser ----------------------------------------
 * cross_entropy on [n,C] rows without temperature (main_moc.py:406,:494):
 * loss[i], optional dlogits[i,:] = grad_scale * (softmax - onehot), optional pred[i]. */
int moc_cross_entropy(const float* bag_logits, const int64_t* labels, int n_slides, int n_classes,
                      float grad_scale, float* loss, float* dlogits, int32_t* pred, void* stream);

/* d(sum_i <dlogits_i, bag_logits_i>) / d(senet parameters), written (not
 * accumulated) to grads laid out [w1 64x512 | b1 64 | w2 4x64 | b2 4] =
 * MOC_NUM_PARAMS floats.  Only the <= topk*C pooled rows of each slide carry
 * gradient (autograd through topk/mean/mul/sigmoid/Linear, main_moc.py:390-409). */
#define MOC_NUM_PARAMS (MOC_HIDDEN * MOC_FEAT_DIM + MOC_HIDDEN + MOC_GATES * MOC_HIDDEN + MOC_GATES)
size_t moc_head_backward_workspace_bytes(int n_slides, int n_classes, int topk);
int moc_head_backward(const float* feat, const float* keys, int64_t key_stride, int n_classes,
                      const int64_t* sel_base, const int32_t* sel_rows, const int32_t* sel_count,
                      int n_slides,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      unsigned active_mask, int topk, const int32_t* pool_pos, const float* dlogits,
                      float* grads, void* workspace, size_t workspace_bytes, void* stream);

/* torch.optim.Adam single step with L2 weight decay folded into the gradient
 * (main_moc.py:316: lr 1e-3, weight_decay 1e-4; betas .9/.999, eps 1e-8).
 * step is the 1-based step count after this update. */
int moc_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int64_t step, float lr, float beta1, float beta2, float eps, float weight_decay,
                  void* stream);

/* The same update with the step count in DEVICE memory, for training steps captured into a CUDA graph (a graph
 * replays fixed kernel arguments, so the bias corrections cannot be host scalars): moc_adam_prepare_dev adds 1 to
 * *step (int64) and writes scalars[0] = lr / (1 - beta1^step), scalars[1] = sqrt(1 - beta2^step); moc_adam_apply_dev
 * updates one parameter tensor with them.  One prepare, then one apply per tensor, per optimizer step. */
int moc_adam_prepare_dev(int64_t* step, float* scalars, float lr, float beta1, float beta2, void* stream);
int moc_adam_apply_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                       const float* scalars, float beta1, float beta2, float eps, float weight_decay, void* stream);

/* dst[i] += src[i]: sums the gate gradients of the slides of one data-parallel micro-batch before the single
 * all-reduce + Adam step (north_star's "small all-reduce of meta-optimizer gradients"; the reference's loop,
 * main_moc.py:406-410, steps once per slide and has no such mode). */
int moc_accumulate(float* dst, const float* src, int64_t n, void* stream);

/* ---- a1: on-disk bags -------------------------------------------------------
 * Native reader for CLAM-style HDF5 bag files, replacing h5py.File(path)['features'][:] /
 * ['coords'][:] (datasets/dataset_generic.py:424-430) for the subset of the format those
 * files use (h5py defaults: superblock v0/v1, version-1 object headers, symbol-table groups;
 * chunked, contiguous or compact layout; no filters).  HOST pointers; no CUDA calls.
 * moc_h5_read writes the dataset densely (row-major, file element type, little endian)
 * into dst_h, e.g. straight into a pinned staging buffer.  type_class: 0 integer, 1 float. */
int moc_h5_open(const char* path, void** handle);
void moc_h5_close(void* handle);
int moc_h5_dataset_info(void* handle, const char* name, int* rank, int64_t* dims4, int* type_class,
                        int* elem_size);
int moc_h5_read(void* handle, const char* name, void* dst_h, size_t dst_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MOC_B200_H */
