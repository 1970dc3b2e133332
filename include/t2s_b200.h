/* libt2s_sm100 -- C ABI of the B200-native forward/grounding path of T2S-QA / M4C.
 *
 * The reference has no FFI for this path: every device operation is a stock ATen call made from
 * pythia/models/t2s.py, pythia/models/m4c.py, pythia/modules/spatio_temporal_grounding.py and
 * pythia/modules/losses.py.  Each entry point below replaces the ATen call group cited beside it
 * (paths relative to /root/reference/pythia); INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises,
 *     nothing allocates (workspaces are passed in, see *_workspace_bytes)
 *   - return value: 0 = ok, < 0 = argument error (T2S_ERR_*), > 0 = cudaError_t of the launch;
 *     t2s_last_error() returns a thread-local description of the last failure
 *   - matrices are row-major; `ld*` are row pitches in ELEMENTS; dims are in elements
 *   - ids / masks coming from the dataset are int64 as in the reference
 */
#ifndef T2S_B200_H
#define T2S_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define T2S_ABI_VERSION 1

/* t2s_gemm_* flags */
#define T2S_GEMM_GELU 1     /* erf-GELU after bias (BertIntermediate, modeling_bert.gelu)           */
#define T2S_GEMM_OUT_F32 2  /* bf16 GEMM: store C as fp32 instead of bf16                           */
#define T2S_GEMM_RES_F32 4  /* bf16 GEMM: residual operand is fp32 instead of bf16                  */
#define T2S_GEMM_OUT_SPLIT 8 /* bf16 GEMM: store C as bf16 hi|lo, hi at [.,0..N), lo at [.,N..2N)       */
#define T2S_GEMM_DGELU 16    /* backward of BertIntermediate: C = (A.W^T) * GELU'(aux), aux = the saved pre-activation passed
                              * through `residual` / `ldr` (bf16, or fp32 with T2S_GEMM_RES_F32); no bias                  */
#define T2S_GEMM_SM_CAP_SHIFT 8 /* bits 8..15 of flags: cap the persistent grid at that many CTAs (0 = every SM), so a
                                 * latency-bound kernel chain on another stream (the greedy decode) finds free SMs    */

int t2s_abi_version(void);
const char* t2s_last_error(void);

/* TMA tensor maps of the GEMM operands are cached per (pointer, geometry, box, device) inside the library;
 * which = 0: hits, 1: misses (encodes) since load */
long long t2s_tmap_cache_stats(int which);
/* K1  C[M,N] = epi(A[M,K] . W[N,K]^T + bias (+ residual)); A, W bf16; tcgen05 + TMA + TMEM.
 * Replaces nn.Linear/addmm of BertSelfAttention.query/key/value, BertSelfOutput.dense,
 * BertIntermediate.dense(+gelu), BertOutput.dense (via models/t2s.py:622), ClassifierLayer
 * (modules/layers.py:101-107), OcrPtrNet.query/key (models/t2s.py:653,659).
 * block_n: 0 = auto, or 64 / 128 / 256.  lda, ldw multiples of 8; bases 16-byte aligned.
 * block_n 256 with M >= 256 launches 2-CTA clusters (one tcgen05.mma.cta_group::2 of M = 256 per pair). */
int t2s_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                  const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K,
                  int flags, int block_n, void* stream);

/* K1x "bf16x3": fp32-class contraction on the bf16 tensor pipe for the grounding chain (TextBert
 * t2s.py:538, obj/OCR encoders t2s.py:211,248, QTV t2s.py:423), whose top-k indices must match the fp32
 * reference.  A [M,2K] and W [N,2K] hold fp32 values split as bf16 hi|lo (lo at column K; see
 * t2s_split_bf16); the kernel accumulates hi.hi + hi.lo + lo.hi in fp32 (error ~2^-17 relative per product).
 * K % 64 == 0.  Same epilogues as t2s_gemm_bf16. */
int t2s_gemm_bf16x3(const void* A, long long lda, const void* W, long long ldw, const float* bias,
                    const void* residual, long long ldr, void* C, long long ldc, int M, int N, int K,
                    int flags, int block_n, void* stream);
/* K8b weight gradient: dW[P,Q] += G[rows,P]^T . X[rows,Q] (fp32 accumulate INTO dW: zero it first).  G = gradient of
 * the layer output, X = the layer input, both bf16 row-major exactly as the forward / backward kernels wrote them
 * (MN-major tcgen05 operands: no transposed copies).  The rows are split into `splits` ranges (0 = auto) whose
 * partial tiles are added with TMA reduce-add, so the summation order -- not the value to fp32 rounding -- can vary
 * from run to run.  Replaces the autograd addmm for every nn.Linear.weight.grad of pythia/models/t2s.py
 * (loss.backward(), trainers/base_trainer.py:264).  ldg, ldx multiples of 8; ldd multiple of 4. */
int t2s_gemm_wgrad_bf16(const void* G, long long ldg, const void* X, long long ldx, float* dW, long long ldd,
                        int rows, int P, int Q, int splits, void* stream);

/* fp32 rows [rows,K] -> bf16 hi|lo rows [rows,2K']: hi = bf16(x) at column c, lo = bf16(x - hi) at column
 * lo_off + c; columns K..lo_off of both halves are zero-filled (lo_off >= K, multiple of 8). */
int t2s_split_bf16(const float* x, long long ldx, int rows, int K, int lo_off, void* out, long long ldo,
                   int rows_per_group, int in_group_rows, int in_row_off, void* stream);
/* rows_per_group > 0 gathers input row r from (r / per) * in_group_rows + in_row_off + r % per */

/* K1f same contraction in fp32 on the FMA pipes (grounding chain: TextBert t2s.py:538, obj/OCR
 * encoders t2s.py:211,248, QTV t2s.py:423, Grounding_Module.q_linear t2s.py:472). K % 4 == 0.
 * a_rows_per_group > 0 gathers A row m from (m / per) * a_group_rows + a_row_off + m % per. */
int t2s_gemm_f32(const float* A, long long lda, const float* W, long long ldw, const float* bias,
                 const float* residual, long long ldr, float* C, long long ldc, int M, int N, int K,
                 int flags, int a_rows_per_group, int a_group_rows, int a_row_off, void* stream);

/* K2  masked multi-head attention over a fused [rows, 3H] q|k|v buffer (head size 64).
 * Replaces BertSelfAttention matmul/+mask/softmax/matmul and the [B,1,L,L] masks of
 * models/t2s.py:413-419,533-534,609-618.  key_idx[b, 0..n_keys[b]) lists the valid key rows. */
int t2s_attn_f32(const float* qkv, long long ld, int B, int L, int H, int heads, const int* key_idx,
                 const int* n_keys, int key_stride, float* out, long long ldo, void* out_split,
                 long long ldo_split, void* stream);   /* out_split: optional bf16 hi|lo copy (lo at column H) */
/* fp32-class attention on the tensor pipe: q|k|v and the output are bf16 hi|lo pairs (lo_off = column offset of
 * the lo half of qkv, >= 3H; output lo half at column H).  Same call sites as t2s_attn_f32. */
int t2s_attn_x3(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads, const int* key_idx,
                const int* n_keys, int key_stride, void* out_split, long long ldo, void* stream);
/* K2t attention on tcgen05/TMEM (see csrc/attn_tc.cu).  lo_off == 0: bf16 q|k|v [rows, 3H], bf16 output;
 * lo_off >= 3H: bf16 hi|lo operands (fp32-class, three products per contraction), output hi|lo (lo at column H). */
int t2s_attn_tc(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads, const int* key_idx,
                const int* n_keys, int key_stride, void* out, long long ldo, void* stream);
int t2s_attn_bf16(const void* qkv, long long ld, int B, int L, int H, int heads, const int* key_idx,
                  const int* n_keys, int key_stride, void* out, long long ldo, void* stream);
/* decoder rows t0..t0+nq-1 (nq <= 16): valid encoder keys + causal decoder keys (t2s.py:574-579,609-615) */
int t2s_attn_dec(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec, long long ld_dec,
                 int T, int B, int H, int heads, const int* key_idx, const int* n_keys, int key_stride,
                 int t0, int nq, void* out, long long ldo, void* stream);

/* K3  BertEmbeddings: LN(word[id] + position[row % L] + token_type[0])  (via models/t2s.py:530) */
int t2s_bert_embed_ln(const long long* ids, int rows, int L, int H, const float* word, const float* pos,
                      const float* type0, const float* gamma, const float* beta, float eps, float* out,
                      long long ldo, void* stream);
/* K3  [normalize(f0) | normalize(f1) | tab0[id0] | tab1[id1] | 0-pad] -> fp32 rows
 * (F.normalize + nn.Embedding + torch.cat of models/t2s.py:195-207 and 223-244) */
int t2s_feat_concat(const float* f0, int d0, const float* f1, int d1, const long long* id0, const float* tab0,
                    const long long* id1, const float* tab1, int id_dim, int rows, float* out, long long ldo,
                    int k_pad, void* out_split, long long ldo_split, void* stream);
/* out (fp32 rows) and/or out_split (bf16 hi|lo rows, lo at column k_pad: operand of t2s_gemm_bf16x3) */
/* K4  y = LN(x (+ res)); optional epilogue y = tanh_base + tanh(y) (QTV residual, t2s.py:430-432);
 * fp32 and/or bf16 output; output rows remapped as (r / rows_per_group) * out_group_rows +
 * out_row_off + r % rows_per_group when rows_per_group > 0 (writes straight into the joint
 * [txt; frames; ocr] buffer instead of torch.cat, t2s.py:392-395). */
int t2s_add_ln(const void* x, int x_bf16, long long ldx, const void* res, int res_bf16, long long ldr,
               const float* gamma, const float* beta, float eps, int rows, int H, const float* tanh_base,
               long long ld_base, float* out32, long long ldo32, void* out16, long long ldo16,
               int rows_per_group, int out_group_rows, int out_row_off, void* stream);
/* same, with out16 written as bf16 hi|lo (lo at column H) -- the operand format of t2s_gemm_bf16x3 */
int t2s_add_ln_split(const void* x, int x_bf16, long long ldx, const void* res, int res_bf16, long long ldr,
                     const float* gamma, const float* beta, float eps, int rows, int H, const float* tanh_base,
                     long long ld_base, float* out32, long long ldo32, void* out16, long long ldo16,
                     int rows_per_group, int out_group_rows, int out_row_off, void* stream);
/* K3  LN(h) + LN(W2 . bbox + b2)  (models/t2s.py:246-252) */
int t2s_ocr_finish(const float* h, long long ldh, const float* bbox, const float* w2, const float* b2,
                   const float* g1, const float* be1, const float* g2, const float* be2, float eps, int rows,
                   int H, float* out, long long ldo, int rows_per_group, int out_group_rows, int out_row_off,
                   void* stream);
/* K3  PrevPredEmbeddings.forward for decoder positions t0..t0+nt-1 (models/t2s.py:690-723) */
int t2s_prev_embed(const long long* prev_inds, int ld_prev, int B, int t0, int nt, int T, int V, int H,
                   const float* ans_w, const float* ocr_emb, long long ocr_batch_stride, long long ld_ocr,
                   const float* pos_emb, const float* type_emb, const float* ans_g, const float* ans_b,
                   const float* ocr_g, const float* ocr_b, const float* emb_g, const float* emb_b, float eps,
                   void* out16, float* out32, long long ldo, int n_ocr, void* stream);
int t2s_cast_rows_bf16(const float* x, long long ldx, int rows, int H, void* out, long long ldo,
                       int rows_per_group, int out_group_rows, int out_row_off, void* stream);

/* K5  grounding (models/t2s.py:453-518, modules/spatio_temporal_grounding.py:15-142, m4c.py:356-422) */
int t2s_mask_prep(const long long* text_len, const long long* frame_mask, const long long* ocr_mask, int B,
                  int Lt, int F, int O, float* joint, void* stream);
int t2s_build_keys(const float* mask, int B, int L, int* key_idx, int* n_keys, int key_stride, void* stream);
int t2s_question_pool(const float* qp, int B, int Lt, int H, const float* w, const float* bw,
                      const float* txt_mask, int mask_stride, float* gq, void* stream);
int t2s_sim_scores(const float* gq, const float* X, long long batch_stride, long long ldx, int row0, int N,
                   int H, int B, float* sim, void* stream);
int t2s_temporal_select(const float* sim, int sim_stride, const float* joint_mask, int B, int Lt, int F,
                        int Of, const float* gumbel, const long long* frame_id, const long long* temporal_id,
                        int topk, const float* pos_override, const float* neg_override,
                        long long* ground_frame, float* pos_joint, float* neg_joint, float* slot_mask,
                        float* dbg_score, void* stream);
int t2s_spatial_select(const float* sim, int sim_stride, int sim_off, const float* slot_mask,
                       const float* joint_mask, int B, int L_joint, int ocr_off, int F, int Of,
                       const float* gumbel, const float* boxes, int topk, int mode, float* ground_box,
                       float* pos_joint, float* neg_joint, float* dbg_score, void* stream);
int t2s_middle_frame_slots(const long long* mid_id, const long long* temporal_id, int B, int O,
                           float* slot_mask, void* stream);
/* Ablation variants (models/t2s_wo_sg.py:496-506, models/t2s_wo_tg.py:483-535).  t2s_spatial_select modes: 0 = T2S,
 * 1 = M4C post-hoc, 2 = "w/o SG" (pos = slot mask, neg = 1 - slot mask, ground_box [B, topk * Of, 4] = boxes of the
 * positive slots, `topk` = number of grounded frames, buffer zeroed by the caller), 3 = "w/o TG" (mode 0, both masks
 * also multiplied by the OCR part of joint_mask), 4 = T5-ViteVQA post-hoc (models/t5vitevqa.py:396-408: the `topk` OCR
 * tokens with the largest masked attention over all frames, ground_box [B, topk, 4] in slot order, padding slots zeroed;
 * joint masks untouched).  t2s_frame_slots: slot_mask [B, O] = 1 where temporal_id equals one
 * of the n_ids ids of the sample (id 0 read as 1).  t2s_frames_from_ocr: frame part of pos / neg joint masks = the
 * first n_pick frames that own a positive / negative OCR slot (fewer: the last frame is set, the reference's index
 * -1), ground_frame [B, n_pick] = positive frame positions, -1 padded; copies the question part from joint_mask. */
int t2s_frame_slots(const long long* ids, int n_ids, const long long* temporal_id, int B, int O, float* slot_mask,
                    void* stream);
int t2s_frames_from_ocr(const float* joint_mask, float* pos_joint, float* neg_joint, int B, int Lt, int F, int Of,
                        int n_pick, long long* ground_frame, void* stream);

/* K6  pointer scores written into scores[:, :, V:], argmax feedback (models/t2s.py:661-666,285,353-354) */
int t2s_ptr_score(const void* q, long long ldq, int B, int T, int t0, int nq, const void* keyp,
                  long long key_batch_stride, long long ldk, int O, int H, const float* mask,
                  long long mask_stride, float* scores, long long ld_scores, int V, void* stream);
int t2s_argmax_feedback(const float* scores, long long ld_scores, int B, int T, int t0, int nt, int N,
                        long long* prev_inds, int ld_prev, long long* argmax_out, void* stream);

/* K7  losses (modules/losses.py:329-343 and 361-385) */
long long t2s_loss_workspace_bytes(int B, int T);
int t2s_pos_bce_loss(const float* scores, const float* targets, const float* loss_mask, int B, int T, int N,
                     void* workspace, float* out, void* stream);
int t2s_info_nce_loss(const float* ref, const float* pos, const float* neg, int B, int T, int N,
                      float temperature, void* workspace, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K8  training step: what autograd runs in the reference for loss.backward(), clip_grad_norm_ and Adam.step()
 * (pythia/trainers/base_trainer.py:262-270, utils/general.py:32-40).  Gradient accumulators (d*) are fp32 and are
 * ADDED to: the caller zeroes the flat gradient buffer once per step.  Activations / activation gradients are bf16
 * unless a *_bf16 flag says otherwise.  The weight-gradient GEMM (t2s_gemm_wgrad_bf16) and T2S_GEMM_DGELU are
 * declared with K1 above.
 * --------------------------------------------------------------------------------------------------------------- */
/* LayerNorm backward: y = LN(h); tanh_out: the forward output was base + tanh(y) (QTV, models/t2s.py:430-432).
 * dh (row-compact) is the gradient of h = of the residual branch and of the preceding Linear's output; dgamma,
 * dbeta and (optional) dbias = column sums of dh are accumulated.  dy rows may be gathered from a larger buffer:
 * row r of dy lives at (r / per) * group + off + r % per when dy_rows_per_group > 0. */
int t2s_ln_bwd(const void* h, int h_bf16, long long ldh, const void* dy, int dy_bf16, long long lddy,
               int dy_rows_per_group, int dy_group_rows, int dy_row_off, const float* gamma, const float* beta,
               float eps, int rows, int H, int tanh_out, void* dh, int dh_bf16, long long lddh, float* dgamma,
               float* dbeta, float* dbias, void* stream);
/* dst[c] += sum_r x[r, c] (bias gradients of q|k|v, intermediate, classifier, pointer-net) */
int t2s_colsum(const void* x, int x_bf16, long long ldx, int rows, int N, float* dst, void* stream);
/* out[map(r), :] (+)= a[r] + b[r] + c[r]; a, b, c bf16 (b, c optional), out fp32 */
int t2s_rows_add(const void* a, const void* b, const void* c, long long ldi, int rows, int H, float* out,
                 long long ldo, int rows_per_group, int out_group_rows, int out_row_off, int accumulate, void* stream);
/* training-mode forward of BertIntermediate's activation: the GEMM stores the pre-activation u (needed by
 * T2S_GEMM_DGELU), this writes gelu(u) as bf16, or as bf16 hi|lo (lo at column lo_off) when lo_off > 0 */
int t2s_gelu_rows(const void* u, int u_bf16, long long ldu, int rows, int N, void* out, long long ldo, int lo_off,
                  void* stream);
/* nn.Embedding backward: table[ids[r], 0..d) += src[r, c0..c0+d); rows with ids[r] == pad_id skipped */
int t2s_embed_scatter_add(const void* src, int src_bf16, long long lds, int c0, int d, const long long* ids, int rows,
                          long long pad_id, float* table, long long ldt, void* stream);
/* OcrPtrNet score backward (models/t2s.py:661-666): dq[b,t] and dkeyp[b,o] from dscores[:, :, V:] */
int t2s_ptr_score_bwd(const float* dscores, long long ld_scores, int B, int T, int V, const void* q, long long ldq,
                      const void* keyp, long long key_batch_stride, long long ldk, int O, int H, void* dq, long long lddq,
                      void* dkeyp, long long dkey_batch_stride, long long lddk, void* stream);
/* PrevPredEmbeddings backward (models/t2s.py:690-723): dx [B*T, H] bf16 -> classifier.weight rows (d_ans_w), OCR rows
 * of the joint embedding gradient (d_ocr_emb, same strides as ocr_emb), position / token_type tables, three LNs */
int t2s_prev_embed_bwd(const void* dx, long long lddx, const long long* prev_inds, int ld_prev, int B, int T, int V,
                       int H, const float* ans_w, const float* ocr_emb, long long ocr_batch_stride, long long ld_ocr,
                       const float* pos_emb, const float* type_emb, const float* ans_g, const float* ocr_g,
                       const float* emb_g, float eps, float* d_ans_w, float* d_ocr_emb, float* d_pos, float* d_type,
                       float* d_ans_g, float* d_ans_b, float* d_ocr_g, float* d_ocr_b, float* d_emb_g, float* d_emb_b,
                       int n_ocr, void* stream);
/* backward of t2s_ocr_finish: dh (bf16) = gradient of linear_ocr_feat_to_mmt_in's output; dc_ws [rows, H] fp32 scratch
 * = gradient of linear_ocr_bbox_to_mmt_in's output; accumulates both LayerNorms, both biases and dW2 [H, 4] */
int t2s_ocr_finish_bwd(const float* h, long long ldh, const float* bbox, const float* w2, const float* b2,
                       const float* g1, const float* g2, float eps, int rows, int H, const float* dout, long long ldd,
                       int dy_rows_per_group, int dy_group_rows, int dy_row_off, void* dh, long long lddh, float* dc_ws,
                       long long lddc, float* dg1, float* db1, float* dg2, float* db2ln, float* dbias1, float* dw2,
                       float* db2, void* stream);
/* BertEmbeddings backward: dy [rows, H] bf16 -> word (padding_idx 0 skipped) / position / token_type tables + LN */
int t2s_bert_embed_bwd(const void* dy, long long lddy, const long long* ids, int rows, int L, int H, const float* word,
                       const float* pos, const float* type0, const float* gamma, float eps, float* d_word, float* d_pos,
                       float* d_type, float* dgamma, float* dbeta, void* stream);
/* attention backward, recompute style (csrc/attn_bwd.cu): encoder rows [B*Le] + optional decoder rows [B*T] with the
 * prefix-LM mask of models/t2s.py:609-618; q|k|v, context, context gradient in; dq|dk|dv out.  max_keys >= max n_keys. */
long long t2s_attn_bwd_workspace_bytes(int B, int Le, int T, int heads);
int t2s_attn_bwd(const void* qkv_enc, long long ld_enc, const void* qkv_dec, long long ld_dec, const void* o_enc,
                 long long ldo_enc, const void* o_dec, long long ldo_dec, const void* do_enc, long long ldg_enc,
                 const void* do_dec, long long ldg_dec, void* dqkv_enc, long long ldq_enc, void* dqkv_dec,
                 long long ldq_dec, int B, int Le, int T, int H, int heads, const int* key_idx, const int* n_keys,
                 int key_stride, int max_keys, void* workspace, void* stream);
/* loss backward (modules/losses.py:329-343, 361-385): grad_out = device scalar dL/dloss (carries the loss weight) */
long long t2s_loss_bwd_workspace_bytes(int B, int T);
int t2s_nce_rowstats(const float* ref, const float* pos, const float* neg, int rows, int N, float* stats, void* stream);
int t2s_pos_bce_loss_bwd(const float* scores, const float* targets, const float* loss_mask, int B, int T, int N,
                         const float* grad_out, float* dscores, int accumulate, void* stream);
int t2s_info_nce_loss_bwd(const float* ref, const float* pos, const float* neg, int B, int T, int N, float temperature,
                          void* workspace, const float* grad_out, float* dref, float* dpos, float* dneg, int accumulate,
                          void* stream);
/* clip_grad_norm_ + torch.optim.Adam over a flat fp32 range: out[0] = sum g^2 (workspace: 1024 doubles); the step
 * scales g by grad_scale * min(1, max_norm / (sqrt(sumsq) * grad_scale + 1e-6)) (max_norm <= 0: no clipping) */
int t2s_sumsq(const float* g, long long n, void* workspace, float* out, void* stream);
/* after the optimizer step: rewrite, in place, every GEMM operand derived from the fp32 parameters (bf16 copies, bf16
 * hi|lo splits, transposed bf16 copies for the dgrad GEMMs) from a device-resident job table of 48-byte records
 * {long long src_off, dst, ld_dst; int rows, cols, mode (0 bf16, 1 hi|lo split, 2 bf16 transposed, 3 fp32), k_pad,
 * tile0, tiles_x}; n_tiles = total 32 x 32 tiles.  Replaces ~300 tensor ops per step (what torch's `.to(bfloat16)`,
 * `torch.cat`, `.t().contiguous()` did after `Adam.step()`, base_trainer.py:269). */
int t2s_repack_weights(const float* flat_param, const void* jobs, int n_jobs, int n_tiles, void* stream);
int t2s_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, int step, const float* sumsq, float max_norm, float grad_scale, void* stream);

/* Dropout of the training step.  The reference trains with p = 0.1 at every site below (configs/t2s_*.yml obj / ocr
 * dropout_prob; BertConfig hidden_dropout_prob / attention_probs_dropout_prob defaults, models/t2s.py:25-28):
 *   BertEmbeddings: LN -> dropout                                  (via t2s.py:530)        t2s_dropout_rows
 *   obj_drop / ocr_drop on the encoded frames / OCR tokens          (t2s.py:95,118,214,253) t2s_dropout_rows
 *   PrevPredEmbeddings.emb_dropout                                  (t2s.py:688,720)        t2s_dropout_rows
 *   BertSelfOutput / BertOutput: LN(dropout(Linear(x)) + input)     (via t2s.py:423,538,622) t2s_add_ln_dropout
 *   BertSelfAttention: dropout(softmax(scores)) . V                 (same)                  t2s_attn_tc_dropout / _dec_
 * No mask is stored: a mask is a pure function of (seed, site, element) -- one 32-bit counter hash per element pair,
 * csrc/common.cuh -- and the backward entry points recompute it from the same (p, seed, site).  p is applied as
 * thr = round(p * 65536) on 16-bit draws; kept values are scaled by 65536 / (65536 - thr).  `site` < 4096.
 * t2s_dropout_rows works in place on rows (r / per) * group + off + r % per (per == 0: r) and is its own backward
 * (apply it to the gradient rows).  t2s_add_ln_dropout = t2s_add_ln / _split with dropout on x before + res; the
 * pre-LayerNorm sum is written to h_out (dtype of x, may alias x) for t2s_ln_bwd_dropout, which returns dh (residual
 * branch) and dh_drop = dh * mask (Linear output; dbias sums it).  t2s_dropout_mask writes a site's multipliers as
 * fp32 (mode 0: [rows, H]; mode 1: attention [BH, n_query, n_key]) -- for tests. */
int t2s_add_ln_dropout(const void* x, int x_bf16, long long ldx, const void* res, int res_bf16, long long ldr,
                       const float* gamma, const float* beta, float eps, int rows, int H, const float* tanh_base,
                       long long ld_base, float* out32, long long ldo32, void* out16, long long ldo16, int split,
                       int rows_per_group, int out_group_rows, int out_row_off, void* h_out, float p,
                       unsigned long long seed, unsigned site, void* stream);
int t2s_dropout_rows(void* x, int x_bf16, long long ldx, int rows, int H, int rows_per_group, int group_rows,
                     int row_off, float p, unsigned long long seed, unsigned site, void* stream);
int t2s_dropout_mask(float* out, int mode, int rows_or_bh, int H, int n_query, int n_key, float p,
                     unsigned long long seed, unsigned site, void* stream);
int t2s_ln_bwd_dropout(const void* h, int h_bf16, long long ldh, const void* dy, int dy_bf16, long long lddy,
                       int dy_rows_per_group, int dy_group_rows, int dy_row_off, const float* gamma, const float* beta,
                       float eps, int rows, int H, int tanh_out, void* dh, int dh_bf16, long long lddh, float* dgamma,
                       float* dbeta, float* dbias, void* dh_drop, float p, unsigned long long seed, unsigned site,
                       void* stream);
/* The attention kernels of the training step also hand the backward what it would otherwise recompute: lse_out != null
 * receives the log2-sum-exp of every query row, at lse_out[((b * heads + h) * rows + i) * 2] with i the row's position in
 * the virtual sequence [encoder rows; decoder rows] (rows = lse_rows for t2s_attn_tc_dropout, L_enc + T for
 * t2s_attn_dec_dropout) -- the {lse2, D} layout t2s_attn_bwd_dropout takes as `stats_lse`.  p = 0 is allowed. */
int t2s_attn_tc_dropout(const void* qkv, long long ld, int lo_off, int B, int L, int H, int heads, const int* key_idx,
                        const int* n_keys, int key_stride, void* out, long long ldo, float p, unsigned long long seed,
                        unsigned site, float* lse_out, int lse_rows, void* stream);
int t2s_attn_dec_dropout(const void* qkv_enc, long long ld_enc, int L_enc, const void* qkv_dec, long long ld_dec, int T,
                         int B, int H, int heads, const int* key_idx, const int* n_keys, int key_stride, int t0, int nq,
                         void* out, long long ldo, float p, unsigned long long seed, unsigned site, float* lse_out,
                         void* stream);
int t2s_attn_bwd_dropout(const void* qkv_enc, long long ld_enc, const void* qkv_dec, long long ld_dec, const void* o_enc,
                         long long ldo_enc, const void* o_dec, long long ldo_dec, const void* do_enc, long long ldg_enc,
                         const void* do_dec, long long ldg_dec, void* dqkv_enc, long long ldq_enc, void* dqkv_dec,
                         long long ldq_dec, int B, int Le, int T, int H, int heads, const int* key_idx,
                         const int* n_keys, int key_stride, int max_keys, void* workspace, float p,
                         unsigned long long seed, unsigned site, float* stats_lse, void* stream);

/* Input featurisation, the step before the path.  PHOC descriptor of OCR tokens: replaces the
 * reference's CPU extension pythia/utils/phoc/src/cphoc.c:12-113 + build_phoc.py:9-14 (lower-case, keep [a-z0-9])
 * + PhocProcessor, datasets/processors.py:904-928.  bytes = the tokens' UTF-8 bytes back to back (non-ASCII tokens
 * lower-cased by the caller; ASCII is lower-cased and filtered on the device), offsets[n_tokens + 1] = byte ranges;
 * out [rows, 604] fp32 0/1 with row stride ldo; rows n_tokens..rows-1 are the processor's zero padding. */
int t2s_phoc_build(const unsigned char* bytes, const int* offsets, int n_tokens, int rows, float* out, long long ldo,
                   void* stream);
/* same descriptor from fixed-width records: token i = bytes[i * width, (i + 1) * width), zero padded (byte 0 is outside
 * the alphabet) -- the batched `ocr_token_bytes` [B, O, width] field of a SampleList, so that the forward takes the OCR
 * token TEXT instead of the 604 fp32 of `context_feature_1` per token (vtextgqa/dataset.py:237, processors.py:904-928) */
int t2s_phoc_build_fixed(const unsigned char* bytes, int width, int n_tokens, float* out, long long ldo, void* stream);
/* Frame sampling + per-frame OCR truncate / pad / pack, the python list work of the dataset in front of the featurisers:
 * replaces vtextgqa/dataset.py:103-158 (uniform frame ids via sample_frames :371-381, per frame the first Of detections,
 * box = min / max of the quadrilateral, "<pad>" slots that keep the frame index), :166-195 (middel_frame_id / _idx),
 * :199-243 (zero-padded id / mask vectors, float64 box normalisation by 1/width, 1/height + CopyProcessor
 * processors.py:932-944).  Detections of all videos of the batch back to back in OCR-info frame order: det_points
 * [n, 8] fp32, det_track [n], det_tokens [n, width] zero-padded UTF-8 records (already through the token processor);
 * frame_ptr = CSR index of the info frames of all videos, info_base[b] = first info frame of video b, n_info[b] =
 * len(ocr_info), n_frames[b] = number of video frames (needs n_frames - 1 <= n_info), vid_w / vid_h as binary64.
 * Outputs as the Sample fields of the reference: ocr_bbox [B, F*Of, 4] fp32, track_id / temporal_id / ocr_mask
 * [B, F*Of] int64, frame_id / frame_mask [B, F] int64, frame_num / mid_frame_id / mid_frame_idx [B] int64, and
 * ocr_token_bytes [B, F*Of, width] (the input of t2s_phoc_build_fixed; empty records on missing frames). */
int t2s_pack_ocr_frames(const float* det_points, const long long* det_track, const unsigned char* det_tokens, int width,
                        const int* frame_ptr, const int* info_base, const int* n_info, const int* n_frames,
                        const double* vid_w, const double* vid_h, int B, int F, int Of, float* ocr_bbox,
                        long long* track_id, long long* temporal_id, long long* ocr_mask, long long* frame_id,
                        long long* frame_mask, long long* frame_num, long long* mid_frame_id, long long* mid_frame_idx,
                        unsigned char* ocr_token_bytes, void* stream);

/* K9  Evaluation step that consumes the forward's outputs.
 * t2s_answer_decode: `pos_scores.argmax(-1)` and the EOS cut of the python loop in modules/metrics.py:186-207 (= 395-416,
 * 498-519).  ans_ids [B, T] int32 = argmax of every decoding row (lowest index on ties); ans_len [B] = number of ids
 * before the first VOCABULARY id equal to eos_idx (ids >= V are OCR copies and never end the answer), T if none. */
int t2s_answer_decode(const float* scores, long long ld_scores, int B, int T, int N, int V, int eos_idx,
                      int* ans_ids, int* ans_len, void* stream);
/* t2s_ground_metrics: BoxGroundAccuracyEvaluator.eval_pred_list at two IoU thresholds and
 * TempGroundAccuracyEvaluator.eval_pred_list (utils/m4c_evaluators.py:301-405; callers modules/metrics.py:233-546) on
 * ground_frame [B, kf] int64 / ground_box [B, n_box, 4] fp32 as the forward returns them.  The ground-truth
 * annotation file is packed once on the host (vitxt_gqa_b200/metrics.py GroundAnnotations): record r owns spans
 * span_ptr[r] .. span_ptr[r+1]; span s = frames [span_st[s], span_ed[s]] (= int(t*fps)+1, computed in python) and the
 * labelled boxes box_ptr[s] .. box_ptr[s+1] (box_frame = int of the dict key, box_xyxy binary64); rec_wh [R, 2] =
 * (width, height); rec_index [B] = record of each sample's question (-1: none).  Outputs: ones / tail_zero [2, B] = per
 * sample and threshold the number of 1s and (0 | 1) trailing 0 the evaluator appends to its score list; t_hit [B];
 * status [B] bit 0 / 1 = the evaluator's box-order assertions on the labelled / predicted box would fail, bit 2 = no
 * annotation; acc [3] = IOU@thr_a, IOU@thr_b, temporal accuracy (sum / len in binary64, then float32);
 * head [2, B] = the first B entries of each concatenated score list (what GQA@x indexes by sample). */
int t2s_ground_metrics(const long long* ground_frame, int kf, const float* ground_box, int n_box, int ocr_topk,
                       const int* rec_index, const int* span_ptr, const long long* span_st, const long long* span_ed,
                       const int* box_ptr, const long long* box_frame, const double* box_xyxy, const double* rec_wh,
                       int B, double thr_a, double thr_b, int* ones, int* tail_zero, int* t_hit, int* status,
                       float* acc, int* head, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2S_B200_H */
