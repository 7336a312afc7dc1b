/* svgb200 — C ABI of the B200-native sparse video-DiT attention engine.
 *
 * The reference (svg-project/Sparse-VideoGen) has no FFI on this path: its boundary is the set of
 * Python operator signatures imported by svg/models/<m>/attention.py (hyvideo/attention.py:12-28)
 * and the svg/kernels/ops API.  This header is the boundary our own Python shim binds with ctypes;
 * each entry point names the reference operator it stands behind.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every function returns 0 on success, <0 on error; svgb_last_error() returns a thread-local
 *     message for the last failure on the calling thread.
 *   - all data pointers are DEVICE pointers unless a parameter is documented "host".
 *   - no function allocates device memory or synchronises the device: the caller provides the
 *     workspace (sizes from the *_bytes queries) and the cudaStream_t (passed as void*).
 *   - integer outputs (labels, permutations, plans, maps) are bit-exact, deterministic.
 *   - dtype: SVGB_BF16 or SVGB_F16 for q/k/v/o.
 */
#ifndef SVGB200_H_
#define SVGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGB_BF16 0
#define SVGB_F16 1
#define SVGB_E4M3 2 /* fp8 e4m3 inputs (svgb_attn_fwd_fp8 only) */
#define SVGB_F32 3  /* fp32 storage (transformer-block glue ops only) */

/* element-mask families for the SVG1 band plan (reference generate_temporal_head_mask_mod) */
#define SVGB_MASK_NONE 0
#define SVGB_MASK_HY 1  /* hyvideo/utils.py:20-44  m0=F*P, m1=F*P+prompt_len, m2=W              */
#define SVGB_MASK_WAN 2 /* wan/utils.py:25-41      m0=P (first-frame sink), m2=W  (|q-kv| <= W) */
#define SVGB_MASK_COG 3 /* cog/utils.py:30-46      m0=first-col limit, m1=prompt_len, m2=W      */

/* ---- library ---------------------------------------------------------------------------- */
int svgb_version(void);
const char* svgb_last_error(void);
/* 0 iff the current CUDA device is sm_100 (B200); fills optional outputs */
int svgb_device_check(int* sm_major, int* sm_minor, int* num_sms);

/* ---- attention plans -------------------------------------------------------------------- */
/* Host-side descriptor of a device-resident plan (filled by the plan functions, read by
 * svgb_attn_fwd).  Plain data; copyable. */
typedef struct svgb_plan {
  int32_t kind;          /* 1 = variable-block (TMA chunks), 2 = band, 3 = variable-block (row gather) */
  int32_t BH, S;
  int32_t max_items;     /* grid.x of the attention launch */
  int32_t items_stride;  /* 0 when one plan serves every head */
  int32_t counts_stride;
  int32_t mask_mode, m0, m1, m2;
  int64_t counts_off, items_off, chunks_off; /* byte offsets inside the plan workspace */
  int64_t bytes;
  int64_t aux_off;       /* kind 3: per-item selected-key totals */
} svgb_plan;
/* A plan built with BH == 1 may serve every head of a launch (one map for all heads, e.g. BSR masks):
 * set items_stride = counts_stride = 0 in the host struct. */

/* SVG2 / BSR / dense: q-block i of head h attends k-block j iff map[h,i,j] != 0.
 * Stands behind dynamic_block_sparse_fwd_flashinfer's wrapper.plan (svg/kmeans_utils.py:1355-1385).
 *   map    uint8 [BH, QC, KC]     row_sz int32 [BH, QC]     col_sz int32 [BH, KC]
 * Each head's row sizes and col sizes must sum to S (not checked on device). */
int svgb_attn_plan_varblock_bytes(int BH, int S, int QC, int KC, size_t* bytes);
int svgb_attn_plan_varblock(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH,
                            int S, int QC, int KC, void* plan_ws, size_t ws_bytes, svgb_plan* plan,
                            void* stream);

/* Same map, lowered for the row-gather kernel path: the selected key ranges are kept as runs and the
 * kernel gathers exactly-full 128-key chunks across run boundaries with cp.async (no padded tail chunk per
 * run), optionally through row-index vectors (svgb_attn_fwd_gather) so the cluster permutation of Q, K, V
 * (permute_tensor_by_labels_triton x3, svg/models/hyvideo/attention.py:651-653) never materialises. */
int svgb_attn_plan_varblock_gather(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH,
                                   int S, int QC, int KC, void* plan_ws, size_t ws_bytes, svgb_plan* plan,
                                   void* stream);

/* SVG1: one element-exact band mask shared by every head.  Stands behind prepare_flexattention /
 * create_block_mask (svg/models/hyvideo/attention.py:527-551). */
int svgb_attn_plan_band_bytes(int S, size_t* bytes);
int svgb_attn_plan_band(int mask_mode, int m0, int m1, int m2, int BH, int S, void* plan_ws,
                        size_t ws_bytes, svgb_plan* plan, void* stream);

/* ---- attention -------------------------------------------------------------------------- */
/* o = softmax(q k^T * sm_scale  [masked by plan]) v, per head.  Stands behind
 * dynamic_block_sparse_fwd_flashinfer (svg/kmeans_utils.py:1319-1392), flex_attention under the
 * SVG1 BlockMask (hyvideo/attention.py:401-403), the BSR ops (svg/kernels/ops/attention_ops.py:
 * 140-197) and the dense fall-backs.
 *   q,k,v : [BH, S, D] with explicit strides (elements): element (h,s,d) at h*head_stride +
 *           s*row_stride + d.  [B,H,S,D] contiguous: row_stride=D, head_stride=S*D.
 *           [S,H,D]: row_stride=H*D, head_stride=D.   D in {64,128}; base pointers 16-B aligned,
 *           strides multiples of 8 elements.
 *   o     : same addressing with its own strides.
 *   o_rows: optional int32 [BH,S]; when non-NULL output row for query row s is o_rows[h,s]
 *           (fuses apply_inverse_permutation_triton, svg/kernels/triton/permute.py:131-170).
 *   lse   : optional float [BH,S], natural-log log-sum-exp of the scaled scores.
 *   plan  : host pointer. */
int svgb_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                  const int32_t* o_rows, int dtype, int BH, int S, int D, long long row_stride,
                  long long head_stride, long long o_row_stride, long long o_head_stride,
                  float sm_scale, const svgb_plan* plan, const void* plan_ws, void* stream);

/* Gather form (plan from svgb_attn_plan_varblock_gather).  Row r of the plan's (cluster-sorted) query order
 * is read from q[h, q_rows[h,r]] (q_rows NULL: identity), key/value row r from k/v[h, kv_rows[h,r]], and the
 * output row is written to o[h, o_rows[h,r]] — pass the argsort of the q-labels as q_rows AND o_rows and the
 * argsort of the k-labels as kv_rows to run SVG2 attention directly on the un-permuted tensors. */
int svgb_attn_fwd_gather(const void* q, const void* k, const void* v, void* o, float* lse,
                         const int32_t* q_rows, const int32_t* kv_rows, const int32_t* o_rows, int dtype,
                         int BH, int S, int D, long long row_stride, long long head_stride,
                         long long o_row_stride, long long o_head_stride, float sm_scale,
                         const svgb_plan* plan, const void* plan_ws, void* stream);

/* FP8 attention (BASELINE config 5; the reference has no FP8 sparse attention, README.md:117).  q8,k8,v8 are
 * e4m3 bytes [BH,S,D] (D = 128) with per-head dequantisation scales (x ~= x8 * scale[h]); QK^T and PV run as
 * tcgen05 kind::f8f6f4 MMAs, P is quantised to e4m3 in TMEM with its range kept in (0, 2^8] (offset 2^4 + lazy-rescale slack 2^4), the output is
 * bf16.  Any plan kind except the row-gather one.  svgb_quantize_e4m3 produces the inputs: per-head absmax
 * scaling (scale = absmax / 448), round-to-nearest, saturating. */
int svgb_quantize_e4m3(const void* x, int dtype, void* x8, float* scale, int BH, int S, int D, void* stream);
int svgb_attn_fwd_fp8(const void* q8, const void* k8, const void* v8, const float* q_scale, const float* k_scale,
                      const float* v_scale, void* o, float* lse, const int32_t* o_rows, int BH, int S, int D,
                      long long row_stride, long long head_stride, long long o_row_stride,
                      long long o_head_stride, float sm_scale, const svgb_plan* plan, const void* plan_ws,
                      void* stream);

/* density of a variable-block map (density_calculation, svg/kmeans_utils.py:13-31) -> float [BH] */
int svgb_density(const uint8_t* map, const int32_t* row_sz, const int32_t* col_sz, int BH, int QC,
                 int KC, float* density, void* stream);

/* ---- layout transforms ------------------------------------------------------------------ */
/* stable ascending argsort of labels in [0,K): perm int32 [BH,S] (sorted_indices of
 * permute_tensor_by_labels_triton, permute.py:113) and counts int32 [BH,K] (cluster sizes). */
int svgb_argsort_labels_bytes(int BH, int S, int K, size_t* bytes);
int svgb_argsort_labels(const int32_t* labels, int BH, int S, int K, int32_t* perm, int32_t* counts,
                        void* ws, size_t ws_bytes, void* stream);
/* out[h,s,:] = in[h,perm[h,s],:]   (_permute_kernel, permute.py:12-43);  16-bit elements */
int svgb_permute_gather(const void* in, const int32_t* perm, void* out, int BH, int S, int D,
                        void* stream);
/* out[h,perm[h,s],:] = in[h,s,:]   (_inverse_permute_kernel, permute.py:46-75) */
int svgb_permute_scatter(const void* in, const int32_t* perm, void* out, int BH, int S, int D,
                         void* stream);
/* SVG1 head placement (hunyuan_sparse_head_placement, hyvideo/placement.py:34-153): for heads with
 * best_mask_idx==1 the video part goes frame-major -> token-major (dst = patch*F + frame), text
 * untouched; other heads are copied.  text_first=1 selects the CogVideoX variant
 * (cog/placement.py).  inverse=1 is hunyuan_hidden_states_placement (placement.py:285-387).
 * n_tensors in {1,2,3}: in/out arrays of that many [BH,S,D] 16-bit tensors. */
int svgb_head_placement(const void* const* in, void* const* out, int n_tensors,
                        const int32_t* best_mask_idx, int BH, int S, int D, int ctx, int F, int P,
                        int text_first, int inverse, void* stream);

/* ---- flash k-means (svg/kmeans_utils.py:464-733) ------------------------------------------ */
/* one workspace serves assign / update / run for a given (BH, N, K, D); K <= 4096 */
int svgb_kmeans_bytes(int BH, int N, int K, int D, size_t* bytes);
/* x_sq[h,n] = sum_d round16(x^2): round_result=1 rounds the sum to the 16-bit dtype as
 * (x**2).sum(-1) does (batch_kmeans_Euclid :704); 0 keeps the fp32 sum (assign kernel :531) */
int svgb_row_sqnorm(const void* x, float* x_sq, int BH, int N, int D, int dtype, int round_result,
                    void* stream);
/* nearest centroid per point: argmin_k max(0, |x|^2 + |c_k|^2 - 2 x.c_k), lowest k wins ties;
 * x [BH,N,D], c [BH,K,D] 16-bit; x_sq float [BH,N]; labels int32 [BH,N].
 * (_euclid_assign_kernel :464-554).  X.C^T runs on tcgen05. */
int svgb_kmeans_assign(const void* x, const void* c, const float* x_sq, int32_t* labels, int BH,
                       int N, int K, int D, int dtype, void* ws, size_t ws_bytes, void* stream);
/* centroid update (triton_centroid_update_sorted_euclid :375-421): fp32 mean of members in token
 * order (deterministic; the reference's atomics are not), empty cluster keeps the old centroid,
 * result rounded to the 16-bit dtype.  counts int32 [BH,K]; shift_max (optional) float[1] =
 * max_k |c_new - c_old|_2 evaluated with the reference's 16-bit tensor roundings (:642). */
int svgb_kmeans_update(const void* x, const int32_t* labels, const void* c_old, void* c_new,
                       int32_t* counts, float* shift_max, int BH, int N, int K, int D, int dtype,
                       void* ws, size_t ws_bytes, void* stream);
/* batch_kmeans_Euclid (:684-733) as one stream-ordered sequence without host syncs: up to
 * max_iters x (assign, update), device-side `break` when shift < tol.  Outputs follow the
 * reference exactly: labels / counts from the last executed assignment; centroids = the updated
 * ones unless the loop broke, then the ones that assignment was made against.  n_iter_out:
 * device int32[1] (optional). */
int svgb_kmeans_run(const void* x, const void* init_centroids, int BH, int N, int K, int D, int dtype,
                    int max_iters, float tol, int32_t* labels, void* centroids_out, int32_t* counts,
                    int32_t* n_iter_out, void* ws, size_t ws_bytes, void* stream);
/* Same loop for the caller the reference has (hyvideo/attention.py:584-626 -> :704-716): x may be a strided view
 * (x_head_stride elements between heads, 0 = N*D; rows stay D apart) so the video part of [H, S, D] is clustered in
 * place (the reference's `[:, :, :-context_length]` slice + `.contiguous()`), and perm_out (optional, int32 [BH,N])
 * receives the stable argsort of the returned labels -- what the reference recomputes with torch.argsort in
 * permute_tensor_by_labels_triton (svg/kernels/triton/permute.py:113; hyvideo/attention.py:651-652) -- for free: the
 * last centroid update already built it. */
int svgb_kmeans_run_sorted(const void* x, long long x_head_stride, const void* init_centroids, int BH, int N, int K,
                           int D, int dtype, int max_iters, float tol, int32_t* labels, void* centroids_out,
                           int32_t* counts, int32_t* n_iter_out, int32_t* perm_out, void* ws, size_t ws_bytes,
                           void* stream);

/* ---- mask selection --------------------------------------------------------------------- */
/* identify_dynamic_map (svg/kmeans_utils.py:864-896): qc [BH,QC,D], kc [BH,KC,D] 16-bit,
 * k_sizes int32 [BH,KC] -> map uint8 [BH,QC,KC].  Ties in the descending sort resolve to the
 * lower column index (stable). */
int svgb_dynamic_map(const void* qc, const void* kc, const int32_t* k_sizes, int BH, int QC, int KC,
                     int D, int dtype, float top_p, int preserve, uint8_t* map, void* stream);
/* sample_mse (hyvideo/attention.py:375-399): MSE between full attention and attention under the two
 * profiling masks (0 spatial, 1 temporal) on `n_rows` sampled query rows -> float [2, BH].
 * layout: 0 = HY (text last, band 1.5*P; hyvideo/utils.py:47-93), 1 = WAN (first-frame sink, band 2*P;
 * wan/utils.py:63-110), 2 = COG (text first, band 1.5*P; cog/utils.py:61-88 -- its temporal mask leaves text rows
 * empty: a sampled text row contributes 0 output here where the reference's softmax yields NaN; the Python mirror
 * CogSVG1Core.sample_mse restores the NaN so argmin picks the same head type). */
int svgb_sample_mse_bytes(int BH, int S, int D, int n_rows, size_t* bytes);
int svgb_sample_mse(const void* q, const void* k, const void* v, const int32_t* rows, int n_rows,
                    int BH, int S, int D, int dtype, int layout, int ctx, int F, int P, float* mse,
                    void* ws, size_t ws_bytes, void* stream);

/* ---- pre-attention elementwise chain (the reference's native `_kernels` module) ------- */
/* rms_norm_forward (svg/kernels/csrc/ops.h:52-75): in place on x [m, n] 16-bit, gamma [n]; n in
 * {32,64,128,256}; y = x * rsqrt(mean(x^2) + eps) * gamma, fp32 inside, one rounding. */
int svgb_rms_norm(void* x, const void* gamma, long long m, int n, float eps, int dtype, void* stream);
/* layer_norm_forward (ops.h:20-44): in place, eps fixed at 1e-5 like the reference kernel. */
int svgb_layer_norm(void* x, const void* gamma, const void* beta, long long m, int n, int dtype, void* stream);
/* apply_qk_rope_inplace_cossin{,_txtlast,_complex} (ops.h:77-260): q [B,Hq,S,D], k [B,Hk,S,D] 16-bit,
 * rotated in place on S - len_text rows of every head; interleaved pairs (2i, 2i+1).
 *   SVGB_ROPE_TXT_FIRST          skip the FIRST len_text rows; cos/sin float [S-len_text, D]
 *   SVGB_ROPE_TXT_LAST           skip the LAST  len_text rows; cos/sin float [S-len_text, D]
 *   SVGB_ROPE_COMPLEX_TXT_FIRST  cos/sin float [S-len_text, D/2] (real / imaginary parts), fp64 arithmetic */
#define SVGB_ROPE_TXT_FIRST 0
#define SVGB_ROPE_TXT_LAST 1
#define SVGB_ROPE_COMPLEX_TXT_FIRST 2
int svgb_qk_rope(void* q, void* k, const float* cos_t, const float* sin_t, int B, int Hq, int Hk, int S, int D,
                 int len_text, int mode, int dtype, void* stream);
/* The whole reference chain in one pass (hyvideo/attention.py:253-295, wan/attention.py:100-135):
 *   q_in,k_in,v_in [B, S_in, H*D] (projection outputs; element strides given, so a packed QKV works)
 *   -> unflatten + transpose + contiguous -> QK norm -> RoPE -> rows [out_row0, out_row0+S_in) of
 *   q_out,k_out,v_out [B, H, S_out, D] (so the video and prompt streams of a double block land in one
 *   tensor without torch.cat).  V is only transposed.
 * norm: SVGB_NORM_NONE | SVGB_NORM_RMS_HEAD (gamma [D]) | SVGB_NORM_LAYER (gamma,beta [D], eps 1e-5) |
 *       SVGB_NORM_RMS_HIDDEN (gamma [H*D]: RMS over the full hidden row before the head split, Wan).
 * rope: 0 none | 1 cos/sin float [rope_n, D] | 2 complex float [rope_n, D/2]; tokens
 *       [rope_lo, rope_lo+rope_n) of this input are rotated with table row (token - rope_lo).
 * The norm result is rounded to the 16-bit dtype before RoPE, exactly like the in-place sequence. */
#define SVGB_NORM_NONE 0
#define SVGB_NORM_RMS_HEAD 1
#define SVGB_NORM_LAYER 2
#define SVGB_NORM_RMS_HIDDEN 3
int svgb_qkv_prep(const void* q_in, const void* k_in, const void* v_in, long long in_token_stride,
                  long long in_batch_stride, void* q_out, void* k_out, void* v_out, long long out_head_stride,
                  long long out_batch_stride, int B, int S_in, int H, int D, int out_row0, int norm,
                  const void* gamma_q, const void* gamma_k, const void* beta_q, const void* beta_k, float eps,
                  int rope, const float* cos_t, const float* sin_t, int rope_lo, int rope_n, int dtype,
                  void* stream);

/* ---- Wan transformer-block glue (svg/kernels/triton/{layernorm,modulate,rmsnorm}.py) ---- */
/* rows of N elements (N % 8 == 0, N <= 8192), contiguous; x/y/w dtypes: SVGB_BF16 | SVGB_F16 | SVGB_F32;
 * scale/shift/gate are float [nb, N], row r uses vector r / rows_per_batch (rows_per_batch = 0: one vector). */
/* y = LayerNorm(x) [* w + b] [* (1 + scale) + shift]: triton_layernorm_forward followed by
 * triton_modulate_shift_forward (custom_models.py:37-56) in one pass; w/b and scale/shift optional. */
int svgb_layernorm_modulate(const void* x, int x_dtype, const void* w, const void* b, int w_dtype, float eps,
                            const float* scale, const float* shift, long long rows_per_batch, void* y, int y_dtype,
                            long long rows, int N, void* stream);
/* y = x * rsqrt(mean(x^2) + eps) * w over the full row (triton_rmsnorm_forward) */
int svgb_rmsnorm_hidden(const void* x, int x_dtype, const void* w, int w_dtype, float eps, void* y, int y_dtype,
                        long long rows, int N, void* stream);
/* y = x * (1 + scale) + shift (triton_modulate_shift_forward) */
int svgb_modulate_shift(const void* x, int x_dtype, const float* scale, const float* shift,
                        long long rows_per_batch, void* y, int y_dtype, long long rows, int N, void* stream);
/* y = residual + x * gate (triton_modulate_gate_residual_forward) */
int svgb_gate_residual(const void* residual, int res_dtype, const void* x, int x_dtype, const float* gate,
                       long long rows_per_batch, void* y, int y_dtype, long long rows, int N, void* stream);

/* ---- self tests (debug; exercised by tests/ on the GPU) ----------------------------------- */
/* One 128x128xD tile through the exact descriptor paths the attention kernel uses:
 * s_out[128,128] = q k^T (fp32), o_out[128,D] = bf16(s*p_scale) v (fp32). */
int svgb_selftest_tile(const void* q, const void* k, const void* v, float* s_out, float* o_out,
                       int D, int dtype, float p_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVGB200_H_ */
