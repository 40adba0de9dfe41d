/* vsn_b200 — C ABI of the B200-native Swin-3D / ViT-3D training hot path.
 *
 * Drop-in boundary.  The reference (EloiNavet/ViT-Stability-Neurodegeneration) has no
 * native layer: everything below `model(x)`, `loss.backward()`, `SAM.first_step/second_step` and
 * `EMAModel.update` is eager ATen.  These entry points are what a binding for that path calls instead;
 * each one names the reference code it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every pointer is a raw DEVICE pointer owned by the caller; the library never allocates, frees or
 *     keeps state between calls;  sizes are element counts, leading dimensions are in elements;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no device synchronisation);
 *   - return 0 on success, non-zero on error with a thread-local message from vsn_last_error();
 *     no C++ exception crosses the boundary;
 *   - re-entrant and thread-safe (forward runs on the caller's thread, backward on the autograd thread);
 *   - sm_100a only: vsn_check_device() fails on anything else and there is no fallback path;
 *   - "bf16" is the 16-bit brain float; fp32 is IEEE binary32.  The precision build libvsn_b200_f16.so (same
 *     sources, -DVSN_F16) exports the same entry points with every "bf16" buffer holding IEEE binary16 instead --
 *     TF32's 11 significant bits, the dtype of the reference's autocast path (train/train_transformer.py:1141-1160);
 *     vsn_precision() tells the two apart.
 */
#ifndef VSN_B200_H_
#define VSN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------- */
int vsn_version(void);
/* 0: the 16-bit buffers are bfloat16 (default build); 1: IEEE binary16 (libvsn_b200_f16.so) */
int vsn_precision(void);
const char* vsn_last_error(void);
int vsn_check_device(void);
/* kernel launches issued by the library since it was loaded (bench.py's gpu_launches counter) */
long long vsn_launch_count(void);

/* ---- tcgen05 GEMM with fused epilogue ---------------------------------------------------------
 * out[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) )
 *   A: a_mn == 0 -> stored [M, K] (lda >= K);  a_mn == 1 -> stored [K, M] (lda >= M)
 *   B: b_mn == 0 -> stored [N, K] (ldb >= K);  b_mn == 1 -> stored [K, N] (ldb >= N)
 *   bias [N] fp32 (nullable); act: 0 none, 1 GELU(erf) with the bf16 pre-activation written to aux,
 *   2 multiply by GELU'(aux); resid fp32 [M, ldr] (nullable): out = resid + rowscale * (acc + bias);
 *   row_scale fp32 [M / rows_per_group] (nullable): per-sample DropPath factor; out_kind: 0 bf16 store,
 *   1 fp32 store, 2 fp32 atomic accumulate (required when split_k > 1; split_k = 0 lets the library choose the split).
 *   rowsum_out fp32 [M] (nullable): += sum_k A(m,k), computed on the tensor cores against an all-ones tile --
 *   in a wgrad GEMM (A = dY read transposed) this is the bias gradient of the Linear, so no separate column
 *   reduction of dY is needed.
 * Replaces F.linear / nn.Linear forward, dgrad and wgrad on the path:
 *   models/swin_transformer_3d.py:52-69 (MLP), :154-156,166-171,197 (qkv, proj), :550,571 (reduction),
 *   :527-529,539 (Conv3d k=s=patch as a GEMM); models/vit_3d.py:59-75,102-105,113,125,372. */
int vsn_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, int M, int N, int K,
                  void* out, long long ldo, int out_kind, const float* bias, int act, void* aux, long long ldaux,
                  const float* resid, long long ldr, const float* row_scale, int rows_per_group, float alpha,
                  int split_k, float* rowsum_out, void* stream);

/* ---- LayerNorm --------------------------------------------------------------------------------
 * nn.LayerNorm(C), eps 1e-5: models/swin_transformer_3d.py:236,255,330,373,541,570,693;
 * models/vit_3d.py:69,98,129,371,373,401.  x fp32 rows; y bf16 (y_bf16=1) or fp32; mean/rstd [rows]. */
int vsn_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy,
                      int y_bf16, float* mean, float* rstd, long long rows, int C, float eps, void* stream);
/* dx = LN'(dy) (+ resid_grad); optional bf16 copy of dx scaled per row group (DropPath of the consumer);
 * dgamma/dbeta [C] fp32 (both or neither): += sum_r dy*xhat, += sum_r dy, fused in the same pass over dy and x. */
int vsn_layernorm_bwd(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx, const float* mean,
                      const float* rstd, const float* gamma, const float* resid_grad, long long ldr, float* dx,
                      long long lddx, void* dx_bf16, long long ldb, const float* row_scale, int rows_per_group,
                      float* dgamma, float* dbeta, long long rows, int C, void* stream);
/* dgamma[c] += sum_r dy*xhat, dbeta[c] += sum_r dy (x != NULL); plain column sum into dbeta when x == NULL
 * (the bias gradients of every Linear / Conv3d on the path). */
int vsn_colreduce(const void* dy, long long lddy, int dy_bf16, const float* x, long long ldx, const float* mean,
                  const float* rstd, float* dgamma, float* dbeta, long long rows, int C, void* stream);

/* ---- attention core ---------------------------------------------------------------------------
 * qkv [T, 3*heads*hd] bf16 (q|k|v blocks, head-major), out [T, heads*hd] bf16, lse [S, heads, ceil64(N)] fp32.
 * win=1: Swin-3D window attention; geom = {B, Dp, Hp, Wp, wd, wh, ww, shift_d, shift_h, shift_w, use_mask};
 *   S = B * (Dp/wd)(Hp/wh)(Wp/ww) windows of N = wd*wh*ww tokens gathered from the padded stage grid with the
 *   cyclic shift, relative_position_bias_table [table_len, heads] fp32 and the {0,-100} region mask applied
 *   in registers.  Replaces torch.roll + window_partition + WindowAttention3D core + window_reverse + roll:
 *   models/swin_transformer_3d.py:72-89,132-152,162-196,330-358,463-492.
 * win=0: dense sequences of N tokens (ViT-3D, models/vit_3d.py:129-141); geom/table NULL. */
int vsn_attn_fwd(const void* qkv, void* out, float* lse, int S, int N, int heads, int hd, int win, const int* geom,
                 const float* table, int table_len, float scale, void* stream);
/* delta [S, heads, ceil64(N)] and dbias_dense are caller scratch: dbias_dense is [heads, ceil64(N), ceil64(N)] fp32,
 * zeroed, for windows the tcgen05 kernels do not cover; for win=0 with hd=64 (ViT-3D) it is the fp32 dQ accumulator
 * [T, heads*hd] (any contents); NULL otherwise.  dqkv [T, 3C] bf16 is fully written for every real token;
 * dtable [table_len, heads] fp32 is accumulated.  lse is whatever vsn_attn_fwd wrote for the same arguments. */
int vsn_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                 float* dbias_dense, float* dtable, int S, int N, int heads, int hd, int win, const int* geom,
                 const float* table, int table_len, float scale, void* stream);

/* ---- layout / small kernels -------------------------------------------------------------------- */
/* Non-overlapping patch gather, zero padded: PatchEmbed3D pad + Conv3d operand
 * (models/swin_transformer_3d.py:532-539) and the ViT Rearrange (models/vit_3d.py:365-370), in_channels = 1.
 * in_dtype 0 fp32 / 1 fp16 / 2 bf16; rows are bf16 (out_bf16=1) or fp32. */
int vsn_patch_gather(const void* vol, int in_dtype, void* out, int out_bf16, int B, int D, int H, int W, int pd,
                     int ph, int pw, void* stream);
/* ViT to_patch_embedding head, Rearrange + LayerNorm(P) in one pass (models/vit_3d.py:364-371): y [B*gd*gh*gw, P]
 * (16-bit operand of the embedding GEMM) and mean / rstd per patch, read straight from the volume -- the fp32 patch
 * rows never exist in HBM.  pd*ph <= 1024 threads, pw <= 16.  vsn_patch_ln_param_grad: dgamma / dbeta [P] += of that
 * LayerNorm from dy [., P] (16-bit), with xhat re-gathered from the volume (its input is data: no dx). */
int vsn_patch_ln_fwd(const void* vol, int in_dtype, int B, int D, int H, int W, int pd, int ph, int pw,
                     const float* gamma, const float* beta, void* y, float* mean, float* rstd, float eps, void* stream);
int vsn_patch_ln_param_grad(const void* dy, const void* vol, int in_dtype, int B, int D, int H, int W, int pd, int ph,
                            int pw, const float* mean, const float* rstd, float* dgamma, float* dbeta, void* stream);
/* Copy the overlap of two channels-last fp32 grids, zero-fill the rest: stage pad / crop
 * (models/swin_transformer_3d.py:457-461,508) and their gradients. */
int vsn_grid_copy(const float* src, int sD, int sH, int sW, float* dst, int dD, int dH, int dW, int B, int C,
                  void* stream);
/* PatchMerging 2x2x2 neighbour gather (scatter=0) or its gradient (scatter=1):
 * models/swin_transformer_3d.py:553-568.  x on the padded grid (pD,pH,pW), real extent (rD,rH,rW). */
int vsn_merge_gather(float* x, int pD, int pH, int pW, int rD, int rH, int rW, float* out, int B, int C, int scatter,
                     void* stream);
/* PatchMerging with the gather fused into its LayerNorm (models/swin_transformer_3d.py:553-572: x0..x7 concat ->
 * norm(8C)): y [B*oD*oH*oW, 8C] bf16 = LN(concat of the 2x2x2 neighbours of x), the operand of the reduction GEMM;
 * mean / rstd per merged row.  The merged fp32 row never exists in HBM.  C = 96 or 192 (other widths: vsn_merge_gather +
 * vsn_layernorm_fwd).  The backward writes dx [B*pD*pH*pW, C] at the source tokens inside the real grid only (zero dx
 * first when the padded grid is larger) and accumulates dgamma / dbeta [8C]; dx_bf16 (optional) = dx * row_scale[b]
 * in the 16-bit operand type, for the block backward that consumes it next. */
int vsn_merge_ln_fwd(const float* x, int pD, int pH, int pW, int rD, int rH, int rW, int B, int C, const float* gamma,
                     const float* beta, void* y, float* mean, float* rstd, float eps, void* stream);
int vsn_merge_ln_bwd(const void* dy, const float* x, int pD, int pH, int pW, int rD, int rH, int rW, int B, int C,
                     const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma, float* dbeta,
                     void* dx_bf16, const float* row_scale, void* stream);
/* MixUp of fp16 volumes on the device (dataset/dataset.py:230-286, `sample1*alpha + sample2*(1-alpha)`):
 * out[b] = fp16(fp16(lam[b]*x[b]) + (1-lam[b])*x[perm[b]]) -- the reference's in-place fp16 mul_ then add_;
 * lam [B] fp32 (1 = sample left alone), perm [B] int32; not in place. */
int vsn_mixup_f16(const void* x, void* out, const float* lam, const int* perm, int B, long long elems_per_sample,
                  void* stream);
/* The input step before the path, on the device: monai NormalizeIntensity() of the (optionally
 * mixed) fp16 volume -- (x - mean) / std over the whole image, population std, no division when std == 0 -- which
 * train/train_transformer.py:1729-1752 puts last in every transform chain, after dataset/dataset.py:230-286's MixUp.
 * vsn_volume_stats_f16: stats[b] = {mean, 1/std} of mix(x[b], x[perm[b]]) (lam/perm nullable: no MixUp);
 * scratch = 2*B doubles, zero on entry, zero again on exit.  vsn_mixup_zscore_f16: out[b] = fp16((mix - mean) * rstd);
 * in place only without MixUp. */
int vsn_volume_stats_f16(const void* x, const float* lam, const int* perm, int B, long long elems_per_sample,
                         double* scratch, float* stats, void* stream);
int vsn_mixup_zscore_f16(const void* x, void* out, const float* lam, const int* perm, const float* stats, int B,
                         long long elems_per_sample, void* stream);
/* All test-time-augmentation views of B fp16 volumes in one launch (eval/test_time_augmentation.py:221-354: identity,
 * flip, RandAffine rotations + translations with bilinear sampling and border padding, centre crop + trilinear resize):
 * out[b][v] = trilinear sample of vol[b] at views[v] = {3x4 row-major matrix: output voxel index (d,h,w,1) -> source
 * voxel coordinates; lo[3], hi[3]: the box coordinates and upper neighbours are clamped to (the volume, or the crop)}.
 * Integer-valued maps (identity, flip) copy bit-exactly. */
int vsn_tta_views_f16(const void* vol, void* out, const float* views, int B, int V, int D, int H, int W, void* stream);
int vsn_cast_rows_bf16(const float* src, void* dst, const float* row_scale, int rows_per_group, long long rows, int C,
                       void* stream);
int vsn_cast_bf16(const float* src, void* dst, long long n, void* stream);
/* AdaptiveAvgPool3d(1) over tokens (models/swin_transformer_3d.py:696): backward=0 x[B,T,C] -> out[B,C];
 * backward=1 x = d(out)[B,C] -> out = dx[B,T,C]. */
int vsn_token_mean(const float* x, float* out, int B, int T, int C, int backward, void* stream);
/* Classifier head Linear(F, K) in fp32 (models/swin_transformer_3d.py:752-760; models/vit_3d.py:400-402). */
int vsn_head_fwd(const float* feat, const float* W, const float* bias, float* logits, int B, int K, int F,
                 void* stream);
int vsn_head_bwd(const float* dlogits, const float* feat, const float* W, float* dfeat, float* dW, float* db, int B,
                 int K, int F, void* stream);
/* ViT token assembly: cls token + positional embedding (models/vit_3d.py:447-449) and its gradient. */
int vsn_vit_assemble(const float* emb, const float* cls, const float* pos, float* x, int B, int T, int C,
                     void* stream);
int vsn_vit_assemble_bwd(const float* dx, float* dcls, float* dpos, int B, int T, int C, void* stream);

/* ---- SAM / EMA multi-tensor kernels -------------------------------------------------------------
 * Tensors are described by device arrays: *_ptrs[t] (addresses as int64), sizes[t], and a chunk table
 * (chunk_tensor[c], chunk_off[c]) cutting every tensor into pieces of vsn_mt_chunk_elems() elements. */
int vsn_mt_chunk_elems(void);
/* sq[t] += ||g_t||^2 (adaptive: ||abs(p_t)*g_t||^2): regularization/sam.py:122-145. */
int vsn_mt_sqnorm(const long long* g_ptrs, const long long* p_ptrs, const long long* sizes, const int* chunk_tensor,
                  const long long* chunk_off, int n_chunks, float* sq, int adaptive, void* stream);
/* out2[0] = rho/(norm+1e-12) (0 => skip), out2[1] = norm; skip_flags[t] = tensor norm not finite:
 * regularization/sam.py:46-55,141-155. */
int vsn_sam_scale(const float* sq, int n_tensors, float rho, float* out2, int* skip_flags, void* stream);
/* old = p; p += (p^2 if adaptive) * g * scale: regularization/sam.py:57-72. */
int vsn_mt_sam_perturb(const long long* p_ptrs, const long long* g_ptrs, const long long* old_ptrs,
                       const long long* sizes, const int* chunk_tensor, const long long* chunk_off, int n_chunks,
                       const float* scale_dev, const int* skip_flags, int adaptive, void* stream);
/* dst = src per tensor (SAM restore, regularization/sam.py:86-90; EMA apply/restore, utils/ema.py:110-142). */
int vsn_mt_copy(const long long* dst_ptrs, const long long* src_ptrs, const long long* sizes, const int* chunk_tensor,
                const long long* chunk_off, int n_chunks, void* stream);
/* new_slot = p; ema = w0*s0 + w1*s1 + w2*p (oldest first, s0/s1 nullable): utils/ema.py:72-108. */
int vsn_mt_ema(const long long* p_ptrs, const long long* new_ptrs, const long long* s0_ptrs, const long long* s1_ptrs,
               const long long* ema_ptrs, const long long* sizes, const int* chunk_tensor, const long long* chunk_off,
               int n_chunks, float w0, float w1, float w2, void* stream);

/* AdamW over all parameter tensors in one pass, replacing torch.optim.AdamW(...).step() + optimizer.zero_grad() of
 * train/train_transformer.py:2125-2147,1225-1260 and the GradScaler's unscale_/found_inf skip (:1203-1232) without a
 * host synchronisation.  vsn_adamw_prepare: step += 1 unless *found_inf (nullable) != 0; ctl4 = {1 - beta1^step,
 * sqrt(1 - beta2^step), skip, *inv_scale (nullable: 1)}.  vsn_mt_adamw: g *= inv_scale; p *= 1 - lr*weight_decay[t];
 * m += (g - m)(1 - beta1); v = beta2 v + (1 - beta2) g^2 (1 - beta passed in, rounded from double as torch does); p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps); g = 0 if
 * zero_grad.  Nothing is touched when ctl4[2] != 0. */
int vsn_adamw_prepare(float* step, const float* found_inf, const float* inv_scale, float beta1, float beta2, float* ctl4,
                      void* stream);
int vsn_mt_adamw(const long long* p_ptrs, const long long* g_ptrs, const long long* m_ptrs, const long long* v_ptrs,
                 const long long* sizes, const int* chunk_tensor, const long long* chunk_off, int n_chunks,
                 const float* weight_decay, const float* ctl4, float lr, float beta2, float one_minus_beta1,
                 float one_minus_beta2, float eps, int zero_grad, void* stream);

/* dst_bf16[t] = bf16(src_fp32[t]): refreshes the bf16 copies of the GEMM weights once per forward. */
int vsn_mt_cast_bf16(const long long* src_ptrs, const long long* dst_ptrs, const long long* sizes,
                     const int* chunk_tensor, const long long* chunk_off, int n_chunks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSN_B200_H_ */
