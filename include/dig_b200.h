/* dig_b200 C-ABI (boundary B3).
 *
 * The reference (ayumiymk/DiG) is pure Python on PyTorch and has no FFI of its own: every entry point
 * below replaces an ATen library call that the reference's pre-training hot path makes, cited as
 *   M = modeling_pretrain_moco_mim_ori.py, V = modeling_pretrain_vit.py, F = modeling_finetune.py,
 *   E = engine_for_pretraining_moco.py, U = utils/utils.py, A = custom_optim/_functional.py.
 *
 * Conventions: plain pointers to DEVICE memory + sizes, no ownership transfer, nothing allocates;
 * every call enqueues on `stream` (a cudaStream_t passed as void*) and returns 0, or a negative code
 * with a message retrievable through dig_last_error().  bf16 = 16-bit brain float, row-major tensors.
 */
#ifndef DIG_B200_H_
#define DIG_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- probes ------------------------------------------------------------------------------------ */
int dig_version(void);            /* ABI version */
int dig_sm(void);                 /* compute capability of the current device * 10 (100 on B200), <0 on error */
const char* dig_last_error(void); /* thread-local message of the last failing call */

/* ---- dense contraction on tcgen05 tensor cores ---------------------------------------------------
 * out[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
 * Replaces F.linear / nn.Linear forward (F:93,119 F:54,58 M:463-482 M:422-426), its dgrad and wgrad
 * (autograd of the same lines), and the 4x4 patch-embed conv as an im2col GEMM (F:188-195).
 * Operands are bf16, accumulation fp32 in TMEM.  An operand is "K-major" when its reduction index is
 * the contiguous one (A stored [M,K], B stored [N,K]) and "MN-major" otherwise (A stored [K,M],
 * B stored [K,N]); leading dimensions are in elements and must be multiples of 8.               */
enum {
  DIG_EPI_LINEAR = 0,   /* v = alpha*acc (+bias[n]) (row-masked replace) (+residual)                       */
  DIG_EPI_GELU = 1,     /* pre = alpha*acc + bias ; aux[m,n] = pre if aux != NULL ; v = gelu_erf(pre) (F:54-55)
                           aux is bf16, or, with aux_q8 != 0, the 8-bit code of pre (see dig_gemm_t.aux_q8)            */
  DIG_EPI_GELU_BWD = 2, /* v = alpha*acc * gelu_erf'(aux[m,n])  (aux: bf16 pre-activation, or its 8-bit code)       */
  DIG_EPI_RELU_MASK = 3,/* v = aux[m,n] > 0 ? alpha*acc : 0     (aux: bf16 post-ReLU activation)           */
  DIG_EPI_ROWDOT = 5    /* v = alpha*acc + bias (bf16 out) ; rowdot[m, n/64] = sum over the 64-column group of v * aux[m,n]
                           (aux: bf16 [M,N]).  Output-projection dgrad: v = dO, aux = O, rowdot = D of the attention backward */
};
typedef struct dig_gemm {
  int64_t M, N, K;
  const void* A; int64_t lda; int32_t a_mn_major;
  const void* B; int64_t ldb; int32_t b_mn_major;
  void* out; int64_t ldo; int32_t out_fp32;        /* 0: bf16, 1: fp32 */
  const float* bias;                               /* [N] or NULL */
  const float* residual; int64_t ldr;              /* fp32 [*, ldr] or NULL; added after everything else */
  int64_t res_row_mod;                             /* >0: residual row index = m % res_row_mod (position table, V:99) */
  const uint8_t* row_mask; const float* row_mask_value; /* if row_mask[m]: v = row_mask_value[n] (mask token, V:95-97) */
  int32_t epilogue;
  void* aux; int64_t ldaux;
  float alpha;
  int32_t split_k;                                 /* >1: K is split over CTAs, fp32 accumulate into out (out_fp32 must be 1,
                                                      DIG_EPI_LINEAR without bias/residual); out must be pre-initialised.
                                                      <0: same, the library picks the tile shape and the number of K slices
                                                      that fill the SMs (weight gradients; out zero- or gradient-initialised) */
  float* colsum;                                   /* optional fp32 [N]: colsum[n] += sum_m out[m,n] (bias gradient of the layer
                                                      that produced the GEMM input), NULL to skip                        */
  float* rowdot; int64_t ldrowdot;                 /* DIG_EPI_ROWDOT: fp32 [M, ldrowdot >= N/64]                            */
  int32_t aux_q8;                                  /* DIG_EPI_GELU / DIG_EPI_GELU_BWD: aux is uint8 [M, ldaux] holding
                                                      code = round((clamp(pre, -4, 4) + 4) * 255 / 8) instead of bf16 pre: the
                                                      backward only needs gelu_erf'(pre), a smooth function of pre (|slope| <=
                                                      0.8), so 256 levels of 0.031 reproduce it to <= 1.3e-2 (3.6e-3 rms) with a
                                                      quarter of the bytes of a bf16 copy written and read (ABI v3)           */
} dig_gemm_t;
int dig_gemm(const dig_gemm_t* g, void* stream);

/* ---- fused attention over the 256 patch tokens, head_dim 64 --------------------------------------
 * Replaces F:97-118 (q*scale, q@k^T, softmax, attn@v) and its autograd backward.
 * qkv  : bf16 [num_seqs*256, 3*heads*64]  (q | k | v; head h = columns h*64..h*64+63 of each third, F:93-95)
 * out  : bf16 [num_seqs*256, heads*64]    context, token-major (the layout F:118 transposes back to)
 * lse  : fp32 [num_seqs, heads, 256]      natural-log-sum-exp of the scaled scores (may be NULL in forward)
 * p_in_smem: 0 = probabilities fed back to the tensor core from TMEM, 1 = from shared memory.      */
int dig_attention_fwd(const void* qkv, void* out, float* lse, int64_t num_seqs, int32_t heads, float scale,
                      int32_t p_in_smem, void* stream);
/* dqkv : bf16 [num_seqs*256, 3*heads*64] gradient w.r.t. qkv given dout (bf16, layout of out).      */
int dig_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      int64_t num_seqs, int32_t heads, float scale, void* stream);
/* Same gradient with D = rowsum(dout o out) per (token, head) supplied by the caller as fp32 [num_seqs*256, heads] (dig_gemm with
 * DIG_EPI_ROWDOT emits it from the output-projection dgrad): persistent kernel, operand tiles prefetched across (sequence, head) items. */
int dig_attention_bwd_d(const void* qkv, const void* dout, const float* lse, const float* dsum, void* dqkv, int64_t num_seqs,
                        int32_t heads, float scale, void* stream);

/* ---- HBM-bound row / column kernels ---------------------------------------------------------------- */
/* 4x4/stride-4 patch extraction for the patch-embed GEMM (F:188-195): images fp32 [n,3,32,128] ->
 * bf16 [n*256, 48], column k = c*16 + kh*4 + kw (the conv-weight order).                           */
int dig_im2col_patch4(const float* images, void* out, int64_t num_images, void* stream);
/* nn.LayerNorm(eps) forward (F:134,140; M:424): x fp32 [rows,d] -> y bf16; optional gelu(y) (M:424-425);
 * mean/rstd fp32 [rows] are saved for the backward (may be NULL).  d % 64 == 0, d <= 512.            */
int dig_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int64_t rows, int32_t d, float eps, int32_t gelu, void* stream);
/* LayerNorm backward: dx = dres + dLN(dy) written as fp32 and/or bf16 (either may be NULL, dx_f32 may alias dres),
 * dgamma/dbeta accumulated (+=); dxsum (optional fp32 [d]) += column sums of dx, i.e. the bias gradient of the
 * Linear that produced the residual branch.  dy is bf16, w.r.t. the LN output (or GELU(LN) output when gelu != 0). */
int dig_layernorm_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const float* beta, const float* dres, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                      float* dxsum, int64_t rows, int32_t d, int32_t gelu, void* stream);
/* The same with the residual gradient carried as a bf16 stream (encoder blocks: the running gradient of the residual stream is the
 * bf16 operand of the next dgrad GEMM anyway, so it is read and written once as bf16 instead of twice as fp32 + once as bf16:
 * 250 instead of 400 bytes per element-row pair).  dx_bf16 = bf16(dres_bf16 + dLN(dy)); statistics and sums stay fp32.      */
int dig_layernorm_bwd_bf16res(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                              const void* dres_bf16, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int64_t rows,
                              int32_t d, void* stream);
/* PatchNet pooling without patch transformer (M:189-193): [S,8,32,d] fp32 -> mean over 8 x (32/num_windows)
 * token windows -> bf16 [S*num_windows, d].  Sequences < split come from x0, the others from x1.     */
int dig_pool_fwd(const float* x0, const float* x1, int64_t split, void* out, int64_t num_seqs, int32_t d,
                 int32_t num_windows, void* stream);
int dig_pool_bwd(const float* dpool, int64_t split, float* dx0, float* dx1, int64_t num_seqs, int32_t d,
                 int32_t num_windows, void* stream);
/* boolean-mask row selection of M:569 as an index gather (fp32 -> bf16) and its backward (+=).       */
int dig_gather_rows(const float* x, const int32_t* idx, void* out, int64_t n, int32_t d, void* stream);
int dig_scatter_add_rows(const float* src, const int32_t* idx, float* dst, int64_t n, int32_t d, void* stream);
/* out[c] += sum_r x[r,c] (bias gradients); with out_sq also out_sq[c] += sum_r x[r,c]^2 (BatchNorm statistics). */
int dig_colsum(const void* x, int32_t x_is_fp32, int64_t ld, float* out, float* out_sq, int64_t rows, int32_t cols,
               void* stream);
/* nn.BatchNorm1d in training mode (M:463-482; SyncBatchNorm under DDP, R:390).  stats = [sum | sumsq] fp32 [2C]
 * over `count` rows -- the caller all-reduces stats across ranks between dig_colsum and dig_bn_apply.       */
int dig_bn_apply(const float* x, const float* stats, float count, const float* gamma, const float* beta, int32_t relu,
                 float eps, void* y_bf16, float* y_f32, int64_t rows, int32_t C, void* stream);
int dig_bn_running(const float* stats, float count, float momentum, float* running_mean, float* running_var,
                   int64_t* num_batches_tracked, int32_t C, void* stream);
/* bstats = [sum dy | sum dy*xhat] (+=), then dx = gamma*rstd*(dy - bstats0/count - xhat*bstats1/count).       */
int dig_bn_bwd_stats(const float* dy, const float* x, const float* stats, float count, float eps, float* bstats,
                     int64_t rows, int32_t C, void* stream);
int dig_bn_bwd_apply(const float* dy, const float* x, const float* stats, const float* bstats, float count,
                     const float* gamma, float eps, void* dx_bf16, float* dx_f32, int64_t rows, int32_t C, void* stream);

/* ---- data-parallel exchanges carried by the kernels over NVLink peer memory (dig_b200/csrc/peer.cu) -------------------
 * Replaces the torch.distributed calls of SyncBatchNorm (R:390: one tiny collective per BatchNorm layer, forward and
 * backward) and concat_all_gather (M:580-591).  Every rank owns a workspace (dig_peer_alloc) whose IPC handle the ranks
 * exchange out of band; `bases` is a HOST array of `world` int64 device addresses: this rank's own workspace at [rank],
 * the mapped peer workspaces elsewhere.  `channel` (0..4) names an independent exchange sequence -- one per stream -- and
 * `epoch` counts the exchanges of that channel from 1; every rank must issue the same sequence per channel.  A wait that
 * times out (a peer died) sets an error word readable with dig_peer_error instead of hanging the GPU.                  */
int dig_peer_workspace_bytes(int64_t key_table_bytes, int64_t* total_out, int64_t* keys_offset_out);
int dig_peer_alloc(int64_t bytes, void** ptr_out, void* handle_out /* 64 bytes */);
int dig_peer_open(const void* handle, void** ptr_out);
int dig_peer_close(void* ptr);
int dig_peer_free(void* ptr);
int dig_peer_error(const int64_t* bases, int32_t world, int32_t rank, int32_t* err_out);
/* buf[0..n) <- sum over ranks, added in rank order (bit-identical on every rank); n <= 8192. */
int dig_peer_allreduce(const int64_t* bases, int32_t world, int32_t rank, int32_t channel, int64_t epoch, float* buf, int32_t n,
                       void* stream);
/* dig_colsum(x fp32 [rows,C]) into stats = [sum | sumsq] (+=) FUSED with the cross-rank sum of stats: the grid's last block
 * pushes the 2C partials to every peer and adds the W partial vectors (SyncBatchNorm forward).                          */
int dig_bn_stats_allreduce(const float* x, float* stats, int64_t rows, int32_t C, const int64_t* bases, int32_t world, int32_t rank,
                           int32_t channel, int64_t epoch, void* stream);
/* dig_bn_bwd_stats fused with the cross-rank sum of bstats (SyncBatchNorm backward); the rank-local sums -- the gradients
 * of the BatchNorm bias and weight -- are written to dbeta_local / dgamma_local (fp32 [C], either may be NULL).         */
int dig_bn_bwd_stats_allreduce(const float* dy, const float* x, const float* stats, float count, float eps, float* bstats,
                               float* dbeta_local, float* dgamma_local, int64_t rows, int32_t C, const int64_t* bases,
                               int32_t world, int32_t rank, int32_t channel, int64_t epoch, void* stream);
/* F.normalize of this rank's keys x [2Q, C] = [k1 ; k2] (M:446-447) FUSED with concat_all_gather (M:580-591): every
 * normalised row is stored into every rank's key table (slot epoch & 1 of the workspace's key area, key_table_bytes per
 * slot) at [half][rank*Q + i][C]; when the kernel completes, the table holds all ranks' keys.  local_copy may be NULL.  */
int dig_peer_l2norm_allgather(const int64_t* bases, int32_t world, int32_t rank, int32_t channel, int64_t epoch, const float* x,
                              float* local_copy, int64_t Q, int32_t C, int64_t key_table_bytes, void* stream);
/* Gradient averaging of the data-parallel step (torch DistributedDataParallel's bucket all-reduce, R:391) over peer memory:
 * grad_bases[r] = device address of rank r's flat fp32 gradient buffer (this rank's own at [rank], the others IPC-mapped;
 * allocate with dig_peer_alloc).  The call averages the `n` floats starting at `offset` (both multiples of 4): rank r reads
 * the r-th slice of that range from all `world` buffers over NVLink, adds them in rank order, scales by 1/world and stores the
 * result into all of them; when the kernel completes, this rank's range holds the average (bit-identical on every rank).
 * world in {2,4,8}; blocks <= 0: one per SM; small_blocks != 0: 128-thread blocks of <= 64 registers that fit next to a
 * resident persistent GEMM / attention CTA (for exchanges overlapped with the backward).  Calls on `channel` (4, used by
 * nothing else) must be stream-ordered and identical on every rank.                                                       */
int dig_peer_grad_allreduce(const int64_t* bases, const int64_t* grad_bases, int32_t world, int32_t rank, int32_t channel,
                            int64_t epoch, int64_t offset, int64_t n, int32_t blocks, int32_t small_blocks, void* stream);

/* elementwise fp32 -> bf16 (n % 4 == 0). */
int dig_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);
/* y (bf16 [rows,d]) = row_mask[r] ? 0 : x[r,:]  -- gradient that reaches the patch-embed GEMM past the mask-token mix (V:95-97). */
int dig_zero_masked_rows(const float* x, const uint8_t* row_mask, void* y, int64_t rows, int32_t d, void* stream);
int dig_zero_masked_rows_bf16(const void* x, const uint8_t* row_mask, void* y, int64_t rows, int32_t d, void* stream);
/* Boolean mask [B,256] -> row indices in boolean-mask order (M:569): idx[b*n_per+j] = b*256 + j-th set position;
 * err[0] = 1 if some sample does not have exactly n_per set bits.                                       */
int dig_mask_to_index(const uint8_t* mask, int32_t* idx, int32_t* err, int32_t B, int32_t n_per, void* stream);

/* ---- GPU-side input stage ------------------------------------------------------------------------
 * transforms.ToTensor + Normalize(0.5, 0.5) of both 32x128 views (dataset/datasets.py:30-37, dataset/dataset_image.py:39-52):
 * uint8 [B,32,128,3] -> fp32 [B,3,32,128] = ((x/255) - 0.5)/0.5, bit-identical to torchvision; RandomGrayscale(gray_p) of the
 * augmented view (dataset_image.py:46, PIL luma) decided per sample by a counter-based draw from (seed, step, sample0 + b).   */
int dig_normalize_views(const uint8_t* img_u8, const uint8_t* aug_u8, float* img_out, float* aug_out, int64_t B, float gray_p,
                        int64_t seed, int64_t step, int64_t sample0, void* stream);
/* RandomMaskingGenerator (masking_generator.py:12-46): per (sample, view) a uniformly random n_mask-subset of the 256 patch
 * positions, as uint8 0/1 [B,V,256] and/or the loader's float64 layout; reproducible from (seed, step, sample0 + b, view).    */
int dig_random_masks(uint8_t* mask_u8, double* mask_f64, int64_t B, int32_t num_view, int32_t n_mask, int64_t seed, int64_t step,
                     int64_t sample0, void* stream);

/* ---- losses (fp32) --------------------------------------------------------------------------------- */
/* F.normalize(x, dim=1) (M:446-447) and its backward, dx = (dy - y<y,dy>) * inv_norm * gscale[0] (gscale may be NULL). */
int dig_l2norm_fwd(const float* x, float* y, float* inv_norm, int64_t rows, int32_t C, void* stream);
int dig_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, const float* gscale, float* dx, int64_t rows,
                   int32_t C, void* stream);
/* Operand split for the InfoNCE logits / gradient on tcgen05 (M:451 torch.einsum('nc,mc->nm')): x fp32 [rows, cols] ->
 * hi = bf16(x), lo = bf16(x - hi).  out_cat bf16 [rows, 3*cols] holds the three parts side by side with lo in part
 * `cat_lo_part` (queries: [hi|hi|lo], lo part 2; keys: [hi|lo|hi], lo part 1), so that dig_gemm over K = 3*cols gives
 * qh.kh + qh.kl + ql.kh = q.k to ~2^-16 relative.  out_stack bf16 [rows/stack_rows, 3, stack_rows, cols] holds the planes
 * (hi, lo, hi) per batch of stack_rows rows: the MN-major B operand of the gradient GEMM.  Either output may be NULL.     */
int dig_split_bf16x3(const float* x, void* out_cat, int32_t cat_lo_part, void* out_stack, int64_t stack_rows, int64_t rows,
                     int32_t cols, void* stream);
/* fp32 CUDA-core GEMM (kept as the in-library reference of the split path; off the step since ABI v2).  Was: InfoNCE logits */
/* fp32 GEMM for the InfoNCE logits einsum('nc,mc->nm') (M:451) and its gradient: C = alpha * A[M,K] . B, with B given
 * as [N,K] (b_is_nk) or [K,N].                                                                        */
int dig_sgemm_f32(const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t b_is_nk, float alpha,
                  void* stream);
/* Row pass of contrastive_loss (M:453-461, M:593-625) on logits[Q,Nk] = q.k^T/T: result[0] += CE*2T (mean over Q),
 * result[1]/[2] += top-1/top-5 accuracy in percent; logits are overwritten with d loss / d (q.k^T).    */
int dig_infonce_rows(float* logits, int32_t Q, int32_t Nk, int64_t label_offset, float T, float* result, void* stream);
/* Masked-pixel MSE (E:85-111 target build + E:141 F.mse_loss): pred fp32 [n_rows,48]; idx[r] = b*256 + token of view 0;
 * images fp32 [B,3,32,128] normalised with mean=std=0.5; loss[0] += mse; dpred (may be NULL) = d mse / d pred.
 * normalize_target != 0: the `normlize_target` branch (E:89-94) -- every colour plane of a patch is standardised over its 16
 * pixels (mean, unbiased variance, (x - mean) / (sqrt(var) + 1e-6)) before the comparison.                                */
int dig_masked_mse(const float* pred, const float* images, const int32_t* idx, float* loss, float* dpred, int64_t n_rows,
                   int32_t normalize_target, void* stream);
int dig_scale_by_device_scalar(const float* x, const float* s, float* y, int64_t n, void* stream);

/* ---- fine-tuning step: transformer decoder pieces ----------------------------------------------------
 * The decoder's Linears run on dig_gemm; these are the small irregular ops around them (fp32 on CUDA cores, T <= 32 queries).
 * x[b*T+t,:] = emb[tok,:] + pos[t,:], tok = <BOS> start_idx at t = 0, targets[b,t-1] after (models/decoder.py:173-178, 212-214).  */
int dig_embed_pos_fwd(const int64_t* targets, const float* emb, const float* pos, float* x, int32_t B, int32_t T, int32_t D,
                      int32_t start_idx, void* stream);
int dig_embed_bwd(const float* dx, const int64_t* targets, float* demb, int32_t B, int32_t T, int32_t D, int32_t start_idx,
                  void* stream);
/* MultiHeadAttention core (models/transformer_layer.py:241-281), head_dim 64, Lq <= 32, Lk <= 256: q/k/v/out are bf16 with
 * head h in columns h*64..h*64+63 (leading dimensions in elements); lse fp32 [B,H,Lq] is saved for the backward; lens != NULL
 * applies get_pad_mask & get_subsequent_mask (transformer_layer.py:433-456; self-attention, Lq == Lk); maps (may be NULL)
 * fp32 [B,Lq,Lk] += attention weights averaged over the heads (vis_attn_maps, :270).                                        */
int dig_dec_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out, int64_t ldo,
                          float* lse, const int64_t* lens, float* maps, int32_t B, int32_t H, int32_t Lq, int32_t Lk, float scale,
                          void* stream);
int dig_dec_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* out,
                          int64_t ldo, const void* dout, int64_t lddo, const float* lse, const int64_t* lens, void* dq, int64_t lddq,
                          void* dk, int64_t lddk, void* dv, int64_t lddv, int32_t B, int32_t H, int32_t Lq, int32_t Lk, float scale,
                          void* stream);
/* SeqCrossEntropyLoss, sample_normalize (loss/seqCrossEntropyLoss.py:47-63): loss[0] += sum over positions t < lens[b] of
 * -log softmax(logits[b,t,:C])[targets[b,t]] / B; dlogits (may be NULL, fp32 [B*T, ldd], zero beyond C and for padded
 * positions) = d loss / d logits; pred (may be NULL) = arg-max class per position (engine_for_finetuning.py:162-164).         */
int dig_seq_cross_entropy(const float* logits, int64_t ld, const int64_t* targets, const int64_t* lens, float* loss, float* dlogits,
                          int64_t ldd, int32_t* pred, int32_t B, int32_t T, int32_t C, void* stream);

/* ---- multi-tensor parameter kernels (pointer tables on the device, see dig_b200/csrc/optim.cu) ------ */
int dig_mt_chunk(void); /* elements per thread block in the blk_* tables */
int dig_mt_cast_bf16(const int64_t* src, const int64_t* dst, const int64_t* numel, const int32_t* blk_tensor,
                     const int32_t* blk_chunk, int32_t num_blocks, void* stream);
int dig_mt_copy_f32(const int64_t* src, const int64_t* dst, const int64_t* numel, const int32_t* blk_tensor,
                    const int32_t* blk_chunk, int32_t num_blocks, void* stream);
/* _update_momentum_encoder (M:428-442): target = m*target + (1-m)*online; shadow (bf16, may be NULL) = target. */
int dig_mt_ema(const int64_t* online, const int64_t* target, const int64_t* shadow, const int64_t* numel,
               const int32_t* blk_tensor, const int32_t* blk_chunk, int32_t num_blocks, float m, void* stream);
/* out[0] += sum of squares over all tensors (get_grad_norm_, U:507-519). */
int dig_mt_sumsq(const int64_t* src, const int64_t* numel, const int32_t* blk_tensor, const int32_t* blk_chunk,
                 int32_t num_blocks, float* out, void* stream);
/* AdamW (custom_optim/_functional.py:115-140) with per-tensor lr / weight_decay, gradient unscale (grad_scale) and
 * optional clip_grad_norm_ (max_norm >= 0, negative = off; sumsq[0] = squared norm of the unscaled grads); refreshes the bf16 shadows
 * (GEMM operands) of the tensors that have one; sumsq_out (may be NULL) += squared norm of the scaled gradients as read
 * by this launch -- get_grad_norm_ (U:507-519) without a second pass when nothing is clipped; guard (may be NULL) is a
 * device scalar (the step's loss): when it is not finite the launch changes nothing (the reference exits before the
 * update, E:148-150).                                                                                             */
int dig_mt_adamw(const int64_t* params, const int64_t* grads, const int64_t* exp_avg, const int64_t* exp_avg_sq,
                 const int64_t* shadow, const int64_t* numel, const float* lr, const float* weight_decay,
                 const int32_t* blk_tensor, const int32_t* blk_chunk, int32_t num_blocks, float beta1, float beta2,
                 float eps, int64_t step, float grad_scale, const float* sumsq, float max_norm, float* sumsq_out,
                 const float* guard, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIG_B200_H_ */
