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
  DIG_EPI_GELU = 1,     /* pre = alpha*acc + bias ; aux[m,n] (bf16) = pre ; v = gelu_erf(pre)   (F:54-55)   */
  DIG_EPI_GELU_BWD = 2, /* v = alpha*acc * gelu_erf'(aux[m,n])  (aux: bf16 pre-activation)                 */
  DIG_EPI_RELU_MASK = 3 /* v = aux[m,n] > 0 ? alpha*acc : 0     (aux: bf16 post-ReLU activation)           */
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
  int32_t split_k;                                 /* >1: K is split over CTAs, fp32 atomicAdd into out (out_fp32 must be 1,
                                                      DIG_EPI_LINEAR without bias/residual); out must be pre-initialised */
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

#ifdef __cplusplus
}
#endif
#endif /* DIG_B200_H_ */
