/*
 * tvit.h -- C ABI of libtvit_b200.so: the sm_100a kernels behind the Temporal 3D ViT hot path.
 *
 * The reference (anthonylu23/neural-vit) has no native code and no FFI: every op below is reached
 * there through torch.nn modules in temporal_vit/models/model.py.  Each entry point names the
 * reference lines whose arithmetic it replaces.  The host side (neural_vit_b200/model.py) mirrors
 * the reference's Python interface (Temporal3DViTConfig / CONFIGS / Temporal3DViT.forward) and
 * calls these functions through ctypes with raw device pointers + the current CUDA stream.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers owned by the caller (PyTorch's caching
 *     allocator).  The library never frees or retains them past the stream-ordered call.
 *   - every function returns TVIT_OK (0) or an error code; tvit_last_error() gives the message.
 *     Nothing falls back to another implementation silently.
 *   - all launches go to `stream` (a cudaStream_t).  No host synchronisation inside the library.
 *   - functions are re-entrant and thread-safe (autograd calls backward from its own thread).
 *   - "act" tensors use `dtype`: TVIT_F32 (fp32 verification path) or TVIT_BF16 (product path).
 *     The residual stream, statistics, parameters and parameter gradients are always fp32.
 *   - matrices are row-major; `ld*` are row strides in ELEMENTS.
 */
#ifndef TVIT_H_
#define TVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tvit_stream_t; /* cudaStream_t */

enum { TVIT_OK = 0, TVIT_ERR_BAD_ARG = 1, TVIT_ERR_CUDA = 2, TVIT_ERR_UNSUPPORTED = 3 };
enum { TVIT_F32 = 0, TVIT_BF16 = 1 };
/* SIMT = fp32-accurate CUDA-core kernels (verification path, any dtype);
 * TCGEN05 = tcgen05/TMEM/TMA tensor-core kernels (bf16 operands, fp32 accumulate). */
enum { TVIT_ENGINE_SIMT = 0, TVIT_ENGINE_TCGEN05 = 1 };

/* Counter-based dropout descriptor: mask(e) is a pure function of (seed, site, element index e),
 * so backward regenerates the mask forward used.  p == 0 disables dropout.
 * One Philox4x32-7 call covers the 16 elements [16 g, 16 g + 16) with 8 random bits each; the 8-bit threshold of
 * each group is dithered (golden-ratio Weyl sequence over g) so that every element is dropped with probability
 * round(65536 p) / 65536 exactly; kept elements are scaled by 1 / (1 - that probability).
 * Replaces nn.Dropout's ATen Philox stream (model.py:102,104,138,140,224,250). */
typedef struct {
  unsigned long long seed;
  unsigned int site;
  float p;
} tvit_dropout;

const char* tvit_last_error(void);
int tvit_version(void);
/* 0 iff `device` is compute capability 10.x (B200).  Called before the first forward. */
int tvit_device_check(int device);

/* ---------------------------------------------------------------------------------------------
 * GEMM with fused epilogues: C[M,N] = op(A) * op(B)^T, fp32 accumulate.
 *   trans_a == 0: A is [M,K] (K contiguous)      trans_a == 1: A is stored [K,M] (M contiguous)
 *   trans_b == 0: B is [N,K] (K contiguous)      trans_b == 1: B is stored [K,N] (N contiguous)
 * nn.Linear forward is (trans_a,trans_b) = (0,0) with B = weight; the input-gradient GEMM is (0,0)
 * with B = weight^T (kept as a bf16 shadow); the weight-gradient GEMM is (1,1) with
 * A = dY stored [tokens, out] and B = X stored [tokens, in].
 * Replaces: nn.Linear qkv/proj (model.py:101,103,108,116), fc1/fc2 (:136,139,143,146),
 * Conv3d patch embed as im2col GEMM (:197-202,300) and their autograd backward (train.py:226).
 * ------------------------------------------------------------------------------------------- */
enum {
  /* out[T][m,n] = acc + bias[n] */
  TVIT_EPI_STORE = 0,
  /* h = acc + bias[n]; out[T][m,n] = dropout(gelu_erf(h))                         (fc1, :143-145)
   * aux[T][m,n] = dropout_mult(m,n) * gelu_erf'(h): d out / d h, i.e. everything the backward of this site needs, so
   * that neither the pre-activation nor the mask has to be re-derived in backward */
  TVIT_EPI_BIAS_GELU = 1,
  /* out_f32[m,n] = resid[m,n] + row_scale[m / rows_per_group] * gamma[n] * dropout(acc + bias[n])
   * (proj/fc2 + proj_drop/drop2 + LayerScale + DropPath + residual, :116-117,:146-147,:82,:67-71,:176-177) */
  TVIT_EPI_RESIDUAL = 2,
  /* out[T][m,n] = acc * aux[T][m,n], aux as written by BIAS_GELU              (backward of fc1's GELU/drop1) */
  TVIT_EPI_GELU_BWD = 3,
  /* out_f32[m,n] += acc  (atomic; caller zero-fills).  Weight gradients, split along K.  */
  TVIT_EPI_ACCUM_F32 = 4,
  /* row m = b * n_patches + i;  out_f32[(b*(n_patches+1) + 1 + i), n] =
   *   dropout(acc + bias[n] + pos_k[k'][n] + pos_f[f'][n] + pos_t[t'][n])        (:300-313) */
  TVIT_EPI_PATCH_EMBED = 5,
  /* out_f32[m,n] = exp(alpha * acc - row_scale[m])  with row_scale = the row log-sum-exp of tvit_attn_fwd:
   * materialised softmax(q k^T * alpha) tile by tile on the tensor cores (get_attention_maps, model.py:325-350) */
  TVIT_EPI_SOFTMAX_PROBS = 6
};

typedef struct {
  int engine; /* TVIT_ENGINE_* */
  int dtype;  /* operand (and act output) element type */
  int trans_a, trans_b;
  int M, N, K;
  const void* A;
  long long lda;
  const void* B;
  long long ldb;
  int epilogue;
  void* out;
  long long ldo;
  const float* bias; /* [N] or NULL */
  void* aux;         /* BIAS_GELU: d out / d pre-activation (out; NULL = not wanted: an inference forward -- the tcgen05
                        engine then takes no dropout); GELU_BWD: the same tensor (in) */
  long long ldaux;
  const float* resid; /* RESIDUAL */
  long long ldres;
  const float* gamma;     /* RESIDUAL: [N] or NULL (no LayerScale) */
  const float* row_scale; /* RESIDUAL: per-sample DropPath multiplier mask/keep, or NULL */
  int rows_per_group;     /* tokens per sample */
  tvit_dropout drop;      /* element index = m * N + n (PATCH_EMBED: out_row * N + n) */
  const float* pos_k;     /* PATCH_EMBED: factorised positional tables */
  const float* pos_f;
  const float* pos_t;
  int Kp, Fp, Tp;
  int split_k; /* ACCUM_F32: number of K splits, 0 = choose */
  float alpha; /* SOFTMAX_PROBS: score scale (head_dim^-0.5) */
  float* colsum; /* GELU_BWD, optional: colsum[n] += sum_m out[m,n] (fp32, before rounding): the bias gradient of the
                    Linear whose pre-activation gradient this GEMM produces, folded into the epilogue. Caller zero-fills. */
} tvit_gemm_args;

int tvit_gemm(const tvit_gemm_args* args, tvit_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head self-attention over the (trial x frequency x time) token volume, flash style:
 * no B*H*N*N tensor is materialised; only the per-row log-sum-exp is kept for backward.
 *   qkv : [B*N, 3*H*hd] act, columns ordered [q(all heads) | k | v], head h = cols h*hd..(h+1)*hd
 *   out : [B*N, H*hd] act (token-major, heads merged)       lse : [B, H, N] fp32 (natural log)
 * Dropout on the probabilities: element index = ((b*H + h)*N + q)*Np + k, Np = N rounded up to 16.
 * Replaces model.py:108-115 (reshape/permute, q@k^T*scale, softmax, attn_drop, @v, transpose).
 * ------------------------------------------------------------------------------------------- */
/* keepbits (optional, tcgen05 engine with dropout only; NULL = off): a cache of the dropout site's keep flags, one bit
 * per element -- tvit_attn_keepbits_bytes() bytes, layout [(b,h)][q tile][key tile][128 rows][8 groups] of 16-bit fields,
 * bit t / 8+t of a field = keep flag of key 2t / 2t+1 of the 16-key group.  tvit_attn_fwd writes it, tvit_attn_bwd of the same
 * (qkv, drop) reads it instead of running the generator again: same masks, ~8 % fewer instructions in the backward
 * kernel.  The cache is a pure function of (drop, shape); passing NULL to either call changes no result. */
size_t tvit_attn_keepbits_bytes(int engine, int B, int N, int H); /* 0 for engines that do not use the cache */
int tvit_attn_fwd(int engine, int dtype, const void* qkv, void* out, float* lse, int B, int N, int H, int hd,
                  const tvit_dropout* drop, void* keepbits, tvit_stream_t stream);
size_t tvit_attn_bwd_workspace_bytes(int engine, int dtype, int B, int N, int H, int hd);
/* dqkv : [B*N, 3*H*hd] act.  workspace must hold tvit_attn_bwd_workspace_bytes() bytes.
 * dqkv_colsum (optional, fp32 [3*H*hd], caller zero-fills): += column sums of dqkv taken in fp32 before rounding --
 * the gradient of the qkv bias (model.py:101), folded into the kernels that produce dqkv. */
int tvit_attn_bwd(int engine, int dtype, const void* qkv, const void* out, const void* dout, const float* lse,
                  void* dqkv, void* workspace, size_t workspace_bytes, int B, int N, int H, int hd,
                  const tvit_dropout* drop, float* dqkv_colsum, const void* keepbits, tvit_stream_t stream);
/* Formulation of the tcgen05 attention backward kernel (bit 0 / 1: transposed form without / with dropout; bit 2 / 3:
 * whole-tile S / dP MMAs with two issuing warps without / with dropout; 0 = key-half pipelined, the default).  All
 * variants compute the same function with the same dropout masks; the switch exists for A-B timing and for the parity
 * tests, which run every variant.  Sets the mask unless mask < 0 and returns the previous one. */
int tvit_attn_bwd_variant(int mask);
/* probs[b,h,q,k] = softmax(q k^T * hd^-0.5) materialised (interpretability API, model.py:325-350) */
int tvit_attn_probs(int dtype, const void* qkv, float* probs, int B, int N, int H, int hd, tvit_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Bandwidth-bound kernels
 * ------------------------------------------------------------------------------------------- */
/* Tubelet im2col + cast: x fp32 (B,K,F,T) -> cols act [B*n, pk*pf*pt], patch order (dk,df,dt),
 * token order k'*F'*T' + f'*T' + t'.  Replaces the data movement of Conv3d (model.py:197-202,300-303). */
int tvit_im2col(const float* x, void* cols, int dtype, int B, int K, int F, int T, int pk, int pf, int pt,
                tvit_stream_t stream);

/* LayerNorm forward over `rows` rows of length D (eps inside sqrt, biased variance): model.py:165,170,244.
 * x row r starts at x + r * x_row_stride.  y is dense [rows, D].  mean/rstd may be NULL (inference). */
int tvit_ln_fwd(const float* x, long long x_row_stride, const float* weight, const float* bias, void* y, int dtype,
                float* mean, float* rstd, long long rows, int D, float eps, tvit_stream_t stream);

/* LayerNorm backward fused with the residual-gradient add and the preparation of the gradient the
 * preceding residual branch consumes:
 *   dx[r,:]   = (g_res ? g_res[r,:] : 0) + LN'(dy)[r,:]                      (fp32, row stride dx_row_stride)
 *   dweight  += sum_r dy * xhat,  dbias += sum_r dy                          (atomic; caller zero-fills)
 *   if gp:  gp[T][r,:] = dx[r,:] * row_scale[r / rows_per_group] * dropout_mult(r*D + c)
 *           gp_colsum[c] += sum_r gp (fp32, before rounding)                  (atomic; caller zero-fills)
 */
int tvit_ln_bwd(const void* dy, int dtype, const float* x, long long x_row_stride, const float* mean,
                const float* rstd, const float* weight, const float* g_res, float* dx, long long dx_row_stride,
                float* dweight, float* dbias, void* gp, const float* row_scale, int rows_per_group,
                const tvit_dropout* drop, float* gp_colsum, long long rows, int D, tvit_stream_t stream);

/* gp[T][r,c] = g[r,c] * row_scale[r / rows_per_group] * dropout_mult(r*D + c);  colsum[c] += sum_r gp.
 * Gradient entering a residual branch (backward of DropPath/LayerScale-free part + proj_drop/drop2). */
int tvit_branch_grad_prep(const float* g, long long rows, int D, const float* row_scale, int rows_per_group,
                          const tvit_dropout* drop, void* gp, int dtype, float* colsum, tvit_stream_t stream);

/* out[c] += sum_r x[r,c]  (bias gradients).  Caller zero-fills. */
int tvit_colsum(const void* x, int dtype, long long rows, int C, long long ld, float* out, tvit_stream_t stream);

/* Parameter shadows for the tensor-core path: out[r,c] = w[r,c]; out_t[c,r] = row_scale[r] * w[r,c].
 * Either output may be NULL; row_scale may be NULL (== 1). */
int tvit_cast_weight(const float* w, int R, int C, const float* row_scale, void* out, void* out_t, int dtype,
                     tvit_stream_t stream);

/* Finish the gradients of  z = gamma (.) (a W^T + b)  from  G = gp^T a  and  cs = colsum(gp):
 *   dW[r,c] = gamma[r] * G[r,c];  dgamma[r] = sum_c W[r,c] G[r,c] + b[r] cs[r];  db[r] = gamma[r] cs[r].
 * gamma == NULL means no LayerScale (dW = G, db = cs, dgamma untouched).  (model.py:74-82 backward)
 * accumulate != 0: the three outputs are added to (gradient buffers that several backward passes accumulate into,
 * e.g. the flat all-reduce buckets) instead of overwritten. */
int tvit_ls_finalize(const float* G, const float* W, const float* gamma, const float* bias, const float* cs,
                     float* dW, float* dgamma, float* dbias, int R, int C, int accumulate, tvit_stream_t stream);

/* h[b,0,:] = dropout(cls[:])  -- CLS prepend (model.py:309-313); element index = (b*N)*D + c. */
int tvit_cls_rows(const float* cls, float* h, int B, int N, int D, const tvit_dropout* drop, tvit_stream_t stream);

/* Backward of embed: g0 fp32 [B, n+1, D] (gradient of the residual stream at block 0's input)
 *   gtok[T][b*n + i, :] = g0[b, 1+i, :] * dropout_mult      (input of the patch-embed weight-gradient GEMM)
 *   R[i,:] = sum_b of the same (fp32)   dcls[:] (+)= sum_b g0[b,0,:] * dropout_mult
 * (accumulate != 0: dcls and the four outputs of tvit_pos_grad_reduce are added to instead of overwritten)
 * then tvit_pos_grad_reduce folds R [n, D] into dpos_k [Kp,D], dpos_f [Fp,D], dpos_t [Tp,D] and dbias [D]. */
int tvit_embed_bwd_prep(const float* g0, int B, int n, int D, const tvit_dropout* drop, void* gtok, int dtype,
                        float* R, float* dcls, int accumulate, tvit_stream_t stream);
int tvit_pos_grad_reduce(const float* R, int Kp, int Fp, int Tp, int D, float* dpos_k, float* dpos_f,
                         float* dpos_t, float* dbias, int accumulate, tvit_stream_t stream);

/* y = x + alpha * y  style helpers are intentionally absent: everything else is fused above. */

/* Fused AdamW step (torch.optim.AdamW semantics, train.py:154-156,227) over one flat fp32
 * parameter/gradient/moment range; step is 1-based.  If shadow_bf16 != NULL the updated parameters are also written
 * there as bf16 (the forward operand copies of the tensor-core path: no separate cast pass after the update). */
int tvit_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, tvit_stream_t stream);

/* One launch that rebuilds the transposed operand shadows out_t[c,r] = row_scale[r] * w[r,c] (row_scale = LayerScale
 * gamma or NULL) of every Linear weight listed in a DEVICE-resident descriptor table; tile_begin is the prefix sum of
 * ceil(R/32)*ceil(C/32) over the table and total_tiles its total. */
typedef struct {
  const float* w;
  const float* row_scale;
  void* out_t;
  int R, C;
  int tile_begin;
  int tiles_x; /* ceil(C / 32) */
} tvit_shadow_desc;
int tvit_shadow_t_multi(const tvit_shadow_desc* descs_device, int count, int total_tiles, int dtype,
                        tvit_stream_t stream);

/* Class-weighted, label-smoothed cross entropy (torch.nn.CrossEntropyLoss(weight, label_smoothing), mean reduction;
 * train.py:167-170,225) forward AND backward in one launch, plus on-device running metrics so the loop needs no host
 * synchronisation per step (train.py:229-235, evaluate :77-105):
 *   loss[0] = the batch loss; dlogits[B,C] = d loss / d logits (NULL in evaluation);
 *   metric_acc[0] += loss * B, metric_acc[1] += #(argmax == label), metric_acc[2] += B   (NULL to skip)
 *   prob_out[i] = softmax(logits_i)[1], label_out[i] = label_i  (NULL to skip; inputs of the epoch-end AUC).
 * labels are int64; class_weight may be NULL. */
int tvit_ce_loss(const float* logits, const long long* labels, const float* class_weight, float label_smoothing,
                 int B, int C, float* loss, float* dlogits, float* metric_acc, float* prob_out, float* label_out,
                 tvit_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TVIT_H_ */
