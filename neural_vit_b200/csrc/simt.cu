// CUDA-core (SIMT) kernels: the fp32 *verification path* of the hot path (north_star: "fp32-path loss
// within 1e-4") and the tiny classifier-head GEMMs.  Same epilogues, same C ABI and same host
// orchestration as the tcgen05 engine -- only the contraction is done with fp32 FMAs.
// They are written for clarity, not speed, and are never used by the bf16 product path for the
// token-level GEMMs or attention.
#include "epilogue.cuh"

namespace tvit {

// ---------------------------------------------------------------------------------------------
// C[M,N] = op(A) op(B)^T with arbitrary strides; 64x64x16 tiles, 256 threads, 4x4 micro-tiles.
// A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk].
// ---------------------------------------------------------------------------------------------
template <typename T, int EPI>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const T* __restrict__ A, long long sam, long long sak,
                                                         const T* __restrict__ B, long long sbn, long long sbk, int M,
                                                         int N, int K, EpiParams ep) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * 256;  // 0..1023 = 64 x 16
      int mm, kk;
      if (sak == 1) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? Act<T>::ld(A + gm * sam + gk * sak) : 0.f;
      int nn, kb;
      if (sbk == 1) { kb = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kb = idx >> 6; }
      const int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < K) ? Act<T>::ld(B + gn * sbn + gkb * sbk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    const int n = n0 + tx * 4;
    if (m < M && n < N) epi_apply4<EPI, T>(ep, m, n, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
  }
}

template <typename T>
static int launch_simt_gemm(const tvit_gemm_args* a, cudaStream_t s) {
  const long long sam = a->trans_a ? 1 : a->lda, sak = a->trans_a ? a->lda : 1;
  const long long sbn = a->trans_b ? 1 : a->ldb, sbk = a->trans_b ? a->ldb : 1;
  dim3 grid((a->N + 63) / 64, (a->M + 63) / 64);
  const EpiParams ep = make_epi_params(a);
#define LAUNCH(E)                                                                                          \
  case E:                                                                                                  \
    simt_gemm_kernel<T, E><<<grid, 256, 0, s>>>((const T*)a->A, sam, sak, (const T*)a->B, sbn, sbk, a->M, \
                                                a->N, a->K, ep);                                           \
    break;
  switch (a->epilogue == TVIT_EPI_BIAS_GELU && !a->aux ? kEpiBiasGeluNoAux : a->epilogue) {
    LAUNCH(kEpiBiasGeluNoAux)  // inference forward: no aux output (epilogue.cuh)
    LAUNCH(TVIT_EPI_STORE)
    LAUNCH(TVIT_EPI_BIAS_GELU)
    LAUNCH(TVIT_EPI_RESIDUAL)
    LAUNCH(TVIT_EPI_GELU_BWD)
    LAUNCH(TVIT_EPI_ACCUM_F32)
    LAUNCH(TVIT_EPI_PATCH_EMBED)
    LAUNCH(TVIT_EPI_SOFTMAX_PROBS)
    default:
      return fail(TVIT_ERR_BAD_ARG, "gemm: unknown epilogue %d", a->epilogue);
  }
#undef LAUNCH
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

int simt_gemm(const tvit_gemm_args* a, cudaStream_t s) {
  if (a->dtype == TVIT_F32) return launch_simt_gemm<float>(a, s);
  if (a->dtype == TVIT_BF16) return launch_simt_gemm<__nv_bfloat16>(a, s);
  return fail(TVIT_ERR_BAD_ARG, "gemm: bad dtype %d", a->dtype);
}

// ---------------------------------------------------------------------------------------------
// SIMT attention (verification path).  One warp per (b, h, q); lanes own head dims d = lane + 32*j.
// Online softmax in fp32; dropout applied to the normalised probabilities exactly like
// attn_drop(softmax(.)) in the reference (model.py:112-113).
// ---------------------------------------------------------------------------------------------
template <typename T, int HPL>
__global__ void __launch_bounds__(128) simt_attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                             float* __restrict__ lse, int B, int N, int H, float scale,
                                                             DropCfg drop) {
  const int hd = HPL * 32;
  const int D = H * hd;
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)B * H * N) return;
  const int q = (int)(w % N);
  const int h = (int)((w / N) % H);
  const int b = (int)(w / ((long long)N * H));
  const T* base = qkv + (long long)b * N * 3 * D;
  float qv[HPL], o[HPL];
#pragma unroll
  for (int j = 0; j < HPL; ++j) {
    qv[j] = Act<T>::ld(base + (long long)q * 3 * D + h * hd + lane + 32 * j) * scale;
    o[j] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  const unsigned long long rowe = attn_drop_row_base(b, H, h, N, q);
  for (int k = 0; k < N; ++k) {
    const T* kr = base + (long long)k * 3 * D + D + h * hd;
    const T* vr = base + (long long)k * 3 * D + 2 * D + h * hd;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < HPL; ++j) s += qv[j] * Act<T>::ld(kr + lane + 32 * j);
    s = warp_sum(s);
    const float mn = fmaxf(mx, s);
    const float corr = __expf(mx - mn);
    const float p = __expf(s - mn);
    l = l * corr + p;
    const float pm = drop_keep(drop, rowe + k) ? p * drop.inv_keep : 0.f;
#pragma unroll
    for (int j = 0; j < HPL; ++j) o[j] = o[j] * corr + pm * Act<T>::ld(vr + lane + 32 * j);
    mx = mn;
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int j = 0; j < HPL; ++j)
    Act<T>::st(out + ((long long)b * N + q) * D + h * hd + lane + 32 * j, o[j] * inv);
  if (lane == 0) lse[((long long)b * H + h) * N + q] = mx + __logf(l);
}

// dq/dk/dv accumulated with fp32 atomics into acc [B*N, 3D] (zero-filled by the caller)
template <typename T, int HPL>
__global__ void __launch_bounds__(128) simt_attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ out,
                                                             const T* __restrict__ dout, const float* __restrict__ lse,
                                                             float* __restrict__ acc, int B, int N, int H, float scale,
                                                             DropCfg drop) {
  const int hd = HPL * 32;
  const int D = H * hd;
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)B * H * N) return;
  const int q = (int)(w % N);
  const int h = (int)((w / N) % H);
  const int b = (int)(w / ((long long)N * H));
  const T* base = qkv + (long long)b * N * 3 * D;
  float* abase = acc + (long long)b * N * 3 * D;
  float qv[HPL], dov[HPL], dq[HPL];
  float dsum = 0.f;
#pragma unroll
  for (int j = 0; j < HPL; ++j) {
    const long long oi = ((long long)b * N + q) * D + h * hd + lane + 32 * j;
    qv[j] = Act<T>::ld(base + (long long)q * 3 * D + h * hd + lane + 32 * j);
    dov[j] = Act<T>::ld(dout + oi);
    dsum += dov[j] * Act<T>::ld(out + oi);
    dq[j] = 0.f;
  }
  dsum = warp_sum(dsum);
  const float L = lse[((long long)b * H + h) * N + q];
  const unsigned long long rowe = attn_drop_row_base(b, H, h, N, q);
  for (int k = 0; k < N; ++k) {
    const T* kr = base + (long long)k * 3 * D + D + h * hd;
    const T* vr = base + (long long)k * 3 * D + 2 * D + h * hd;
    float s = 0.f, dpd = 0.f;
    float kv[HPL];
#pragma unroll
    for (int j = 0; j < HPL; ++j) {
      kv[j] = Act<T>::ld(kr + lane + 32 * j);
      s += qv[j] * kv[j];
      dpd += dov[j] * Act<T>::ld(vr + lane + 32 * j);
    }
    s = warp_sum(s) * scale;
    dpd = warp_sum(dpd);
    const float p = __expf(s - L);
    const float mlt = drop_keep(drop, rowe + k) ? drop.inv_keep : 0.f;
    const float ds = p * (dpd * mlt - dsum) * scale;
    const float pd = p * mlt;
#pragma unroll
    for (int j = 0; j < HPL; ++j) {
      dq[j] += ds * kv[j];
      atomicAdd(abase + (long long)k * 3 * D + D + h * hd + lane + 32 * j, ds * qv[j]);
      atomicAdd(abase + (long long)k * 3 * D + 2 * D + h * hd + lane + 32 * j, pd * dov[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < HPL; ++j) abase[(long long)q * 3 * D + h * hd + lane + 32 * j] = dq[j];
}

template <typename T>
__global__ void cast_from_f32_kernel(const float* __restrict__ in, T* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Act<T>::st(out + i, in[i]);
}

// probs[b,h,q,:] = softmax(q k^T * scale): one warp per row, lanes stride over keys
template <typename T>
__global__ void __launch_bounds__(128) attn_probs_kernel(const T* __restrict__ qkv, float* __restrict__ probs, int B,
                                                          int N, int H, int hd, float scale) {
  const int D = H * hd;
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (w >= (long long)B * H * N) return;
  const int q = (int)(w % N);
  const int h = (int)((w / N) % H);
  const int b = (int)(w / ((long long)N * H));
  const T* base = qkv + (long long)b * N * 3 * D;
  const T* qr = base + (long long)q * 3 * D + h * hd;
  float* pr = probs + w * (long long)N;
  float mx = -INFINITY;
  for (int k = lane; k < N; k += 32) {
    const T* kr = base + (long long)k * 3 * D + D + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s += Act<T>::ld(qr + d) * Act<T>::ld(kr + d);
    s *= scale;
    pr[k] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float l = 0.f;
  for (int k = lane; k < N; k += 32) {
    const float e = __expf(pr[k] - mx);
    pr[k] = e;
    l += e;
  }
  l = warp_sum(l);
  const float inv = 1.0f / l;
  for (int k = lane; k < N; k += 32) pr[k] *= inv;
}

template <typename T>
static int simt_attn_fwd_t(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, DropCfg dc,
                           cudaStream_t s) {
  const long long warps = (long long)B * H * N;
  const int grid = (int)((warps + 3) / 4);
  const float scale = 1.0f / sqrtf((float)hd);
  switch (hd / 32) {
    case 1: simt_attn_fwd_kernel<T, 1><<<grid, 128, 0, s>>>((const T*)qkv, (T*)out, lse, B, N, H, scale, dc); break;
    case 2: simt_attn_fwd_kernel<T, 2><<<grid, 128, 0, s>>>((const T*)qkv, (T*)out, lse, B, N, H, scale, dc); break;
    case 4: simt_attn_fwd_kernel<T, 4><<<grid, 128, 0, s>>>((const T*)qkv, (T*)out, lse, B, N, H, scale, dc); break;
    default: return fail(TVIT_ERR_UNSUPPORTED, "attention: head_dim %d not in {32,64,128}", hd);
  }
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

template <typename T>
static int simt_attn_bwd_t(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                           float* acc, int B, int N, int H, int hd, DropCfg dc, cudaStream_t s) {
  const long long warps = (long long)B * H * N;
  const int grid = (int)((warps + 3) / 4);
  const float scale = 1.0f / sqrtf((float)hd);
  const long long total = (long long)B * N * 3 * H * hd;
  TVIT_CUDA_OK(cudaMemsetAsync(acc, 0, total * sizeof(float), s));
  switch (hd / 32) {
    case 1: simt_attn_bwd_kernel<T, 1><<<grid, 128, 0, s>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, acc, B, N, H, scale, dc); break;
    case 2: simt_attn_bwd_kernel<T, 2><<<grid, 128, 0, s>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, acc, B, N, H, scale, dc); break;
    case 4: simt_attn_bwd_kernel<T, 4><<<grid, 128, 0, s>>>((const T*)qkv, (const T*)out, (const T*)dout, lse, acc, B, N, H, scale, dc); break;
    default: return fail(TVIT_ERR_UNSUPPORTED, "attention: head_dim %d not in {32,64,128}", hd);
  }
  TVIT_LAUNCH_OK();
  int g2 = (int)((total + 255) / 256);
  if (g2 > num_sms() * 16) g2 = num_sms() * 16;
  cast_from_f32_kernel<T><<<g2, 256, 0, s>>>(acc, (T*)dqkv, total);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

int simt_attn_fwd(int dtype, const void* qkv, void* out, float* lse, int B, int N, int H, int hd,
                  const tvit_dropout* drop, cudaStream_t s) {
  if (hd % 32 != 0) return fail(TVIT_ERR_UNSUPPORTED, "attention: head_dim %d must be a multiple of 32", hd);
  const DropCfg dc = make_drop(drop);
  if (dtype == TVIT_F32) return simt_attn_fwd_t<float>(qkv, out, lse, B, N, H, hd, dc, s);
  if (dtype == TVIT_BF16) return simt_attn_fwd_t<__nv_bfloat16>(qkv, out, lse, B, N, H, hd, dc, s);
  return fail(TVIT_ERR_BAD_ARG, "attention: bad dtype %d", dtype);
}

size_t simt_attn_bwd_workspace(int B, int N, int H, int hd) { return (size_t)B * N * 3 * H * hd * sizeof(float); }

int simt_attn_bwd(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                  void* ws, size_t ws_bytes, int B, int N, int H, int hd, const tvit_dropout* drop, cudaStream_t s) {
  if (hd % 32 != 0) return fail(TVIT_ERR_UNSUPPORTED, "attention: head_dim %d must be a multiple of 32", hd);
  if (ws_bytes < simt_attn_bwd_workspace(B, N, H, hd) || !ws)
    return fail(TVIT_ERR_BAD_ARG, "attn_bwd: workspace too small (%zu bytes)", ws_bytes);
  const DropCfg dc = make_drop(drop);
  if (dtype == TVIT_F32) return simt_attn_bwd_t<float>(qkv, out, dout, lse, dqkv, (float*)ws, B, N, H, hd, dc, s);
  if (dtype == TVIT_BF16)
    return simt_attn_bwd_t<__nv_bfloat16>(qkv, out, dout, lse, dqkv, (float*)ws, B, N, H, hd, dc, s);
  return fail(TVIT_ERR_BAD_ARG, "attention: bad dtype %d", dtype);
}

int attn_probs(int dtype, const void* qkv, float* probs, int B, int N, int H, int hd, cudaStream_t s) {
  const long long warps = (long long)B * H * N;
  const int grid = (int)((warps + 3) / 4);
  const float scale = 1.0f / sqrtf((float)hd);
  if (dtype == TVIT_F32)
    attn_probs_kernel<float><<<grid, 128, 0, s>>>((const float*)qkv, probs, B, N, H, hd, scale);
  else if (dtype == TVIT_BF16)
    attn_probs_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)qkv, probs, B, N, H, hd, scale);
  else
    return fail(TVIT_ERR_BAD_ARG, "attn_probs: bad dtype %d", dtype);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

}  // namespace tvit
