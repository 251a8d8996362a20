// Kernels right after the hot path in the training step:
//   * fused AdamW over one flat fp32 range that also re-casts the bf16 operand shadow of the range
//     (torch.optim.AdamW semantics; reference train.py:154-156,227, train_hptune.py:321-325),
//   * multi-tensor builder of the transposed, LayerScale-prescaled bf16 shadows (gamma (.) W)^T used by the
//     input-gradient GEMMs: one launch for every Linear of the model,
//   * class-weighted, label-smoothed cross entropy forward + backward + on-device running metrics in one launch
//     (reference train.py:167-170,225,229-235 and evaluate :77-105): no host synchronisation per step.
// All HBM-bound; AdamW moves 16 B read + 12 B (+2 B shadow) written per parameter.
#include "common.cuh"

namespace tvit {

template <bool kShadow>
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          __nv_bfloat16* __restrict__ shadow, long long n4, long long n,
                                                          float lr, float b1, float b2, float eps, float wd, float inv_bc1,
                                                          float inv_bc2_sqrt, float gscale) {
  const float decay = 1.0f - lr * wd, step = lr * inv_bc1;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gscale;
    mi = b1 * mi + (1.0f - b1) * gi;
    vi = b2 * vi + (1.0f - b2) * gi * gi;
    pi = pi * decay - step * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
  };
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nth) {
    float4 pv = ld4(p + 4 * i), mv = ld4(m + 4 * i), vv = ld4(v + 4 * i);
    const float4 gv = ld4(g + 4 * i);
    upd(pv.x, gv.x, mv.x, vv.x);
    upd(pv.y, gv.y, mv.y, vv.y);
    upd(pv.z, gv.z, mv.z, vv.z);
    upd(pv.w, gv.w, mv.w, vv.w);
    st4(p + 4 * i, pv);
    st4(m + 4 * i, mv);
    st4(v + 4 * i, vv);
    if (kShadow) st4(shadow + 4 * i, pv);
  }
  for (long long i = 4 * n4 + tid; i < n; i += nth) {  // tail (n % 4)
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
    if (kShadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

// out_t[c, r] = row_scale[r] * w[r, c] for every tensor of a descriptor table; one 32x32 tile per block.
template <typename T>
__global__ void shadow_t_multi_kernel(const tvit_shadow_desc* __restrict__ descs, int count) {
  __shared__ float tile[32][33];
  __shared__ int s_t;
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int lo = 0, hi = count - 1;  // last descriptor with tile_begin <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (descs[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_t = lo;
  }
  __syncthreads();
  const tvit_shadow_desc d = descs[s_t];
  const int tix = (int)blockIdx.x - d.tile_begin;
  const int c0 = (tix % d.tiles_x) * 32, r0 = (tix / d.tiles_x) * 32;
  const float* w = d.w;
  const float* rsc = d.row_scale;
  T* out_t = (T*)d.out_t;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float val = 0.f;
    if (r < d.R && c < d.C) {
      val = w[(long long)r * d.C + c];
      if (rsc) val *= rsc[r];
    }
    tile[j][threadIdx.x] = val;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < d.R && c < d.C) Act<T>::st(out_t + (long long)c * d.R + r, tile[threadIdx.x][j]);
  }
}

// One block.  loss = sum_i l_i / sum_i w[y_i],
//   l_i = (1 - eps) w[y_i] (-log p_i[y_i]) + (eps / C) sum_c w[c] (-log p_i[c])          (torch CrossEntropyLoss)
//   dlogits[i, c] = ((1 - eps) w[y_i] (p_ic - [c == y_i]) + (eps / C) (p_ic W - w[c])) / sum_i w[y_i],  W = sum_c w[c]
// metrics (optional, accumulated across calls): macc[0] += loss * B, macc[1] += #(argmax == y), macc[2] += B;
// prob_out[i] = softmax(logits_i)[1], label_out[i] = y_i  (inputs of the epoch-end AUC).
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                       const float* __restrict__ cw, float ls, int B, int C,
                                                       float* __restrict__ loss, float* __restrict__ dlogits,
                                                       float* __restrict__ macc, float* __restrict__ prob_out,
                                                       float* __restrict__ label_out) {
  __shared__ float red[3][8];
  __shared__ float tot[3];
  float W = 0.f;
  for (int c = 0; c < C; ++c) W += cw ? cw[c] : 1.0f;
  float num = 0.f, den = 0.f, correct = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float* z = logits + (long long)i * C;
    const int y = (int)labels[i];
    float mx = z[0];
    int am = 0;
    for (int c = 1; c < C; ++c)
      if (z[c] > mx) { mx = z[c]; am = c; }
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
    const float lse = mx + logf(se);
    const float wy = cw ? cw[y] : 1.0f;
    float smooth = 0.f;
    for (int c = 0; c < C; ++c) smooth += (cw ? cw[c] : 1.0f) * (lse - z[c]);
    num += (1.0f - ls) * wy * (lse - z[y]) + (ls / (float)C) * smooth;
    den += wy;
    correct += (am == y) ? 1.0f : 0.f;
    if (prob_out) prob_out[i] = C > 1 ? expf(z[1] - lse) : 1.0f;
    if (label_out) label_out[i] = (float)y;
  }
  num = warp_sum(num); den = warp_sum(den); correct = warp_sum(correct);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; red[2][threadIdx.x >> 5] = correct;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; c += red[2][i]; }
    tot[0] = a; tot[1] = b; tot[2] = c;
    const float l = a / b;
    if (loss) *loss = l;
    if (macc) { macc[0] += l * (float)B; macc[1] += c; macc[2] += (float)B; }
  }
  __syncthreads();
  if (!dlogits) return;
  const float inv_den = 1.0f / tot[1];
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float* z = logits + (long long)i * C;
    const int y = (int)labels[i];
    float mx = z[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
    const float inv_se = 1.0f / se;
    const float wy = cw ? cw[y] : 1.0f;
    for (int c = 0; c < C; ++c) {
      const float pc = expf(z[c] - mx) * inv_se, wc = cw ? cw[c] : 1.0f;
      dlogits[(long long)i * C + c] =
          ((1.0f - ls) * wy * (pc - (c == y ? 1.0f : 0.f)) + (ls / (float)C) * (pc * W - wc)) * inv_den;
    }
  }
}

}  // namespace tvit

using namespace tvit;

extern "C" int tvit_adamw(float* p, const float* g, float* m, float* v, void* shadow_bf16, long long n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                          tvit_stream_t stream) {
  TVIT_CHECK_ARG(p && g && m && v && step >= 1, "adamw: bad argument");
  if (n == 0) return TVIT_OK;
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0) &&
                   (((uintptr_t)shadow_bf16 & 7u) == 0);
  const long long n4 = vec ? n / 4 : 0;
  const float inv_bc1 = (float)(1.0 / (1.0 - pow((double)beta1, (double)step)));
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  cudaStream_t s = (cudaStream_t)stream;
  if (shadow_bf16)
    adamw_flat_kernel<true><<<(int)blocks, 256, 0, s>>>(p, g, m, v, (__nv_bfloat16*)shadow_bf16, n4, n, lr, beta1, beta2,
                                                        eps, weight_decay, inv_bc1, inv_bc2_sqrt, grad_scale);
  else
    adamw_flat_kernel<false><<<(int)blocks, 256, 0, s>>>(p, g, m, v, nullptr, n4, n, lr, beta1, beta2, eps, weight_decay,
                                                         inv_bc1, inv_bc2_sqrt, grad_scale);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_shadow_t_multi(const tvit_shadow_desc* descs_device, int count, int total_tiles, int dtype,
                                   tvit_stream_t stream) {
  TVIT_CHECK_ARG(descs_device && count > 0 && total_tiles > 0, "shadow_t_multi: bad argument");
  dim3 block(32, 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == TVIT_F32)
    shadow_t_multi_kernel<float><<<total_tiles, block, 0, s>>>(descs_device, count);
  else if (dtype == TVIT_BF16)
    shadow_t_multi_kernel<__nv_bfloat16><<<total_tiles, block, 0, s>>>(descs_device, count);
  else
    return fail(TVIT_ERR_BAD_ARG, "shadow_t_multi: bad dtype %d", dtype);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_ce_loss(const float* logits, const long long* labels, const float* class_weight,
                            float label_smoothing, int B, int C, float* loss, float* dlogits, float* metric_acc,
                            float* prob_out, float* label_out, tvit_stream_t stream) {
  TVIT_CHECK_ARG(logits && labels && B > 0 && C > 0, "ce_loss: bad argument");
  TVIT_CHECK_ARG(loss || metric_acc, "ce_loss: nothing to compute");
  ce_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, labels, class_weight, label_smoothing, B, C, loss, dlogits,
                                                      metric_acc, prob_out, label_out);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}
