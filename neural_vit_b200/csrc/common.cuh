// Shared device/host helpers for the Temporal 3D ViT sm_100a kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/tvit.h"

namespace tvit {

// ---------------------------------------------------------------------------------------------
// error reporting (tvit_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define TVIT_CHECK_ARG(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return ::tvit::fail(TVIT_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define TVIT_CUDA_OK(expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::tvit::fail(TVIT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                                 \
  } while (0)

#define TVIT_LAUNCH_OK()                                                                          \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess)                                                                        \
      return ::tvit::fail(TVIT_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                                    \
  } while (0)

int num_sms();

// ---------------------------------------------------------------------------------------------
// activation element types: float (verification path) and bf16 (product path)
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Act;
template <>
struct Act<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <>
struct Act<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 4 consecutive elements (16 B fp32 / 8 B bf16)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  uint2 r;
  r.x = pack_bf16(v.x, v.y);
  r.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// single-MUFU 2^x (ex2.approx.ftz): softmax exponentials of the attention kernels
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exact (erf) GELU and its derivative -- nn.GELU() default, reference model.py:137
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast GELU pair for the bf16 tensor-core epilogues: erfc by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7,
// two MUFU ops) -- far inside bf16 resolution; the fp32 verification path keeps erff above.
//   phi_cdf(x) = 0.5 erfc(-x / sqrt 2);   gelu = x * cdf;   gelu' = cdf + x * pdf
__device__ __forceinline__ void gelu_fast_parts(float x, float& cdf, float& pdf) {
  // t = 1 / (1 + p |x| / sqrt2);  e = exp(-x^2/2);  0.5 erfc(|x|/sqrt2) = t (a1/2 + t (a2/2 + ...)) e
  const float t = rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.0f));
  const float e = ex2_approx(x * x * (-0.5f * 1.4426950408889634f));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float half_erfc = poly * t * e;
  cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  pdf = 0.39894228040143268f * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float c, p;
  gelu_fast_parts(x, c, p);
  return x * c;
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float c, p;
  gelu_fast_parts(x, c, p);
  return fmaf(x, p, c);
}
template <typename T> __device__ __forceinline__ float gelu_t(float x) { return sizeof(T) == 2 ? gelu_fast(x) : gelu_f(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) {
  return sizeof(T) == 2 ? gelu_grad_fast(x) : gelu_grad_f(x);
}

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG: Philox4x32-7, 16 random bits per element, 8 elements per call.
// keep(e) <=> bits16(e) >= thr16, thr16 = round(p * 65536).  A mask is a pure function of
// (seed, site, element index), so forward and backward regenerate it without storing it.
// ---------------------------------------------------------------------------------------------
struct DropCfg {
  unsigned long long seed;
  unsigned int site;   // unique per dropout call site within one forward
  unsigned int thr16;  // 0 => dropout disabled (identity)
  float inv_keep;      // 1 / (1 - p)
  // Philox round keys (k0 + r W0, k1 + r W1), r = 0..6, filled on the host: the struct is a kernel parameter,
  // so the rounds read them as constant-bank operands instead of recomputing the key schedule per call
  unsigned int rk0[7], rk1[7];
};

__host__ __device__ __forceinline__ DropCfg make_drop(const tvit_dropout* d) {
  DropCfg c;
  c.seed = d ? d->seed : 0ull;
  c.site = d ? d->site : 0u;
  for (int r = 0; r < 7; ++r) {
    c.rk0[r] = (unsigned int)c.seed + (unsigned int)r * 0x9E3779B9u;
    c.rk1[r] = (unsigned int)(c.seed >> 32) + (unsigned int)r * 0xBB67AE85u;
  }
  float p = d ? d->p : 0.f;
  if (p <= 0.f) {
    c.thr16 = 0;
    c.inv_keep = 1.f;
  } else {
    unsigned int t = (unsigned int)(p * 65536.0f + 0.5f);
    if (t > 65535u) t = 65535u;
    c.thr16 = t;
    c.inv_keep = 1.0f / (1.0f - (float)t * (1.0f / 65536.0f));
  }
  return c;
}

__device__ __forceinline__ uint4 philox4x32_7(unsigned long long seed, unsigned long long ctr, unsigned int site) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = site, c3 = 0x5eed5eedu;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// same generator with the host-precomputed round keys of a DropCfg kernel parameter
__device__ __forceinline__ uint4 philox4x32_7(const DropCfg& c, unsigned long long ctr) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = c.site, c3 = 0x5eed5eedu;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ c.rk0[r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ c.rk1[r];
    c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// random 16-bit lanes for the 8 consecutive elements [8*g, 8*g+8)
__device__ __forceinline__ void drop_bits8(const DropCfg& c, unsigned long long group, uint32_t out[4]) {
  uint4 r = philox4x32_7(c, group);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  out[3] = r.w;
}
// multiplier (0 or 1/(1-p)) for a single element index
__device__ __forceinline__ float drop_mult(const DropCfg& c, unsigned long long e) {
  if (c.thr16 == 0) return 1.0f;
  uint32_t w[4];
  drop_bits8(c, e >> 3, w);
  const unsigned int j = (unsigned int)(e & 7ull);
  const unsigned int k = j >> 1;  // selects instead of a dynamic index: keeps w[] out of local memory
  const uint32_t ww = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
  const uint32_t bits = (ww >> ((j & 1u) * 16u)) & 0xffffu;
  return bits >= c.thr16 ? c.inv_keep : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// Attention-probability dropout (the N x N site): 16 elements per Philox call, 8 random bits per element.
// An 8-bit compare alone would quantise p to 1/256, so the threshold of each 16-element group is dithered:
//   T_g = (thr16 >> 8) + [weyl8(g, seed) < (thr16 & 255)],   keep(e) <=> byte(e) >= T_g,
// where weyl8 is the top byte of a golden-ratio Weyl sequence over the group index.  With the dither counted as
// part of the generator, every element is dropped with probability exactly thr16 / 65536 (the same marginal as the
// 16-bit scheme used by the GEMM epilogues); elements of one group share T_g, a correlation of ~4e-5.
// Element index: row-major over (b, h, q, k) with the k extent padded to a multiple of 16 so that groups never
// straddle rows; forward, backward and the SIMT verification path all use these helpers.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long attn_drop_row_base(int b, int H, int h, int N, int q) {
  const unsigned long long npad = (unsigned long long)((N + 15) & ~15);
  return (((unsigned long long)b * H + h) * N + q) * npad;
}
__device__ __forceinline__ uint32_t attn_drop_thr8(const DropCfg& c, unsigned long long group16) {
  const uint32_t weyl = ((uint32_t)group16 * 0x9E3779B1u + (uint32_t)c.seed) >> 24;
  return (c.thr16 >> 8) + (weyl < (c.thr16 & 255u) ? 1u : 0u);
}
// random bytes for the 16 consecutive elements [16 g, 16 g + 16): element j is byte (j & 3) of word j >> 2
__device__ __forceinline__ void attn_drop_bits16(const DropCfg& c, unsigned long long group16, uint32_t out[4]) {
  uint4 r = philox4x32_7(c, group16);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  out[3] = r.w;
}
// keep flag of a single element (SIMT verification path)
__device__ __forceinline__ bool attn_drop_keep(const DropCfg& c, unsigned long long e) {
  if (c.thr16 == 0) return true;
  uint32_t w[4];
  attn_drop_bits16(c, e >> 4, w);
  const unsigned int j = (unsigned int)(e & 15ull), k = j >> 2;
  const uint32_t ww = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
  return ((ww >> ((j & 3u) * 8u)) & 0xffu) >= attn_drop_thr8(c, e >> 4);
}

}  // namespace tvit
