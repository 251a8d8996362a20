// Shared device/host helpers for the Temporal 3D ViT sm_100a kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
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

// Fast GELU pair for the bf16 tensor-core epilogues.  The normal CDF is evaluated without any MUFU operation as
//   Phi(x) = sat(0.5 + x P(min(x^2, 16))),  P = degree-7 minimax fit of (Phi(x) - 0.5) / x on |x| <= 4
// (|Phi error| <= 5.4e-5, relative Frobenius error of gelu over [-6, 6] 6e-5: far inside bf16 resolution; beyond
// |x| = 4 the saturating FMA clamps to 0 / 1).  The GELU epilogues are bound by issue slots and the 16-lane MUFU
// pipe, which the earlier erfc form (rcp + ex2 per element) loaded twice per element.  The fp32 verification path
// keeps erff above.
__device__ __forceinline__ float phi_cdf_fast(float x, float t) {
  const float tc = fminf(t, 16.0f);
  float p = fmaf(-1.5809006326250596e-09f, tc, 1.2171747698630497e-07f);
  p = fmaf(p, tc, -4.10100710723782e-06f);
  p = fmaf(p, tc, 8.066896407399327e-05f);
  p = fmaf(p, tc, -0.00104821368586272f);
  p = fmaf(p, tc, 0.009664901532232761f);
  p = fmaf(p, tc, -0.0661754161119461f);
  p = fmaf(p, tc, 0.3988475203514099f);
  return __saturatef(fmaf(x, p, 0.5f));
}
__device__ __forceinline__ float gelu_fast(float x) { return x * phi_cdf_fast(x, x * x); }
// gelu'(x) = Phi(x) + x pdf(x),  pdf = exp(-x^2 / 2) / sqrt(2 pi): one ex2
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float t = x * x;
  const float e = ex2_approx(t * (-0.5f * 1.4426950408889634f));
  return fmaf(x * e, 0.39894228040143268f, phi_cdf_fast(x, t));
}
// ---------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 -- two IEEE fp32 operations per issue slot).  The GEMM
// epilogues at K = 384 have a budget of ~12 issued instructions per output element; the GELU
// polynomial alone is 8 FMAs per element in scalar form and 4 in packed form.  Results are bit-identical to the
// scalar fmaf / fmul / fadd per lane.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void up2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint32_t pack_bf16_2(f32x2 v) {  // bf16x2 of an fp32 pair (lo -> low half)
  float lo, hi;
  up2(v, lo, hi);
  return pack_bf16(lo, hi);
}
#ifndef TVIT_SCALAR_F32X2  // A-B builds only: -DTVIT_SCALAR_F32X2 emits the scalar instructions instead
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#else
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  float a0, a1, b0, b1, c0, c1;
  up2(a, a0, a1); up2(b, b0, b1); up2(c, c0, c1);
  return pk2(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  float a0, a1, b0, b1;
  up2(a, a0, a1); up2(b, b0, b1);
  return pk2(a0 * b0, a1 * b1);
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  float a0, a1, b0, b1;
  up2(a, a0, a1); up2(b, b0, b1);
  return pk2(a0 + b0, a1 + b1);
}
#endif
// gelu_fast / gelu_grad_fast of two values at once (same polynomial), both scaled by k (the dropout 1/(1-p)):
// y = k x Phi(x);  d = k (Phi(x) + x pdf(x)) (only when kGrad).
template <bool kGrad, bool kScale>
__device__ __forceinline__ void gelu_fast2(float x0, float x1, float k, f32x2& y, f32x2& d) {
  const f32x2 x = pk2(x0, x1);
  const f32x2 t = mul2(x, x);
  float t0, t1;
  up2(t, t0, t1);
  const f32x2 tc = pk2(fminf(t0, 16.0f), fminf(t1, 16.0f));
  f32x2 p = fma2(pk2(-1.5809006326250596e-09f, -1.5809006326250596e-09f), tc,
                 pk2(1.2171747698630497e-07f, 1.2171747698630497e-07f));
  p = fma2(p, tc, pk2(-4.10100710723782e-06f, -4.10100710723782e-06f));
  p = fma2(p, tc, pk2(8.066896407399327e-05f, 8.066896407399327e-05f));
  p = fma2(p, tc, pk2(-0.00104821368586272f, -0.00104821368586272f));
  p = fma2(p, tc, pk2(0.009664901532232761f, 0.009664901532232761f));
  p = fma2(p, tc, pk2(-0.0661754161119461f, -0.0661754161119461f));
  p = fma2(p, tc, pk2(0.3988475203514099f, 0.3988475203514099f));
  float p0, p1;
  up2(p, p0, p1);
  // k Phi: the scale rides on the saturated CDF, so neither output needs a multiply of its own
  f32x2 phi = pk2(__saturatef(fmaf(x0, p0, 0.5f)), __saturatef(fmaf(x1, p1, 0.5f)));
  if (kScale) phi = mul2(phi, pk2(k, k));
  y = mul2(x, phi);
  if (kGrad) {
    const f32x2 a = mul2(t, pk2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
    float a0, a1;
    up2(a, a0, a1);
    const f32x2 xe = mul2(x, pk2(ex2_approx(a0), ex2_approx(a1)));
    const float ck = kScale ? 0.39894228040143268f * k : 0.39894228040143268f;
    d = fma2(xe, pk2(ck, ck), phi);
  }
}

template <typename T> __device__ __forceinline__ float gelu_t(float x) { return sizeof(T) == 2 ? gelu_fast(x) : gelu_f(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_t(float x) {
  return sizeof(T) == 2 ? gelu_grad_fast(x) : gelu_grad_f(x);
}

// ---------------------------------------------------------------------------------------------
// Counter-based dropout RNG.  A mask is a pure function of (seed, site, element index), so forward and backward
// regenerate it without storing it.  One Philox4x32-7 call serves 16 consecutive elements with 8 random bits
// each.  An 8-bit compare alone would quantise p to 1/256, so the threshold of each 16-element group is dithered:
//   T_g = (thr16 >> 8) + [weyl8(g, seed) < (thr16 & 255)],   keep(e) <=> byte(e) >= T_g,   thr16 = round(65536 p),
// where weyl8 is the top byte of a golden-ratio Weyl sequence over the group index.  With the dither counted as
// part of the generator every element is dropped with probability exactly thr16 / 65536; the elements of one
// group share T_g, a correlation of about 4e-5.
// ---------------------------------------------------------------------------------------------
#ifndef TVIT_PHILOX_ROUNDS
#define TVIT_PHILOX_ROUNDS 7
#endif
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
  for (int r = 0; r < TVIT_PHILOX_ROUNDS; ++r) {
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
  for (int r = 0; r < TVIT_PHILOX_ROUNDS; ++r) {
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
  for (int r = 0; r < TVIT_PHILOX_ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ c.rk0[r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ c.rk1[r];
    c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// group threshold T_g in [0, 256]
__device__ __forceinline__ uint32_t drop_thr8(const DropCfg& c, unsigned long long group16) {
  const uint32_t weyl = ((uint32_t)group16 * 0x9E3779B1u + c.rk0[0]) >> 24;
  return (c.thr16 >> 8) + (weyl < (c.thr16 & 255u) ? 1u : 0u);
}
// random bytes for the 16 consecutive elements [16 g, 16 g + 16): element j is byte (j & 3) of word j >> 2
__device__ __forceinline__ void drop_bits16(const DropCfg& c, unsigned long long group16, uint32_t out[4]) {
#ifdef TVIT_EXPERIMENT_NOPHILOX  // timing experiment only (wrong masks): what the generator itself costs a kernel
  out[0] = (uint32_t)group16 * 0x9E3779B1u; out[1] = out[0] ^ c.rk0[1]; out[2] = out[0] + c.rk1[2]; out[3] = ~out[0];
  return;
#endif
  uint4 r = philox4x32_7(c, group16);
  out[0] = r.x;
  out[1] = r.y;
  out[2] = r.z;
  out[3] = r.w;
}
// keep flag / multiplier (0 or 1/(1-p)) of a single element
__device__ __forceinline__ bool drop_keep(const DropCfg& c, unsigned long long e) {
  if (c.thr16 == 0) return true;
  uint32_t w[4];
  drop_bits16(c, e >> 4, w);
  const unsigned int j = (unsigned int)(e & 15ull), k = j >> 2;
  const uint32_t ww = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));  // selects keep w[] in registers
  return ((ww >> ((j & 3u) * 8u)) & 0xffu) >= drop_thr8(c, e >> 4);
}
__device__ __forceinline__ float drop_mult(const DropCfg& c, unsigned long long e) {
  return drop_keep(c, e) ? c.inv_keep : 0.0f;
}
// multipliers of the four elements e..e+3, e % 4 == 0 (one word of the group)
__device__ __forceinline__ void drop_mult4(const DropCfg& c, unsigned long long e, float m[4]) {
  if (c.thr16 == 0) {
    m[0] = m[1] = m[2] = m[3] = 1.0f;
    return;
  }
  uint32_t w[4];
  drop_bits16(c, e >> 4, w);
  const uint32_t t = drop_thr8(c, e >> 4);
  const unsigned int k = (unsigned int)(e >> 2) & 3u;
  const uint32_t ww = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = (((ww >> (8 * j)) & 0xffu) >= t) ? c.inv_keep : 0.f;
}
// multipliers of the 16 elements e..e+15, e % 16 == 0: one Philox call
__device__ __forceinline__ void drop_mult16(const DropCfg& c, unsigned long long e, float m[16]) {
  uint32_t w[4];
  drop_bits16(c, e >> 4, w);
  const uint32_t t = drop_thr8(c, e >> 4);
#pragma unroll
  for (int j = 0; j < 16; ++j) m[j] = (((w[j >> 2] >> (8 * (j & 3))) & 0xffu) >= t) ? c.inv_keep : 0.f;
}

// prmt.b32 in its generic mode: selector nibble bit 3 replicates the sign bit of the selected byte.  The selector is
// a template argument so that it is encoded as an immediate (as a register operand ptxas re-materialised the
// constant with a UMOV + move in front of every use).
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}
// Keep mask (0xffff per kept bf16 lane) for two dropout elements whose random bytes are bytes (2 HI, 2 HI + 1) of
// w: the bytes are spread into two 16-bit lanes, tgc = 0x80008000 - T_g * 0x10001 is added to both lanes (T_g <= 256,
// so no carry crosses lanes) and bit 15 of a lane ends up set iff byte >= T_g; prmt then smears it over the lane.
__device__ __forceinline__ uint32_t drop_tgc(uint32_t thr8) { return 0x80008000u - thr8 * 0x10001u; }
template <int HI>
__device__ __forceinline__ uint32_t drop_keep_mask2(uint32_t w, uint32_t tgc) {
  const uint32_t x = HI ? prmt<0x4342u>(w, 0u) : prmt<0x4140u>(w, 0u);
  return prmt<0xBB99u>(x + tgc, 0u);
}
// pair masks for the 16 elements of one group: mk[j] covers elements (2j, 2j+1); one Philox call
__device__ __forceinline__ void drop_keep_masks16(const DropCfg& c, unsigned long long e, uint32_t (&mk)[8]) {
  uint32_t w[4];
  drop_bits16(c, e >> 4, w);
  const uint32_t tgc = drop_tgc(drop_thr8(c, e >> 4));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mk[2 * j] = drop_keep_mask2<0>(w[j], tgc);
    mk[2 * j + 1] = drop_keep_mask2<1>(w[j], tgc);
  }
}

// Ragged tail of the tcgen05 attention kernels: when N = 128 m + 1 (the CLS token: N = 2049, 16385, ...), the last key
// (forward) / query (backward) can be handled on the CUDA cores instead of through an almost empty 128-wide tile.
// Measured at the bench shape: forward -4 % without dropout, -2 % with it (profiles/r2_kern_v13_tail.log); backward
// -3.3 % / -3.7 % once the tail work runs on the two idle warps (profiles/r2_kern_v14_tail_bwd.log).  Both default ON.
// TVIT_ATTN_TAIL = bit 0: forward, bit 1: backward (A-B timing / debugging); 0 disables both.
constexpr int kMaxAttnTail = 1;
inline int attn_tail_mask() {
  static const int m = [] { const char* e = getenv("TVIT_ATTN_TAIL"); return e ? atoi(e) : 3; }();
  return m;
}
inline int attn_tail(int N, int pass_bit) {
  const int t = N % 128;
  return ((attn_tail_mask() & pass_bit) && N > 128 && t >= 1 && t <= kMaxAttnTail) ? t : 0;
}

// Attention-probability dropout (the N x N site) element index: row-major over (b, h, q, k) with the k extent
// padded to a multiple of 16 so that groups never straddle rows (forward, backward and the SIMT path agree).
__host__ __device__ __forceinline__ unsigned long long attn_drop_row_base(int b, int H, int h, int N, int q) {
  const unsigned long long npad = (unsigned long long)((N + 15) & ~15);
  return (((unsigned long long)b * H + h) * N + q) * npad;
}

}  // namespace tvit
