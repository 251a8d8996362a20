// Fused GEMM epilogues shared by the tcgen05 GEMM (tc_gemm.cu) and the SIMT verification GEMM
// (simt.cu).  One call handles 4 consecutive output columns of one row from fp32 accumulators.
#pragma once
#include "common.cuh"

namespace tvit {

// Internal epilogue id: BIAS_GELU without the aux output (tvit_gemm_args.aux == NULL, an inference forward).  A template
// instantiation of its own on the tcgen05 path, so that the training kernel keeps its unconditional aux store (a
// predicated store in the shared kernel cost the training fc1 GEMM 2.5 %).
constexpr int kEpiBiasGeluNoAux = 101;
__host__ __device__ constexpr bool is_bias_gelu(int epi) { return epi == TVIT_EPI_BIAS_GELU || epi == kEpiBiasGeluNoAux; }

struct EpiParams {
  int M, N;
  void* out;
  long long ldo;
  const float* bias;
  void* aux;
  long long ldaux;
  const float* resid;
  long long ldres;
  const float* gamma;
  const float* row_scale;
  int rpg;
  DropCfg drop;
  const float* pos_k;
  const float* pos_f;
  const float* pos_t;
  int Kp, Fp, Tp;
  float alpha;  // SOFTMAX_PROBS score scale
  float* colsum;  // GELU_BWD: optional column sums of the output (bias gradient)
  int vec_ok;   // N % 4 == 0 and every leading dimension % 4 == 0 -> 4-wide vector path is legal
  int vec8_ok;  // same with 8 (16-byte bf16 accesses)
  int vec16_ok; // N, leading dimensions % 16 == 0 and 32-byte aligned bases: 256-bit (full-sector) accesses
};

inline EpiParams make_epi_params(const tvit_gemm_args* a) {
  EpiParams p;
  p.M = a->M;
  p.N = a->N;
  p.out = a->out;
  p.ldo = a->ldo;
  p.bias = a->bias;
  p.aux = a->aux;
  p.ldaux = a->ldaux;
  p.resid = a->resid;
  p.ldres = a->ldres;
  p.gamma = a->gamma;
  p.row_scale = a->row_scale;
  p.rpg = a->rows_per_group > 0 ? a->rows_per_group : 1;
  p.drop = make_drop(&a->drop);
  p.pos_k = a->pos_k;
  p.pos_f = a->pos_f;
  p.pos_t = a->pos_t;
  p.Kp = a->Kp;
  p.Fp = a->Fp;
  p.Tp = a->Tp;
  p.alpha = a->alpha;
  p.colsum = a->colsum;
  p.vec_ok = (a->N % 4 == 0) && (a->ldo % 4 == 0) && (a->aux == nullptr || a->ldaux % 4 == 0) &&
             (a->resid == nullptr || a->ldres % 4 == 0);
  p.vec8_ok = (a->N % 8 == 0) && (a->ldo % 8 == 0) && (a->aux == nullptr || a->ldaux % 8 == 0);
  p.vec16_ok = (a->N % 16 == 0) && (a->ldo % 16 == 0) && (a->aux == nullptr || a->ldaux % 16 == 0) &&
               (a->resid == nullptr || a->ldres % 16 == 0) && (((uintptr_t)a->out & 31u) == 0) &&
               (((uintptr_t)a->aux & 31u) == 0) && (((uintptr_t)a->resid & 31u) == 0);
  return p;
}

// four multipliers for e..e+3 with arbitrary alignment
__device__ __forceinline__ void drop_mult4e(const DropCfg& c, unsigned long long e, float m[4]) {
  if (c.thr16 == 0) {
    m[0] = m[1] = m[2] = m[3] = 1.0f;
  } else if ((e & 3ull) == 0ull) {
    drop_mult4(c, e, m);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = drop_mult(c, e + j);
  }
}

// 8 multipliers for e..e+7, e % 8 == 0
__device__ __forceinline__ void drop_mult8(const DropCfg& c, unsigned long long e, float m[8]) {
  if (c.thr16 == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = 1.0f;
    return;
  }
  uint32_t w[4];
  drop_bits16(c, e >> 4, w);
  const uint32_t t = drop_thr8(c, e >> 4);
  const bool hi = ((e >> 3) & 1ull) != 0ull;  // second half of the 16-element group
  const uint32_t w0 = hi ? w[2] : w[0], w1 = hi ? w[3] : w[1];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[j] = (((w0 >> (8 * j)) & 0xffu) >= t) ? c.inv_keep : 0.f;
    m[4 + j] = (((w1 >> (8 * j)) & 0xffu) >= t) ? c.inv_keep : 0.f;
  }
}

// scalar element (used for ragged N, e.g. the 2-class logits)
template <int EPI, typename T>
__device__ __forceinline__ void epi_apply1(const EpiParams& p, int m, int n, float v) {
  if (EPI == TVIT_EPI_STORE) {
    if (p.bias) v += p.bias[n];
    Act<T>::st((T*)p.out + m * p.ldo + n, v);
  } else if (is_bias_gelu(EPI)) {
    if (p.bias) v += p.bias[n];
    const float mlt = drop_mult(p.drop, (unsigned long long)m * p.N + n);
    if (EPI == TVIT_EPI_BIAS_GELU) Act<T>::st((T*)p.aux + m * p.ldaux + n, gelu_grad_t<T>(v) * mlt);
    Act<T>::st((T*)p.out + m * p.ldo + n, gelu_t<T>(v) * mlt);
  } else if (EPI == TVIT_EPI_RESIDUAL) {
    if (p.bias) v += p.bias[n];
    v *= drop_mult(p.drop, (unsigned long long)m * p.N + n);
    if (p.gamma) v *= p.gamma[n];
    if (p.row_scale) v *= p.row_scale[m / p.rpg];
    ((float*)p.out)[m * p.ldo + n] = p.resid[m * p.ldres + n] + v;
  } else if (EPI == TVIT_EPI_GELU_BWD) {
    const float o = v * Act<T>::ld((const T*)p.aux + m * p.ldaux + n);
    Act<T>::st((T*)p.out + m * p.ldo + n, o);
    if (p.colsum) atomicAdd(p.colsum + n, o);  // generic path only; the tcgen05 fast path reduces per warp first
  } else if (EPI == TVIT_EPI_ACCUM_F32) {
    atomicAdd((float*)p.out + m * p.ldo + n, v);
  } else if (EPI == TVIT_EPI_SOFTMAX_PROBS) {
    ((float*)p.out)[m * p.ldo + n] = __expf(fmaf(v, p.alpha, -p.row_scale[m]));
  } else if (EPI == TVIT_EPI_PATCH_EMBED) {
    const int npatch = p.Kp * p.Fp * p.Tp;
    const int b = m / npatch, i = m - b * npatch;
    const int kp = i / (p.Fp * p.Tp), fp = (i / p.Tp) % p.Fp, tp = i % p.Tp;
    const long long orow = (long long)b * (npatch + 1) + 1 + i;
    if (p.bias) v += p.bias[n];
    v += p.pos_k[(long long)kp * p.N + n] + p.pos_f[(long long)fp * p.N + n] + p.pos_t[(long long)tp * p.N + n];
    v *= drop_mult(p.drop, (unsigned long long)orow * p.N + n);
    ((float*)p.out)[orow * p.ldo + n] = v;
  }
}

// 4 consecutive columns n0..n0+3 of row m (n0 % 4 == 0).  Falls back to scalars at a ragged edge.
template <int EPI, typename T>
__device__ __forceinline__ void epi_apply4(const EpiParams& p, int m, int n0, float4 v) {
  if (!(p.vec_ok && n0 + 4 <= p.N)) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + j < p.N) epi_apply1<EPI, T>(p, m, n0 + j, vv[j]);
    return;
  }
  if (EPI == TVIT_EPI_STORE) {
    if (p.bias) {
      const float4 b = ld4(p.bias + n0);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    st4((T*)p.out + m * p.ldo + n0, v);
  } else if (is_bias_gelu(EPI)) {
    if (p.bias) {
      const float4 b = ld4(p.bias + n0);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    float mlt[4];
    drop_mult4e(p.drop, (unsigned long long)m * p.N + n0, mlt);
    if (EPI == TVIT_EPI_BIAS_GELU)
      st4((T*)p.aux + m * p.ldaux + n0,
          make_float4(gelu_grad_t<T>(v.x) * mlt[0], gelu_grad_t<T>(v.y) * mlt[1], gelu_grad_t<T>(v.z) * mlt[2],
                      gelu_grad_t<T>(v.w) * mlt[3]));
    st4((T*)p.out + m * p.ldo + n0,
        make_float4(gelu_t<T>(v.x) * mlt[0], gelu_t<T>(v.y) * mlt[1], gelu_t<T>(v.z) * mlt[2], gelu_t<T>(v.w) * mlt[3]));
  } else if (EPI == TVIT_EPI_RESIDUAL) {
    if (p.bias) {
      const float4 b = ld4(p.bias + n0);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    float mlt[4];
    drop_mult4e(p.drop, (unsigned long long)m * p.N + n0, mlt);
    float rs = p.row_scale ? p.row_scale[m / p.rpg] : 1.0f;
    float4 g = p.gamma ? ld4(p.gamma + n0) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 r = ld4(p.resid + m * p.ldres + n0);
    st4((float*)p.out + m * p.ldo + n0,
        make_float4(r.x + rs * g.x * v.x * mlt[0], r.y + rs * g.y * v.y * mlt[1], r.z + rs * g.z * v.z * mlt[2],
                    r.w + rs * g.w * v.w * mlt[3]));
  } else if (EPI == TVIT_EPI_GELU_BWD) {
    const float4 h = ld4((const T*)p.aux + m * p.ldaux + n0);
    const float4 o = make_float4(v.x * h.x, v.y * h.y, v.z * h.z, v.w * h.w);
    st4((T*)p.out + m * p.ldo + n0, o);
    if (p.colsum) {
      atomicAdd(p.colsum + n0, o.x); atomicAdd(p.colsum + n0 + 1, o.y);
      atomicAdd(p.colsum + n0 + 2, o.z); atomicAdd(p.colsum + n0 + 3, o.w);
    }
  } else if (EPI == TVIT_EPI_ACCUM_F32) {
    float* o = (float*)p.out + m * p.ldo + n0;
    atomicAdd(o + 0, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w);
  } else if (EPI == TVIT_EPI_SOFTMAX_PROBS) {
    const float l = p.row_scale[m];
    st4((float*)p.out + m * p.ldo + n0, make_float4(__expf(fmaf(v.x, p.alpha, -l)), __expf(fmaf(v.y, p.alpha, -l)),
                                                     __expf(fmaf(v.z, p.alpha, -l)), __expf(fmaf(v.w, p.alpha, -l))));
  } else if (EPI == TVIT_EPI_PATCH_EMBED) {
    const int npatch = p.Kp * p.Fp * p.Tp;
    const int b = m / npatch, i = m - b * npatch;
    const int kp = i / (p.Fp * p.Tp), fp = (i / p.Tp) % p.Fp, tp = i % p.Tp;
    const long long orow = (long long)b * (npatch + 1) + 1 + i;
    if (p.bias) {
      const float4 bb = ld4(p.bias + n0);
      v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
    }
    const float4 a = ld4(p.pos_k + (long long)kp * p.N + n0);
    const float4 c = ld4(p.pos_f + (long long)fp * p.N + n0);
    const float4 d = ld4(p.pos_t + (long long)tp * p.N + n0);
    float mlt[4];
    drop_mult4e(p.drop, (unsigned long long)orow * p.N + n0, mlt);
    st4((float*)p.out + orow * p.ldo + n0,
        make_float4((v.x + a.x + c.x + d.x) * mlt[0], (v.y + a.y + c.y + d.y) * mlt[1],
                    (v.z + a.z + c.z + d.z) * mlt[2], (v.w + a.w + c.w + d.w) * mlt[3]));
  }
}

// 8 consecutive columns (n0 % 8 == 0) of row m: 16-byte bf16 accesses for the act-output epilogues of
// the tensor-core path; everything else is two 4-wide calls.
template <int EPI, typename T>
__device__ __forceinline__ void epi_apply8(const EpiParams& p, int m, int n0, const float (&v)[8]) {
  constexpr bool kWide = (sizeof(T) == 2) &&
                         (EPI == TVIT_EPI_STORE || is_bias_gelu(EPI) || EPI == TVIT_EPI_GELU_BWD);
  if (kWide && p.vec8_ok && n0 + 8 <= p.N && !(EPI == TVIT_EPI_GELU_BWD && p.colsum)) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = v[j];
    if (EPI != TVIT_EPI_GELU_BWD && p.bias) {
      const float4 b0 = ld4(p.bias + n0), b1 = ld4(p.bias + n0 + 4);
      x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
      x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
    }
    if (EPI == TVIT_EPI_STORE) {
      uint4 o = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      *reinterpret_cast<uint4*>((__nv_bfloat16*)p.out + m * p.ldo + n0) = o;
      return;
    }
    if (is_bias_gelu(EPI)) {
      float ml[8];
      drop_mult8(p.drop, (unsigned long long)m * p.N + n0, ml);  // vec8_ok: N % 8 == 0 and n0 % 8 == 0
      float d[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = gelu_grad_t<T>(x[j]) * ml[j];
      uint4 h = make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
      if (EPI == TVIT_EPI_BIAS_GELU) *reinterpret_cast<uint4*>((__nv_bfloat16*)p.aux + m * p.ldaux + n0) = h;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = gelu_t<T>(x[j]) * ml[j];
    } else {  // GELU_BWD
      const uint4 h = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)p.aux + m * p.ldaux + n0);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
        x[2 * j] *= f.x;
        x[2 * j + 1] *= f.y;
      }
    }
    uint4 o = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    *reinterpret_cast<uint4*>((__nv_bfloat16*)p.out + m * p.ldo + n0) = o;
    return;
  }
  epi_apply4<EPI, T>(p, m, n0, make_float4(v[0], v[1], v[2], v[3]));
  if (n0 + 4 < p.N) epi_apply4<EPI, T>(p, m, n0 + 4, make_float4(v[4], v[5], v[6], v[7]));
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core epilogue fast path: one thread owns 16 consecutive columns of one row (a tcgen05.ld 32x32b.x16
// sub-chunk).  Everything the epilogue reads is requested BEFORE the TMEM load is waited for, so all latencies
// overlap: bias / LayerScale gamma come from shared memory (staged once per tile by the GEMM kernel, zero /
// one filled when absent) through ld.shared, the per-thread operands of RESIDUAL (fp32 residual row) and
// GELU_BWD (bf16 pre-activation) through 16-byte global loads.
// Requires p.vec8_ok, nc + 16 <= p.N, bf16 activations.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// 256-bit (one full 32-byte sector per thread) global accesses -- sm_100 STG/LDG.E.ENL2.256.  Sixteen-byte stores
// from threads that own different rows reach L2 as half-written sectors and cap the output stream near 1.8 TB/s.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st_bf16x16(__nv_bfloat16* o, const float (&x)[16]) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = pack_bf16(x[2 * j], x[2 * j + 1]);
  st_global_v8(o, v);
}

// Per-thread epilogue operands that come from global memory for one 16-column chunk: the fp32 residual row
// (RESIDUAL, 16 words) or the bf16 pre-activations (GELU_BWD, 8 words).  Split from tc_epi16 so that the caller can
// request chunk u+1 before it processes chunk u (two chunks of loads in flight per thread).
template <int EPI>
__device__ __forceinline__ void tc_epi16_load(const EpiParams& p, int m, int nc, bool row_ok, uint32_t (&ext)[16]) {
  if (EPI == TVIT_EPI_RESIDUAL && row_ok) {
    const float* r = p.resid + m * p.ldres + nc;
    ld_global_v8(r, *reinterpret_cast<uint32_t(*)[8]>(&ext[0]));
    ld_global_v8(r + 8, *reinterpret_cast<uint32_t(*)[8]>(&ext[8]));
  }
  if (EPI == TVIT_EPI_GELU_BWD && row_ok) {
    ld_global_v8((const __nv_bfloat16*)p.aux + m * p.ldaux + nc, *reinterpret_cast<uint32_t(*)[8]>(&ext[0]));
  }
}

// Sum 16 per-thread values over the 32 lanes of the warp with a halving butterfly (62 instructions instead of 16 x 5
// shuffle-adds): afterwards the EVEN lane L holds the total of value index L >> 1 in v[0].
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {  // exchange partner: lane ^ (2 w); keeps w values
    const bool hi = (lane & (2 * w)) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = hi ? v[i] : v[i + w], keep = hi ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * w);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <int EPI, bool kDrop>
__device__ __forceinline__ void tc_epi16(const EpiParams& p, uint32_t s_bias, uint32_t s_gamma, float row_scale, int m,
                                         int nc, uint32_t taddr, bool row_ok, const uint32_t (&ext)[16],
                                         float (&colv)[16]) {
  constexpr bool kBias = (EPI == TVIT_EPI_STORE || is_bias_gelu(EPI) || EPI == TVIT_EPI_RESIDUAL);
  uint32_t acc[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
        "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
        "=r"(acc[14]), "=r"(acc[15])
      : "r"(taddr)
      : "memory");
  float4 b[4], g[4];
  if (kBias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = ld_shared_f4(s_bias + 16 * j);
  }
  if (EPI == TVIT_EPI_RESIDUAL) {
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = ld_shared_f4(s_gamma + 16 * j);
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (EPI == TVIT_EPI_GELU_BWD) {
#pragma unroll
    for (int j = 0; j < 16; ++j) colv[j] = 0.f;  // rows past M contribute nothing to the column sums
  }
  if (!row_ok) return;

  // All arithmetic below is packed fp32x2 (FFMA2 / FMUL2 / FADD2: two fp32 results per issue slot, bit-identical per
  // lane to the scalar instructions): at K = 384 the epilogue has ~12 issue slots per output element before it,
  // not the MMA, sets the pace of the kernel.
  f32x2 x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = pk2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
  if (kBias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[2 * j] = add2(x[2 * j], pk2(b[j].x, b[j].y));
      x[2 * j + 1] = add2(x[2 * j + 1], pk2(b[j].z, b[j].w));
    }
  }
  if (EPI == TVIT_EPI_STORE) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo, hi;
      up2(x[j], lo, hi);
      v[j] = pack_bf16(lo, hi);
    }
    st_global_v8((__nv_bfloat16*)p.out + m * p.ldo + nc, v);
    return;
  }
  if (is_bias_gelu(EPI)) {
    constexpr bool kAux = EPI == TVIT_EPI_BIAS_GELU;  // kEpiBiasGeluNoAux: inference forward, `out` only
    // out = drop(gelu(h)),  aux = dropmask/(1-p) * gelu'(h): the factor the backward GEMM's epilogue multiplies by, so
    // GELU_BWD needs neither the activation derivative nor the mask generator.  bf16 outputs: the keep masks are
    // applied to packed bf16 pairs (one Philox call, then two prmt and an integer subtract per pair).
    uint32_t mk[8];
    if (kDrop) drop_keep_masks16(p.drop, (unsigned long long)m * p.N + nc, mk);  // vec16_ok: N, nc % 16 == 0
    const float ks = kDrop ? p.drop.inv_keep : 1.0f;
    uint32_t v[8], dv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x0, x1;
      up2(x[j], x0, x1);
      f32x2 y, d;
      gelu_fast2<kAux, kDrop>(x0, x1, ks, y, d);
      float y0, y1, d0, d1;
      up2(y, y0, y1);
      up2(d, d0, d1);
      v[j] = pack_bf16(y0, y1);
      dv[j] = pack_bf16(d0, d1);
      if (kDrop) {
        v[j] &= mk[j];
        dv[j] &= mk[j];
      }
    }
    if (kAux) st_global_v8((__nv_bfloat16*)p.aux + m * p.ldaux + nc, dv);
    st_global_v8((__nv_bfloat16*)p.out + m * p.ldo + nc, v);
    return;
  }
  if (EPI == TVIT_EPI_GELU_BWD) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ext[j]));
      float o0, o1;
      up2(mul2(x[j], pk2(f.x, f.y)), o0, o1);
      v[j] = pack_bf16(o0, o1);
      colv[2 * j] = o0;
      colv[2 * j + 1] = o1;
    }
    st_global_v8((__nv_bfloat16*)p.out + m * p.ldo + nc, v);
    return;
  }
  if (EPI == TVIT_EPI_RESIDUAL) {
    // out = resid + (row_scale * gamma) * (dropout multiplier * (acc + bias))
    float ml[16];
    if (kDrop) drop_mult16(p.drop, (unsigned long long)m * p.N + nc, ml);  // vec16_ok: N % 16 == 0 and nc % 16 == 0
    float* o = (float*)p.out + m * p.ldo + nc;
    const f32x2 rs2 = pk2(row_scale, row_scale);
    const float gg[16] = {g[0].x, g[0].y, g[0].z, g[0].w, g[1].x, g[1].y, g[1].z, g[1].w,
                          g[2].x, g[2].y, g[2].z, g[2].w, g[3].x, g[3].y, g[3].z, g[3].w};
    uint32_t ov[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f32x2 v = x[j];
      if (kDrop) v = mul2(v, pk2(ml[2 * j], ml[2 * j + 1]));
      const f32x2 r = fma2(mul2(rs2, pk2(gg[2 * j], gg[2 * j + 1])), v,
                           pk2(__uint_as_float(ext[2 * j]), __uint_as_float(ext[2 * j + 1])));
      float r0, r1;
      up2(r, r0, r1);
      ov[2 * j] = __float_as_uint(r0);
      ov[2 * j + 1] = __float_as_uint(r1);
    }
    st_global_v8(o, *reinterpret_cast<uint32_t(*)[8]>(&ov[0]));
    st_global_v8(o + 8, *reinterpret_cast<uint32_t(*)[8]>(&ov[8]));
  }
}

}  // namespace tvit
