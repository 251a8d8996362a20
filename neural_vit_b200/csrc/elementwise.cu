// Bandwidth-bound kernels of the Temporal 3D ViT hot path (sm_100a):
// tubelet im2col + cast, LayerNorm fwd/bwd (warp-shuffle reductions, 16-byte accesses), residual-branch
// gradient preparation, bias-gradient column sums, weight shadows, LayerScale gradient finalisation,
// CLS / positional-embedding gradients.  (AdamW, multi-tensor shadows and the loss kernel: optim_loss.cu)
//
// Each kernel is HBM-bound; the algorithmic bytes per element are listed in DESIGN.md section 4.
#include <cstdlib>

#include "common.cuh"

namespace tvit {

static inline int blocks_for(long long work, int per_block, int max_blocks) {
  long long b = (work + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// ---------------------------------------------------------------------------------------------
// im2col: x (B,K,F,T) fp32 -> cols [B*n, P] act.  V = 4 elements per thread when pt % 4 == 0.
// Threads are ordered by OUTPUT element so stores are fully coalesced; each load is a whole
// 16-byte piece of a 32-byte sector run along T.
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void im2col_kernel(const float* __restrict__ x, T* __restrict__ cols, long long total_v, int K, int F,
                              int Tt, int pk, int pf, int pt, int Fp, int Tp, int n, int P) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total_v;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx * V;
    const long long r = e / P;
    const int c = (int)(e - r * P);
    const int dk = c / (pf * pt);
    const int df = (c / pt) % pf;
    const int dt = c % pt;
    const int b = (int)(r / n);
    const int i = (int)(r - (long long)b * n);
    const int kp = i / (Fp * Tp);
    const int fp = (i / Tp) % Fp;
    const int tp = i % Tp;
    const long long src = (((long long)b * K + kp * pk + dk) * F + fp * pf + df) * Tt + tp * pt + dt;
    if (V == 4) {
      st4(cols + e, ld4(x + src));
    } else {
      Act<T>::st(cols + e, x[src]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, VPL float4 vectors per lane held in registers (two-pass
// statistics from registers: mean, then centred variance -- matches ATen's numerics closely).
// ---------------------------------------------------------------------------------------------
template <typename T, int VPL>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long xrs,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      T* __restrict__ y, float* __restrict__ mean,
                                                      float* __restrict__ rstd, long long rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float invD = 1.0f / (float)D;
  for (long long r = warp0; r < rows; r += nwarps) {
    const float* xr = x + r * xrs;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      v[j] = (vi < nvec) ? ld4(xr + 4 * vi) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mu = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
        const float a0 = v[j].x - mu, a1 = v[j].y - mu, a2 = v[j].z - mu, a3 = v[j].w - mu;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invD + eps);
    if (lane == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = rs;
    }
    T* yr = y + r * (long long)D;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
        const float4 ww = ld4(w + 4 * vi), bb = ld4(b + 4 * vi);
        float4 o;
        o.x = (v[j].x - mu) * rs * ww.x + bb.x;
        o.y = (v[j].y - mu) * rs * ww.y + bb.y;
        o.z = (v[j].z - mu) * rs * ww.z + bb.z;
        o.w = (v[j].w - mu) * rs * ww.w + bb.w;
        st4(yr + 4 * vi, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (+ residual add, + branch-gradient emission, + dgamma/dbeta/colsum partials)
// ---------------------------------------------------------------------------------------------
// MINB = resident blocks per SM the register allocation is held to: the kernel is latency-bound (ncu: 123 registers ->
// 2 blocks = 16 warps per SM, 14 long-scoreboard stalls per issue, 4.7 TB/s), so more resident warps can pay for a few
// spilled accumulators: 4 blocks (64 registers, ~0.4 KB of L1-resident spills) run the plain variant at 5.6 TB/s.
// TVIT_LN_MINB=2|3 selects the other instantiations for A/B timing.
template <typename T, int VPL, int MINB>
__global__ void __launch_bounds__(256, MINB) ln_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ x, long long xrs,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      const float* __restrict__ w, const float* __restrict__ gres,
                                                      float* __restrict__ dx, long long dxrs, float* __restrict__ dw,
                                                      float* __restrict__ db, T* __restrict__ gp,
                                                      const float* __restrict__ row_scale, int rpg, DropCfg drop,
                                                      float* __restrict__ gpcs, long long rows, int D) {
  extern __shared__ float red[];  // [3][D]
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 2;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float invD = 1.0f / (float)D;
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();

  float4 aw[VPL], ab[VPL], ac[VPL];
  float4 wv[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    aw[j] = ab[j] = ac[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int vi = lane + 32 * j;
    wv[j] = (vi < nvec) ? ld4(w + 4 * vi) : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (long long r = warp0; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    const float* xr = x + r * xrs;
    const T* dyr = dy + r * (long long)D;
    float4 xh[VPL], g[VPL], gr[VPL];
    float s1 = 0.f, s2 = 0.f;
    // all three streams (x, dy, residual gradient) are requested up front: one row = one round trip to HBM
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      gr[j] = (gres && vi < nvec) ? ld4(gres + r * (long long)D + 4 * vi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
        const float4 xv = ld4(xr + 4 * vi);
        const float4 d = ld4(dyr + 4 * vi);
        xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        aw[j].x += d.x * xh[j].x; aw[j].y += d.y * xh[j].y; aw[j].z += d.z * xh[j].z; aw[j].w += d.w * xh[j].w;
        ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
        g[j] = make_float4(d.x * wv[j].x, d.y * wv[j].y, d.z * wv[j].z, d.w * wv[j].w);
        s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
        s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
      } else {
        xh[j] = g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float c1 = warp_sum(s1) * invD, c2 = warp_sum(s2) * invD;
    const float rsc = (gp && row_scale) ? row_scale[r / rpg] : 1.0f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
        float4 o;
        o.x = (g[j].x - c1 - xh[j].x * c2) * rs;
        o.y = (g[j].y - c1 - xh[j].y * c2) * rs;
        o.z = (g[j].z - c1 - xh[j].z * c2) * rs;
        o.w = (g[j].w - c1 - xh[j].w * c2) * rs;
        o.x += gr[j].x; o.y += gr[j].y; o.z += gr[j].z; o.w += gr[j].w;
        st4(dx + r * dxrs + 4 * vi, o);
        if (gp) {
          float m[4];
          drop_mult4(drop, (unsigned long long)r * D + 4 * vi, m);
          float4 p = make_float4(o.x * rsc * m[0], o.y * rsc * m[1], o.z * rsc * m[2], o.w * rsc * m[3]);
          st4(gp + r * (long long)D + 4 * vi, p);
          ac[j].x += p.x; ac[j].y += p.y; ac[j].z += p.z; ac[j].w += p.w;
        }
      }
    }
  }
  // block reduction through shared memory, then one atomic per column per block
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int vi = lane + 32 * j;
    if (vi < nvec) {
      const int c = 4 * vi;
      atomicAdd(&red[c + 0], aw[j].x); atomicAdd(&red[c + 1], aw[j].y);
      atomicAdd(&red[c + 2], aw[j].z); atomicAdd(&red[c + 3], aw[j].w);
      atomicAdd(&red[D + c + 0], ab[j].x); atomicAdd(&red[D + c + 1], ab[j].y);
      atomicAdd(&red[D + c + 2], ab[j].z); atomicAdd(&red[D + c + 3], ab[j].w);
      if (gp) {
        atomicAdd(&red[2 * D + c + 0], ac[j].x); atomicAdd(&red[2 * D + c + 1], ac[j].y);
        atomicAdd(&red[2 * D + c + 2], ac[j].z); atomicAdd(&red[2 * D + c + 3], ac[j].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    if (dw) atomicAdd(dw + i, red[i]);
    if (db) atomicAdd(db + i, red[D + i]);
    if (gp && gpcs) atomicAdd(gpcs + i, red[2 * D + i]);
  }
}

// ---------------------------------------------------------------------------------------------
// gp = g * row_scale * dropout_mult, colsum += sum_r gp
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) branch_grad_prep_kernel(const float* __restrict__ g, long long rows, int D,
                                                                const float* __restrict__ row_scale, int rpg,
                                                                DropCfg drop, T* __restrict__ gp,
                                                                float* __restrict__ colsum) {
  // thread owns vector-column vc = blockIdx.x*blockDim.x + threadIdx.x; blockIdx.y strides rows
  const int nvec = D >> 2;
  const int vc = blockIdx.x * blockDim.x + threadIdx.x;
  if (vc >= nvec) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // four rows per iteration with all loads issued first: one 16-byte load in flight per thread left the kernel
  // latency-bound at ~55 % of the HBM rate
  constexpr int U = 4;
  for (long long r0 = blockIdx.y; r0 < rows; r0 += (long long)U * gridDim.y) {
    float4 v[U];
    float rsc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * gridDim.y;
      if (r < rows) {
        v[u] = ld4(g + r * (long long)D + 4 * vc);
        rsc[u] = row_scale ? row_scale[r / rpg] : 1.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + (long long)u * gridDim.y;
      if (r < rows) {
        float m[4];
        drop_mult4(drop, (unsigned long long)r * D + 4 * vc, m);
        float4 p = make_float4(v[u].x * rsc[u] * m[0], v[u].y * rsc[u] * m[1], v[u].z * rsc[u] * m[2],
                               v[u].w * rsc[u] * m[3]);
        st4(gp + r * (long long)D + 4 * vc, p);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
    }
  }
  if (colsum) {
    atomicAdd(colsum + 4 * vc + 0, acc.x); atomicAdd(colsum + 4 * vc + 1, acc.y);
    atomicAdd(colsum + 4 * vc + 2, acc.z); atomicAdd(colsum + 4 * vc + 3, acc.w);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long rows, int C, long long ld,
                                                      float* __restrict__ out) {
  const int nvec = C >> 2;
  const int vc = blockIdx.x * blockDim.x + threadIdx.x;
  if (vc >= nvec) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    const float4 v = ld4(x + r * ld + 4 * vc);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  atomicAdd(out + 4 * vc + 0, acc.x); atomicAdd(out + 4 * vc + 1, acc.y);
  atomicAdd(out + 4 * vc + 2, acc.z); atomicAdd(out + 4 * vc + 3, acc.w);
}

// generic (unaligned / tiny C) column sum: one thread per column, blockIdx.y strides rows
template <typename T>
__global__ void colsum_scalar_kernel(const T* __restrict__ x, long long rows, int C, long long ld,
                                     float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) acc += Act<T>::ld(x + r * ld + c);
  atomicAdd(out + c, acc);
}

// ---------------------------------------------------------------------------------------------
// weight shadows: out = w (cast), out_t[c,r] = row_scale[r] * w[r,c]   (32x32 smem tile transpose)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cast_weight_kernel(const float* __restrict__ w, int R, int C, const float* __restrict__ row_scale,
                                   T* __restrict__ out, T* __restrict__ out_t) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = w[(long long)r * C + c];
      if (out) Act<T>::st(out + (long long)r * C + c, v);
      if (row_scale) v *= row_scale[r];
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  if (out_t) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int c = c0 + j, r = r0 + threadIdx.x;
      if (r < R && c < C) Act<T>::st(out_t + (long long)c * R + r, tile[threadIdx.x][j]);
    }
  }
}

__global__ void ls_finalize_kernel(const float* __restrict__ G, const float* __restrict__ W,
                                   const float* __restrict__ gamma, const float* __restrict__ bias,
                                   const float* __restrict__ cs, float* __restrict__ dW, float* __restrict__ dgamma,
                                   float* __restrict__ dbias, int R, int C, int accumulate) {
  __shared__ float part[32];
  const int r = blockIdx.x;
  const float gm = gamma ? gamma[r] : 1.0f;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float gv = G[(long long)r * C + c];
    float* d = dW + (long long)r * C + c;
    *d = accumulate ? *d + gm * gv : gm * gv;
    if (gamma) s += W[(long long)r * C + c] * gv;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    const float c = cs[r];
    if (gamma && dgamma) dgamma[r] = (accumulate ? dgamma[r] : 0.f) + t + (bias ? bias[r] : 0.f) * c;
    if (dbias) dbias[r] = (accumulate ? dbias[r] : 0.f) + gm * c;
  }
}

__global__ void cls_rows_kernel(const float* __restrict__ cls, float* __restrict__ h, int B, int N, int D, DropCfg drop) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, c = idx - b * D;
  const unsigned long long e = (unsigned long long)b * N * D + c;
  h[e] = cls[c] * drop_mult(drop, e);
}

// block (i, vector-column chunk): i in [0, n] ; i == n handles the CLS row
template <typename T>
__global__ void embed_bwd_prep_kernel(const float* __restrict__ g0, int B, int n, int D, DropCfg drop,
                                      T* __restrict__ gtok, float* __restrict__ R, float* __restrict__ dcls,
                                      int accumulate) {
  const int nvec = D >> 2;
  const int vc = blockIdx.y * blockDim.x + threadIdx.x;
  if (vc >= nvec) return;
  const int i = blockIdx.x;
  const int N = n + 1;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int tok = (i == n) ? 0 : 1 + i;
  for (int b = 0; b < B; ++b) {
    const unsigned long long e = ((unsigned long long)b * N + tok) * D + 4 * vc;
    const float4 v = ld4(g0 + e);
    float m[4];
    drop_mult4(drop, e, m);
    float4 p = make_float4(v.x * m[0], v.y * m[1], v.z * m[2], v.w * m[3]);
    if (i < n) st4(gtok + ((long long)b * n + i) * D + 4 * vc, p);
    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
  }
  float* dst = (i == n) ? dcls : (R + (long long)i * D);
  if (i == n && accumulate) {
    const float4 o = ld4(dst + 4 * vc);
    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
  }
  st4(dst + 4 * vc, acc);
}

// block r in [0, Kp+Fp+Tp]: one table row (or the bias row when r == Kp+Fp+Tp)
__global__ void pos_grad_reduce_kernel(const float* __restrict__ R, int Kp, int Fp, int Tp, int D,
                                       float* __restrict__ dpk, float* __restrict__ dpf, float* __restrict__ dpt,
                                       float* __restrict__ dbias, int accumulate) {
  const int r = blockIdx.x;
  const int n = Kp * Fp * Tp;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
    if (r < Kp) {
      for (int j = 0; j < Fp * Tp; ++j) s += R[((long long)r * Fp * Tp + j) * D + c];
      dpk[(long long)r * D + c] = accumulate ? dpk[(long long)r * D + c] + s : s;
    } else if (r < Kp + Fp) {
      const int f = r - Kp;
      for (int k = 0; k < Kp; ++k)
        for (int t = 0; t < Tp; ++t) s += R[(((long long)k * Fp + f) * Tp + t) * D + c];
      dpf[(long long)f * D + c] = accumulate ? dpf[(long long)f * D + c] + s : s;
    } else if (r < Kp + Fp + Tp) {
      const int t = r - Kp - Fp;
      for (int j = 0; j < Kp * Fp; ++j) s += R[((long long)j * Tp + t) * D + c];
      dpt[(long long)t * D + c] = accumulate ? dpt[(long long)t * D + c] + s : s;
    } else {
      for (int j = 0; j < n; ++j) s += R[(long long)j * D + c];
      dbias[c] = accumulate ? dbias[c] + s : s;
    }
  }
}

}  // namespace tvit

using namespace tvit;

#define DISPATCH_T(dtype, ...)                      \
  if ((dtype) == TVIT_F32) {                        \
    using T = float;                                \
    __VA_ARGS__                                     \
  } else if ((dtype) == TVIT_BF16) {                \
    using T = __nv_bfloat16;                        \
    __VA_ARGS__                                     \
  } else {                                          \
    return fail(TVIT_ERR_BAD_ARG, "bad dtype %d", (int)(dtype)); \
  }

#define DISPATCH_VPL(D, ...)                                          \
  {                                                                   \
    const int _nv = ((D) / 4 + 31) / 32;                              \
    if (_nv <= 1) { constexpr int VPL = 1; __VA_ARGS__ }              \
    else if (_nv <= 2) { constexpr int VPL = 2; __VA_ARGS__ }         \
    else if (_nv <= 3) { constexpr int VPL = 3; __VA_ARGS__ }         \
    else if (_nv <= 4) { constexpr int VPL = 4; __VA_ARGS__ }         \
    else if (_nv <= 6) { constexpr int VPL = 6; __VA_ARGS__ }         \
    else if (_nv <= 8) { constexpr int VPL = 8; __VA_ARGS__ }         \
    else return fail(TVIT_ERR_UNSUPPORTED, "embed_dim %d > 1024 not supported", (int)(D)); \
  }

extern "C" int tvit_im2col(const float* x, void* cols, int dtype, int B, int K, int F, int T_, int pk, int pf, int pt,
                           tvit_stream_t stream) {
  TVIT_CHECK_ARG(x && cols, "im2col: null pointer");
  TVIT_CHECK_ARG(K % pk == 0 && F % pf == 0 && T_ % pt == 0, "im2col: dims not divisible by patch");
  cudaStream_t s = (cudaStream_t)stream;
  const int Kp = K / pk, Fp = F / pf, Tp = T_ / pt;
  const int n = Kp * Fp * Tp, P = pk * pf * pt;
  const long long total = (long long)B * n * P;
  const bool vec = (pt % 4 == 0);
  const long long total_v = vec ? total / 4 : total;
  const int grid = blocks_for(total_v, 256, num_sms() * 16);
  DISPATCH_T(dtype, {
    if (vec)
      im2col_kernel<T, 4><<<grid, 256, 0, s>>>(x, (T*)cols, total_v, K, F, T_, pk, pf, pt, Fp, Tp, n, P);
    else
      im2col_kernel<T, 1><<<grid, 256, 0, s>>>(x, (T*)cols, total_v, K, F, T_, pk, pf, pt, Fp, Tp, n, P);
  })
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_ln_fwd(const float* x, long long xrs, const float* weight, const float* bias, void* y, int dtype,
                           float* mean, float* rstd, long long rows, int D, float eps, tvit_stream_t stream) {
  TVIT_CHECK_ARG(x && weight && bias && y, "ln_fwd: null pointer");
  TVIT_CHECK_ARG(D % 4 == 0 && xrs % 4 == 0, "ln_fwd: D and row stride must be multiples of 4 (D=%d)", D);
  if (rows == 0) return TVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = blocks_for(rows, 8, num_sms() * 8);
  DISPATCH_T(dtype, DISPATCH_VPL(D, {
    ln_fwd_kernel<T, VPL><<<grid, 256, 0, s>>>(x, xrs, weight, bias, (T*)y, mean, rstd, rows, D, eps);
  }))
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_ln_bwd(const void* dy, int dtype, const float* x, long long xrs, const float* mean,
                           const float* rstd, const float* weight, const float* g_res, float* dx, long long dxrs,
                           float* dweight, float* dbias, void* gp, const float* row_scale, int rows_per_group,
                           const tvit_dropout* drop, float* gp_colsum, long long rows, int D, tvit_stream_t stream) {
  TVIT_CHECK_ARG(dy && x && mean && rstd && weight && dx, "ln_bwd: null pointer");
  TVIT_CHECK_ARG(D % 4 == 0 && xrs % 4 == 0 && dxrs % 4 == 0, "ln_bwd: D/strides must be multiples of 4");
  TVIT_CHECK_ARG(!row_scale || rows_per_group > 0, "ln_bwd: rows_per_group must be > 0");
  if (rows == 0) return TVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  static const int minb = [] {
    const char* e = getenv("TVIT_LN_MINB");
    const int v = e ? atoi(e) : 4;  // measured (B200, M = 524544, D = 384): 0.607 / 0.552 / 0.507 ms for 2 / 3 / 4
    return (v == 2 || v == 3) ? v : 4;
  }();
  const int grid = blocks_for(rows, 8 * 4, num_sms() * (minb > 2 ? 2 * minb : 4));
  const DropCfg dc = make_drop(drop);
  const size_t smem = 3 * (size_t)D * sizeof(float);
#define LN_BWD_LAUNCH(MB)                                                                                            \
  ln_bwd_kernel<T, VPL, MB><<<grid, 256, smem, s>>>((const T*)dy, x, xrs, mean, rstd, weight, g_res, dx, dxrs, dweight, \
                                                    dbias, (T*)gp, row_scale, rows_per_group > 0 ? rows_per_group : 1, \
                                                    dc, gp_colsum, rows, D)
  DISPATCH_T(dtype, DISPATCH_VPL(D, {
    if (minb == 3) LN_BWD_LAUNCH(3);
    else if (minb == 4) LN_BWD_LAUNCH(4);
    else LN_BWD_LAUNCH(2);
  }))
#undef LN_BWD_LAUNCH
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

// Variant for D % 16 == 0 with dropout: a thread owns one 16-element dropout group of a row (four consecutive
// float4 loads, ONE Philox call, one 32-byte bf16 / 64-byte fp32 store) instead of one 4-element vector (one
// Philox call per vector, which left the kernel ALU-bound at 62 % of the HBM rate).  blockDim = (32, 8): x indexes
// the group within the row, y the row; column sums are reduced over y in shared memory first.
template <typename T>
__global__ void __launch_bounds__(256) branch_grad_prep16_kernel(const float* __restrict__ g, long long rows, int D,
                                                                  const float* __restrict__ row_scale, int rpg,
                                                                  DropCfg drop, T* __restrict__ gp,
                                                                  float* __restrict__ colsum) {
  __shared__ float red[8][32][17];
  const int ngrp = D >> 4;
  const int gc = blockIdx.x * 32 + threadIdx.x;  // 16-element group within the row
  const bool active = gc < ngrp;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const long long rstride = (long long)gridDim.y * 8;
  if (active) {
    for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += rstride) {
      const float* src = g + r * (long long)D + 16 * gc;
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ld4(src + 4 * k);
      const float rsc = row_scale ? row_scale[r / rpg] : 1.0f;
      float m[16];
      drop_mult16(drop, (unsigned long long)r * D + 16 * gc, m);
      float p[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        p[4 * k] = v[k].x * rsc * m[4 * k];
        p[4 * k + 1] = v[k].y * rsc * m[4 * k + 1];
        p[4 * k + 2] = v[k].z * rsc * m[4 * k + 2];
        p[4 * k + 3] = v[k].w * rsc * m[4 * k + 3];
      }
      T* dst = gp + r * (long long)D + 16 * gc;
#pragma unroll
      for (int k = 0; k < 4; ++k) st4(dst + 4 * k, make_float4(p[4 * k], p[4 * k + 1], p[4 * k + 2], p[4 * k + 3]));
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += p[j];
    }
  }
  if (colsum) {
#pragma unroll
    for (int j = 0; j < 16; ++j) red[threadIdx.y][threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.y == 0 && active) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x][j];
        atomicAdd(colsum + 16 * gc + j, t);
      }
    }
  }
}

extern "C" int tvit_branch_grad_prep(const float* g, long long rows, int D, const float* row_scale, int rows_per_group,
                                     const tvit_dropout* drop, void* gp, int dtype, float* colsum,
                                     tvit_stream_t stream) {
  TVIT_CHECK_ARG(g && gp, "branch_grad_prep: null pointer");
  TVIT_CHECK_ARG(D % 4 == 0, "branch_grad_prep: D must be a multiple of 4");
  TVIT_CHECK_ARG(!row_scale || rows_per_group > 0, "branch_grad_prep: rows_per_group must be > 0");
  if (rows == 0) return TVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int nvec = D / 4;
  const int bx = nvec < 256 ? ((nvec + 31) / 32) * 32 : 256;
  dim3 grid((nvec + bx - 1) / bx, 1);
  long long gy = (long long)num_sms() * 8 / grid.x;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  grid.y = (unsigned)gy;
  const DropCfg dc = make_drop(drop);
  if (dc.thr16 != 0 && D % 16 == 0) {
    const int ngrp = D / 16;
    dim3 grid16((ngrp + 31) / 32, 1), block16(32, 8);
    long long gy16 = (long long)num_sms() * 8 / grid16.x;
    if (gy16 > (rows + 7) / 8) gy16 = (rows + 7) / 8;
    if (gy16 < 1) gy16 = 1;
    grid16.y = (unsigned)gy16;
    DISPATCH_T(dtype, {
      branch_grad_prep16_kernel<T><<<grid16, block16, 0, s>>>(g, rows, D, row_scale, rows_per_group > 0 ? rows_per_group : 1,
                                                               dc, (T*)gp, colsum);
    })
    TVIT_LAUNCH_OK();
    return TVIT_OK;
  }
  DISPATCH_T(dtype, {
    branch_grad_prep_kernel<T><<<grid, bx, 0, s>>>(g, rows, D, row_scale, rows_per_group > 0 ? rows_per_group : 1, dc,
                                                   (T*)gp, colsum);
  })
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_colsum(const void* x, int dtype, long long rows, int C, long long ld, float* out,
                           tvit_stream_t stream) {
  TVIT_CHECK_ARG(x && out, "colsum: null pointer");
  if (rows == 0) return TVIT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (C % 4 != 0 || ld % 4 != 0) {
    dim3 g((C + 63) / 64, (unsigned)(rows < 64 ? rows : 64));
    DISPATCH_T(dtype, { colsum_scalar_kernel<T><<<g, 64, 0, s>>>((const T*)x, rows, C, ld, out); })
    TVIT_LAUNCH_OK();
    return TVIT_OK;
  }
  const int nvec = C / 4;
  const int bx = nvec < 256 ? ((nvec + 31) / 32) * 32 : 256;
  dim3 grid((nvec + bx - 1) / bx, 1);
  long long gy = (long long)num_sms() * 8 / grid.x;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  grid.y = (unsigned)gy;
  DISPATCH_T(dtype, { colsum_kernel<T><<<grid, bx, 0, s>>>((const T*)x, rows, C, ld, out); })
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_cast_weight(const float* w, int R, int C, const float* row_scale, void* out, void* out_t, int dtype,
                                tvit_stream_t stream) {
  TVIT_CHECK_ARG(w && (out || out_t), "cast_weight: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  DISPATCH_T(dtype, { cast_weight_kernel<T><<<grid, block, 0, s>>>(w, R, C, row_scale, (T*)out, (T*)out_t); })
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_ls_finalize(const float* G, const float* W, const float* gamma, const float* bias, const float* cs,
                                float* dW, float* dgamma, float* dbias, int R, int C, int accumulate,
                                tvit_stream_t stream) {
  TVIT_CHECK_ARG(G && cs && dW, "ls_finalize: null pointer");
  TVIT_CHECK_ARG(!gamma || W, "ls_finalize: W required with gamma");
  ls_finalize_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(G, W, gamma, bias, cs, dW, dgamma, dbias, R, C, accumulate);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_cls_rows(const float* cls, float* h, int B, int N, int D, const tvit_dropout* drop,
                             tvit_stream_t stream) {
  TVIT_CHECK_ARG(cls && h, "cls_rows: null pointer");
  const int total = B * D;
  cls_rows_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cls, h, B, N, D, make_drop(drop));
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_embed_bwd_prep(const float* g0, int B, int n, int D, const tvit_dropout* drop, void* gtok,
                                   int dtype, float* R, float* dcls, int accumulate, tvit_stream_t stream) {
  TVIT_CHECK_ARG(g0 && gtok && R && dcls, "embed_bwd_prep: null pointer");
  TVIT_CHECK_ARG(D % 4 == 0, "embed_bwd_prep: D must be a multiple of 4");
  const int nvec = D / 4;
  const int bx = nvec < 128 ? ((nvec + 31) / 32) * 32 : 128;
  dim3 grid(n + 1, (nvec + bx - 1) / bx);
  const DropCfg dc = make_drop(drop);
  DISPATCH_T(dtype, {
    embed_bwd_prep_kernel<T><<<grid, bx, 0, (cudaStream_t)stream>>>(g0, B, n, D, dc, (T*)gtok, R, dcls, accumulate);
  })
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

extern "C" int tvit_pos_grad_reduce(const float* R, int Kp, int Fp, int Tp, int D, float* dpos_k, float* dpos_f,
                                    float* dpos_t, float* dbias, int accumulate, tvit_stream_t stream) {
  TVIT_CHECK_ARG(R && dpos_k && dpos_f && dpos_t && dbias, "pos_grad_reduce: null pointer");
  pos_grad_reduce_kernel<<<Kp + Fp + Tp + 1, 128, 0, (cudaStream_t)stream>>>(R, Kp, Fp, Tp, D, dpos_k, dpos_f, dpos_t,
                                                                           dbias, accumulate);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}
