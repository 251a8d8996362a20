// Flash-attention backward for sm_100a (head_dim 64, bf16 operands, fp32 accumulate in TMEM).
//
// One CTA per (128-key tile j, head, sample) loops over the 128-query tiles i.  Per (i, j):
//   S  = Q_i K_j^T          tcgen05.mma SS  (A = Q_i  K-major,  B = K_j  K-major)   -> TMEM tS  [128 x 128]
//   dP = dO_i V_j^T         tcgen05.mma SS  (A = dO_i K-major,  B = V_j  K-major)   -> TMEM tDP [128 x 128]
//   softmax warps (thread == query row): P = exp2(S*c - lse2), dS = P o (dP o mask - D) * scale,
//       written as bf16 into two 128B-swizzled smem tiles sP / sDS laid out [q rows][kv contiguous]
//   dV_j += P^T  dO_i       A = sP  read MN-major (M = kv),  B = dO_i read MN-major (N = hd)  -> TMEM tDV
//   dK_j += dS^T Q_i        A = sDS read MN-major,           B = Q_i  read MN-major           -> TMEM tDK
//   dQ_i  = dS   K_j        A = sDS read K-major (M = q),    B = K_j  read MN-major           -> TMEM tDQ
// The same smem tile serves as a K-major and as an MN-major UMMA operand (the 128B swizzle is a pure
// address function), so neither P nor dS is ever transposed.  dQ_i partial tiles are drained by four
// dedicated warps with coalesced 16-byte fp32 reductions into a tile-chunked accumulator and converted
// to bf16 by a finishing kernel; dK_j / dV_j stay in TMEM across the whole query loop.
// D = rowsum(dO o O) and lse reach the softmax threads as ready-made per-query records [-lse2 | -D'] written by the
// prep kernel (one pair of scalars per thread and tile, loaded one tile ahead).  With dropout the keep flags are either
// regenerated (Philox, one call per 16 keys) or -- kBits -- read from the cache the forward kernel wrote.
//
// Pipelining.  TMEM (512 columns: S 128, dP 128, dV 64, dK 64, dQ 64) has no room for a second S / dP tile, so
// the tile is pipelined by key halves instead: S and dP are issued as two N = 64 MMA groups (keys [0,64) and
// [64,128)), every softmax warp processes half a then half b and releases each half's TMEM columns right after
// its tcgen05.ld, and the MMA warp refills half a of tile i+1 while the softmax warps are still working on half
// b of tile i.  Q/dO tiles are triple-buffered, dS double-buffered, P single-buffered (dV_i is issued first and
// frees it long before the first P store of tile i+1).  Each CTA starts its query loop at its own key-tile index
// (i -> (i + j) mod nq) so that the CTAs of one (sample, head) never reduce into the same dQ tile at once.
//
// Ragged tail: when N = 128 m + t with a small t (N = 2049 = 16 * 128 + 1 at the bench shape), the last t QUERIES are
// not given a query tile of their own (one more trip through the five-MMA loop for one valid row: 1/17 of the kernel).
// Every key-tile CTA handles the tail query on the CUDA cores, thread == key row with K_j / V_j in shared memory:
// s = q_t . k, dP = dO_t . v, dS as in the loop.  The two otherwise idle warps 22-23 do this concurrently with the tile
// loop: they leave dS / (P mask) per key row in shared memory and add the warp-reduced partial dQ_t = sum_k dS k to the
// same fp32 accumulator the drained dQ tiles go to; the softmax warps only apply the rank-1 updates dK_j += dS q_t,
// dV_j += (P mask) dO_t right before they store dK_j / dV_j.  (Doing it in the epilogue cost as much as the tile it
// saved; doing it on the dQ-drain warps delayed dq_free and with it the MMA pipeline: +7 %.)
#include <cstdlib>

#include "tc_common.cuh"

namespace tvit {

#ifndef TVIT_ATTN_BWD_TSDQ
#define TVIT_ATTN_BWD_TSDQ 1  // 0: dQ reads dS from shared memory in the dropout instantiation too (A-B builds)
#endif
#ifndef TVIT_ATTN_BWD_TSDQ_BITS
#define TVIT_ATTN_BWD_TSDQ_BITS 0  // 1: the keep-flag-cache instantiation takes dS for dQ from TMEM as well (A-B builds)
#endif
#ifndef TVIT_ATTN_BWD_T_DEFAULT
#define TVIT_ATTN_BWD_T_DEFAULT 0
#endif
constexpr int kHdB = 64;
constexpr int kTileB = 128;
constexpr int kTileBytesB = kTileB * kHdB * 2;  // 16384: one [128 x 64] bf16 operand tile
constexpr int kPBytes = kTileB * kTileB * 2;    // 32768: one [128 x 128] bf16 P / dS tile (two 64-wide blocks)
constexpr int kAttnBwdThreads = 768;  // warps 0-15 softmax, 16-19 dQ drain, 20 TMA, 21 MMA (+TMEM alloc), 22-23 tail query
constexpr int kSoftmaxWarps = 16;

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b, float acc) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[i]));
    acc = fmaf(fa.x, fb.x, acc);
    acc = fmaf(fa.y, fb.y, acc);
  }
  return acc;
}
__device__ __forceinline__ uint4 ld_shared_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// Column sums over the 32 lanes of a warp with halving butterflies: lane L ends up with the total of v[L] (32 values)
// resp. the even lane L with the total of v[L >> 1] (16 values).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool hi = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = hi ? v[i] : v[i + w], keep = hi ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return v[0];
}
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
    const bool hi = (lane & (2 * w)) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = hi ? v[i] : v[i + w], keep = hi ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * w);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// 1-D bulk copy global -> shared (16-byte aligned, size % 16 == 0), completion counted on an mbarrier like a TMA load
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t prmt_r(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// Dropout keep flags of one Philox group as sign bits: bit 7 of byte j of w[] ends up set iff byte j >= thr8
// (thr8 in [0, 256]); the other bits are don't-care.  Four bytes per SWAR step: r = (w | 0x80..) - low7(T) has bit 7 of a
// byte set iff low7(byte) >= low7(T) (no borrow can cross bytes), and byte >= T <=> T < 128 ? (b7 | r7) : (b7 & r7);
// T = 256 (keep nothing) is mapped to low7 = 0x80, which clears r7.
__device__ __forceinline__ void drop_keep_sign16(uint32_t (&w)[4], uint32_t thr8) {
  const uint32_t tl = ((thr8 & 0x7fu) | ((thr8 >> 1) & 0x80u)) * 0x01010101u;
  const uint32_t tm = thr8 >= 128u ? 0xffffffffu : 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t r = (w[k] | 0x80808080u) - tl;
    w[k] = (tm & w[k] & r) | (~tm & (w[k] | r));
  }
}
// In-register transpose of a 16 x 16 byte matrix held by the 16 lanes of each half-warp (lane i: row i, byte j of
// w[j >> 2] = element (i, j)): two word-granular butterflies (lane bits 3, 2 <-> word index bits) and two byte-granular
// ones (lane bits 1, 0 <-> byte index bits, prmt picks own / received bytes).  12 shuffles, 20 ALU instructions.
__device__ __forceinline__ void transpose16x16_bytes(uint32_t (&w)[4], int lane) {
  {
    const bool hi = (lane & 8) != 0;
    const uint32_t s0 = __shfl_xor_sync(0xffffffffu, hi ? w[0] : w[2], 8);
    const uint32_t s1 = __shfl_xor_sync(0xffffffffu, hi ? w[1] : w[3], 8);
    w[0] = hi ? s0 : w[0]; w[1] = hi ? s1 : w[1]; w[2] = hi ? w[2] : s0; w[3] = hi ? w[3] : s1;
  }
  {
    const bool hi = (lane & 4) != 0;
    const uint32_t s0 = __shfl_xor_sync(0xffffffffu, hi ? w[0] : w[1], 4);
    const uint32_t s1 = __shfl_xor_sync(0xffffffffu, hi ? w[2] : w[3], 4);
    w[0] = hi ? s0 : w[0]; w[2] = hi ? s1 : w[2]; w[1] = hi ? w[1] : s0; w[3] = hi ? w[3] : s1;
  }
  {
    const uint32_t sel = (lane & 2) ? 0x3276u : 0x5410u;
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = prmt_r(w[k], __shfl_xor_sync(0xffffffffu, w[k], 2), sel);
  }
  {
    const uint32_t sel = (lane & 1) ? 0x3715u : 0x6240u;
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = prmt_r(w[k], __shfl_xor_sync(0xffffffffu, w[k], 1), sel);
  }
}

// Optional event trace (-DTVIT_ATTN_TRACE builds only; development aid): CTA (0,0,0) records clock64 at pipeline events,
// g_attn_trace[(slot * 32 + tile) * 8 + event]; read back with tvit_attn_bwd_trace().
#ifdef TVIT_ATTN_TRACE
__device__ long long g_attn_trace[8 * 32 * 8];
#define TVIT_TRACE(slot, tile, ev)                                                                   \
  do {                                                                                               \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && (tile) < 32) \
      g_attn_trace[((slot) * 32 + (tile)) * 8 + (ev)] = clock64();                                   \
  } while (0)
#else
#define TVIT_TRACE(slot, tile, ev) do { } while (0)
#endif

constexpr int kQStages = 3;  // Q / dO tiles in flight
struct AttnBwdSmem {
  uint64_t kv_full, qdo_full[kQStages], qdo_empty[kQStages], s_full[2], s_free[2], p_full, p_free, ds_free[2], dq_full,
      dq_free, tds_free, tail_ready, pt_full[2];
  uint32_t tmem_base;
  uint32_t pad_;
  float tail_ds[kTileB], tail_pm[kTileB];  // tail query (see header): dS and masked P of this CTA's 128 key rows
};
static_assert(sizeof(AttnBwdSmem) <= 1280, "AttnBwdSmem must fit the 1.25 KB tail of the dynamic smem block");

// Dvec[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d], and the per-query statistics in the form the main kernel consumes: one
// 1 KB record per (b, h, query tile), [-lse2 x 128 | -D' x 128] with lse2 = lse log2(e) - log2(1/(1-p)) and
// D' = D scale (1-p); rows past N get -inf / 0 so that their P is exactly 0.  One thread per (b, padded q, h).
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                     const float* __restrict__ lse, float* __restrict__ dvec, float* __restrict__ stat,
                                     int B, int N, int H, int nqt, float log2_inv_keep, float d_scale) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int npadq = nqt * kTileB;
  const long long total = (long long)B * npadq * H;
  if (idx >= total) return;
  const int h = (int)(idx % H);
  const long long prow = idx / H;  // b * npadq + q
  const int b = (int)(prow / npadq), q = (int)(prow % npadq);
  float nl = -INFINITY, nd = 0.f;
  if (q < N) {
    const long long row = (long long)b * N + q;
    const __nv_bfloat16* po = o + row * (long long)(H * kHdB) + h * kHdB;
    const __nv_bfloat16* pd = dout + row * (long long)(H * kHdB) + h * kHdB;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kHdB / 8; ++c) {
      const uint4 a = *reinterpret_cast<const uint4*>(po + 8 * c);
      const uint4 d = *reinterpret_cast<const uint4*>(pd + 8 * c);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[t]));
        const float2 fd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&dw[t]));
        s += fa.x * fd.x + fa.y * fd.y;
      }
    }
    const long long bhq = ((long long)b * H + h) * N + q;
    dvec[bhq] = s;
    nl = log2_inv_keep - lse[bhq] * 1.4426950408889634f;
    nd = -s * d_scale;
  }
  float* rec = stat + (((long long)b * H + h) * nqt + (q >> 7)) * 256 + (q & (kTileB - 1));
  rec[0] = nl;
  rec[128] = nd;
}

// dq accumulator layout: [(b*H+h)][q tile][16 chunks][128 rows][4 fp32]
// thread = (row r, group of 4 chunks): four coalesced float4 reads, one full-sector 32-byte bf16 store.  A 512-thread
// block walks the query tiles i = blockIdx.y, blockIdx.y + gridDim.y, ... of one (b, h), so that the dq part of the
// qkv-bias gradient (column sums) is accumulated in registers over the tiles and leaves the block as 64 atomics
// (one warp-level butterfly and one atomic per 32 rows and tile put 6.7 M same-address atomics on 384 words at C2).
__global__ void __launch_bounds__(512)
attn_bwd_dq_finish_kernel(const float* __restrict__ dqacc, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ colsum,
                          int N, int H, int nq) {
  __shared__ float red[4][64];
  const int r = threadIdx.x & 127, cg = threadIdx.x >> 7;
  const long long bh = blockIdx.x;
  const int h = (int)(bh % H), b = (int)(bh / H);
  float cs[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) cs[k] = 0.f;
  for (int i = blockIdx.y; i < nq; i += gridDim.y) {
    const int q = i * kTileB + r;
    if (q >= N) continue;
    const float* src = dqacc + (((bh * nq + i) * 16 + 4 * cg) * 128 + r) * 4;
    float f[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 x = *reinterpret_cast<const float4*>(src + (long long)k * 512);
      f[4 * k] = x.x; f[4 * k + 1] = x.y; f[4 * k + 2] = x.z; f[4 * k + 3] = x.w;
    }
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = pack_bf16(f[2 * k], f[2 * k + 1]);
    st_global_b32x8(dqkv + ((long long)b * N + q) * (3LL * H * kHdB) + h * kHdB + 16 * cg, v[0], v[1], v[2], v[3], v[4],
                    v[5], v[6], v[7]);
#pragma unroll
    for (int k = 0; k < 16; ++k) cs[k] += f[k];
  }
  if (colsum) {  // a warp holds 32 rows of the same 16 columns; the four warps of a column group meet in smem
    const int lane = threadIdx.x & 31, wq = (threadIdx.x >> 5) & 3;
    const float tot = warp_colsum16(cs, lane);
    if ((lane & 1) == 0) red[wq][16 * cg + (lane >> 1)] = tot;
    __syncthreads();
    if (threadIdx.x < 64)
      atomicAdd(colsum + h * kHdB + threadIdx.x,
                (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]));
  }
}

// kT selects the transposed formulation (see the header): S^T = K_j Q_i^T and dP^T = V_j dO_i^T land in TMEM with
// thread == key row, so that P^T and dS^T -- written back in place as bf16 -- are the TMEM A operands of the dV / dK
// MMAs, and only dS^T goes through shared memory (for dQ).  `stat` holds the per-query column statistics
// [(b,h)][q tile][-lse2 x 128 | -D' x 128] produced by attn_bwd_prep_kernel.
// kFull (non-transposed only): S and dP are issued as whole N = 128 MMAs (4 x 107 clk instead of 8 x 75 clk,
// profiles/r2_mma_rate.txt) into the single-buffered S / dP columns as soon as every softmax warp has loaded tile i's
// scores -- i.e. half-way through the softmax of tile i -- instead of half by half, and dQ always takes dS from TMEM.
// kBits (default formulation with dropout only): the keep flags come from the cache the forward kernel wrote
// (include/tvit.h: tvit_attn_keepbits_bytes) instead of from the generator -- two 32-bit loads and a shift + prmt per
// element pair replace two Philox calls and the byte compares per tile and thread.
template <bool kDrop, bool kT, bool kFull, bool kBits = false>
__global__ void __launch_bounds__(kAttnBwdThreads, 1)
tc_attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                   const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                   const float* __restrict__ lse, const float* __restrict__ dvec, const float* __restrict__ stat,
                   float* __restrict__ dqacc, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ colsum, int N,
                   int tail_arg, int H, float scale, DropCfg drop, const uint32_t* __restrict__ keepbits) {
  static_assert(!kBits || (kDrop && !kT && !kFull), "the keep-bit cache feeds the default dropout instantiation only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTileBytesB;
  uint8_t* sQdO = smem + 2 * kTileBytesB;  // stage s: Q at +s*32K, dO at +s*32K+16K
  uint8_t* sP = sQdO + kQStages * 2 * kTileBytesB;   // P  [2 key blocks][128 q rows][64 keys], single buffer
  uint8_t* sDS = sP + kPBytes;                       // dS, same layout, buffer u (= tile parity) at +u*32K
  AttnBwdSmem* sm = reinterpret_cast<AttnBwdSmem*>(sDS + 2 * kPBytes);
  uint8_t* sStat = sP;  // kT: stage s holds [-lse2 x 128 | -D' x 128] fp32 at +s*1024 (there is no sP tile)

  // dQ_i = dS K_j with A = dS read from TMEM (a bf16 copy written by the softmax warps into the 64 spare columns)
  // instead of from the K-major smem tile: -12 % shared-memory traffic per tile pair, at the price of a single-
  // buffered TMEM operand that must be free again before the next tile's first half is written.  With dropout the
  // softmax halves are long enough to hide that (7.9 -> 7.7 ms per launch); without dropout they are not (6.0 ->
  // 6.6 ms), so the variant is tied to the dropout instantiation.
  // With the keep-flag cache (kBits) the softmax halves are short again and the smem operand wins (6.5 vs 7.0 ms).
  constexpr bool kTsDq = ((kDrop && (!kBits || TVIT_ATTN_BWD_TSDQ_BITS) && TVIT_ATTN_BWD_TSDQ) || kFull) && !kT;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * kHdB;
  const int kv0 = jt * kTileB;
  const int dbg = tail_arg >> 8, tail = tail_arg & 0xff;  // dbg: timing experiments only (TVIT_ATTN_TAIL_DBG)
  const int nqt = (N + kTileB - 1) / kTileB;  // query tiles of the dQ accumulator
  const int Nq = N - tail;                    // queries that go through the tile loop
  const int nq = (Nq + kTileB - 1) / kTileB;  // tiles in the loop (== nqt unless there is a tail)

  if (threadIdx.x == 0) {
    mbar_init(&sm->kv_full, 1);
    for (int i = 0; i < kQStages; ++i) {
      mbar_init(&sm->qdo_full[i], 1);
      mbar_init(&sm->qdo_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm->s_full[i], 1);
      mbar_init(&sm->s_free[i], kSoftmaxWarps);  // one elected arrive per softmax warp
      mbar_init(&sm->ds_free[i], 1);
    }
    mbar_init(&sm->p_full, kSoftmaxWarps);
    mbar_init(&sm->pt_full[0], kSoftmaxWarps);
    mbar_init(&sm->pt_full[1], kSoftmaxWarps);
    mbar_init(&sm->p_free, 1);
    mbar_init(&sm->tds_free, 1);
    mbar_init(&sm->dq_full, 1);
    mbar_init(&sm->dq_free, 128);
    mbar_init(&sm->tail_ready, kFull ? 32 : 64);
    fence_barrier_init();
  }
  if (warp == 21) {
    tmem_alloc(&sm->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;
  const uint32_t tDS = tmem + 448;  // dS as bf16 [128 q lanes x 128 keys] = 64 columns: A operand of the dQ MMA

  // Role code below is warp-uniform (all 32 lanes run the loops and the barrier waits); only the TMA / MMA /
  // commit instructions themselves are predicated on one elected lane.  Issuing them from divergent code
  // (`if (lane == 0)`) makes ptxas wrap every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST loop (~100 clk each).
  if (warp == 20) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      tma_prefetch_desc(&tm_qkv);
      tma_prefetch_desc(&tm_do);
      mbar_expect_tx(&sm->kv_full, 2 * kTileBytesB);
      tma_load_3d(sK, &tm_qkv, &sm->kv_full, D + h * kHdB, kv0, b);
      tma_load_3d(sV, &tm_qkv, &sm->kv_full, 2 * D + h * kHdB, kv0, b);
    }
    __syncwarp();
    for (int i = 0; i < nq; ++i) {
      const int st = i % kQStages;
      const int qt = (i + jt) % nq;  // staggered start, see header
      mbar_wait_backoff(&sm->qdo_empty[st], (((uint32_t)i / kQStages) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&sm->qdo_full[st], 2 * kTileBytesB + (kT ? 1024 : 0));
        uint8_t* sQ = sQdO + st * 2 * kTileBytesB;
        tma_load_3d(sQ, &tm_qkv, &sm->qdo_full[st], h * kHdB, qt * kTileB, b);
        tma_load_3d(sQ + kTileBytesB, &tm_do, &sm->qdo_full[st], h * kHdB, qt * kTileB, b);
        if (kT)
          bulk_load(sStat + st * 1024, stat + (((long long)b * H + h) * nqt + qt) * 256, 1024, &sm->qdo_full[st]);
      }
      __syncwarp();
    }
  } else if (warp == 21) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_h = umma_idesc_bf16(128, 64, 0, 0);    // S, dP halves: both operands K-major
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);    // dV, dK: A (sP/sDS) MN-major, B MN-major
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);    // dQ: A (sDS) K-major, B (K_j) MN-major
    // Descriptor low words (start >> 4 | LBO field) are formed once per tile and stepped by constants, and the high
    // word is a compile-time constant: the issue loops must stay short because this warp competes for issue slots
    // with four busy softmax warps on its scheduler.
    constexpr uint32_t kHi = umma_desc_hi(1024);
    constexpr uint32_t kMn = (16384u >> 4) << 16;                 // LBO of the MN-major [2 x 64] tiles
    const uint32_t dK = umma_desc_lo(smem_u32(sK), 0), dV = umma_desc_lo(smem_u32(sV), 0);
    const uint32_t dQ0 = umma_desc_lo(smem_u32(sQdO), 0), dP = umma_desc_lo(smem_u32(sP), 0) | kMn;
    const uint32_t dDS0 = umma_desc_lo(smem_u32(sDS), 0);
    // S / dP of a tile for the keys [64 hf, 64 hf + 64): rows [64 hf, ..) of the K-major K_j / V_j tiles
    auto issue_half = [&](int st, int hf) {
      const uint32_t dQ = dQ0 + (uint32_t)st * (2 * kTileBytesB >> 4), dDO = dQ + (kTileBytesB >> 4);
      const uint32_t dKh = dK + (uint32_t)hf * (kTileBytesB / 2 >> 4), dVh = dV + (uint32_t)hf * (kTileBytesB / 2 >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHdB / 16; ++k)
          umma_ss(tS + hf * 64, umma_desc(dQ + 2 * k, kHi), umma_desc(dKh + 2 * k, kHi), idesc_h, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kHdB / 16; ++k)
          umma_ss(tDP + hf * 64, umma_desc(dDO + 2 * k, kHi), umma_desc(dVh + 2 * k, kHi), idesc_h, k > 0 ? 1u : 0u);
        tc_commit(&sm->s_full[hf]);
      }
      __syncwarp();
    };
    // ---- transposed formulation (kT) ----
    constexpr uint32_t idesc_a = umma_idesc_bf16(128, 64, 0, 1);  // dV, dK: A = P^T / dS^T in TMEM, B (dO_i / Q_i) MN-major
    // S^T / dP^T for the queries [64 hf, 64 hf + 64): A = all 128 rows of K_j / V_j, B = rows [64 hf, ..) of Q_i / dO_i
    auto issue_half_t = [&](int st, int hf) {
      const uint32_t dQ = dQ0 + (uint32_t)st * (2 * kTileBytesB >> 4) + (uint32_t)hf * (kTileBytesB / 2 >> 4);
      const uint32_t dDO = dQ + (kTileBytesB >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHdB / 16; ++k)
          umma_ss(tS + hf * 64, umma_desc(dK + 2 * k, kHi), umma_desc(dQ + 2 * k, kHi), idesc_h, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kHdB / 16; ++k)
          umma_ss(tDP + hf * 64, umma_desc(dV + 2 * k, kHi), umma_desc(dDO + 2 * k, kHi), idesc_h, k > 0 ? 1u : 0u);
        tc_commit(&sm->s_full[hf]);
      }
      __syncwarp();
    };
    // dV_j += P^T dO_i, dK_j += dS^T Q_i over the 64 queries of half hf: four k-steps of 16 queries whose bf16 A
    // operands sit in place in the first 8 of the 16 S^T / dP^T columns they were computed from
    auto issue_dvdk = [&](int st, int hf, uint32_t accum) {
      const uint32_t dQm = (dQ0 + (uint32_t)st * (2 * kTileBytesB >> 4)) | kMn, dDOm = dQm + (kTileBytesB >> 4);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ts(tDV, tS + 16 * (hf * 4 + kk), umma_desc(dDOm + 128 * (hf * 4 + kk), kHi), idesc_a, kk > 0 ? 1u : accum);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ts(tDK, tDP + 16 * (hf * 4 + kk), umma_desc(dQm + 128 * (hf * 4 + kk), kHi), idesc_a, kk > 0 ? 1u : accum);
      }
      __syncwarp();
    };
    mbar_wait(&sm->kv_full, 0);
    mbar_wait(&sm->qdo_full[0], 0);
    tc_fence_after();
    if constexpr (kT) {
      // Tensor-pipe order per tile: [dV_a dK_a | S^T_a' dP^T_a'] [dV_b dK_b | S^T_b' dP^T_b' | dQ]  (' = tile i+1).  The MMAs of
      // one thread execute in issue order, so refilling a half's S^T / dP^T columns right behind the dV / dK MMAs that
      // read P^T / dS^T from them needs no barrier.
      issue_half_t(0, 0);
      issue_half_t(0, 1);
      int st = 0;
      for (int i = 0; i < nq; ++i) {
        const int stn = st + 1 == kQStages ? 0 : st + 1;
        const bool more = i + 1 < nq;
        mbar_wait(&sm->pt_full[0], (uint32_t)i & 1u);  // P^T / dS^T of the first 64 queries written
        tc_fence_after();
        issue_dvdk(st, 0, i > 0 ? 1u : 0u);
        if (more) {
          mbar_wait(&sm->qdo_full[stn], ((uint32_t)(i + 1) / kQStages) & 1u);
          tc_fence_after();
          issue_half_t(stn, 0);
        }
        mbar_wait(&sm->pt_full[1], (uint32_t)i & 1u);
        tc_fence_after();
        issue_dvdk(st, 1, 1u);
        if (elect_one()) tc_commit(&sm->qdo_empty[st]);  // Q_i / dO_i / statistics of tile i are consumed
        __syncwarp();
        if (more) issue_half_t(stn, 1);
        if (i > 0) {
          mbar_wait(&sm->dq_free, (uint32_t)(i - 1) & 1u);  // drain warps have read dQ_{i-1}
          tc_fence_after();
        }
        // dQ_i = dS K_j: A = the [kv rows][q contiguous] dS^T tile read MN-major (M = q), B = K_j MN-major
        const uint32_t dDSm = (dDS0 + (uint32_t)(i & 1) * (kPBytes >> 4)) | kMn, dKm = dK | kMn;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileB / 16; ++k)
            umma_ss(tDQ, umma_desc(dDSm + 128 * k, kHi), umma_desc(dKm + 128 * k, kHi), idesc_t, k > 0 ? 1u : 0u);
          tc_commit(&sm->dq_full);
          tc_commit(&sm->ds_free[i & 1]);
        }
        __syncwarp();
        st = stn;
      }
    } else {
    constexpr uint32_t idesc_f = umma_idesc_bf16(128, 128, 0, 0);  // S, dP whole tiles (kFull)
    auto issue_full = [&](int st) {
      const uint32_t dQ = dQ0 + (uint32_t)st * (2 * kTileBytesB >> 4), dDO = dQ + (kTileBytesB >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHdB / 16; ++k)
          umma_ss(tS, umma_desc(dQ + 2 * k, kHi), umma_desc(dK + 2 * k, kHi), idesc_f, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kHdB / 16; ++k)
          umma_ss(tDP, umma_desc(dDO + 2 * k, kHi), umma_desc(dV + 2 * k, kHi), idesc_f, k > 0 ? 1u : 0u);
        tc_commit(&sm->s_full[0]);
        tc_commit(&sm->s_full[1]);
      }
      __syncwarp();
    };
    if constexpr (kFull) {
      // MMA issue is split over two warps (on different schedulers): measured with the event trace, the issuing warp --
      // which shares its scheduler with four busy softmax warps -- needs ~70 clk per tcgen05.mma, i.e. ~2800 clk per
      // tile pair for 40 MMAs: as long as the tensor pipe itself.  This warp issues S / dP of the next tile and dQ;
      // warp 22 issues dV and dK (below).
      issue_full(0);
      int st = 0;
      for (int i = 0; i < nq; ++i) {
        const int stn = st + 1 == kQStages ? 0 : st + 1;
        if (i + 1 < nq) {  // refill S / dP as soon as every softmax warp holds tile i's second half in registers
          mbar_wait(&sm->qdo_full[stn], ((uint32_t)(i + 1) / kQStages) & 1u);
          mbar_wait(&sm->s_free[1], (uint32_t)i & 1u);
          tc_fence_after();
          TVIT_TRACE(0, i, 0);
          issue_full(stn);
        }
        mbar_wait(&sm->p_full, (uint32_t)i & 1u);  // P / dS of tile i written (smem + the TMEM copy of dS)
        if (i > 0) mbar_wait(&sm->dq_free, (uint32_t)(i - 1) & 1u);  // drain warps have read dQ_{i-1}
        tc_fence_after();
        TVIT_TRACE(0, i, 1);
        const uint32_t dKm = dK | kMn;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kTileB / 16; ++k)
            umma_ts(tDQ, tDS + k * 8, umma_desc(dKm + 128 * k, kHi), idesc_q, k > 0 ? 1u : 0u);
          tc_commit(&sm->dq_full);
          tc_commit(&sm->tds_free);
        }
        __syncwarp();
        TVIT_TRACE(0, i, 2);
        st = stn;
      }
    } else {
    issue_half(0, 0);
    issue_half(0, 1);
    int st = 0;
    for (int i = 0; i < nq; ++i) {
      const int stn = st + 1 == kQStages ? 0 : st + 1;
      const bool more = i + 1 < nq;
      if (more) {  // refill half a while the softmax warps work on half b of tile i
        mbar_wait(&sm->qdo_full[stn], ((uint32_t)(i + 1) / kQStages) & 1u);
        mbar_wait(&sm->s_free[0], (uint32_t)i & 1u);
        tc_fence_after();
        issue_half(stn, 0);
      }
      mbar_wait(&sm->p_full, (uint32_t)i & 1u);  // sP / sDS written for tile i
      tc_fence_after();
      TVIT_TRACE(0, i, 1);
      // MN-major operands (LBO 16384): Q_i, dO_i as B; sP, sDS as A.  K-major sDS as A of the dQ MMA.
      const uint32_t dQm = (dQ0 + (uint32_t)st * (2 * kTileBytesB >> 4)) | kMn, dDOm = dQm + (kTileBytesB >> 4);
      const uint32_t dDS = dDS0 + (uint32_t)(i & 1) * (kPBytes >> 4), dDSm = dDS | kMn, dKm = dK | kMn;
      const uint32_t accum = i > 0 ? 1u : 0u;
      if (elect_one()) {
        // reduction over the 128 query rows in 8 steps of 16 (2048 B per step in the MN-major tiles)
#pragma unroll
        for (int k = 0; k < kTileB / 16; ++k)
          umma_ss(tDV, umma_desc(dP + 128 * k, kHi), umma_desc(dDOm + 128 * k, kHi), idesc_t, k > 0 ? 1u : accum);
        tc_commit(&sm->p_free);  // sP may be overwritten by tile i+1
      }
      __syncwarp();
      auto issue_dq = [&]() {
        if (i > 0) {
          mbar_wait(&sm->dq_free, (uint32_t)(i - 1) & 1u);  // drain warps have read dQ_{i-1}
          tc_fence_after();
        }
        TVIT_TRACE(0, i, 2);
        if (elect_one()) {
          // reduction over the 128 keys, B = K_j MN-major; A = dS from TMEM (8 columns per 16 keys) or from the
          // K-major smem tile (two 64-key blocks 16 KB apart)
#pragma unroll
          for (int k = 0; k < kTileB / 16; ++k) {
            if (kTsDq)
              umma_ts(tDQ, tDS + k * 8, umma_desc(dKm + 128 * k, kHi), idesc_q, k > 0 ? 1u : 0u);
            else
              umma_ss(tDQ, umma_desc(dDS + (k >> 2) * (16384 >> 4) + (k & 3) * 2, kHi), umma_desc(dKm + 128 * k, kHi),
                      idesc_q, k > 0 ? 1u : 0u);
          }
          tc_commit(&sm->dq_full);
          if (kTsDq) tc_commit(&sm->tds_free);
        }
        __syncwarp();
      };
      if (kTsDq) issue_dq();  // second, so that the TMEM copy of dS is free before tile i+1's first half is written
      if (more) {
        mbar_wait(&sm->s_free[1], (uint32_t)i & 1u);
        tc_fence_after();
        issue_half(stn, 1);
      }
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kTileB / 16; ++k)
          umma_ss(tDK, umma_desc(dDSm + 128 * k, kHi), umma_desc(dQm + 128 * k, kHi), idesc_t, k > 0 ? 1u : accum);
        if (kTsDq) {
          tc_commit(&sm->qdo_empty[st]);
          tc_commit(&sm->ds_free[i & 1]);
        }
      }
      __syncwarp();
      if (!kTsDq) {
        issue_dq();
        if (elect_one()) {
          tc_commit(&sm->qdo_empty[st]);
          tc_commit(&sm->ds_free[i & 1]);
        }
        __syncwarp();
      }
      st = stn;
    }
    }
    }
  } else if (kFull && warp == 22) {
    // ============================ second MMA issuer (kFull): dV_j += P^T dO_i, dK_j += dS^T Q_i ============================
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);
    constexpr uint32_t kHi = umma_desc_hi(1024);
    constexpr uint32_t kMn = (16384u >> 4) << 16;
    const uint32_t dQ0 = umma_desc_lo(smem_u32(sQdO), 0), dP = umma_desc_lo(smem_u32(sP), 0) | kMn;
    const uint32_t dDS0 = umma_desc_lo(smem_u32(sDS), 0);
    int st = 0;
    for (int i = 0; i < nq; ++i) {
      mbar_wait(&sm->p_full, (uint32_t)i & 1u);
      tc_fence_after();
      TVIT_TRACE(4, i, 0);
      const uint32_t dQm = (dQ0 + (uint32_t)st * (2 * kTileBytesB >> 4)) | kMn, dDOm = dQm + (kTileBytesB >> 4);
      const uint32_t dDSm = (dDS0 + (uint32_t)(i & 1) * (kPBytes >> 4)) | kMn;
      const uint32_t accum = i > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kTileB / 16; ++k)
          umma_ss(tDV, umma_desc(dP + 128 * k, kHi), umma_desc(dDOm + 128 * k, kHi), idesc_t, k > 0 ? 1u : accum);
        tc_commit(&sm->p_free);  // sP may be overwritten by tile i+1
      }
      __syncwarp();
      TVIT_TRACE(4, i, 1);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kTileB / 16; ++k)
          umma_ss(tDK, umma_desc(dDSm + 128 * k, kHi), umma_desc(dQm + 128 * k, kHi), idesc_t, k > 0 ? 1u : accum);
        tc_commit(&sm->qdo_empty[st]);
        tc_commit(&sm->ds_free[i & 1]);
      }
      __syncwarp();
      TVIT_TRACE(4, i, 2);
      st = st + 1 == kQStages ? 0 : st + 1;
    }
  } else if (warp < kSoftmaxWarps) {
    // ===== softmax warps: thread == query row (TMEM lane quarter warp % 4), 16 keys (warp / 4) of each key half =====
    // (kT: thread == key row, 16 queries (warp / 4) of each query half)
    const int qd4 = warp & 3, chunk = warp >> 2;
    const int r = qd4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const float c_log2 = scale * 1.4426950408889634f;
    const f32x2 c_log2_2 = pk2(c_log2, c_log2), scale_2 = pk2(scale, scale);
    if constexpr (kT) {
      // Dropout: lane L generates the keep bytes of (query q0 + (L & 15), key group L >> 4 of this warp's 32 key rows);
      // a 16 x 16 byte transpose inside each half-warp then hands every lane the bytes of its own key row for the 16
      // queries -- the masks stay exactly the forward kernel's (one Philox group = 16 consecutive keys of one query).
      const unsigned long long bh_row0 = ((unsigned long long)b * H + h) * (unsigned long long)N;
      const unsigned long long npad16 = (unsigned long long)(((N + 15) & ~15) >> 4);
      const unsigned long long kgrp = (unsigned long long)((kv0 + qd4 * 32 + (lane >> 4) * 16) >> 4);
      int st3 = 0, qt = jt % nq;
      for (int i = 0; i < nq; ++i) {
        const uint32_t aDSbuf = smem_u32(sDS) + (uint32_t)(i & 1) * kPBytes;
        const uint32_t aStat = smem_u32(sStat) + (uint32_t)st3 * 1024u + (uint32_t)chunk * 64u;
        mbar_wait(&sm->qdo_full[st3], ((uint32_t)i / kQStages) & 1u);  // column statistics of tile i (already landed)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {  // query halves
          uint32_t sv[16], dp[16];
          mbar_wait(&sm->s_full[hf], (uint32_t)i & 1u);
          tc_fence_after();
          tmem_ld16(tS + lane_off + hf * 64 + chunk * 16, sv);
          tmem_ld16(tDP + lane_off + hf * 64 + chunk * 16, dp);
          uint32_t mk[4] = {0, 0, 0, 0};
          if (kDrop) {
            const int ql = qt * kTileB + hf * 64 + chunk * 16 + (lane & 15);
            const unsigned long long grp = (bh_row0 + (unsigned long long)(ql < N ? ql : 0)) * npad16 + kgrp;
            drop_bits16(drop, grp, mk);
            drop_keep_sign16(mk, drop_thr8(drop, grp));
            transpose16x16_bytes(mk, lane);
          }
          tmem_ld_wait();
          uint32_t pk[8], dk[8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {  // queries 4u .. 4u+3 of this thread's 16
            const float4 nl = ld_shared_f4(aStat + (uint32_t)(hf * 256 + u * 16));        // -lse2 (+inf rows: -inf)
            const float4 nd = ld_shared_f4(aStat + 512u + (uint32_t)(hf * 256 + u * 16));  // -D'
            const float nls[4] = {nl.x, nl.y, nl.z, nl.w}, nds[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              const int e = 4 * u + 2 * v, t = 2 * u + v;
              float e0, e1;
              const f32x2 nl2 = pk2(nls[2 * v], nls[2 * v + 1]), nd2 = pk2(nds[2 * v], nds[2 * v + 1]);
              up2(fma2(pk2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), c_log2_2, nl2), e0, e1);
              const f32x2 p = pk2(ex2_approx(e0), ex2_approx(e1));
              const f32x2 a = fma2(pk2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), scale_2, nd2);
              if (kDrop) {
                const uint32_t m = v ? prmt<0xBBAAu>(mk[u], 0u) : prmt<0x9988u>(mk[u], 0u);
                const uint32_t kept = pack_bf16_2(mul2(p, a)), dropped = pack_bf16_2(mul2(p, nd2));
                pk[t] = pack_bf16_2(p) & m;
                dk[t] = (kept & m) | (dropped & ~m);
              } else {
                pk[t] = pack_bf16_2(p);
                dk[t] = pack_bf16_2(mul2(p, a));
              }
            }
          }
          if (hf == 0 && i >= 2)
            mbar_wait(&sm->ds_free[i & 1], (((uint32_t)i >> 1) - 1u) & 1u);  // dQ of tile i-2 has read this dS^T buffer
          tmem_st8(tS + lane_off + hf * 64 + chunk * 16, pk);   // P^T, in place: A operand of dV
          tmem_st8(tDP + lane_off + hf * 64 + chunk * 16, dk);  // dS^T, in place: A operand of dK
          // key row r of 64-query block hf: 16-byte pieces chunk * 2 + g, XOR-swizzled with (r & 7)
          const uint32_t row_off = (uint32_t)hf * 16384u + (uint32_t)r * 128u;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t piece = (uint32_t)((chunk * 2 + g) ^ (r & 7)) * 16u;
            st_shared_v4(aDSbuf + row_off + piece, dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
          }
          tmem_st_wait();
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm->pt_full[hf]);
        }
        qt = (qt + 1 == nq) ? 0 : qt + 1;
        st3 = st3 + 1 == kQStages ? 0 : st3 + 1;
      }
    } else {
    // Per-query statistics (-lse2, -D') come ready-made from the prep kernel's records; with dropout (keep mask k,
    // m = k / (1-p)): P m = P' k and dS = P (m dP scale - D scale) = P' (k dP scale - D'), where P' = P / (1-p) comes
    // for free from the exponent (lse2 = lse log2(e) - log2(1/(1-p))) and D' = D scale (1-p).  Rows past N: -inf / 0.
    const float* stat_row = stat + ((long long)b * H + h) * nqt * 256 + r;
    int qt = jt % nq;
    float nl_next = stat_row[qt * 256], nd_next = stat_row[qt * 256 + 128];
    // Dropout group (16 consecutive keys of one query row) of this thread's keys: a CTA-uniform 64-bit base plus a
    // 32-bit row term (q * npad / 16 < 2^32 for every N this kernel accepts)
    const uint32_t npad16 = (uint32_t)((N + 15) >> 4);
    const unsigned long long grp_cta = (attn_drop_row_base(b, H, h, N, 0) + (unsigned long long)kv0) >> 4;
    // keep-bit records of this key tile: [(b,h)][q tile][key tile][row][8 groups] 16-bit fields (bits 0-7: even keys,
    // 8-15: odd keys); this thread reads groups chunk and 4 + chunk = half (chunk & 1) of words chunk >> 1 and 2 + (chunk >> 1)
    const uint32_t* kb_row =
        kBits ? keepbits + ((((long long)b * H + h) * nqt * nqt + jt) * kTileB + r) * 4 + (chunk >> 1) : nullptr;
    const int kb_qstride = nqt * kTileB * 4;
    const uint32_t kb_sel = (chunk & 1) ? 0x4342u : 0x4140u;  // prmt: field -> [even byte, 0, odd byte, 0]
    uint32_t kb_next0 = 0, kb_next1 = 0;
    if (kBits) {
      kb_next0 = __ldg(kb_row + qt * kb_qstride);
      kb_next1 = __ldg(kb_row + qt * kb_qstride + 2);
    }
    for (int i = 0; i < nq; ++i) {
      const float nlse2 = nl_next, negDq = nd_next;
      const f32x2 nlse2_2 = pk2(nlse2, nlse2), negDq_2 = pk2(negDq, negDq);
      const unsigned long long grp_row = grp_cta + (unsigned long long)((uint32_t)(qt * kTileB + r) * npad16);
      const uint32_t kb0 = kb_next0, kb1 = kb_next1;
      {  // prefetch the next tile's row statistics so the global-load latency is off the critical path
        qt = (qt + 1 == nq) ? 0 : qt + 1;
        nl_next = stat_row[qt * 256];
        nd_next = stat_row[qt * 256 + 128];
        if (kBits) {
          kb_next0 = __ldg(kb_row + qt * kb_qstride);
          kb_next1 = __ldg(kb_row + qt * kb_qstride + 2);
        }
      }
      const uint32_t aDSbuf = smem_u32(sDS) + (uint32_t)(i & 1) * kPBytes;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // key halves; 16 keys per thread per half keep the live register set small
        uint32_t sv[16], dp[16];
        mbar_wait(&sm->s_full[hf], (uint32_t)i & 1u);
        tc_fence_after();
        TVIT_TRACE(1 + (warp == 15), i, hf * 4 + 0);
        tmem_ld16(tS + lane_off + hf * 64 + chunk * 16, sv);
        tmem_ld16(tDP + lane_off + hf * 64 + chunk * 16, dp);
        tmem_ld_wait();
        TVIT_TRACE(1 + (warp == 15), i, hf * 4 + 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->s_free[hf]);  // this half's S / dP columns may be refilled
        uint32_t pk[8], dk[8];
        uint32_t w[4] = {0, 0, 0, 0}, tg2 = 0;
        const uint32_t kbw = kBits ? prmt_r(hf == 0 ? kb0 : kb1, 0u, kb_sel) : 0u;  // bit t / 16 + t: element 2t / 2t + 1
        if (kDrop && !kBits) {  // this thread's 16 keys are exactly one Philox group (common.cuh)
          const unsigned long long grp = grp_row + (unsigned long long)(hf * 4 + chunk);
          drop_bits16(drop, grp, w);
          tg2 = drop_tgc(drop_thr8(drop, grp));
        }
        // packed fp32x2 arithmetic (FFMA2 / FMUL2: two IEEE operations per issue slot, bit-identical per lane to the
        // scalar form): the TMEM loads deliver each element pair in an aligned register pair
#pragma unroll
        for (int t = 0; t < 8; ++t) {  // element pairs (2t, 2t+1)
          float e0, e1;
          up2(fma2(pk2(__uint_as_float(sv[2 * t]), __uint_as_float(sv[2 * t + 1])), c_log2_2, nlse2_2), e0, e1);
          const f32x2 p = pk2(ex2_approx(e0), ex2_approx(e1));
          const f32x2 a = fma2(pk2(__uint_as_float(dp[2 * t]), __uint_as_float(dp[2 * t + 1])), scale_2, negDq_2);
          if (kDrop) {
            const uint32_t m = kBits   ? prmt<0xBB99u>(kbw << (15 - t), 0u)
                               : (t & 1) ? drop_keep_mask2<1>(w[t >> 1], tg2)
                                         : drop_keep_mask2<0>(w[t >> 1], tg2);
            const uint32_t kept = pack_bf16_2(mul2(p, a)), dropped = pack_bf16_2(mul2(p, negDq_2));
            pk[t] = pack_bf16_2(p) & m;
            dk[t] = (kept & m) | (dropped & ~m);
          } else {
            pk[t] = pack_bf16_2(p);
            dk[t] = pack_bf16_2(mul2(p, a));
          }
        }
        TVIT_TRACE(1 + (warp == 15), i, hf * 4 + 2);
        if (hf == 0) {
          // One wait covers all three operand buffers.  A tcgen05.commit completes when ALL earlier MMAs of the issuing
          // thread have, and that thread issues dK_{i-2} (last reader of this sDS buffer) before dV_{i-1} (sP) before,
          // with the TMEM dS operand, dQ_{i-1} (tDS): the latest of the commits implies the others.  Each extra wait
          // on an already completed barrier cost ~130 clk here (a shared-memory round trip behind the UMMA operand
          // traffic), ~300 clk per tile pair.  With two issuing warps (kFull) dV / dK and dQ are committed separately.
          if (kTsDq && i >= 1) mbar_wait(&sm->tds_free, (uint32_t)(i - 1) & 1u);  // dQ_{i-1} has read tDS
          if ((!kTsDq || kFull) && i >= 1) mbar_wait(&sm->p_free, (uint32_t)(i - 1) & 1u);  // dV_{i-1} has read sP
        }
        // (Deferring half a's TMEM store to the end of the tile, where the tensor pipe is idle and a tcgen05.st does not
        // stall for 300-800 clk, lengthens the p_full -> dV -> dQ chain instead: 7.25 -> 7.56 ms.  Not kept.)
        if (kTsDq) tmem_st8(tDS + lane_off + hf * 32 + chunk * 8, dk);  // dS for the dQ MMA (A operand from TMEM)
        TVIT_TRACE(1 + (warp == 15), i, hf * 4 + 3);
        // row r of 64-key block hf: 16-byte pieces chunk * 2 + g, XOR-swizzled with (r & 7)
        const uint32_t row_off = (uint32_t)hf * 16384u + (uint32_t)r * 128u;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t piece = (uint32_t)((chunk * 2 + g) ^ (r & 7)) * 16u;
          st_shared_v4(smem_u32(sP) + row_off + piece, pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          st_shared_v4(aDSbuf + row_off + piece, dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
        }
        if (warp == 0) TVIT_TRACE(5 + hf, i, hf == 0 ? 5 : 1);
      }
      if (kTsDq) {
        tmem_st_wait();
        tc_fence_before();
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->p_full);
    }
    }
    // ---- epilogue: dK_j, dV_j from TMEM -> bf16 rows of dqkv ----
    mbar_wait(&sm->ds_free[(nq - 1) & 1], ((uint32_t)(nq - 1) >> 1) & 1u);
    tc_fence_after();
    const int kv = kv0 + r;
    __nv_bfloat16* drow = dqkv + ((long long)b * N + kv) * (3LL * D) + h * kHdB;
    {  // warps with chunk 0/1 store dK columns [0,32)/[32,64) -> [D, 2D); chunk 2/3 store dV -> [2D, 3D)
      const int which = chunk >> 1, c = chunk & 1;
      const uint32_t tsrc = which == 0 ? tDK : tDV;
      uint32_t o[32];
      tmem_ld32(tsrc + lane_off + c * 32, o);
      tmem_ld_wait();
      if (tail > 0 && !(dbg & 2)) {  // ---- tail query (see header): rank-1 updates with the values the drain warps left in smem ----
        const int qi = Nq;
        mbar_wait(&sm->tail_ready, 0);
        const float f = which == 0 ? sm->tail_ds[r] : sm->tail_pm[r];
        const __nv_bfloat16* src = (which == 0 ? qkv + ((long long)b * N + qi) * (3LL * D) + h * kHdB
                                               : dout + ((long long)b * N + qi) * D + h * kHdB) + c * 32;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const uint4 vv = __ldg(reinterpret_cast<const uint4*>(src + 8 * cc));
          const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vw[i]));
            o[8 * cc + 2 * i] = __float_as_uint(fmaf(f, g.x, __uint_as_float(o[8 * cc + 2 * i])));
            o[8 * cc + 2 * i + 1] = __float_as_uint(fmaf(f, g.y, __uint_as_float(o[8 * cc + 2 * i + 1])));
          }
        }
      }
      if (kv < N) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {  // 16 bf16 = one full 32-byte sector per store
          uint32_t v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            v[u] = pack_bf16(__uint_as_float(o[16 * t + 2 * u]), __uint_as_float(o[16 * t + 2 * u + 1]));
          st_global_b32x8(drow + (which + 1) * D + c * 32 + 16 * t, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        }
      }
      if (colsum) {  // dk / dv parts of the qkv-bias gradient: this warp's 32 key rows x 32 columns (once per CTA)
        float cv[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) cv[u] = kv < N ? __uint_as_float(o[u]) : 0.f;
        const float tot = warp_colsum32(cv, lane);
        atomicAdd(colsum + (which + 1) * D + h * kHdB + c * 32 + lane, tot);
      }
    }
  } else if (warp < 20) {
    // ============================ dQ drain warps (16-19) ============================
    const int qd = warp & 3;  // TMEM lane quarter (warp % 4)
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float* acc_bh = dqacc + ((long long)b * H + h) * nqt * (16LL * 128 * 4);
    for (int i = 0; i < nq; ++i) {
      mbar_wait_backoff(&sm->dq_full, (uint32_t)i & 1u);
      tc_fence_after();
      TVIT_TRACE(3, i, 0);
      float* tile = acc_bh + (long long)((i + jt) % nq) * (16 * 128 * 4) + r * 4;
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {  // 32 columns at a time: this warpgroup runs with 64 registers
        uint32_t o[32];
        tmem_ld32(tDQ + lane_off + hc * 32, o);
        tmem_ld_wait();
        if (hc == 1) {
          tc_fence_before();
          mbar_arrive(&sm->dq_free);
          TVIT_TRACE(3, i, 1);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          atomicAdd(reinterpret_cast<float4*>(tile + (hc * 8 + c) * 512),
                    make_float4(__uint_as_float(o[4 * c]), __uint_as_float(o[4 * c + 1]), __uint_as_float(o[4 * c + 2]),
                                __uint_as_float(o[4 * c + 3])));
        }
      }
      TVIT_TRACE(3, i, 2);
    }
  } else if (warp >= 22) {
    // ============================ tail-query warps (22-23; see header) ============================
    // Two otherwise idle warps: thread == key rows kv0 + r and kv0 + r + 64.  They are not part of the tile pipeline,
    // so the latency of their global loads costs nothing (on the dQ-drain warps the same work delayed dq_free of the
    // following tile and cost 7 % of the kernel: profiles/r2_kern_v14_tail_bwd.log).
    if (tail > 0 && !(dbg & 1)) {
      const int qi = Nq;
      const float c_log2 = scale * 1.4426950408889634f;
      const __nv_bfloat16* qrow = qkv + ((long long)b * N + qi) * (3LL * D) + h * kHdB;
      const __nv_bfloat16* dorow = dout + ((long long)b * N + qi) * D + h * kHdB;
      const float lse_t = lse[((long long)b * H + h) * N + qi], d_t = dvec[((long long)b * H + h) * N + qi];
      mbar_wait(&sm->kv_full, 0);
      constexpr int kTailRows = kFull ? 4 : 2;  // kFull: warp 22 issues MMAs, warp 23 takes all 128 key rows
      float dsr[kTailRows];
#pragma unroll
      for (int rr = 0; rr < kTailRows; ++rr) {
        const int r = kFull ? lane + 32 * rr : (warp - 22) * 32 + lane + 64 * rr, kv = kv0 + r;
        const uint32_t krow = smem_u32(sK) + (uint32_t)r * 128u, vrow = smem_u32(sV) + (uint32_t)r * 128u;
        float sc = 0.f, dpv = 0.f;
#pragma unroll 4
        for (int cc = 0; cc < 8; ++cc) {  // 16-byte pieces of the 128B-swizzled K / V rows
          const uint32_t sw = (uint32_t)((cc ^ (r & 7)) * 16);
          sc = dot8_bf16(ld_shared_u4(krow + sw), __ldg(reinterpret_cast<const uint4*>(qrow + 8 * cc)), sc);
          dpv = dot8_bf16(ld_shared_u4(vrow + sw), __ldg(reinterpret_cast<const uint4*>(dorow + 8 * cc)), dpv);
        }
        const float pt = ex2_approx(fmaf(sc, c_log2, -lse_t * 1.4426950408889634f));
        float mlt = 1.0f;
        if (kDrop) mlt = drop_keep(drop, attn_drop_row_base(b, H, h, N, qi) + (unsigned long long)kv) ? drop.inv_keep : 0.f;
        float dst = pt * (dpv * mlt - d_t) * scale, pmt = pt * mlt;
        if (kv >= N) dst = pmt = 0.f;
        sm->tail_ds[r] = dst;
        sm->tail_pm[r] = pmt;
        dsr[rr] = dst;
      }
      mbar_arrive(&sm->tail_ready);
      // dQ_t += sum over this warp's 64 key rows of dS k  (accumulator tile nq, row 0)
      float* trow = dqacc + (((long long)b * H + h) * nqt + nq) * (16LL * 128 * 4);
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        float v8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = 0.f;
#pragma unroll
        for (int rr = 0; rr < kTailRows; ++rr) {
          const int r = kFull ? lane + 32 * rr : (warp - 22) * 32 + lane + 64 * rr;
          const uint4 kk = ld_shared_u4(smem_u32(sK) + (uint32_t)r * 128u + (uint32_t)((cc ^ (r & 7)) * 16));
          const uint32_t kw[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kw[i]));
            v8[2 * i] = fmaf(dsr[rr], g.x, v8[2 * i]);
            v8[2 * i + 1] = fmaf(dsr[rr], g.y, v8[2 * i + 1]);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
          for (int i = 0; i < 8; ++i) v8[i] += __shfl_xor_sync(0xffffffffu, v8[i], off);
        float val = v8[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) val = (lane == i) ? v8[i] : val;
        // column d = 8 cc + lane of accumulator row 0: chunk d / 4, element d % 4
        if (lane < 8) atomicAdd(trow + (2 * cc + (lane >> 2)) * 512 + (lane & 3), val);
      }
    } else if (tail > 0) {  // timing experiment: no tail work
      for (int r = kFull ? lane : (warp - 22) * 32 + lane; r < kTileB; r += kFull ? 32 : 64)
        sm->tail_ds[r] = sm->tail_pm[r] = 0.f;
      mbar_arrive(&sm->tail_ready);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 21) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int make_tok_tmap(CUtensorMap* tm, const void* base, int B, int N, int cols) {
  const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)cols * 2, (uint64_t)N * cols * 2};
  const uint32_t box[3] = {64, (uint32_t)kTileB, 1};
  return make_tmap_bf16(tm, base, 3, dims, strides, box);
}

// workspace: Dvec [B,H,N] fp32 | dQ accumulator [B*H][nq][16][128][4] fp32 | column statistics [B*H][nq][256] fp32
static size_t ws_dvec_bytes(int B, int N, int H) { return (((size_t)B * H * N * 4) + 255) & ~(size_t)255; }
static size_t ws_dq_bytes(int B, int N, int H) {
  return (size_t)B * H * ((N + kTileB - 1) / kTileB) * 16 * 128 * 4 * sizeof(float);
}
static size_t ws_stat_bytes(int B, int N, int H) { return (size_t)B * H * ((N + kTileB - 1) / kTileB) * 256 * sizeof(float); }

size_t tc_attn_bwd_workspace(int B, int N, int H, int hd) {
  (void)hd;
  return ws_dvec_bytes(B, N, H) + ws_dq_bytes(B, N, H) + ws_stat_bytes(B, N, H);
}

// Formulation of the kernel per instantiation (bit mask; DESIGN.md section 4.3 has the measurements):
//   bit 0 / 1: transposed (S^T = K Q^T, P^T / dS^T as TMEM operands of dV / dK) without / with dropout
//   bit 2 / 3: whole-tile S / dP MMAs + two issuing warps without / with dropout
// 0 = key-half pipelined (default: fastest on B200 at the bench shapes).  TVIT_ATTN_BWD_T sets the initial value,
// tvit_attn_bwd_variant() changes it at run time (tests run every variant against the oracle).
static int& attn_bwd_t_mask() {
  static int m = [] { const char* e = getenv("TVIT_ATTN_BWD_T"); return e ? atoi(e) : TVIT_ATTN_BWD_T_DEFAULT; }();
  return m;
}
int tc_attn_bwd_variant(int mask) {
  const int old = attn_bwd_t_mask();
  if (mask >= 0) attn_bwd_t_mask() = mask & 15;
  return old;
}

int tc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* ws,
                size_t ws_bytes, int B, int N, int H, int hd, const tvit_dropout* drop, float* dqkv_colsum,
                const void* keepbits, cudaStream_t s) {
  if (hd != kHdB) return fail(TVIT_ERR_UNSUPPORTED, "tcgen05 attention supports head_dim 64 only (got %d)", hd);
  if (!ws || ws_bytes < tc_attn_bwd_workspace(B, N, H, hd))
    return fail(TVIT_ERR_BAD_ARG, "attn_bwd: workspace too small (%zu < %zu)", ws_bytes,
                tc_attn_bwd_workspace(B, N, H, hd));
  const int D = H * hd;
  const int nq = (N + kTileB - 1) / kTileB;
  float* dvec = reinterpret_cast<float*>(ws);
  float* dqacc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + ws_dvec_bytes(B, N, H));
  const size_t dq_bytes = ws_dq_bytes(B, N, H);
  float* stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + ws_dvec_bytes(B, N, H) + dq_bytes);

  constexpr int smem_bytes = 2 * kTileBytesB + kQStages * 2 * kTileBytesB + 3 * kPBytes + 1024 + 1280;  // 226.25 KB
  int rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<false, false, false>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<true, false, false>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<false, false, true>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<true, false, true>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<false, true, false>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<true, true, false>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_bwd_kernel<true, false, false, true>, smem_bytes)) != TVIT_OK) return rc;

  TVIT_CUDA_OK(cudaMemsetAsync(dqacc, 0, dq_bytes, s));
  CUtensorMap tm_qkv, tm_do;
  if ((rc = make_tok_tmap(&tm_qkv, qkv, B, N, 3 * D)) != TVIT_OK) return rc;
  if ((rc = make_tok_tmap(&tm_do, dout, B, N, D)) != TVIT_OK) return rc;
  dim3 grid(nq, H, B);
  const float scale = 1.0f / sqrtf((float)hd);
  const DropCfg dc = make_drop(drop);
  int tail = attn_tail(N, 2);
  if (const char* e = getenv("TVIT_ATTN_TAIL_DBG")) tail |= atoi(e) << 8;
  const __nv_bfloat16* qp = (const __nv_bfloat16*)qkv;
  const __nv_bfloat16* dop = (const __nv_bfloat16*)dout;
  const bool drop_on = dc.thr16 != 0;
  const bool transposed = (attn_bwd_t_mask() >> (drop_on ? 1 : 0)) & 1;
  const bool full = !transposed && ((attn_bwd_t_mask() >> (drop_on ? 3 : 2)) & 1);
  {
    const long long total = (long long)B * nq * kTileB * H;
    attn_bwd_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, lse, dvec, stat, B, N, H, nq,
        drop_on ? log2f(dc.inv_keep) : 0.f, scale / dc.inv_keep);
    TVIT_LAUNCH_OK();
  }
#define TVIT_BWD_LAUNCH(DROP, T, F)                                                                               \
  tc_attn_bwd_kernel<DROP, T, F><<<grid, kAttnBwdThreads, smem_bytes, s>>>(                                        \
      tm_qkv, tm_do, qp, dop, lse, dvec, stat, dqacc, (__nv_bfloat16*)dqkv, dqkv_colsum, N, tail, H, scale, dc,   \
      (const uint32_t*)keepbits)
  if (drop_on) {
    if (transposed) TVIT_BWD_LAUNCH(true, true, false);
    else if (full) TVIT_BWD_LAUNCH(true, false, true);
    else if (keepbits)
      tc_attn_bwd_kernel<true, false, false, true><<<grid, kAttnBwdThreads, smem_bytes, s>>>(
          tm_qkv, tm_do, qp, dop, lse, dvec, stat, dqacc, (__nv_bfloat16*)dqkv, dqkv_colsum, N, tail, H, scale, dc,
          (const uint32_t*)keepbits);
    else TVIT_BWD_LAUNCH(true, false, false);
  } else {
    if (transposed) TVIT_BWD_LAUNCH(false, true, false);
    else if (full) TVIT_BWD_LAUNCH(false, false, true);
    else TVIT_BWD_LAUNCH(false, false, false);
  }
#undef TVIT_BWD_LAUNCH
  TVIT_LAUNCH_OK();
  {
    const int split = nq >= 4 ? 4 : 1;  // B H x 4 blocks of 512 threads: several waves even at small B H
    attn_bwd_dq_finish_kernel<<<dim3((unsigned)(B * H), split), 512, 0, s>>>(dqacc, (__nv_bfloat16*)dqkv, dqkv_colsum, N, H,
                                                                           nq);
    TVIT_LAUNCH_OK();
  }
  return TVIT_OK;
}

}  // namespace tvit

#ifdef TVIT_ATTN_TRACE
extern "C" int tvit_attn_bwd_trace(long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, tvit::g_attn_trace, (size_t)n * sizeof(long long));
}
#endif
