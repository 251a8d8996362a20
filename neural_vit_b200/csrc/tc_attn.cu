// Fused flash-style multi-head self-attention for sm_100a (head_dim 64, bf16 operands, fp32 softmax).
// Forward: one CTA per (128-query tile, head, sample).
//   warp 4   TMA producer: Q tile once, then K_j / V_j tiles (128 x 64 bf16, 128B swizzle) into a 2-stage ring
//   warp 5   MMA issuer  : S_j = Q K_j^T (tcgen05.mma SS, 128x128x64) and O += P_j V_j (tcgen05.mma TS: P from
//                          TMEM, V as an MN-major smem operand, 128x64x128); also owns the TMEM allocation
//   warps 0-3 softmax    : thread == query row; S row from TMEM (tcgen05.ld), online softmax with running
//                          max / sum in registers, P -> bf16 -> TMEM (tcgen05.st), conditional O rescale
// The S_{j+1} MMA is issued as soon as S_j has been read into registers, so it overlaps softmax_j; two
// CTAs per SM (80 KB smem, 256 TMEM columns each) overlap one CTA's exponentials with the other's MMAs.
// Ragged tail: when N = 128 m + t with a small t (the CLS token makes N = 2049 = 16 * 128 + 1 at the bench shape), the
// last t keys are NOT given a 17th, almost empty 128-key tile (it cost 1/17 of the kernel); their scores are t dot
// products per query row, computed by the softmax threads on the CUDA cores from the Q tile in shared memory: they
// initialise the running max / sum before the tile loop and their P V contribution is added in the epilogue.
// Only the row log-sum-exp is kept for backward.  qkv is read in place through a 3-D tensor map
// {3D, N, B} (rows past N are zero-filled by TMA), the output is written token-major [B*N, H*64].
#include "tc_common.cuh"

#ifndef TVIT_ATTN_FWD_PASSW
#define TVIT_ATTN_FWD_PASSW 16
#endif

namespace tvit {

constexpr int kHd = 64;
constexpr int kTile = 128;
constexpr int kAttnFwdThreads = 320;  // warps 0-7 softmax (2 threads per query row), 8 TMA, 9 MMA
constexpr int kTileBytes = kTile * kHd * 2;  // 16384

// 8 bf16 (one 16-byte piece) -> fp32 dot-product accumulation helpers
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

struct AttnFwdSmem {
  uint64_t q_full, kv_full[2], kv_empty[2], s_full, s_free, p_full, pv_done, tail_full;
  uint32_t tmem_base;
  uint32_t pad_;
  float xchg[2][2][128];  // [tile parity][key half][row]: partner exchange of row maxima (and final row sums)
  uint4 tail_kv[16];      // the tail key's K row (pieces 0-7) and V row (pieces 8-15), staged once per CTA
};
static_assert(sizeof(AttnFwdSmem) <= 256 + 2048 + 512, "AttnFwdSmem must fit its slot of the dynamic smem block");

template <bool kDrop>
__global__ void __launch_bounds__(kAttnFwdThreads, 2)
tc_attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __nv_bfloat16* __restrict__ qkv,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N, int tail, int H, float scale_log2,
                   DropCfg drop, uint32_t* __restrict__ keepbits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTileBytes;  // stage s: K at sKV + s*2*kTileBytes, V right after K
  AttnFwdSmem* sm = reinterpret_cast<AttnFwdSmem*>(smem + 5 * kTileBytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * kHd;
  const int q0 = qt * kTile;
  const int Nk = N - tail;  // keys covered by 128-key tiles; the remaining `tail` keys go through the CUDA cores
  const int nkv = (Nk + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    mbar_init(&sm->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm->kv_full[i], 1);
      mbar_init(&sm->kv_empty[i], 1);
    }
    mbar_init(&sm->s_full, 1);
    mbar_init(&sm->s_free, 8);  // one elected arrive per softmax warp
    mbar_init(&sm->p_full, 8);
    mbar_init(&sm->pv_done, 1);
    mbar_init(&sm->tail_full, 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(&sm->tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t tS = tmem, tO = tmem + 128, tP = tmem + 192;

  if (warp == 8) {
    // ============================ TMA producer ============================
    // warp-uniform loop; TMA instructions predicated on one elected lane (see the note in tc_gemm.cu)
    if (elect_one()) {
      tma_prefetch_desc(&tm_qkv);
      mbar_expect_tx(&sm->q_full, kTileBytes);
      tma_load_3d(sQ, &tm_qkv, &sm->q_full, h * kHd, q0, b);
    }
    __syncwarp();
    if (tail > 0) {  // stage the tail key's K and V rows (2 x 128 B) so that no softmax thread waits on global memory
      if (lane < 16) {
        const __nv_bfloat16* row = qkv + ((long long)b * N + Nk) * (3LL * D) + (lane < 8 ? D : 2 * D) + h * kHd;
        sm->tail_kv[lane] = __ldg(reinterpret_cast<const uint4*>(row + 8 * (lane & 7)));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->tail_full);
    }
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1;
      mbar_wait_backoff(&sm->kv_empty[st], (((uint32_t)j >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&sm->kv_full[st], 2 * kTileBytes);
        uint8_t* sK = sKV + st * 2 * kTileBytes;
        tma_load_3d(sK, &tm_qkv, &sm->kv_full[st], D + h * kHd, j * kTile, b);
        tma_load_3d(sK + kTileBytes, &tm_qkv, &sm->kv_full[st], 2 * D + h * kHd, j * kTile, b);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B = V: MN-major
    // descriptor low words formed once and stepped by constants (short issue loops: this warp shares its
    // scheduler with busy softmax warps)
    constexpr uint32_t kHi = umma_desc_hi(1024);
    const uint32_t dQ = umma_desc_lo(smem_u32(sQ), 0), dKV0 = umma_desc_lo(smem_u32(sKV), 0);
    mbar_wait(&sm->q_full, 0);
    auto issue_s = [&](int j) {
      const int st = j & 1;
      mbar_wait(&sm->kv_full[st], ((uint32_t)j >> 1) & 1u);
      tc_fence_after();
      const uint32_t dK = dKV0 + (uint32_t)st * (2 * kTileBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_ss(tS, umma_desc(dQ + 2 * k, kHi), umma_desc(dK + 2 * k, kHi), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(&sm->s_full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1;
      if (j + 1 < nkv) {
        mbar_wait(&sm->s_free, (uint32_t)j & 1u);  // softmax_j holds S_j in registers
        tc_fence_after();
        issue_s(j + 1);
      }
      mbar_wait(&sm->p_full, (uint32_t)j & 1u);  // P_j in TMEM, O rescaled
      tc_fence_after();
      const uint32_t dVm = (dKV0 + (uint32_t)st * (2 * kTileBytes >> 4) + (kTileBytes >> 4)) | ((16384u >> 4) << 16);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)
          umma_ts(tO, tP + k * 8, umma_desc(dVm + 128 * k, kHi), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        tc_commit(&sm->kv_empty[st]);
        tc_commit(&sm->pv_done);
      }
      __syncwarp();
    }
  } else {
    // ============================ softmax warps ============================
    // Two threads per query row: warp w owns TMEM lane quarter (w & 3) and key half (w >> 2), i.e. 64 of the
    // tile's 128 keys.  Pass A reads the 64 scores only to find the row maximum (partners exchange it through
    // smem, one 256-thread named barrier per tile); pass B re-reads them from TMEM 32 at a time and turns them
    // into bf16 probabilities.  Nothing but 32 scores is ever live in registers, so 16 softmax warps per SM
    // (2 CTAs) fit without spilling and hide each other's MUFU / TMEM latencies.  Each thread keeps a partial
    // row sum; the partners' sums are combined once at the end.
    const int qd4 = warp & 3, hf = warp >> 2;
    const int r = qd4 * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
    const int q = q0 + r;
    float m2 = -INFINITY, l = 0.f;
    const unsigned long long rowe = attn_drop_row_base(b, H, h, N, q < N ? q : 0);
    // raw scores q . k of the tail keys (both partner threads compute all of them).  They are computed twice -- here
    // and again in the epilogue -- rather than kept in registers across the tile loop (the Q tile stays in smem).
    auto tail_score = [&]() {  // q . k_tail from the Q tile and the staged K row, both in shared memory
      float st = 0.f;
      const uint32_t qrow = smem_u32(sQ) + (uint32_t)r * 128u, krow = smem_u32(&sm->tail_kv[0]);
#pragma unroll
      for (int c = 0; c < 8; ++c)  // 16-byte pieces of the 128B-swizzled Q row
        st = dot8_bf16(ld_shared_u4(qrow + (uint32_t)((c ^ (r & 7)) * 16)), ld_shared_u4(krow + 16 * c), st);
      return st;
    };
    if (tail > 0) {  // (finishes while the first S tile is still being loaded / multiplied)
      mbar_wait(&sm->q_full, 0);     // the Q tile has landed (TMA complete_tx)
      mbar_wait(&sm->tail_full, 0);  // ... and the tail key's rows
      m2 = tail_score() * scale_log2;
      if (hf == 0) l = 1.0f;  // ex2(0); the partners' row sums are added at the end: only one of them counts the tail
    }
    // keep-bit records of this CTA's query tile: [(b,h)][q tile][key tile][row][8 groups] 16-bit fields
    const int nkt_all = (N + kTile - 1) / kTile;
    uint32_t* kb_tile = keepbits + ((((long long)b * H + h) * gridDim.x + blockIdx.x) * nkt_all * kTile + r) * 4 + hf * 2;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&sm->s_full, (uint32_t)j & 1u);
      tc_fence_after();
      const int valid = Nk - j * kTile - hf * 64;  // keys beyond the tiled range are masked (last tile only)
      // ---- pass A: row maximum of this thread's 64 scores ----
      float mx;
      {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t sv[32];
          tmem_ld32(tS + lane_off + hf * 64 + c2 * 32, sv);
          tmem_ld_wait();
          if (valid < 64) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c2 * 32 + c >= valid) sv[c] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(sv[c]));
            mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
          }
        }
        mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      sm->xchg[j & 1][hf][r] = mx;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float mxo = sm->xchg[j & 1][hf ^ 1][r];
      const float m_new = fmaxf(m2, fmaxf(mx, mxo) * scale_log2);  // running max in the scaled log2 domain
      const float corr = ex2_approx(m2 - m_new);

      if (j > 0) {
        mbar_wait(&sm->pv_done, (uint32_t)(j - 1) & 1u);  // O and the P buffer are free again
        tc_fence_after();
        if (__any_sync(0xffffffffu, m_new > m2)) {  // warp-uniform: rescale this thread's 32 columns of O
          uint32_t o[32];
          tmem_ld32(tO + lane_off + hf * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * corr);
          tmem_st32(tO + lane_off + hf * 32, o);
        }
      }
      // ---- pass B: exponentials, partial row sum (of the un-dropped probabilities), dropout mask, bf16 pack ----
      // (the 1/(1-p) factor of kept elements is applied once to O in the epilogue)
      f32x2 rs01 = pk2(0.f, 0.f), rs23 = rs01;
      uint32_t kbw[4] = {0, 0, 0, 0};  // keep bits of this thread's four 16-key groups of the tile
      const f32x2 sl2 = pk2(scale_log2, scale_log2), nm2 = pk2(-m_new, -m_new);
      const unsigned long long g0 = (rowe + (unsigned long long)j * kTile + hf * 64) >> 4;  // 16-element groups
      // kPassW scores per TMEM load: 32 without dropout; 16 with it, where the Philox state, the masks and the keep-bit
      // words compete for the 96 registers that two CTAs per SM allow (TVIT_ATTN_FWD_PASSW: A-B builds)
      constexpr int kPassW = kDrop ? TVIT_ATTN_FWD_PASSW : 32;
#pragma unroll
      for (int cw = 0; cw < 64 / kPassW; ++cw) {
        uint32_t sv[kPassW];
        if constexpr (kPassW == 32) tmem_ld32(tS + lane_off + hf * 64 + cw * 32, sv);
        else tmem_ld16(tS + lane_off + hf * 64 + cw * 16, sv);
        tmem_ld_wait();
        if (cw == 64 / kPassW - 1) {  // all of S is now in registers / consumed: release the S buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm->s_free);
        }
        if (valid < 64) {
#pragma unroll
          for (int c = 0; c < kPassW; ++c)
            if (cw * kPassW + c >= valid) sv[c] = 0xff800000u;  // -inf
        }
        uint32_t pk[kPassW / 2];
#pragma unroll
        for (int g = 0; g < kPassW / 16; ++g) {  // one Philox call per 16 keys (common.cuh: attention-probability dropout)
          const int gi = cw * (kPassW / 16) + g;  // 16-key group of this thread's 64 keys
          uint32_t w[4] = {0, 0, 0, 0}, tg2 = 0, kb = 0;
          if (kDrop) {
            drop_bits16(drop, g0 + gi, w);
            tg2 = drop_tgc(drop_thr8(drop, g0 + gi));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = g * 16 + u * 4;
            // packed fp32x2 (FFMA2 / FADD2; bit-identical per lane to the scalar form): exponent arguments, row sums
            float e0, e1, e2, e3;
            up2(fma2(pk2(__uint_as_float(sv[c]), __uint_as_float(sv[c + 1])), sl2, nm2), e0, e1);
            up2(fma2(pk2(__uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3])), sl2, nm2), e2, e3);
            const f32x2 p01 = pk2(ex2_approx(e0), ex2_approx(e1)), p23 = pk2(ex2_approx(e2), ex2_approx(e3));
            rs01 = add2(rs01, p01);
            rs23 = add2(rs23, p23);
            uint32_t v01 = pack_bf16_2(p01), v23 = pack_bf16_2(p23);
            if (kDrop) {
              const uint32_t m01 = drop_keep_mask2<0>(w[u], tg2), m23 = drop_keep_mask2<1>(w[u], tg2);
              v01 &= m01;
              v23 &= m23;
              // keep-bit record of the group (include/tvit.h: tvit_attn_keepbits_bytes): one LOP3 per element pair
              kb |= (m01 & (0x00010001u << (2 * u))) | (m23 & (0x00010001u << (2 * u + 1)));
            }
            pk[g * 8 + u * 2] = v01;
            pk[g * 8 + u * 2 + 1] = v23;
          }
          if (kDrop) kbw[gi] = kb;
        }
        if constexpr (kPassW == 32) tmem_st16(tP + lane_off + hf * 32 + cw * 16, pk);
        else tmem_st8(tP + lane_off + hf * 32 + cw * 8, pk);
      }
      float rs0, rs1, rs2, rs3;
      up2(rs01, rs0, rs1);
      up2(rs23, rs2, rs3);
      l = l * corr + ((rs0 + rs1) + (rs2 + rs3));
      m2 = m_new;
      if (kDrop && keepbits)  // record (q tile, key tile j): [row r][8 groups] 16-bit fields; this thread owns groups
        st_global_v2(kb_tile + (long long)j * (kTile * 4),  // hf*4 .. +3: bytes 0 / 2 of each kbw = even / odd keys
                     prmt<0x6420u>(kbw[0], kbw[1]), prmt<0x6420u>(kbw[2], kbw[3]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->p_full);
    }
    // ---- epilogue: combine the partners' row sums; O / l -> bf16, token-major store; LSE ----
    sm->xchg[nkv & 1][hf][r] = l;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    l += sm->xchg[nkv & 1][hf ^ 1][r];
    float pt = 0.f;  // tail key's probability, computed while the last P V MMA is still in flight
    if (tail > 0) {
      pt = ex2_approx(fmaf(tail_score(), scale_log2, -m2));
      if (kDrop) {
        const bool keep_t = drop_keep(drop, rowe + (unsigned long long)Nk);
        if (!keep_t) pt = 0.f;
        // the tail key is element 0 of group 0 of key tile nkv (the backward CTA of that key tile reads it there)
        if (keepbits && hf == 0) kb_tile[(long long)nkv * (kTile * 4)] = keep_t ? 1u : 0u;
      }
    }
    mbar_wait(&sm->pv_done, (uint32_t)(nkv - 1) & 1u);
    tc_fence_after();
    const float inv = (kDrop ? drop.inv_keep : 1.0f) / l;
    __nv_bfloat16* orow = out + ((long long)b * N + q) * D + h * kHd + hf * 32;
    {
      uint32_t o[32];
      tmem_ld32(tO + lane_off + hf * 32, o);
      tmem_ld_wait();
      if (tail > 0 && q < N) {  // P V of the tail key (dropout keep factor included; 1/(1-p) rides on `inv`)
        const uint32_t vrow = smem_u32(&sm->tail_kv[8]) + (uint32_t)hf * 64u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 vv = ld_shared_u4(vrow + 16 * c);
          const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vw[i]));
            o[8 * c + 2 * i] = __float_as_uint(fmaf(pt, f.x, __uint_as_float(o[8 * c + 2 * i])));
            o[8 * c + 2 * i + 1] = __float_as_uint(fmaf(pt, f.y, __uint_as_float(o[8 * c + 2 * i + 1])));
          }
        }
      }
      if (q < N) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {  // 16 bf16 = one full 32-byte sector per store (D % 16 == 0)
          uint32_t v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            v[u] = pack_bf16(__uint_as_float(o[16 * t + 2 * u]) * inv, __uint_as_float(o[16 * t + 2 * u + 1]) * inv);
          st_global_b32x8(orow + 16 * t, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        }
      }
    }
    if (hf == 0 && q < N) lse[((long long)b * H + h) * N + q] = (m2 + log2f(l)) * 0.69314718055994531f;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

static int make_qkv_tmap(CUtensorMap* tm, const void* qkv, int B, int N, int D3, int box_rows) {
  const uint64_t dims[3] = {(uint64_t)D3, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)D3 * 2, (uint64_t)N * D3 * 2};
  const uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tmap_bf16(tm, qkv, 3, dims, strides, box);
}

// keep-bit cache of the attention-probability dropout site (include/tvit.h): 16 bits per 16-key group
size_t tc_attn_keepbits_bytes(int B, int N, int H) {
  const size_t nt = (size_t)((N + kTile - 1) / kTile);
  return (size_t)B * H * nt * nt * kTile * 8 * sizeof(uint16_t);
}

int tc_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, const tvit_dropout* drop,
                void* keepbits, cudaStream_t s) {
  if (hd != kHd) return fail(TVIT_ERR_UNSUPPORTED, "tcgen05 attention supports head_dim 64 only (got %d)", hd);
  const int D = H * hd;
  if (D % 8 != 0) return fail(TVIT_ERR_BAD_ARG, "attention: embed dim must be a multiple of 8");
  constexpr int smem_bytes = 5 * kTileBytes + 1024 + 256 + 2048 + 512;  // tiles + alignment + barriers + exchange + tail
  int rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_fwd_kernel<false>, smem_bytes)) != TVIT_OK) return rc;
  if ((rc = ensure_dynamic_smem((const void*)tc_attn_fwd_kernel<true>, smem_bytes)) != TVIT_OK) return rc;
  CUtensorMap tm;
  if ((rc = make_qkv_tmap(&tm, qkv, B, N, 3 * D, kTile)) != TVIT_OK) return rc;
  dim3 grid((N + kTile - 1) / kTile, H, B);
  const float scale_log2 = (1.0f / sqrtf((float)hd)) * 1.4426950408889634f;
  const DropCfg dc = make_drop(drop);
  const int tail = attn_tail(N, 1);
  const __nv_bfloat16* qp = (const __nv_bfloat16*)qkv;
  if (dc.thr16 != 0)
    tc_attn_fwd_kernel<true><<<grid, kAttnFwdThreads, smem_bytes, s>>>(tm, qp, (__nv_bfloat16*)out, lse, N, tail, H, scale_log2, dc,
                                                                       (uint32_t*)keepbits);
  else
    tc_attn_fwd_kernel<false><<<grid, kAttnFwdThreads, smem_bytes, s>>>(tm, qp, (__nv_bfloat16*)out, lse, N, tail, H, scale_log2, dc,
                                                                        nullptr);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

// backward: see tc_attn_bwd.cu

}  // namespace tvit
