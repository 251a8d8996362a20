// tcgen05 GEMM kernel and its launchers (see tc_gemm.cu for the design notes).  This header is compiled once per BN in
// tc_gemm_bn{128,192,256}.cu -- one translation unit with all ~40 instantiations took 3.5 minutes of the 3.8-minute
// build -- and tc_gemm.cu only dispatches to the explicit instantiations of dispatch_epi<BN>.
#pragma once
#include <cstdlib>

#include "epilogue.cuh"
#include "tc_common.cuh"

namespace tvit {


// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
constexpr int kBM = 128;
constexpr int kBK = 64;
// epilogue warps: 4 TMEM lane quarters x column groups (BN = 192 -> 3 groups of 64 columns, else 2 halves)
template <int BN> struct EpiCfg {
  static constexpr int kGroups = (BN == 192) ? 3 : 2;
  static constexpr int kEpiThreads = 128 * kGroups;
  static constexpr int kThreads = 128 + kEpiThreads;
};
constexpr int kMaxResKBlocks = 6;  // weight-stationary variant: K <= 384

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                                    8192 /*bias + gamma staging, 1 KB per epilogue warp*/;
};

struct GemmShape {
  int M, N, K;          // logical GEMM shape (K = reduction)
  int m_tiles, n_tiles; // output tiles
  int k_blocks;         // ceil(K / 64)
  int splits;           // K splits (>= 1), every split non-empty
  int kb_per_split;
  // UMMA smem-descriptor geometry (bytes); runtime so a debug override can sweep it (TVIT_MN_DESC)
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;
};

template <int BN, bool A_MN, bool B_MN, int EPI, bool kDrop, bool B_RES>
__global__ void __launch_bounds__(EpiCfg<BN>::kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape sh,
               EpiParams ep) {
  // B_RES ("weight-stationary", K <= 384): this CTA keeps one n-tile of B (all k-blocks) resident in shared
  // memory and streams only A, which removes ~60 % of the L2 -> SM operand traffic that bounds the K = 384 GEMMs.
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = B_RES ? 4 : Cfg::kStages;
  constexpr int kStageBytes = B_RES ? Cfg::kABytes : Cfg::kStageBytes;
  constexpr int kBResBytes = B_RES ? kMaxResKBlocks * Cfg::kBBytes : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_all = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_bres = smem_all;            // [k_blocks][BN x 64] bf16, only with B_RES
  uint8_t* smem = smem_all + kBResBytes;  // operand ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + 2;
  uint64_t* bres_bar = bars + 2 * kStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 5);
  float* s_cols = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [epilogue warp][bias W | gamma W]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(bres_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EpiCfg<BN>::kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = sh.m_tiles * sh.n_tiles * sh.splits;
  // work-item sequence of this CTA: round-robin over all items, or (B_RES) a fixed n-tile and strided m-tiles
  int w_first = blockIdx.x, w_step = gridDim.x;
  if (B_RES) {
    const int n_fixed = blockIdx.x % sh.n_tiles, g = blockIdx.x / sh.n_tiles;
    const int G = ((int)gridDim.x - n_fixed + sh.n_tiles - 1) / sh.n_tiles;  // CTAs sharing this n-tile
    w_first = g * sh.n_tiles + n_fixed;
    w_step = G * sh.n_tiles;
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (warp-uniform loop; only the TMA instructions are predicated on one elected lane -- issuing UTMALDG /
    //  UTCHMMA from divergent `if (lane == 0)` code makes ptxas wrap each one in an ELECT + R2UR loop)
    {
      int stage = 0;
      uint32_t phase = 0;
      if (B_RES && w_first < total_work) {
        if (elect_one()) {
          const int n0 = (w_first % sh.n_tiles) * BN;
          mbar_expect_tx(bres_bar, (uint32_t)(sh.k_blocks * Cfg::kBBytes));
          for (int kb = 0; kb < sh.k_blocks; ++kb)
            tma_load_2d(s_bres + kb * Cfg::kBBytes, &tmB, bres_bar, kb * kBK, n0);
        }
        __syncwarp();
      }
      for (int w = w_first; w < total_work; w += w_step) {
        const int tile = w / sh.splits, split = w - tile * sh.splits;
        const int m0 = (tile / sh.n_tiles) * kBM, n0 = (tile % sh.n_tiles) * BN;
        const int kb0 = split * sh.kb_per_split;
        const int kb1 = min(kb0 + sh.kb_per_split, sh.k_blocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < kBM / 64; ++j)
              tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, kb * kBK);
          }
          if (B_RES) {
            // B is resident
          } else if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * kBK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, kb * kBK);
          }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint32_t a_lbo = sh.a_lbo, b_lbo = sh.b_lbo, a_sbo = sh.a_sbo, b_sbo = sh.b_sbo;
      const uint32_t a_kq = sh.a_kstep >> 4, b_kq = sh.b_kstep >> 4;  // descriptor units (16 B) per UMMA_K = 16
      const uint32_t a_hi = umma_desc_hi(a_sbo), b_hi = umma_desc_hi(b_sbo);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (B_RES && w_first < total_work) mbar_wait(bres_bar, 0);
      for (int w = w_first; w < total_work; w += w_step, ++it) {
        const int tile = w / sh.splits, split = w - tile * sh.splits;
        const int kb0 = split * sh.kb_per_split;
        const int kb1 = min(kb0 + sh.kb_per_split, sh.k_blocks);
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = B_RES ? smem_u32(s_bres + kb * Cfg::kBBytes) : sa + Cfg::kABytes;
          // descriptor low words once per k-block, stepped by (k-step >> 4); high words loop-invariant
          const uint32_t da_lo = umma_desc_lo(sa, a_lbo), db_lo = umma_desc_lo(sb, b_lbo);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_ss(tmem_d, umma_desc(da_lo + k * a_kq, a_hi), umma_desc(db_lo + k * b_kq, b_hi), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
            tc_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) tc_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;            // TMEM lane quarter == warp_id % 4
    const int half = (warp - 4) >> 2;  // column group of this warp
    constexpr int kGroups = EpiCfg<BN>::kGroups;
    constexpr int kChunks = BN / 32, kHalfChunks = kChunks / kGroups;
    // GELU_BWD with ep.colsum: per-column sums of the output (the bias gradient of the preceding Linear) are
    // accumulated in this warp's private smem slice across tiles (it is not needed for bias / gamma staging in this
    // epilogue) and flushed with one atomic per column when the n-tile changes -- with the weight-stationary
    // schedule (B_RES) that is once per CTA.
    constexpr int kWc = BN / kGroups;
    const bool do_colsum = (EPI == TVIT_EPI_GELU_BWD) && ep.colsum != nullptr && ep.vec16_ok;
    float* cs_slice = s_cols + (warp - 4) * (2 * kWc);
    int cs_n0 = -1;
    auto flush_colsum = [&]() {
      if (cs_n0 < 0) return;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kWc / 32; ++j) {
        const int e = j * 32 + lane, col = cs_n0 + half * kWc + e;
        if (col < sh.N) atomicAdd(ep.colsum + col, cs_slice[e]);
        cs_slice[e] = 0.f;
      }
      __syncwarp();
    };
    if (do_colsum) {
#pragma unroll
      for (int j = 0; j < kWc / 32; ++j) cs_slice[j * 32 + lane] = 0.f;
      __syncwarp();
    }
    int it = 0;
    for (int w = w_first; w < total_work; w += w_step, ++it) {
      const int tile = w / sh.splits;
      const int m0 = (tile / sh.n_tiles) * kBM, n0 = (tile % sh.n_tiles) * BN;
      if (do_colsum && n0 != cs_n0) {
        flush_colsum();
        cs_n0 = n0;
      }
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      // GELU_BWD: request the NEXT tile's bf16 pre-activations (this thread's row segment) into L2 now, one epilogue
      // period ahead of the loads that consume them (-0.8 ms per step).  Not done for the fp32 residual rows of
      // RESIDUAL: with the register double-buffering below it bought nothing there and ncu showed the rows being
      // fetched from DRAM twice (+30 % read traffic).
      if (EPI == TVIT_EPI_GELU_BWD && ep.vec16_ok && w + w_step < total_work) {
        const int tn = (w + w_step) / sh.splits;
        const int mn = (tn / sh.n_tiles) * kBM + q * 32 + lane, nn = (tn % sh.n_tiles) * BN + half * (BN / kGroups);
        if (mn < sh.M && nn < sh.N) l2_prefetch_line((const __nv_bfloat16*)ep.aux + (long long)mn * ep.ldaux + nn);
      }
      // Stage the per-column vectors (bias, LayerScale gamma) of this warp's column group in a warp-private slice of
      // shared memory: only __syncwarp is needed, so the epilogue warps never wait for each other.
      constexpr int kW = BN / kGroups;  // columns per warp
      float* sb = s_cols + (warp - 4) * (2 * kW);
      if (EPI != TVIT_EPI_GELU_BWD) {  // (GELU_BWD reads neither; its slice holds the column sums)
        __syncwarp();  // the previous tile's reads of this slice are done
#pragma unroll
        for (int j = 0; j < kW / 32; ++j) {
          const int e = j * 32 + lane, col = n0 + half * kW + e;
          sb[e] = (ep.bias && col < sh.N) ? ep.bias[col] : 0.f;
          sb[kW + e] = (ep.gamma && col < sh.N) ? ep.gamma[col] : 1.f;
        }
        __syncwarp();
      }
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < sh.M;
      // global epilogue operands are double-buffered in registers: chunk u+1 is requested before chunk u is
      // processed (and the first chunk before the accumulator is even complete), so every thread keeps two chunks
      // of loads in flight
      constexpr bool kExt = (EPI == TVIT_EPI_RESIDUAL || EPI == TVIT_EPI_GELU_BWD);
      constexpr int kHalfSub = BN / 16 / kGroups;
      // GELU_BWD reads 8 words per chunk: ALL chunks of the tile are requested up front (32 registers), before the
      // accumulator is even complete, so their global-load latency overlaps the main loop instead of stalling every
      // chunk (ncu: 12.5 long-scoreboard stalls per issue with the one-chunk-ahead scheme).  RESIDUAL needs 16 words
      // per chunk and keeps the two-deep register ring.
      constexpr bool kAllUpfront = (EPI == TVIT_EPI_GELU_BWD) && kHalfSub <= 4;
      constexpr int kBuf = kAllUpfront ? kHalfSub : 2;
      uint32_t ext[kBuf][16];
      if (kExt && ep.vec16_ok) {
        if (kAllUpfront) {
#pragma unroll
          for (int uu = 0; uu < kHalfSub; ++uu) {
            const int nc = n0 + (half * kHalfSub + uu) * 16;
            if (nc + 16 <= sh.N) tc_epi16_load<EPI>(ep, m, nc, row_ok, ext[uu]);
          }
        } else if (n0 + half * kHalfSub * 16 + 16 <= sh.N) {
          tc_epi16_load<EPI>(ep, m, n0 + half * kHalfSub * 16, row_ok, ext[0]);
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      constexpr bool kFast = (EPI == TVIT_EPI_STORE || is_bias_gelu(EPI) || EPI == TVIT_EPI_RESIDUAL ||
                              EPI == TVIT_EPI_GELU_BWD);
      if (kFast && ep.vec16_ok) {
        const float rsc = (EPI == TVIT_EPI_RESIDUAL && ep.row_scale && row_ok) ? ep.row_scale[m / ep.rpg] : 1.0f;
        const uint32_t sb_addr = smem_u32(sb);
#pragma unroll
        for (int uu = 0; uu < kHalfSub; ++uu) {
          const int u = half * kHalfSub + uu;
          const int nc = n0 + u * 16;
          if (nc >= sh.N) break;  // warp-uniform (N % 16 == 0 on this path is implied by vec8_ok only for N % 8;
                                  // a trailing 8-column piece falls to the generic path below)
          if (nc + 16 <= sh.N) {
            if (kExt && !kAllUpfront && uu + 1 < kHalfSub && nc + 32 <= sh.N)
              tc_epi16_load<EPI>(ep, m, nc + 16, row_ok, ext[(uu + 1) % kBuf]);
            const uint32_t so = sb_addr + (uint32_t)(uu * 64);
            float colv[16];
            tc_epi16<EPI, kDrop>(ep, so, so + (uint32_t)(kW * 4), rsc, m, nc, taddr + (uint32_t)(u * 16), row_ok,
                                 ext[uu % kBuf], colv);
            if (do_colsum) {  // warp-uniform
              const float tot = warp_colsum16(colv, lane);
              if ((lane & 1) == 0) cs_slice[uu * 16 + (lane >> 1)] += tot;
            }
          } else {
            uint32_t r[16];
            tmem_ld16(taddr + (uint32_t)(u * 16), r);
            tmem_ld_wait();
            if (row_ok) {
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[t]);
              epi_apply8<EPI, __nv_bfloat16>(ep, m, nc, v);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c = half * kHalfChunks; c < (half + 1) * kHalfChunks; ++c) {
          const int nc = n0 + c * 32;
          if (nc >= sh.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float v[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[8 * j + t]);
              epi_apply8<EPI, __nv_bfloat16>(ep, m, nc + 8 * j, v);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
    if (do_colsum) flush_colsum();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host launch
// ------------------------------------------------------------------------------------------
template <int BN, bool MN, int EPI, bool kDrop, bool B_RES = false>
static int launch_tc_d(const tvit_gemm_args* a, const GemmShape& sh, const EpiParams& ep, cudaStream_t s) {
  using Cfg = GemmCfg<BN>;
  auto kern = tc_gemm_kernel<BN, MN, MN, EPI, kDrop, B_RES>;
  constexpr int kSmem = B_RES ? (kMaxResKBlocks * Cfg::kBBytes + 4 * Cfg::kABytes + 1024 + 256 + 8192) : Cfg::kSmemBytes;
  int rc;
  if ((rc = ensure_dynamic_smem((const void*)kern, kSmem)) != TVIT_OK) return rc;
  CUtensorMap tmA, tmB;
  if (!MN) {
    // A [M,K] K contiguous; B [N,K] K contiguous
    const uint64_t da[2] = {(uint64_t)a->K, (uint64_t)a->M}, sa[1] = {(uint64_t)a->lda * 2};
    const uint32_t ba[2] = {kBK, kBM};
    if ((rc = make_tmap_bf16(&tmA, a->A, 2, da, sa, ba)) != TVIT_OK) return rc;
    const uint64_t db[2] = {(uint64_t)a->K, (uint64_t)a->N}, sb[1] = {(uint64_t)a->ldb * 2};
    const uint32_t bb[2] = {kBK, (uint32_t)BN};
    if ((rc = make_tmap_bf16(&tmB, a->B, 2, db, sb, bb)) != TVIT_OK) return rc;
  } else {
    // A stored [K, M] (M contiguous); B stored [K, N] (N contiguous)
    const uint64_t da[2] = {(uint64_t)a->M, (uint64_t)a->K}, sa[1] = {(uint64_t)a->lda * 2};
    const uint32_t ba[2] = {64, kBK};
    if ((rc = make_tmap_bf16(&tmA, a->A, 2, da, sa, ba)) != TVIT_OK) return rc;
    const uint64_t db[2] = {(uint64_t)a->N, (uint64_t)a->K}, sb[1] = {(uint64_t)a->ldb * 2};
    const uint32_t bb[2] = {64, kBK};
    if ((rc = make_tmap_bf16(&tmB, a->B, 2, db, sb, bb)) != TVIT_OK) return rc;
  }
  const int total = sh.m_tiles * sh.n_tiles * sh.splits;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, EpiCfg<BN>::kThreads, kSmem, s>>>(tmA, tmB, sh, ep);
  TVIT_LAUNCH_OK();
  return TVIT_OK;
}

template <int BN, bool MN, int EPI>
static int launch_tc(const tvit_gemm_args* a, const GemmShape& sh, const EpiParams& ep, cudaStream_t s) {
  constexpr bool kCanDrop = (EPI == TVIT_EPI_BIAS_GELU || EPI == TVIT_EPI_RESIDUAL || EPI == TVIT_EPI_PATCH_EMBED);
  // weight-stationary variant: K-major, K <= 384, BN = 192 divides N, enough m-tiles to keep every CTA busy
  constexpr bool kResOk = (BN == 192) && !MN &&
                          (EPI == TVIT_EPI_STORE || is_bias_gelu(EPI) || EPI == TVIT_EPI_RESIDUAL ||
                           EPI == TVIT_EPI_GELU_BWD);
  if (kResOk && sh.k_blocks <= kMaxResKBlocks && sh.N % 192 == 0 && sh.m_tiles >= 2 * num_sms() &&
      getenv("TVIT_NO_BRES") == nullptr) {
    if (kCanDrop && ep.drop.thr16 != 0) return launch_tc_d<BN, MN, EPI, kCanDrop, kResOk>(a, sh, ep, s);
    return launch_tc_d<BN, MN, EPI, false, kResOk>(a, sh, ep, s);
  }
  if (kCanDrop && ep.drop.thr16 != 0) return launch_tc_d<BN, MN, EPI, kCanDrop>(a, sh, ep, s);
  return launch_tc_d<BN, MN, EPI, false>(a, sh, ep, s);
}

template <int BN>
int dispatch_epi(const tvit_gemm_args* a, const GemmShape& sh, const EpiParams& ep, cudaStream_t s) {
  const bool mn = a->trans_a != 0;
  if (mn) {
    if (a->epilogue != TVIT_EPI_ACCUM_F32)
      return fail(TVIT_ERR_UNSUPPORTED, "tcgen05 gemm: transposed operands only with the ACCUM_F32 epilogue");
    return launch_tc<BN, true, TVIT_EPI_ACCUM_F32>(a, sh, ep, s);
  }
  switch (a->epilogue) {
    case TVIT_EPI_STORE: return launch_tc<BN, false, TVIT_EPI_STORE>(a, sh, ep, s);
    case TVIT_EPI_BIAS_GELU:
      if (a->aux) return launch_tc<BN, false, TVIT_EPI_BIAS_GELU>(a, sh, ep, s);
      if (ep.drop.thr16 != 0)
        return fail(TVIT_ERR_UNSUPPORTED, "tcgen05 gemm: BIAS_GELU without aux (inference) does not take dropout");
      return launch_tc<BN, false, kEpiBiasGeluNoAux>(a, sh, ep, s);
    case TVIT_EPI_RESIDUAL: return launch_tc<BN, false, TVIT_EPI_RESIDUAL>(a, sh, ep, s);
    case TVIT_EPI_GELU_BWD: return launch_tc<BN, false, TVIT_EPI_GELU_BWD>(a, sh, ep, s);
    case TVIT_EPI_PATCH_EMBED: return launch_tc<BN, false, TVIT_EPI_PATCH_EMBED>(a, sh, ep, s);
    case TVIT_EPI_ACCUM_F32: return launch_tc<BN, false, TVIT_EPI_ACCUM_F32>(a, sh, ep, s);
    case TVIT_EPI_SOFTMAX_PROBS: return launch_tc<BN, false, TVIT_EPI_SOFTMAX_PROBS>(a, sh, ep, s);
    default: return fail(TVIT_ERR_BAD_ARG, "gemm: unknown epilogue %d", a->epilogue);
  }
}

}  // namespace tvit
