// tcgen05 GEMM for sm_100a: C[M,N] = op(A) op(B)^T (bf16 operands, fp32 accumulate in TMEM) with the
// fused epilogues of epilogue.cuh.
//
// Persistent, warp-specialised:
//   warp 0   TMA producer    (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1   MMA issuer      (one elected thread: tcgen05.mma cta_group::1, 128 x BN x 16 per instruction)
//   warp 2   TMEM allocator  (2 accumulator stages x BN fp32 columns)
//   warps 4-11 epilogue      (8 warps: TMEM lane quarter = warp % 4, column half = (warp - 4) / 4;
//                            tcgen05.ld 32x32b -> registers -> fused epilogue -> global), overlapped with
//                            the next tile's main loop through the double-buffered TMEM accumulator.
// Operand layouts:
//   K-major (trans == 0): box {64 k, rows} -> rows of 128 B, 8-row swizzle atoms (SBO 1024).
//   MN-major (trans == 1, weight-gradient GEMMs with K = tokens): boxes {64 mn, 64 k}; 64-wide MN blocks
//   are LBO = 8192 B apart, 8-k-row groups SBO = 1024 B apart.
// Work items = m_tile x n_tile x k_split, static round-robin over the persistent CTAs.
#include <cstdlib>
#include <cstring>
#include <set>
#include <unordered_map>
#include <utility>

#include "epilogue.cuh"
#include "tc_common.cuh"

namespace tvit {

// ------------------------------------------------------------------------------------------
// tensor-map creation through the driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

// Encoded tensor maps are cached, keyed by (device, base pointer, rank, dims, strides, box): activations come from
// PyTorch's caching allocator, so from the second training step on every launch finds its descriptors here instead
// of re-encoding them ("TMA descriptor cache keyed by ptr/shape").  A tensor map only describes
// addresses and extents -- it never caches data -- so reuse after the allocator hands the same block to a different
// tensor of the same geometry is correct by construction.
struct TmapKey {
  uint64_t v[11];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t x : k.v) h = (h ^ x) * 0x100000001b3ull;
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t kTmapCacheMax = 8192;  // entries (128 B each); cleared wholesale when full

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(TVIT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (((uintptr_t)base & 15u) != 0) return fail(TVIT_ERR_BAD_ARG, "TMA base address %p not 16-byte aligned", base);
  if (rank < 1 || rank > 3) return fail(TVIT_ERR_BAD_ARG, "tensor map rank %d not in [1,3]", rank);
  int dev = 0;
  cudaGetDevice(&dev);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.v[0] = (uint64_t)(uintptr_t)base;
  key.v[1] = ((uint64_t)(uint32_t)dev << 32) | (uint32_t)rank;
  for (int i = 0; i < rank; ++i) {
    key.v[2 + i] = dims[i];
    key.v[8 + i] = box[i];
    if (i > 0) key.v[5 + i] = strides_bytes[i - 1];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return TVIT_OK;
    }
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      if (strides_bytes[i - 1] % 16 != 0)
        return fail(TVIT_ERR_BAD_ARG, "TMA stride %llu bytes not a multiple of 16 (leading dimension %% 8 != 0)",
                    (unsigned long long)strides_bytes[i - 1]);
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TVIT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return TVIT_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember (kernel, device) pairs
// so that a process driving several GPUs sets it on each of them (a process-wide once-flag left the second device
// at the 48 KB default).
int ensure_dynamic_smem(const void* kern, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(TVIT_ERR_CUDA, "cudaGetDevice failed: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({kern, dev})) return TVIT_OK;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess)
    return fail(TVIT_ERR_CUDA, "cudaFuncSetAttribute(smem=%d) failed: %s", bytes, cudaGetErrorString(e));
  done.insert({kern, dev});
  return TVIT_OK;
}
}  // namespace tvit

#include "tc_gemm_impl.cuh"

namespace tvit {

extern template int dispatch_epi<128>(const tvit_gemm_args*, const GemmShape&, const EpiParams&, cudaStream_t);
extern template int dispatch_epi<192>(const tvit_gemm_args*, const GemmShape&, const EpiParams&, cudaStream_t);
extern template int dispatch_epi<256>(const tvit_gemm_args*, const GemmShape&, const EpiParams&, cudaStream_t);

int tc_gemm(const tvit_gemm_args* a, cudaStream_t s) {
  if (a->dtype != TVIT_BF16) return fail(TVIT_ERR_BAD_ARG, "tcgen05 gemm needs bf16 operands");
  if (a->trans_a != a->trans_b)
    return fail(TVIT_ERR_UNSUPPORTED, "tcgen05 gemm: supported layouts are (trans_a,trans_b) = (0,0) or (1,1)");
  if (a->lda % 8 != 0 || a->ldb % 8 != 0)
    return fail(TVIT_ERR_BAD_ARG, "tcgen05 gemm: lda/ldb must be multiples of 8 (lda=%lld ldb=%lld)", a->lda, a->ldb);

  // pick BN in {128,192,256} minimising padded N; ties -> larger tile
  int best_bn = 128;
  long long best_pad = -1;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long pad = (long long)((a->N + bn - 1) / bn) * bn;
    if (best_pad < 0 || pad < best_pad) {
      best_pad = pad;
      best_bn = bn;
    }
  }
  if (!a->trans_a && a->K <= kMaxResKBlocks * kBK && a->N % 192 == 0) best_bn = 192;  // weight-stationary path
  GemmShape sh;
  sh.M = a->M;
  sh.N = a->N;
  sh.K = a->K;
  sh.m_tiles = (a->M + kBM - 1) / kBM;
  sh.n_tiles = (a->N + best_bn - 1) / best_bn;
  sh.k_blocks = (a->K + kBK - 1) / kBK;
  int splits = 1;
  if (a->epilogue == TVIT_EPI_ACCUM_F32) {
    splits = a->split_k > 0 ? a->split_k : num_sms() / (sh.m_tiles * sh.n_tiles);
    if (splits < 1) splits = 1;
    if (splits > sh.k_blocks) splits = sh.k_blocks;
  }
  sh.kb_per_split = (sh.k_blocks + splits - 1) / splits;
  sh.splits = (sh.k_blocks + sh.kb_per_split - 1) / sh.kb_per_split;
  // K-major: 8-row atoms 1024 B apart (SBO), 32 B per 16-element k step inside the 128 B swizzle row.
  // MN-major: 64-wide MN blocks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), 16 k rows = 2048 B.
  const bool mn = a->trans_a != 0;
  sh.a_lbo = sh.b_lbo = mn ? 8192u : 0u;
  sh.a_sbo = sh.b_sbo = 1024u;
  sh.a_kstep = sh.b_kstep = mn ? 2048u : 32u;
  if (mn) {
    if (const char* dbg = getenv("TVIT_MN_DESC")) {  // debug only: "lbo,sbo,kstep"
      unsigned l = 0, sb = 0, ks = 0;
      if (sscanf(dbg, "%u,%u,%u", &l, &sb, &ks) == 3) {
        sh.a_lbo = sh.b_lbo = l;
        sh.a_sbo = sh.b_sbo = sb;
        sh.a_kstep = sh.b_kstep = ks;
      }
    }
  }
  const EpiParams ep = make_epi_params(a);
  switch (best_bn) {
    case 256: return dispatch_epi<256>(a, sh, ep, s);
    case 192: return dispatch_epi<192>(a, sh, ep, s);
    default: return dispatch_epi<128>(a, sh, ep, s);
  }
}

}  // namespace tvit
