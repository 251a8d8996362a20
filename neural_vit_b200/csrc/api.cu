// C-ABI plumbing: error state, device check, engine dispatch.  See include/tvit.h.
#include <cstdarg>

#include "common.cuh"

namespace tvit {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// engines (simt.cu, tc_gemm.cu, tc_attn.cu)
int simt_gemm(const tvit_gemm_args* a, cudaStream_t s);
int tc_gemm(const tvit_gemm_args* a, cudaStream_t s);
int simt_attn_fwd(int dtype, const void* qkv, void* out, float* lse, int B, int N, int H, int hd,
                  const tvit_dropout* drop, cudaStream_t s);
size_t simt_attn_bwd_workspace(int B, int N, int H, int hd);
int simt_attn_bwd(int dtype, const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                  void* ws, size_t ws_bytes, int B, int N, int H, int hd, const tvit_dropout* drop, cudaStream_t s);
int attn_probs(int dtype, const void* qkv, float* probs, int B, int N, int H, int hd, cudaStream_t s);
int tc_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int hd, const tvit_dropout* drop,
                void* keepbits, cudaStream_t s);
size_t tc_attn_keepbits_bytes(int B, int N, int H);
size_t tc_attn_bwd_workspace(int B, int N, int H, int hd);
int tc_attn_bwd_variant(int mask);
int tc_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* ws,
                size_t ws_bytes, int B, int N, int H, int hd, const tvit_dropout* drop, float* dqkv_colsum,
                const void* keepbits, cudaStream_t s);

}  // namespace tvit

using namespace tvit;

extern "C" const char* tvit_last_error(void) { return g_last_error.c_str(); }

extern "C" int tvit_version(void) { return 100; }

extern "C" int tvit_device_check(int device) {
  int major = 0, minor = 0;
  TVIT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  TVIT_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10)
    return fail(TVIT_ERR_UNSUPPORTED,
                "device %d has compute capability %d.%d; libtvit_b200 is built for sm_100a (B200) only", device, major,
                minor);
  return TVIT_OK;
}

extern "C" int tvit_gemm(const tvit_gemm_args* a, tvit_stream_t stream) {
  TVIT_CHECK_ARG(a != nullptr, "gemm: null args");
  TVIT_CHECK_ARG(a->A && a->B && a->out, "gemm: null operand");
  TVIT_CHECK_ARG(a->M >= 0 && a->N > 0 && a->K > 0, "gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  if (a->M == 0) return TVIT_OK;
  switch (a->epilogue) {
    case TVIT_EPI_GELU_BWD:  // (BIAS_GELU: aux is optional -- an inference forward has no use for it)
      TVIT_CHECK_ARG(a->aux != nullptr, "gemm: epilogue %d needs aux", a->epilogue);
      break;
    case TVIT_EPI_RESIDUAL:
      TVIT_CHECK_ARG(a->resid != nullptr, "gemm: RESIDUAL needs resid");
      TVIT_CHECK_ARG(!a->row_scale || a->rows_per_group > 0, "gemm: rows_per_group must be > 0");
      break;
    case TVIT_EPI_PATCH_EMBED:
      TVIT_CHECK_ARG(a->pos_k && a->pos_f && a->pos_t && a->Kp > 0 && a->Fp > 0 && a->Tp > 0,
                     "gemm: PATCH_EMBED needs positional tables");
      TVIT_CHECK_ARG(a->M % (a->Kp * a->Fp * a->Tp) == 0, "gemm: PATCH_EMBED M must be B * n_patches");
      break;
    case TVIT_EPI_SOFTMAX_PROBS:
      TVIT_CHECK_ARG(a->row_scale != nullptr, "gemm: SOFTMAX_PROBS needs the row log-sum-exp in row_scale");
      break;
    default:
      break;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (a->engine == TVIT_ENGINE_SIMT) return simt_gemm(a, s);
  if (a->engine == TVIT_ENGINE_TCGEN05) return tc_gemm(a, s);
  return fail(TVIT_ERR_BAD_ARG, "gemm: unknown engine %d", a->engine);
}

extern "C" int tvit_attn_fwd(int engine, int dtype, const void* qkv, void* out, float* lse, int B, int N, int H,
                             int hd, const tvit_dropout* drop, void* keepbits, tvit_stream_t stream) {
  TVIT_CHECK_ARG(qkv && out && lse, "attn_fwd: null pointer");
  TVIT_CHECK_ARG(B > 0 && N > 0 && H > 0 && hd > 0, "attn_fwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  if (engine == TVIT_ENGINE_SIMT) return simt_attn_fwd(dtype, qkv, out, lse, B, N, H, hd, drop, s);
  if (engine == TVIT_ENGINE_TCGEN05) {
    TVIT_CHECK_ARG(dtype == TVIT_BF16, "attn_fwd: tcgen05 engine needs bf16");
    return tc_attn_fwd(qkv, out, lse, B, N, H, hd, drop, keepbits, s);
  }
  return fail(TVIT_ERR_BAD_ARG, "attn_fwd: unknown engine %d", engine);
}

extern "C" size_t tvit_attn_keepbits_bytes(int engine, int B, int N, int H) {
  return engine == TVIT_ENGINE_TCGEN05 ? tc_attn_keepbits_bytes(B, N, H) : 0;
}

extern "C" size_t tvit_attn_bwd_workspace_bytes(int engine, int dtype, int B, int N, int H, int hd) {
  (void)dtype;
  if (engine == TVIT_ENGINE_TCGEN05) return tc_attn_bwd_workspace(B, N, H, hd);
  return simt_attn_bwd_workspace(B, N, H, hd);
}

extern "C" int tvit_attn_bwd(int engine, int dtype, const void* qkv, const void* out, const void* dout,
                             const float* lse, void* dqkv, void* workspace, size_t workspace_bytes, int B, int N, int H,
                             int hd, const tvit_dropout* drop, float* dqkv_colsum, const void* keepbits,
                             tvit_stream_t stream) {
  TVIT_CHECK_ARG(qkv && out && dout && lse && dqkv, "attn_bwd: null pointer");
  TVIT_CHECK_ARG(B > 0 && N > 0 && H > 0 && hd > 0, "attn_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  if (engine == TVIT_ENGINE_SIMT) {
    int rc = simt_attn_bwd(dtype, qkv, out, dout, lse, dqkv, workspace, workspace_bytes, B, N, H, hd, drop, s);
    if (rc != TVIT_OK || !dqkv_colsum) return rc;
    return tvit_colsum(dqkv, dtype, (long long)B * N, 3 * H * hd, 3LL * H * hd, dqkv_colsum, stream);
  }
  if (engine == TVIT_ENGINE_TCGEN05) {
    TVIT_CHECK_ARG(dtype == TVIT_BF16, "attn_bwd: tcgen05 engine needs bf16");
    return tc_attn_bwd(qkv, out, dout, lse, dqkv, workspace, workspace_bytes, B, N, H, hd, drop, dqkv_colsum, keepbits,
                       s);
  }
  return fail(TVIT_ERR_BAD_ARG, "attn_bwd: unknown engine %d", engine);
}

extern "C" int tvit_attn_bwd_variant(int mask) { return tc_attn_bwd_variant(mask); }

extern "C" int tvit_attn_probs(int dtype, const void* qkv, float* probs, int B, int N, int H, int hd,
                               tvit_stream_t stream) {
  TVIT_CHECK_ARG(qkv && probs, "attn_probs: null pointer");
  return attn_probs(dtype, qkv, probs, B, N, H, hd, (cudaStream_t)stream);
}
