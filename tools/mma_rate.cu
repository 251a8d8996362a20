// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16, M = 128, K = 16) for the operand flavours the
// attention kernels use, one CTA per SM, `reps` x 8 back-to-back MMAs issued by one thread, timed with clock64 between
// the first issue and the commit's mbarrier completion.  Development aid (run under gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/mma_rate tools/mma_rate.cu
#include <cstdio>

#include "../neural_vit_b200/csrc/tc_common.cuh"

using namespace tvit;

struct Flavour {
  const char* name;
  int ts, a_mn, b_mn, n;
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int ts, int a_mn, int b_mn, int n, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;            // 32 KB: [128 x 128] bf16 worth of operand data
  uint8_t* sB = smem + 32768;    // 64 KB: up to [256 x 128]
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (warp == 0) {
    const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
                            ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
    constexpr uint32_t kHi = umma_desc_hi(1024);
    const uint32_t kMn = (16384u >> 4) << 16;
    const uint32_t dA = umma_desc_lo(smem_u32(sA), 0) | (a_mn ? kMn : 0u);
    const uint32_t dB = umma_desc_lo(smem_u32(sB), 0) | (b_mn ? kMn : 0u);
    const uint32_t stepA = a_mn ? 128u : 2u, stepB = b_mn ? 128u : 2u;
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ka = a_mn ? k : (k & 3), kb = b_mn ? k : (k & 3);
            if (ts)
              umma_ts(tmem, tmem + 256 + 8 * k, umma_desc(dB + stepB * kb, kHi), idesc, 1u);
            else
              umma_ss(tmem, umma_desc(dA + stepA * ka, kHi), umma_desc(dB + stepB * kb, kHi), idesc, 1u);
          }
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)pass & 1u);
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  const Flavour fl[] = {
      {"SS  A K-major,  B K-major,  N=64  (S / dP half tiles)", 0, 0, 0, 64},
      {"SS  A K-major,  B K-major,  N=128 (S full tile, forward)", 0, 0, 0, 128},
      {"SS  A K-major,  B K-major,  N=256", 0, 0, 0, 256},
      {"SS  A MN-major, B MN-major, N=64  (dV / dK from smem P / dS; transposed dQ)", 0, 1, 1, 64},
      {"SS  A K-major,  B MN-major, N=64  (dQ from K-major dS)", 0, 0, 1, 64},
      {"TS  A TMEM,     B MN-major, N=64  (P V forward; transposed dV / dK)", 1, 0, 1, 64},
      {"TS  A TMEM,     B K-major,  N=64", 1, 0, 0, 64},
      {"TS  A TMEM,     B K-major,  N=128", 1, 0, 0, 128},
  };
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem_bytes = 97 * 1024 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  long long* out;
  cudaMallocManaged(&out, sms * sizeof(long long));
  const int reps = 256;
  for (const Flavour& f : fl) {
    for (int grid : {1, sms}) {
      mma_rate_kernel<<<grid, 128, smem_bytes>>>(f.ts, f.a_mn, f.b_mn, f.n, reps, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: %s\n", f.name, cudaGetErrorString(e));
        return 1;
      }
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
      printf("%-80s grid %3d: %7.1f clk / MMA (ideal %d)\n", f.name, grid, (double)mx / (reps * 8), f.n / 2);
    }
  }
  return 0;
}
