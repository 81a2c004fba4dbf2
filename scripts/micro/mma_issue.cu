// Micro-benchmark: how fast can ONE thread issue small tcgen05.mma instructions (M = 128, K = 16, N = 16 .. 256, bf16, operands in
// shared memory), into one accumulator (a dependent chain, what a GEMM's K loop is) or round-robin into several, and from one or
// two issuing warps?  The decode-batch Linears (N = tokens <= 64) and the codec's narrow convolutions (N = 16 .. 64 channels) are
// made of exactly such instructions.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csm_mlx_b200/csrc -o mma_issue
// scripts/micro/mma_issue.cu csm_mlx_b200/csrc/gemm_tc.cu (tc.cuh helpers); run on a B200.
#include <cstdio>
#include <cstdlib>

#include "tc.cuh"

using namespace csmb;

// A operand from tensor memory (lane = row of A, a 32-bit column = two consecutive K elements, 8 columns per K = 16 step)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ts = 1: every instruction takes A from TMEM columns [256, 288) (whatever bits are there: timing only); accumulators stay
// below column 256
__global__ void __launch_bounds__(192, 1) k_mma_issue(int N, int count, int nacc, int issuers, int ts, unsigned long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t done[4];
  __shared__ uint32_t tmem_base_s;
  // warp index through a shuffle so that the compiler keeps the tcgen05.mma operands in uniform registers (as the chain does)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tc_mbar_init(&done[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  int err = 0;
  if (lane == 0 && warp >= 1 && warp <= issuers) {
    const uint32_t idesc = umma_idesc(N);
    const uint64_t da = umma_desc(s32(smem)), db = umma_desc(s32(smem + 16384));
    const int w = warp - 1;
    const uint32_t colstride = N < 32 ? 32 : N;
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t acc = (uint32_t)(w * nacc + (i % nacc));
      const uint64_t koff = (uint64_t)(((i & 3) * 32) >> 4);
      if (ts) umma_f16_ts(tmem_base + acc * colstride, tmem_base + 256u + (uint32_t)((i & 3) * 8), db + koff, idesc, 1u);
      else umma_f16(tmem_base + acc * colstride, da + koff, db + koff, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(&done[w]);
    tc_mbar_wait(&done[w], 0, &err);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[w * 2] = (unsigned long long)(t1 - t0);
      out[w * 2 + 1] = (unsigned long long)(t2 - t0);
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k_mma_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int count = 4096;
  printf("M=128 K=16 bf16, %d MMAs per issuer, 148 CTAs; cycles per MMA: issue loop only / until all complete\n", count);
  for (int N : {16, 32, 64, 128, 256}) {
    for (int issuers : {1, 2, 4}) {
      for (int nacc : {1, 2}) {
        if ((N < 32 ? 32 : N) * nacc * issuers > 512) continue;
        cudaMemset(d, 0, 64);
        k_mma_issue<<<148, 192, 64 * 1024>>>(N, count, nacc, issuers, 0, d);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned long long h[4];
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("N=%3d issuers=%d accumulators/issuer=%d : issue %.1f  complete %.1f  (per issuer; tensor work = %.0f cycles)  %s\n", N, issuers,
               nacc, (double)h[0] / count, (double)h[1] / count, 128.0 * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  printf("the same with the A operand in tensor memory (tcgen05.mma [d], [a_tmem], b_desc):\n");
  for (int N : {16, 32, 64, 128}) {
    for (int issuers : {1, 2}) {
      if ((N < 32 ? 32 : N) * issuers > 256) continue;
      cudaMemset(d, 0, 64);
      k_mma_issue<<<148, 192, 64 * 1024>>>(N, count, 1, issuers, 1, d);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h[4];
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("A in TMEM N=%3d issuers=%d : issue %.1f  complete %.1f  (per issuer)  %s\n", N, issuers, (double)h[0] / count,
             (double)h[1] / count, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
