// Microbenchmark: HBM -> shared-memory streaming rate of the frame kernel's weight ring (cp.async.bulk + mbarrier),
// with NO arithmetic: G CTAs (one per SM) each stream a contiguous slice of a large buffer through a ring of NST stages
// of SB bytes; one producer thread issues the copies, NCW consumer warps only wait and release the stages.
// Answers: how much of the measured HBM peak can this structure pull, and with which ring geometry?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/bulkstream scripts/micro/bulkstream.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol, int hint) {
  if (hint)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int NCW = 8;
constexpr int MAXST = 32;

// touch: consumers read every TOUCHth 16-byte word of a stage (0 = none) to emulate LDS traffic
__global__ void __launch_bounds__((NCW + 1) * 32, 1) k_stream(const unsigned char* buf, size_t bytes_per_cta, int nst, int sb, int hint, int touch, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[MAXST], empty[MAXST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nst; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned char* src = buf + (size_t)blockIdx.x * bytes_per_cta;
  const int n = (int)(bytes_per_cta / sb);
  if (warp == NCW) {
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      for (int q = 0; q < n; ++q) {
        const int s = q % nst;
        mbar_wait(&empty[s], ((q / nst) & 1) ^ 1);
        mbar_expect(&full[s], sb);
        bulk_g2s(smem + (size_t)s * sb, src + (size_t)q * sb, sb, &full[s], pol, hint);
      }
    }
  } else {
    float acc = 0.f;
    for (int q = 0; q < n; ++q) {
      const int s = q % nst;
      mbar_wait(&full[s], (q / nst) & 1);
      if (touch) {
        const float4* p = reinterpret_cast<const float4*>(smem + (size_t)s * sb);
        for (int i = warp * 32 + lane; i < sb / 16; i += NCW * 32 * touch) { const float4 v = p[i]; acc += v.x + v.w; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 1.2345e30f) sink[0] = acc;
  }
}

int main(int argc, char** argv) {
  const size_t total = (size_t)6 << 30;  // 6 GiB >> 126 MB L2
  unsigned char* buf;
  float* sink;
  if (cudaMalloc(&buf, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&sink, 16);
  cudaMemset(buf, 1, total);
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int G, nst, sb, hint, touch; };
  const Cfg cfgs[] = {
      {128, 8, 16384, 1, 0}, {148, 8, 16384, 1, 0}, {128, 8, 16384, 0, 0}, {128, 12, 16384, 1, 0}, {148, 12, 16384, 1, 0},
      {128, 4, 32768, 1, 0}, {128, 6, 32768, 1, 0}, {148, 6, 32768, 1, 0}, {128, 16, 8192, 1, 0}, {128, 24, 8192, 1, 0},
      {128, 3, 65536, 1, 0}, {128, 4, 16384, 1, 0}, {128, 2, 16384, 1, 0}, {128, 8, 16384, 1, 1}, {148, 12, 16384, 1, 1},
      {64, 8, 16384, 1, 0}, {32, 8, 16384, 1, 0}, {64, 12, 16384, 1, 0},
  };
  printf("%5s %4s %6s %5s %5s | %9s %9s\n", "CTAs", "nst", "stage", "hint", "touch", "GB/s", "GB/s/SM");
  for (const Cfg& c : cfgs) {
    size_t per = (total / c.G) / c.sb * c.sb;
    per = per > ((size_t)40 << 20) ? ((size_t)40 << 20) / c.sb * c.sb : per;  // 40 MiB per CTA per launch
    const size_t smem = (size_t)c.nst * c.sb;
    float best = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_stream<<<c.G, (NCW + 1) * 32, smem>>>(buf, per, c.nst, c.sb, c.hint, c.touch, sink);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const float gbs = (float)((double)per * c.G / (ms * 1e-3) / 1e9);
      best = gbs > best ? gbs : best;
    }
    printf("%5d %4d %6d %5d %5d | %9.1f %9.2f\n", c.G, c.nst, c.sb, c.hint, c.touch, best, best / c.G);
  }
  return 0;
}
