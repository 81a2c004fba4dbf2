// Microbenchmark: latency of a batch of 8 independent 16-byte loads from L2-resident data, per load flavour,
// and cost of the release/acquire primitives used by the frame kernel's grid barrier.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float4 ld_weak(const float* p) { float4 v; asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_na(const float* p) { float4 v; asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_cg(const float* p) { float4 v; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_vol(const float* p) { float4 v; asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_rlx(const float* p) { float4 v; asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float4 ld_cv(const float* p) { float4 v; asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w) : "l"(p) : "memory"); return v; }

template <int FL> __device__ __forceinline__ float4 ld(const float* p) {
  if (FL == 0) return ld_weak(p); if (FL == 1) return ld_na(p); if (FL == 2) return ld_cg(p);
  if (FL == 3) return ld_vol(p); if (FL == 4) return ld_rlx(p); return ld_cv(p);
}
template <int FL, int NB>
__global__ void k_lat(const float* buf, size_t nfloats, long long* out, float* sink, int iters) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long tot = 0; float acc = 0.f;
  uint32_t rng = blockIdx.x * 977 + warp * 131 + 7;
  for (int it = 0; it < iters; ++it) {
    rng = rng * 1664525u + 1013904223u;
    const size_t base = ((size_t)(rng >> 8) % (nfloats / 4096)) * 4096;   // random 16 KB-aligned window
    __syncwarp();
    const long long t0 = clock64();
    float4 v[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) v[j] = ld<FL>(buf + base + j * 128 + lane * 4);
#pragma unroll
    for (int j = 0; j < NB; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    // force completion
    if (acc == 123456.789f) sink[0] = acc;
    const long long t1 = clock64();
    tot += t1 - t0;
  }
  if (lane == 0) out[blockIdx.x * (blockDim.x / 32) + warp] = tot / iters;
  if (acc == 1e30f) sink[1] = acc;
}
__global__ void k_fill(float* buf, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = 1.0f; }
// fence / atomic costs after a few stores
__global__ void k_sync_cost(float* buf, unsigned* ctr, long long* out, int iters) {
  long long t_fence = 0, t_red = 0, t_redrel = 0, t_acq = 0, t_st8 = 0;
  for (int it = 0; it < iters; ++it) {
    buf[blockIdx.x * 1024 + threadIdx.x] = (float)it;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t0 = clock64();
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      long long t1 = clock64();
      asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(ctr) : "memory");
      long long t2 = clock64();
      buf[blockIdx.x * 1024 + 512] = (float)it;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(ctr + 32) : "memory");
      long long t3 = clock64();
      unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v == 0xffffffffu) buf[0] = 1.f;
      long long t4 = clock64();
      t_fence += t1 - t0; t_red += t2 - t1; t_redrel += t3 - t2; t_acq += t4 - t3;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[blockIdx.x * 8 + 0] = t_fence / iters; out[blockIdx.x * 8 + 1] = t_red / iters; out[blockIdx.x * 8 + 2] = t_redrel / iters; out[blockIdx.x * 8 + 3] = t_acq / iters; }
}
template <int FL, int NB> void run(const char* name, const float* buf, size_t n, long long* out, float* sink, int grid, int block) {
  k_lat<FL, NB><<<grid, block>>>(buf, n, out, sink, 200);
  cudaDeviceSynchronize();
  long long h[148 * 8]; cudaMemcpy(h, out, sizeof(long long) * grid * (block / 32), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid * (block / 32); ++i) s += h[i];
  printf("%-22s batch %2d: %7.0f cycles\n", name, NB, s / (grid * (block / 32)));
}
int main() {
  const size_t n = 16 << 20;  // 64 MB of floats: fits L2 (126 MB)
  float *buf, *sink; long long* out; unsigned* ctr;
  cudaMalloc(&buf, n * 4); cudaMalloc(&sink, 64); cudaMalloc(&out, 8 * 148 * 8 * 2); cudaMalloc(&ctr, 1024); cudaMemset(ctr, 0, 1024);
  k_fill<<<1184, 256>>>(buf, n); cudaDeviceSynchronize();
  for (int rep = 0; rep < 2; ++rep) {
    printf("--- 148 CTAs x 8 warps, random 16 KB windows in a 64 MB (L2-resident) buffer, rep %d\n", rep);
    run<0, 8>("weak", buf, n, out, sink, 148, 256);      run<0, 16>("weak", buf, n, out, sink, 148, 256);
    run<1, 8>("weak L1::no_allocate", buf, n, out, sink, 148, 256); run<1, 16>("weak L1::no_allocate", buf, n, out, sink, 148, 256);
    run<2, 8>("ld.cg", buf, n, out, sink, 148, 256);     run<2, 16>("ld.cg", buf, n, out, sink, 148, 256);
    run<3, 8>("ld.volatile", buf, n, out, sink, 148, 256); run<3, 16>("ld.volatile", buf, n, out, sink, 148, 256);
    run<4, 8>("ld.relaxed.gpu", buf, n, out, sink, 148, 256); run<4, 16>("ld.relaxed.gpu", buf, n, out, sink, 148, 256);
    run<5, 8>("ld.cv", buf, n, out, sink, 148, 256);
    run<0, 1>("weak", buf, n, out, sink, 148, 256); run<2, 1>("ld.cg", buf, n, out, sink, 148, 256); run<4, 1>("ld.relaxed.gpu", buf, n, out, sink, 148, 256);
    run<0, 2>("weak", buf, n, out, sink, 148, 32); run<4, 2>("ld.relaxed.gpu 1 warp", buf, n, out, sink, 148, 32);
  }
  k_sync_cost<<<148, 256>>>(buf, ctr, out, 200); cudaDeviceSynchronize();
  long long h[148 * 8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double s[4] = {0, 0, 0, 0}; for (int i = 0; i < 148; ++i) for (int j = 0; j < 4; ++j) s[j] += h[i * 8 + j];
  printf("after 256 stores + bar.sync: fence.acq_rel.gpu %.0f | red.relaxed %.0f | store+red.release %.0f | ld.acquire %.0f cycles\n", s[0] / 148, s[1] / 148, s[2] / 148, s[3] / 148);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
