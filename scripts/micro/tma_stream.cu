// Micro-benchmark: HBM read bandwidth of the chain Linear's weight stream, no arithmetic.  Every CTA streams its 128-row
// weight tile K block by K block (TMA box 128 rows x 64 bf16 = 16 KiB, 128B swizzle, 6-stage ring), (a) from the checkpoint's
// row-major [N][K] layout — each box is 128 separate 128-byte pieces, 2*K bytes apart —, (b) from a tile-major copy in which
// every box is one contiguous 16 KiB.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csm_mlx_b200/csrc
// scripts/micro/tma_stream.cu csm_mlx_b200/csrc/gemm_tc.cu csm_mlx_b200/csrc/ops.cu -lcuda
#include <cstdio>
#include <cstdlib>

#include "tc.cuh"

using namespace csmb;
constexpr int NSTG = 6;

__global__ void __launch_bounds__(64, 1) k_stream(const __grid_constant__ CUtensorMap map, int nkb, int tiled, int reps) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[NSTG];
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTG; ++i) tc_mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  int err = 0;
  const int total = nkb * reps;
  for (int it = 0; it < total + NSTG; ++it) {
    if (it >= NSTG) tc_mbar_wait(&full[(it - NSTG) % NSTG], ((it - NSTG) / NSTG) & 1, &err);   // the stage is "consumed" at once
    if (it < total) {
      const int kb = it % nkb, s = it % NSTG;
      tc_mbar_expect_tx(&full[s], 16384);
      if (tiled) tma_load_2d(smem + s * 16384, &map, 0, (blockIdx.x * nkb + kb) * 128, &full[s]);
      else tma_load_2d(smem + s * 16384, &map, kb * 64, blockIdx.x * 128, &full[s]);
    }
  }
}

int main() {
  struct Shape { int N, K; const char* name; } shapes[] = {{16384, 1024, "decoder gate|up"}, {1024, 8192, "decoder down (8 tiles x K split 18 -> 144 CTAs)"},
                                                           {16384, 2048, "backbone gate|up"}};
  for (auto& sh : shapes) {
    const size_t bytes = (size_t)sh.N * sh.K * 2;
    void* w;
    cudaMalloc(&w, bytes * 4);   // 4 distinct copies so that successive repetitions do not hit L2
    cudaMemset(w, 0, bytes * 4);
    for (int tiled = 0; tiled < 2; ++tiled) {
      int ctas = sh.N / 128, nkb = sh.K / 64, split = 1;
      if (ctas < 100) { split = 144 / ctas; ctas *= split; nkb /= split; }   // K split like the chain
      CUtensorMap map;
      bool ok;
      // emulate the split by treating each (tile, split) as its own "tile" of nkb K blocks
      if (tiled) ok = tc_make_map_ld(&map, w, (long long)4 * sh.N * (sh.K / 64), 64, 64, 128);
      else ok = tc_make_map_ld(&map, w, (long long)4 * sh.N, sh.K, sh.K, 128);
      if (!ok) { printf("map failed\n"); return 1; }
      if (!tiled && split > 1) { printf("%-50s row-major with K split: skipped in this micro-benchmark\n", sh.name); continue; }
      cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, NSTG * 16384 + 1024);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int reps = tiled && split > 1 ? 1 : 1;
      k_stream<<<ctas, 64, NSTG * 16384 + 1024>>>(map, nkb, tiled, reps);
      cudaDeviceSynchronize();
      float best = 1e9f;
      for (int r = 0; r < 5; ++r) {
        cudaMemset(w, r, bytes * 4);   // flush L2 (126 MB) with a larger write
        cudaEventRecord(e0);
        k_stream<<<ctas, 64, NSTG * 16384 + 1024>>>(map, nkb, tiled, reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      const double moved = (double)ctas * nkb * 16384;
      printf("%-50s %-10s %4d CTAs x %3d K blocks: %7.1f us  %6.0f GB/s  (%s)\n", sh.name, tiled ? "tile-major" : "row-major", ctas, nkb, best * 1e3,
             moved / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(w);
  }
  return 0;
}
