// Shared device/host helpers for libcsm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/csm_b200.h"

namespace csmb {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* what);
void count_launch();  // diagnostic counter behind csmb_debug_launch_count()

#define CSMB_CUDA(expr)                                  \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) {                             \
      ::csmb::set_cuda_error(_e, #expr);                 \
      return CSMB_ERR_CUDA;                              \
    }                                                    \
  } while (0)

#define CSMB_LAUNCH_CHECK()                              \
  do {                                                   \
    ::csmb::count_launch();                              \
    cudaError_t _e = cudaGetLastError();                 \
    if (_e != cudaSuccess) {                             \
      ::csmb::set_cuda_error(_e, "kernel launch");       \
      return CSMB_ERR_CUDA;                              \
    }                                                    \
  } while (0)

#define CSMB_REQUIRE(cond)                               \
  do {                                                   \
    if (!(cond)) return CSMB_ERR_INVALID;                \
  } while (0)

// Makes `device` current for the duration of a call and restores the caller's device afterwards, so the
// library never changes the calling thread's CUDA state (callers hop threads between frames).
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      changed = (err == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

#define CSMB_ENTER(device)                               \
  ::csmb::DeviceGuard _guard(device);                    \
  if (_guard.err != cudaSuccess) {                       \
    ::csmb::set_cuda_error(_guard.err, "cudaSetDevice"); \
    return CSMB_ERR_CUDA;                                \
  }

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__

// bf16 pair packed in a 32-bit word (little endian: element 0 in the low half) -> two fp32, exactly.
__device__ __forceinline__ float bf16lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
__device__ __forceinline__ float bf16_to_f32(uint16_t v) { return __uint_as_float(((uint32_t)v) << 16); }

// Streaming 128-bit load of read-once weights: read-only path, do not allocate in L1.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// (value, index) argmax with lowest-index tie-break.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) {
    v = ov;
    i = oi;
  }
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    argmax_combine(v, i, ov, oi);
  }
}

// Philox4x32-10 (Salmon et al. 2011), counter c[4], key k[2].
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// uniform in (0,1) from 32 random bits: 24-bit mantissa, never 0 or 1.
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Gumbel(0,1) noise of vocabulary entry idx for draw (draw_lo, draw_hi) of sequence word `seq`: the sampling definition
// shared by every sampler of the library and by oracle/sampling.py (argmax of logit/T + gumbel).
__device__ __forceinline__ float gumbel_for(int idx, uint32_t draw_lo, uint32_t draw_hi, uint32_t seq,
                                            uint32_t k0, uint32_t k1) {
  uint32_t c[4] = {(uint32_t)(idx >> 2), draw_lo, draw_hi, seq};
  philox4x32_10(c, k0, k1);
  const float u = u01(c[idx & 3]);
  return -logf(-logf(u));
}

// block-wide argmax over score(i), 256 threads; result broadcast to every thread.
template <typename F>
__device__ int block_argmax(int V, F score, float* red_v, int* red_i) {
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 256) argmax_combine(bv, bi, score(i), i);
  warp_argmax(bv, bi);
  if ((threadIdx.x & 31) == 0) {
    red_v[threadIdx.x >> 5] = bv;
    red_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  bv = red_v[0];
  bi = red_i[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) argmax_combine(bv, bi, red_v[w], red_i[w]);
  __syncthreads();
  return bi == 0x7fffffff ? 0 : bi;
}

#endif  // __CUDACC__

}  // namespace csmb
