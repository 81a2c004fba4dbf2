// Mimi codec kernels of libcsm_b200 (sm_100a).
//
// The reference reaches the codec through moshi_mlx.models.mimi.Mimi (csm_mlx/tokenizers.py:14-21,70,150;
// generation.py:224-225,251,258).  Layout choice: activations are TIME-MAJOR, [batch][time][channels] fp32, so
//   * a causal Conv1d(k, stride s) is a GEMM whose A-row for output step t is the contiguous span of k*Cin
//     floats starting at row t*s of the left-padded input  (row stride s*Cin < K: rows overlap), with the
//     weight re-laid as [Cout][k][Cin];
//   * a causal ConvTranspose1d(k=2s, stride s) is a GEMM with K = 2*Cin over rows (t-1, t) and N = s*Cout
//     outputs (phase-major), whose result is already the time-major output [t*s + r][Cout];
// i.e. every SEANet layer and every transformer Linear is the same "strided-row GEMM" with fused prologue
// (ELU on A) and epilogue (bias, GELU, LayerScale, residual).  Small kernels cover LayerNorm, RoPE + ring
// KV cache, windowed attention, the depthwise x2 upsampler, RVQ gather and RVQ nearest-neighbour search.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace csmb {

// ------------------------------------------------------------------------------------------------
// Strided-row fp32 GEMM:  Y[b][t][n] = epi( sum_j actA(A[b][t*lda + j]) * W[n][j] ),  j < K
struct GemmArgs {
  const float* A; long long a_batch; int lda;   // element strides
  const float* W;                                // [N][K]
  float* Y; long long y_batch; int ldy;
  const float* bias;                             // [N] or null
  const float* scale;                            // [N] or null  (LayerScale)
  const float* Rsd; long long r_batch; int ldr;  // residual or null
  int T, N, K, B;
  int act_in;   // 0 none, 1 ELU
  int act_out;  // 0 none, 1 GELU(erf)
};

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

template <bool VEC>
__global__ void __launch_bounds__(256) k_gemm_f32(GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // row tiles on x: 60 s x batch at 24 kHz exceeds gridDim.y
  const int M = g.B * g.T;
  // loader mapping: thread -> (row lr, k-quad lk)
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int am = m0 + lr;
  const float* arow = nullptr;
  if (am < M) arow = g.A + (long long)(am / g.T) * g.a_batch + (long long)(am % g.T) * g.lda;
  const int bn = n0 + lr;
  const float* brow = bn < g.N ? g.W + (size_t)bn * g.K : nullptr;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // software pipeline: the global loads of tile k0+BK are in flight while tile k0 is multiplied out of shared memory
  auto load_tile = [&](int k0, float4& av, float4& bv) {
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = k0 + lk;
    if (VEC) {
      if (arow && k < g.K) av = *reinterpret_cast<const float4*>(arow + k);
      if (brow && k < g.K) bv = *reinterpret_cast<const float4*>(brow + k);
    } else {
      if (arow) {
        if (k < g.K) av.x = arow[k];
        if (k + 1 < g.K) av.y = arow[k + 1];
        if (k + 2 < g.K) av.z = arow[k + 2];
        if (k + 3 < g.K) av.w = arow[k + 3];
      }
      if (brow) {
        if (k < g.K) bv.x = brow[k];
        if (k + 1 < g.K) bv.y = brow[k + 1];
        if (k + 2 < g.K) bv.z = brow[k + 2];
        if (k + 3 < g.K) bv.w = brow[k + 3];
      }
    }
    if (g.act_in == 1) {
      av.x = elu1(av.x); av.y = elu1(av.y); av.z = elu1(av.z); av.w = elu1(av.w);
    }
  };
  float4 av, bv;
  load_tile(0, av, bv);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    __syncthreads();
    As[lk][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Bs[lk][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
    __syncthreads();
    if (k0 + BK < g.K) load_tile(k0 + BK, av, bv);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = m / g.T, t = m % g.T;
    float* yrow = g.Y + (long long)b * g.y_batch + (long long)t * g.ldy;
    const float* rrow = g.Rsd ? g.Rsd + (long long)b * g.r_batch + (long long)t * g.ldr : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.act_out == 1) v = gelu_erf(v);
      if (g.scale) v *= g.scale[n];
      if (rrow) v += rrow[n];
      yrow[n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Small-M variant of the strided-row GEMM (streaming decode: M = B*T is 1..32 while the weights are MBs): the
// weights are streamed once per block of RB rows by warps that split both N and K — 8 warps per CTA =
// (8/KSPLIT) outputs x KSPLIT K-slices, partials reduced through shared memory in a fixed order — so even
// N = 512 launches >= 296 CTAs' worth of warps.  Same prologue / epilogue as k_gemm_f32.
template <int RB, int KSPLIT>
__global__ void __launch_bounds__(256) k_gemv_f32(GemmArgs g) {
  constexpr int NPC = 8 / KSPLIT;
  __shared__ float red[8][RB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * NPC + warp / KSPLIT, ks = warp % KSPLIT;
  const int m0 = blockIdx.y * RB;
  const int M = g.B * g.T;
  const int kper = ((g.K / 4 + KSPLIT - 1) / KSPLIT) * 4;
  const int kbeg = ks * kper, kend = min(g.K, kbeg + kper);
  float acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.f;
  if (n < g.N) {
    const float* wrow = g.W + (size_t)n * g.K;
    const float* arow[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int m = min(m0 + r, M - 1);
      arow[r] = g.A + (long long)(m / g.T) * g.a_batch + (long long)(m % g.T) * g.lda;
    }
#pragma unroll 4
    for (int k = kbeg + lane * 4; k < kend; k += 128) {
      const float4 w = ldg_stream_f4(wrow + k);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float4 a = __ldg(reinterpret_cast<const float4*>(arow[r] + k));
        if (g.act_in == 1) {
          a.x = elu1(a.x); a.y = elu1(a.y); a.z = elu1(a.z); a.w = elu1(a.w);
        }
        acc[r] = fmaf(w.x, a.x, acc[r]);
        acc[r] = fmaf(w.y, a.y, acc[r]);
        acc[r] = fmaf(w.z, a.z, acc[r]);
        acc[r] = fmaf(w.w, a.w, acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const float v = warp_sum(acc[r]);
    if (lane == 0) red[warp][r] = v;
  }
  __syncthreads();
  if (ks == 0 && n < g.N && lane < RB && m0 + lane < M) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < KSPLIT; ++j) v += red[warp + j][lane];
    const int m = m0 + lane, b = m / g.T, t = m % g.T;
    if (g.bias) v += g.bias[n];
    if (g.act_out == 1) v = gelu_erf(v);
    if (g.scale) v *= g.scale[n];
    if (g.Rsd) v += g.Rsd[(long long)b * g.r_batch + (long long)t * g.ldr + n];
    g.Y[(long long)b * g.y_batch + (long long)t * g.ldy + n] = v;
  }
}

template <int RB>
static void launch_gemv(const GemmArgs& g, cudaStream_t st) {
  const int M = g.B * g.T;
  const int gy = cdiv(M, RB);
  // smallest K-split that still gives every SM a couple of CTAs
  int ksplit = 1;
  while (ksplit < 8 && (long long)cdiv(g.N, 8 / ksplit) * gy < 296 && g.K / (ksplit * 2) >= 128) ksplit *= 2;
  dim3 grid(cdiv(g.N, 8 / ksplit), gy);
  if (ksplit == 1) k_gemv_f32<RB, 1><<<grid, 256, 0, st>>>(g);
  else if (ksplit == 2) k_gemv_f32<RB, 2><<<grid, 256, 0, st>>>(g);
  else if (ksplit == 4) k_gemv_f32<RB, 4><<<grid, 256, 0, st>>>(g);
  else k_gemv_f32<RB, 8><<<grid, 256, 0, st>>>(g);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (with bias), one warp per row; the row (d <= 1024, d % 128 == 0) stays in registers.
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ x, long long x_batch,
                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                   float* __restrict__ y, int R, int T, int d, float eps) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const float* xr = x + (long long)(r / T) * x_batch + (long long)(r % T) * d;
  float4 v[8];
  const int nv = d / 128;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      v[j] = *reinterpret_cast<const float4*>(xr + j * 128 + lane * 4);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
      q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float4 g = *reinterpret_cast<const float4*>(w + j * 128 + lane * 4);
      const float4 bb = *reinterpret_cast<const float4*>(b + j * 128 + lane * 4);
      *reinterpret_cast<float4*>(y + (size_t)r * d + j * 128 + lane * 4) =
          make_float4((v[j].x - mean) * rstd * g.x + bb.x, (v[j].y - mean) * rstd * g.y + bb.y,
                      (v[j].z - mean) * rstd * g.z + bb.z, (v[j].w - mean) * rstd * g.w + bb.w);
    }
}

// ------------------------------------------------------------------------------------------------
// Mimi transformer attention.  qkv [B][T][3][H][64] (q,k,v planes).  Absolute position of step t of this call
// is pos0 + t (pos0 read from device memory so a captured graph can be replayed).  K (rotated) and V go to a
// ring cache [B][cap][2][H][64] at slot pos % cap; then each (b,t,h) attends to positions
// max(0,pos-ctx+1)..pos.  RoPE: adjacent pairs, theta_i given as freqs[32].
__global__ void __launch_bounds__(256) k_mimi_rope_cache(float* __restrict__ qkv, float* __restrict__ cache,
                                                         const float* __restrict__ freqs,
                                                         const int* __restrict__ pos0_ptr, int T, int H, int cap) {
  const int bt = blockIdx.x, b = bt / T, t = bt % T;
  const int pos = *pos0_ptr + t;
  float* row = qkv + (size_t)bt * 3 * H * 64;
  float* crow = cache + ((size_t)b * cap + (pos % cap)) * 2 * H * 64;
  for (int i = threadIdx.x; i < H * 32; i += blockDim.x) {
    const int h = i >> 5, p = i & 31;
    float s, c;
    sincosf((float)pos * freqs[p], &s, &c);
    float2 q = *reinterpret_cast<float2*>(row + h * 64 + 2 * p);
    *reinterpret_cast<float2*>(row + h * 64 + 2 * p) = make_float2(q.x * c - q.y * s, q.x * s + q.y * c);
    const float2 k = *reinterpret_cast<const float2*>(row + (H + h) * 64 + 2 * p);
    *reinterpret_cast<float2*>(crow + h * 64 + 2 * p) = make_float2(k.x * c - k.y * s, k.x * s + k.y * c);
  }
  for (int i = threadIdx.x; i < H * 64; i += blockDim.x) crow[H * 64 + i] = row[2 * H * 64 + i];
}

__global__ void __launch_bounds__(128) k_mimi_attention(const float* __restrict__ qkv,
                                                        const float* __restrict__ cache, float* __restrict__ out,
                                                        const int* __restrict__ pos0_ptr, int B, int T, int H,
                                                        int cap, int ctx) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= B * T * H) return;
  const int h = item % H, bt = item / H, b = bt / T, t = bt % T;
  const int pos = *pos0_ptr + t;
  const int first = pos - ctx + 1 > 0 ? pos - ctx + 1 : 0;
  const int S = pos - first + 1;
  float* sq = smem + warp * (64 + ctx);
  float* sc = sq + 64;
  const float* q = qkv + (size_t)bt * 3 * H * 64 + h * 64;
  sq[lane] = q[lane];
  sq[lane + 32] = q[lane + 32];
  __syncwarp();
  const float* cb = cache + (size_t)b * cap * 2 * H * 64;
  float m = -INFINITY;
  for (int j = lane; j < S; j += 32) {
    const float* kp = cb + (size_t)((first + j) % cap) * 2 * H * 64 + h * 64;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 64; c += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kp + c);
      dot = fmaf(kv.x, sq[c], dot);
      dot = fmaf(kv.y, sq[c + 1], dot);
      dot = fmaf(kv.z, sq[c + 2], dot);
      dot = fmaf(kv.w, sq[c + 3], dot);
    }
    dot *= 0.125f;
    sc[j] = dot;
    m = fmaxf(m, dot);
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int j = lane; j < S; j += 32) {
    const float e = expf(sc[j] - m);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  constexpr int PG = 32;  // keys per group: 2 x PG independent loads in flight (the values are L2-resident, not L1), so a
                          // 250-key window costs 8 L2 round trips instead of 32; the accumulation order stays j ascending
  for (int j0 = 0; j0 < S; j0 += PG) {
    float v0[PG], v1[PG];
#pragma unroll
    for (int u = 0; u < PG; ++u) {
      const int j = min(j0 + u, S - 1);
      const float* vp = cb + (size_t)((first + j) % cap) * 2 * H * 64 + H * 64 + h * 64;
      v0[u] = vp[lane];
      v1[u] = vp[lane + 32];
    }
#pragma unroll
    for (int u = 0; u < PG; ++u)
      if (j0 + u < S) {
        const float p = sc[j0 + u];
        a0 = fmaf(p, v0[u], a0);
        a1 = fmaf(p, v1[u], a1);
      }
  }
  const float inv = 1.f / sum;
  float* o = out + (size_t)bt * H * 64 + h * 64;
  o[lane] = a0 * inv;
  o[lane + 32] = a1 * inv;
}

// ------------------------------------------------------------------------------------------------
// RVQ decode gather: codes [B][K][F] int32 -> sem [B][F][D] = C0[c0], ac [B][F][D] = sum_{k>=1} Ck[ck].
// codebooks [K][bins][D] fp32 (already embedding_sum / max(usage, eps)).  Out-of-range ids are clamped
// (CSM heads emit ids up to 2050 while Mimi has 2048 bins; SURVEY.md hazard H3).
__global__ void __launch_bounds__(256) k_rvq_gather(const int32_t* __restrict__ codes,
                                                    const float* __restrict__ codebooks, float* __restrict__ sem,
                                                    float* __restrict__ ac, int K, int F, int bins, int D) {
  const int bf = blockIdx.x, b = bf / F, f = bf % F;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < K; ++k) {
      int id = codes[((size_t)b * K + k) * F + f];
      id = id < 0 ? 0 : (id >= bins ? bins - 1 : id);
      const float v = codebooks[((size_t)k * bins + id) * D + c];
      if (k == 0) sem[(size_t)bf * D + c] = v; else a += v;
    }
    ac[(size_t)bf * D + c] = a;
  }
}

// depthwise ConvTranspose1d(k=4, s=2, groups=C, no bias), causal: y[2t+r][c] = x[t][c]*w[c][r] + x[t-1][c]*w[c][r+2]
// x has one context row before t=0 (xprev: [B][C] carried between streaming calls, zeros at start).
__global__ void __launch_bounds__(256) k_upsample_dw(const float* __restrict__ x, const float* __restrict__ xprev,
                                                     const float* __restrict__ w, float* __restrict__ y, int T, int C,
                                                     size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const size_t bt = i / C;
  const int t = (int)(bt % T);
  const size_t b = bt / T;
  const float cur = x[i];
  const float prev = t > 0 ? x[i - C] : xprev[b * C + c];
  const float4 wv = *reinterpret_cast<const float4*>(w + (size_t)c * 4);
  float* yo = y + ((b * T + t) * 2) * C + c;
  yo[0] = cur * wv.x + prev * wv.z;
  yo[C] = cur * wv.y + prev * wv.w;
}

// RVQ encode step: dots [M][bins] = r . C^T (from the GEMM), c2 [bins] = |C_j|^2.
// idx = argmin_j (c2[j] - 2 dots[j]) (first index on ties); r -= C[idx]; codes[b][k][f] = idx.
__global__ void __launch_bounds__(256) k_rvq_argmin_update(const float* __restrict__ dots,
                                                           const float* __restrict__ c2,
                                                           const float* __restrict__ codebook, float* __restrict__ r,
                                                           int32_t* __restrict__ codes, int bins, int D, int K, int k,
                                                           int F) {
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ int s_idx;
  const int m = blockIdx.x;
  const float* dr = dots + (size_t)m * bins;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < bins; j += 256) argmax_combine(bv, bi, -(c2[j] - 2.f * dr[j]), j);
  warp_argmax(bv, bi);
  if ((threadIdx.x & 31) == 0) {
    red_v[threadIdx.x >> 5] = bv;
    red_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) argmax_combine(bv, bi, red_v[w], red_i[w]);
    s_idx = bi;
    const int b = m / F, f = m % F;
    codes[((size_t)b * K + k) * F + f] = bi;
  }
  __syncthreads();
  const float* crow = codebook + (size_t)s_idx * D;
  for (int c = threadIdx.x; c < D; c += 256) r[(size_t)m * D + c] -= crow[c];
}

// copy rows: dst[b][t][c] = src[b][src_t0 + t][c]  (conv context carry / padding fill), and replicate fill.
__global__ void __launch_bounds__(256) k_copy_rows3(const float* __restrict__ src, long long s_batch, int src_t0,
                                                    float* __restrict__ dst, long long d_batch, int dst_t0, int T,
                                                    int C, int replicate, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const size_t bt = i / C;
  const int t = (int)(bt % T);
  const size_t b = bt / T;
  const int st = replicate ? src_t0 : src_t0 + t;
  dst[b * d_batch + (size_t)(dst_t0 + t) * C + c] = src[b * s_batch + (size_t)st * C + c];
}

// streaming context carry: buf[b][0:pad] = buf[b][T:T+pad]  (ranges may overlap: load everything, sync, store)
__global__ void __launch_bounds__(1024) k_shift_rows(float* __restrict__ buf, long long batch, int T, int pad, int C) {
  float* p = buf + (long long)blockIdx.x * batch;
  const int n = pad * C;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = threadIdx.x + i * 1024;
    v[i] = j < n ? p[(size_t)T * C + j] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int j = threadIdx.x + i * 1024;
    if (j < n) p[j] = v[i];
  }
}

__global__ void k_add_int(int* p, int v) { *p += v; }

}  // namespace csmb

using namespace csmb;

extern "C" {

/* Strided-row fp32 GEMM (see top of file): the Conv1d / ConvTranspose1d / Linear workhorse of the codec
 * (moshi SEANet + transformer; reached from csm_mlx/tokenizers.py:70,150 and generation.py:251). */
int csmb_gemm_f32(const float* A, long long a_batch, int lda, const float* W, float* Y, long long y_batch, int ldy,
                  const float* bias, const float* scale, const float* residual, long long r_batch, int ldr, int B,
                  int T, int N, int K, int act_in, int act_out, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(B > 0 && T >= 0 && N > 0 && K > 0);
  if (T == 0) return CSMB_OK;
  GemmArgs g{A, a_batch, lda, W, Y, y_batch, ldy, bias, scale, residual, r_batch, ldr, T, N, K, B, act_in, act_out};
  const long long M = (long long)B * T;
  dim3 grid((unsigned)((M + 63) / 64), cdiv(N, 64));
  const bool vec = (K % 4 == 0) && (lda % 4 == 0) && (a_batch % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (vec && M <= 32) {
    // the residual may alias Y (in-place layer update): each output element is read then written by one lane
    if (M == 1) launch_gemv<1>(g, (cudaStream_t)stream);
    else if (M == 2) launch_gemv<2>(g, (cudaStream_t)stream);
    else if (M <= 4) launch_gemv<4>(g, (cudaStream_t)stream);
    else launch_gemv<8>(g, (cudaStream_t)stream);
  } else if (vec) {
    k_gemm_f32<true><<<grid, 256, 0, (cudaStream_t)stream>>>(g);
  } else {
    k_gemm_f32<false><<<grid, 256, 0, (cudaStream_t)stream>>>(g);
  }
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

/* nn.LayerNorm (eps, affine) of the Mimi transformer layers: x [B][T][d] with batch stride x_batch -> y dense. */
int csmb_layernorm(const float* x, long long x_batch, const float* w, const float* b, float* y, int B, int T, int d,
                   float eps, int device, void* stream) {
  CSMB_ENTER(device);
  const int R = B * T;
  if (R == 0) return CSMB_OK;
  CSMB_REQUIRE(d % 128 == 0 && d <= 1024 && x_batch % 4 == 0);
  k_layernorm<<<cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(x, x_batch, w, b, y, R, T, d, eps);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

/* RoPE (adjacent pairs) on q,k in place + ring KV-cache write + causal windowed attention of the Mimi
 * transformers (8 heads x 64, context `ctx`).  pos0 is a DEVICE int: absolute position of step 0. */
int csmb_mimi_attention(float* qkv, float* cache, const float* freqs, const int* pos0, float* out, int B, int T,
                        int H, int cap, int ctx, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(B > 0 && T > 0 && H > 0 && cap >= ctx + T - 1 && ctx > 0);
  cudaStream_t st = (cudaStream_t)stream;
  k_mimi_rope_cache<<<B * T, 256, 0, st>>>(qkv, cache, freqs, pos0, T, H, cap);
  CSMB_LAUNCH_CHECK();
  const size_t smem = (size_t)4 * (64 + ctx) * sizeof(float);
  k_mimi_attention<<<cdiv(B * T * H, 4), 128, smem, st>>>(qkv, cache, out, pos0, B, T, H, cap, ctx);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_rvq_gather(const int32_t* codes, const float* codebooks, float* sem, float* ac, int B, int K, int F,
                    int bins, int D, int device, void* stream) {
  CSMB_ENTER(device);
  if (B * F == 0) return CSMB_OK;
  k_rvq_gather<<<B * F, 256, 0, (cudaStream_t)stream>>>(codes, codebooks, sem, ac, K, F, bins, D);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_upsample_dw(const float* x, const float* xprev, const float* w, float* y, int B, int T, int C, int device,
                     void* stream) {
  CSMB_ENTER(device);
  const size_t total = (size_t)B * T * C;
  if (total == 0) return CSMB_OK;
  k_upsample_dw<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, xprev, w, y, T, C, total);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_rvq_argmin_update(const float* dots, const float* c2, const float* codebook, float* r, int32_t* codes,
                           int M, int bins, int D, int K, int k, int F, int device, void* stream) {
  CSMB_ENTER(device);
  if (M == 0) return CSMB_OK;
  k_rvq_argmin_update<<<M, 256, 0, (cudaStream_t)stream>>>(dots, c2, codebook, r, codes, bins, D, K, k, F);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_copy_rows(const float* src, long long s_batch, int src_t0, float* dst, long long d_batch, int dst_t0, int B,
                   int T, int C, int replicate, int device, void* stream) {
  CSMB_ENTER(device);
  const size_t total = (size_t)B * T * C;
  if (total == 0) return CSMB_OK;
  k_copy_rows3<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, s_batch, src_t0, dst, d_batch,
                                                                                  dst_t0, T, C, replicate, total);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_shift_rows(float* buf, long long batch, int B, int T, int pad, int C, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(pad * C <= 8192);
  if (B == 0 || pad == 0) return CSMB_OK;
  k_shift_rows<<<B, 1024, 0, (cudaStream_t)stream>>>(buf, batch, T, pad, C);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_add_int(int* p, int v, int device, void* stream) {
  CSMB_ENTER(device);
  k_add_int<<<1, 1, 0, (cudaStream_t)stream>>>(p, v);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

}  // extern "C"
