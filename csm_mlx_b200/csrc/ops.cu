// Elementary LM kernels of libcsm_b200 (sm_100a): embedding gather-sum, RMSNorm, weight-streaming
// linear (bf16 weights, fp32 activations and accumulation), SwiGLU, RoPE + paged KV append, paged GQA
// attention, sampling.  These are the row-based building blocks used for prefill, for batched decode and
// as the per-op parity surface; the B=1 latency path runs the fused persistent kernel in frame_kernel.cu.
//
// Reference behaviour restated per kernel (paths relative to /root/reference):
//   k_embed_sum      csm_mlx/models.py:82-92 + generation.py:32-36
//   k_rmsnorm        mlx nn.RMSNorm inside mlx_lm TransformerBlock (models.py:50-51)
//   k_linear         nn.Linear(bias=False): attention.py:216-218,253, mlx_lm MLP, generation.py:42,75,79
//   k_rope_append    attention.py:119-177 (adjacent-pair RoPE, fp32) + :236-237 (KVCache append)
//   k_attention      attention.py:242-249 (mx.repeat GQA + scaled_dot_product_attention, causal)
//   k_sample*        generation.py:51-54,81-84 + mlx_lm.sample_utils (cli/generate.py:168-174)
#include <float.h>
#include <math.h>

#include <atomic>
#include <mutex>
#include <string>

#include "ops.cuh"

namespace csmb {

// ------------------------------------------------------------------------------------------------
static std::mutex g_err_mu;
static char g_err[512] = "";
void set_cuda_error(cudaError_t e, const char* what) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// embed_sum: one block per row; thread t handles 8 consecutive channels per iteration.
__global__ void __launch_bounds__(256) k_embed_sum(const int32_t* __restrict__ tokens,
                                                   const uint8_t* __restrict__ mask,
                                                   const uint16_t* __restrict__ text_emb,
                                                   const uint16_t* __restrict__ audio_emb,
                                                   float* __restrict__ out, int d, int ncb, int audio_vocab) {
  const int r = blockIdx.x;
  const int32_t* tk = tokens + (size_t)r * (ncb + 1);
  const uint8_t* mk = mask + (size_t)r * (ncb + 1);
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int s = 0; s <= ncb; ++s) {
      if (!mk[s]) continue;
      const uint16_t* row = (s < ncb) ? audio_emb + ((size_t)tk[s] + (size_t)s * audio_vocab) * d
                                      : text_emb + (size_t)tk[s] * d;
      uint4 w = *reinterpret_cast<const uint4*>(row + c);
      acc[0] += bf16lo(w.x); acc[1] += bf16hi(w.x);
      acc[2] += bf16lo(w.y); acc[3] += bf16hi(w.y);
      acc[4] += bf16lo(w.z); acc[5] += bf16hi(w.z);
      acc[6] += bf16lo(w.w); acc[7] += bf16hi(w.w);
    }
    float4* o = reinterpret_cast<float4*>(out + (size_t)r * d + c);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

int launch_embed_sum(const int32_t* tokens, const uint8_t* mask, const uint16_t* text_emb,
                     const uint16_t* audio_emb, float* out, int R, int d, int ncb, int audio_vocab,
                     cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && d > 0 && d % 8 == 0 && ncb > 0);
  if (R == 0) return CSMB_OK;
  k_embed_sum<<<R, 256, 0, st>>>(tokens, mask, text_emb, audio_emb, out, d, ncb, audio_vocab);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

__global__ void __launch_bounds__(256) k_embed_audio(const int32_t* __restrict__ tokens, int tok_stride,
                                                     const uint16_t* __restrict__ audio_emb,
                                                     float* __restrict__ out, int ldo, int d, int codebook,
                                                     int audio_vocab) {
  const int r = blockIdx.x;
  int t = tokens[(size_t)r * tok_stride];
  t = t < 0 ? 0 : (t >= audio_vocab ? audio_vocab - 1 : t);
  const uint16_t* row = audio_emb + ((size_t)t + (size_t)codebook * audio_vocab) * d;
  for (int c = threadIdx.x * 8; c < d; c += blockDim.x * 8) {
    uint4 w = *reinterpret_cast<const uint4*>(row + c);
    float4* o = reinterpret_cast<float4*>(out + (size_t)r * ldo + c);
    o[0] = make_float4(bf16lo(w.x), bf16hi(w.x), bf16lo(w.y), bf16hi(w.y));
    o[1] = make_float4(bf16lo(w.z), bf16hi(w.z), bf16lo(w.w), bf16hi(w.w));
  }
}

int launch_embed_audio(const int32_t* tokens, int tok_stride, const uint16_t* audio_emb, float* out, int ldo,
                       int R, int d, int codebook, int audio_vocab, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && d % 8 == 0 && ldo % 4 == 0);
  if (R == 0) return CSMB_OK;
  k_embed_audio<<<R, 256, 0, st>>>(tokens, tok_stride, audio_emb, out, ldo, d, codebook, audio_vocab);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// ------------------------------------------------------------------------------------------------
// rmsnorm: one block of 256 threads per row.
__global__ void __launch_bounds__(256) k_rmsnorm(const float* __restrict__ x, int ldx,
                                                 const float* __restrict__ w, float* __restrict__ y, int ldy,
                                                 int d, float eps, const int32_t* __restrict__ row_idx) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float* xr = x + (size_t)(row_idx ? row_idx[r] : r) * ldx;
  float ss = 0.f;
  for (int c = threadIdx.x * 4; c < d; c += 1024) {
    float4 v = *reinterpret_cast<const float4*>(xr + c);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float scale = rsqrtf(tot / (float)d + eps);
  for (int c = threadIdx.x * 4; c < d; c += 1024) {
    float4 v = *reinterpret_cast<const float4*>(xr + c);
    float4 g = *reinterpret_cast<const float4*>(w + c);
    *reinterpret_cast<float4*>(y + (size_t)r * ldy + c) =
        make_float4(v.x * scale * g.x, v.y * scale * g.y, v.z * scale * g.z, v.w * scale * g.w);
  }
}

int launch_rmsnorm(const float* x, int ldx, const float* w, float* y, int ldy, int R, int d, float eps,
                   const int32_t* row_idx, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && d % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0);
  if (R == 0) return CSMB_OK;
  k_rmsnorm<<<R, 256, 0, st>>>(x, ldx, w, y, ldy, d, eps, row_idx);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// ------------------------------------------------------------------------------------------------
// linear: y[R][N] (+)= x[R][K] . W[N][K]^T, bf16 W streamed once per block of RB rows with 128-bit
// no-allocate loads, fp32 x through the read-only path (L1-resident, every warp re-reads it), fp32
// accumulation.  One warp owns NB consecutive output features for RB rows; lanes split K in 8-element
// (16-byte) pieces, 256 elements per warp iteration; a shuffle tree finishes each dot product.
template <int RB, int NB>
__global__ void __launch_bounds__(256) k_linear(const float* __restrict__ x, int ldx,
                                                const uint16_t* __restrict__ W, float* __restrict__ y,
                                                int ldy, int R, int N, int K, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 8 + warp) * NB;
  const int r0 = blockIdx.y * RB;
  if (n0 >= N) return;
  float acc[RB][NB];
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j] = 0.f;
  const uint16_t* wrow[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) wrow[j] = W + (size_t)min(n0 + j, N - 1) * K;
  const float* xrow[RB];
#pragma unroll
  for (int i = 0; i < RB; ++i) xrow[i] = x + (size_t)min(r0 + i, R - 1) * ldx;

#pragma unroll 2
  for (int k = lane * 8; k < K; k += 256) {
    uint4 w[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) w[j] = ldg_stream(wrow[j] + k);
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const float4 xa = __ldg(reinterpret_cast<const float4*>(xrow[i] + k));
      const float4 xb = __ldg(reinterpret_cast<const float4*>(xrow[i] + k + 4));
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float a = acc[i][j];
        a = fmaf(bf16lo(w[j].x), xa.x, a);
        a = fmaf(bf16hi(w[j].x), xa.y, a);
        a = fmaf(bf16lo(w[j].y), xa.z, a);
        a = fmaf(bf16hi(w[j].y), xa.w, a);
        a = fmaf(bf16lo(w[j].z), xb.x, a);
        a = fmaf(bf16hi(w[j].z), xb.y, a);
        a = fmaf(bf16lo(w[j].w), xb.z, a);
        a = fmaf(bf16hi(w[j].w), xb.w, a);
        acc[i][j] = a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float v = warp_sum(acc[i][j]);
      if (lane == i * NB + j && r0 + i < R && n0 + j < N) {
        float* dst = y + (size_t)(r0 + i) * ldy + n0 + j;
        *dst = accumulate ? *dst + v : v;
      }
    }
}

template <int RB, int NB>
static int launch_linear_t(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
                           int accumulate, cudaStream_t st) {
  dim3 grid(cdiv(N, 8 * NB), cdiv(R, RB));
  k_linear<RB, NB><<<grid, 256, 0, st>>>(x, ldx, W, y, ldy, R, N, K, accumulate);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// The same Linear on a weight-only FP8 blob (CSMB_WEIGHTS_E4M3): one warp per output feature and RB rows, lanes split K in
// 16-element (16-byte) pieces, e4m3 -> f16 pairs with the hardware conversion (exact), f16 -> f32, fp32 FMA, shuffle tree,
// then the per-output-channel scale on the finished dot product: y = scale[n] * sum_k x[k] * q[n][k].
__device__ __forceinline__ float2 e4m3x2_to_f32(uint16_t v) {
  uint32_t h2;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(v));
  float2 f;
  asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(f.x), "=f"(f.y) : "r"(h2));
  return f;
}
template <int RB, int NB>
__global__ void __launch_bounds__(256) k_linear_e4m3(const float* __restrict__ x, int ldx, const float* __restrict__ scale,
                                                     const uint8_t* __restrict__ Q, float* __restrict__ y, int ldy, int R, int N,
                                                     int K, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 8 + warp) * NB;
  const int r0 = blockIdx.y * RB;
  if (n0 >= N) return;
  float acc[RB][NB];
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j] = 0.f;
  const uint8_t* qrow[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) qrow[j] = Q + (size_t)min(n0 + j, N - 1) * K;
  const float* xrow[RB];
#pragma unroll
  for (int i = 0; i < RB; ++i) xrow[i] = x + (size_t)min(r0 + i, R - 1) * ldx;
#pragma unroll 2
  for (int k = lane * 16; k < K; k += 512) {
    uint4 q[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) q[j] = ldg_stream(qrow[j] + k);   // NB independent 16-byte loads in flight
    float4 xv[RB][4];
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) xv[i][e] = __ldg(reinterpret_cast<const float4*>(xrow[i] + k + 4 * e));
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const uint32_t qw[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = e4m3x2_to_f32((uint16_t)(qw[e] & 0xffffu)), b = e4m3x2_to_f32((uint16_t)(qw[e] >> 16));
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          float s = acc[i][j];
          s = fmaf(a.x, xv[i][e].x, s);
          s = fmaf(a.y, xv[i][e].y, s);
          s = fmaf(b.x, xv[i][e].z, s);
          s = fmaf(b.y, xv[i][e].w, s);
          acc[i][j] = s;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const float v = warp_sum(acc[i][j]);
      if (lane == i * NB + j && r0 + i < R && n0 + j < N) {
        float* dst = y + (size_t)(r0 + i) * ldy + n0 + j;
        const float t = v * __ldg(scale + n0 + j);
        *dst = accumulate ? *dst + t : t;
      }
    }
}

int launch_linear_e4m3(const float* x, int ldx, const void* blob, float* y, int ldy, int R, int N, int K, int accumulate,
                       cudaStream_t st) {
  CSMB_REQUIRE(x && blob && y && R > 0 && N > 0 && K > 0 && K % 16 == 0 && ldx % 4 == 0);
  CSMB_REQUIRE((reinterpret_cast<uintptr_t>(blob) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const float* scale = static_cast<const float*>(blob);
  const uint8_t* Q = static_cast<const uint8_t*>(blob) + e4m3_scale_bytes(N);
  // the same sums for a row whatever R: one accumulation order (k ascending per lane, then the shuffle tree)
  if (R <= 2) {
    k_linear_e4m3<2, 4><<<dim3(cdiv(N, 32), cdiv(R, 2)), 256, 0, st>>>(x, ldx, scale, Q, y, ldy, R, N, K, accumulate);
  } else {
    k_linear_e4m3<4, 4><<<dim3(cdiv(N, 32), cdiv(R, 4)), 256, 0, st>>>(x, ldx, scale, Q, y, ldy, R, N, K, accumulate);
  }
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// Tiled variant for many rows (prefill, R > 8): 64 x 64 output tile per CTA, BK = 16, 4 x 4 outputs per thread, bf16
// weights widened to fp32 on the way into shared memory; every weight tile is read once per 64 rows instead of once
// per 4.  (CUDA-core fp32 FMA: prefill is ~0.3 TFLOP for a 164-row prompt; a tcgen05 path is the next step.)
__global__ void __launch_bounds__(256) k_linear_tiled(const float* __restrict__ x, int ldx,
                                                      const uint16_t* __restrict__ W, float* __restrict__ y, int ldy,
                                                      int R, int N, int K, int accumulate) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const float* arow = (m0 + lr < R) ? x + (size_t)(m0 + lr) * ldx : nullptr;
  const uint16_t* brow = (n0 + lr < N) ? W + (size_t)(n0 + lr) * K : nullptr;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 bw = make_uint2(0u, 0u);
    if (arow) av = *reinterpret_cast<const float4*>(arow + k0 + lk);
    if (brow) bw = *reinterpret_cast<const uint2*>(brow + k0 + lk);
    __syncthreads();
    As[lk][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Bs[lk][lr] = bf16lo(bw.x); Bs[lk + 1][lr] = bf16hi(bw.x); Bs[lk + 2][lr] = bf16lo(bw.y); Bs[lk + 3][lr] = bf16hi(bw.y);
    __syncthreads();
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
    if (m >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* dst = y + (size_t)m * ldy + n;
      *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
    }
  }
}

int launch_linear(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
                  int accumulate, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && N > 0 && K > 0 && K % 8 == 0 && ldx % 4 == 0);
  CSMB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (R == 0) return CSMB_OK;
  if (R > 8 && K % 16 == 0) {
    dim3 grid(cdiv(N, 64), cdiv(R, 64));
    k_linear_tiled<<<grid, 256, 0, st>>>(x, ldx, W, y, ldy, R, N, K, accumulate);
    CSMB_LAUNCH_CHECK();
    return CSMB_OK;
  }
  // enough warps to keep every SM's memory pipe full: prefer >= 148*16 warps
  const int nb = (N >= 9472) ? 4 : (N >= 4736 ? 2 : 1);
#define CSMB_LIN(RB)                                                                              \
  (nb == 4   ? launch_linear_t<RB, 4>(x, ldx, W, y, ldy, R, N, K, accumulate, st)                 \
   : nb == 2 ? launch_linear_t<RB, 2>(x, ldx, W, y, ldy, R, N, K, accumulate, st)                 \
             : launch_linear_t<RB, 1>(x, ldx, W, y, ldy, R, N, K, accumulate, st))
  if (R == 1) return CSMB_LIN(1);
  if (R == 2) return CSMB_LIN(2);
  return CSMB_LIN(4);
#undef CSMB_LIN
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_swiglu(const float* __restrict__ gu, float* __restrict__ out, int F,
                                                size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t r = i / F, f = i % F;
  float g = gu[r * 2 * F + f], u = gu[r * 2 * F + F + f];
  out[i] = (g / (1.f + expf(-g))) * u;
}

int launch_swiglu(const float* gu, float* out, int R, int F, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && F > 0);
  if (R == 0) return CSMB_OK;
  size_t total = (size_t)R * F;
  k_swiglu<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gu, out, F, total);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// ------------------------------------------------------------------------------------------------
// rope + append: one block per row.
__global__ void __launch_bounds__(256) k_rope_append(float* __restrict__ qkv, const float* __restrict__ rope,
                                                     float* __restrict__ kv_pool,
                                                     const int32_t* __restrict__ block_table, int max_pages,
                                                     const int32_t* __restrict__ row_seq,
                                                     const int32_t* __restrict__ row_pos, int H, int Hkv,
                                                     int hd) {
  const int r = blockIdx.x;
  const int pos = row_pos[r], seq = row_seq[r];
  const int half = hd >> 1;
  float* row = qkv + (size_t)r * (H + 2 * Hkv) * hd;
  const float* rc = rope + (size_t)pos * half * 2;
  const int page = block_table[(size_t)seq * max_pages + pos / CSMB_PAGE];
  const int slot = pos % CSMB_PAGE;
  float* kdst = kv_pool + ((size_t)page * 2 + 0) * Hkv * CSMB_PAGE * hd + (size_t)slot * hd;
  float* vdst = kv_pool + ((size_t)page * 2 + 1) * Hkv * CSMB_PAGE * hd + (size_t)slot * hd;
  // q pairs (in place)
  for (int i = threadIdx.x; i < H * half; i += blockDim.x) {
    const int h = i / half, p = i % half;
    float2 v = *reinterpret_cast<float2*>(row + h * hd + 2 * p);
    const float2 cs = *reinterpret_cast<const float2*>(rc + 2 * p);
    *reinterpret_cast<float2*>(row + h * hd + 2 * p) =
        make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
  }
  // k pairs -> cache
  for (int i = threadIdx.x; i < Hkv * half; i += blockDim.x) {
    const int h = i / half, p = i % half;
    const float2 v = *reinterpret_cast<const float2*>(row + (H + h) * hd + 2 * p);
    const float2 cs = *reinterpret_cast<const float2*>(rc + 2 * p);
    *reinterpret_cast<float2*>(kdst + (size_t)h * CSMB_PAGE * hd + 2 * p) =
        make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
  }
  // v -> cache
  for (int i = threadIdx.x; i < Hkv * hd; i += blockDim.x) {
    const int h = i / hd, c = i % hd;
    vdst[(size_t)h * CSMB_PAGE * hd + c] = row[(H + Hkv + h) * hd + c];
  }
}

int launch_rope_kv_append(float* qkv, const float* rope, float* kv_pool, const int32_t* block_table,
                          int max_pages, const int32_t* row_seq, const int32_t* row_pos, int R, int H, int Hkv,
                          int hd, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && hd % 2 == 0 && H % Hkv == 0);
  if (R == 0) return CSMB_OK;
  k_rope_append<<<R, 256, 0, st>>>(qkv, rope, kv_pool, block_table, max_pages, row_seq, row_pos, H, Hkv, hd);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// ------------------------------------------------------------------------------------------------
// attention: one warp per (row, head); 4 warps per block.  Pass 1: lane-per-key scores into shared
// memory + running max; pass 2: exp and sum; pass 3: lanes split head_dim, loop over keys.
template <int HD>
__global__ void __launch_bounds__(128) k_attention(const float* __restrict__ qkv, int ldq,
                                                   const float* __restrict__ kv_pool,
                                                   const int32_t* __restrict__ block_table, int max_pages,
                                                   const int32_t* __restrict__ row_seq,
                                                   const int32_t* __restrict__ row_pos, float* __restrict__ out,
                                                   int R, int H, int Hkv, int max_pos) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= R * H) return;
  const int r = item / H, h = item % H;
  const int kvh = h / (H / Hkv);
  float* sq = smem + warp * (HD + max_pos);
  float* sc = sq + HD;
  const int S = row_pos[r] + 1;
  const int32_t* bt = block_table + (size_t)row_seq[r] * max_pages;
  const float* q = qkv + (size_t)r * ldq + h * HD;
  for (int c = lane; c < HD; c += 32) sq[c] = q[c];
  __syncwarp();
  const float scale = rsqrtf((float)HD);
  const size_t head_off = (size_t)kvh * CSMB_PAGE * HD;
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD;
  float m = -INFINITY;
  for (int j = lane; j < S; j += 32) {
    const float* kp = kv_pool + (size_t)bt[j / CSMB_PAGE] * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kp + c);
      dot = fmaf(kv.x, sq[c], dot);
      dot = fmaf(kv.y, sq[c + 1], dot);
      dot = fmaf(kv.z, sq[c + 2], dot);
      dot = fmaf(kv.w, sq[c + 3], dot);
    }
    dot *= scale;
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
  const float inv = 1.f / sum;
  constexpr int PER = HD / 32;
  float acc[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) acc[i] = 0.f;
  for (int j = 0; j < S; ++j) {
    const float* vp = kv_pool + (size_t)bt[j / CSMB_PAGE] * page_stride + (size_t)Hkv * CSMB_PAGE * HD + head_off +
                      (size_t)(j % CSMB_PAGE) * HD;
    const float p = sc[j];
#pragma unroll
    for (int i = 0; i < PER; ++i) acc[i] = fmaf(p, vp[lane + 32 * i], acc[i]);
  }
  float* o = out + (size_t)r * H * HD + h * HD;
#pragma unroll
  for (int i = 0; i < PER; ++i) o[lane + 32 * i] = acc[i] * inv;
}

int launch_attention(const float* qkv, int ldq, const float* kv_pool, const int32_t* block_table, int max_pages,
                     const int32_t* row_seq, const int32_t* row_pos, float* out, int R, int H, int Hkv, int hd,
                     int max_pos, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && H % Hkv == 0 && max_pos > 0);
  if (R == 0) return CSMB_OK;
  const int grid = cdiv(R * H, 4);
  const size_t smem = (size_t)4 * (hd + max_pos) * sizeof(float);
  if (hd == 32) {
    k_attention<32><<<grid, 128, smem, st>>>(qkv, ldq, kv_pool, block_table, max_pages, row_seq, row_pos, out, R, H, Hkv, max_pos);
  } else if (hd == 64) {
    k_attention<64><<<grid, 128, smem, st>>>(qkv, ldq, kv_pool, block_table, max_pages, row_seq, row_pos, out, R, H, Hkv, max_pos);
  } else if (hd == 128) {
    k_attention<128><<<grid, 128, smem, st>>>(qkv, ldq, kv_pool, block_table, max_pages, row_seq, row_pos, out, R, H, Hkv, max_pos);
  } else {
    return CSMB_ERR_UNSUPPORTED;
  }
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

// ------------------------------------------------------------------------------------------------
// sampling.  One block of 256 threads per row.
struct SampleArgs {
  float inv_temp;   // 0 => greedy
  int top_k;
  float top_p, min_p;
  int min_keep;
  uint32_t seed_lo, seed_hi;
  uint64_t draw_base;
  uint32_t draw_pos_mul;
};

__global__ void __launch_bounds__(256) k_sample(const float* __restrict__ logits, int ldl,
                                                int32_t* __restrict__ out, int out_stride, int V, SampleArgs a,
                                                const int32_t* __restrict__ row_pos,
                                                const int32_t* __restrict__ forced, int forced_stride) {
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  const int r = blockIdx.x;
  const float* lg = logits + (size_t)r * ldl;
  int tok;
  if (forced) {
    tok = forced[(size_t)r * forced_stride];
  } else if (a.inv_temp == 0.f) {
    tok = block_argmax(V, [&](int i) { return lg[i]; }, red_v, red_i);
  } else {
    const uint64_t draw = a.draw_base + (uint64_t)(row_pos ? row_pos[r] : 0) * a.draw_pos_mul;
    const uint32_t dlo = (uint32_t)draw, dhi = (uint32_t)(draw >> 32);
    tok = block_argmax(
        V, [&](int i) { return lg[i] * a.inv_temp + gumbel_for(i, dlo, dhi, (uint32_t)r, a.seed_lo, a.seed_hi); },
        red_v, red_i);
  }
  if (threadIdx.x == 0) out[(size_t)r * out_stride] = tok;
}

// Filtered sampling (top-k / top-p / min-p): 1024 threads per row, V <= 4096.  Sorts the softmax
// probabilities (descending, bitonic in shared memory), derives ONE probability threshold from the three
// filters, then Gumbel-argmax over the tokens at or above it.
__global__ void __launch_bounds__(1024) k_sample_filtered(const float* __restrict__ logits, int ldl,
                                                          int32_t* __restrict__ out, int out_stride, int V,
                                                          SampleArgs a, const int32_t* __restrict__ row_pos) {
  __shared__ float sp[4096];
  __shared__ float red[32];
  __shared__ int redi[32];
  __shared__ float s_thresh;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lg = logits + (size_t)r * ldl;
  // max
  float m = -INFINITY;
  for (int i = tid; i < V; i += 1024) m = fmaxf(m, lg[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  // exp + sum
  float s = 0.f;
  for (int i = tid; i < 4096; i += 1024) {
    float e = (i < V) ? expf(lg[i] - m) : -1.f;  // padding sorts last
    sp[i] = e;
    if (i < V) s += e;
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 32; ++w) tot += red[w];
  (void)tot;  // the filters work on unnormalised e (top-p: on integer masses, below)
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= 4096; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < 4096; i += 1024) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float x = sp[i], y = sp[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (x < y) : (x > y)) {
            sp[i] = y;
            sp[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    // thresholds are on UNNORMALISED e = p * tot.  Sequential: V <= 4096 adds, rare path.
    int n1 = (a.top_k > 0 && a.top_k < V) ? a.top_k : V;
    float t = sp[n1 - 1];
    if (a.top_p > 0.f && a.top_p < 1.f) {
      // the nucleus is defined on exact integer masses q = floor(e * 2^32), so that the result does not depend on a
      // summation order and the fused samplers (frame_kernel.cu, batch_frame.cu), which find the same threshold by
      // bisection, agree with this sorter bit for bit: token i stays while the mass of the strictly more likely
      // tokens is below top_p * total mass
      unsigned long long Q = 0;
      for (int i = 0; i < V; ++i) Q += __float2ull_rz(sp[i] * 4294967296.f);
      const double need = (double)a.top_p * (double)Q;
      unsigned long long c = 0;
      int n2 = 0;
      for (int i = 0; i < n1; ++i) {
        if ((double)c < need) n2 = i + 1; else break;
        c += __float2ull_rz(sp[i] * 4294967296.f);
      }
      t = fmaxf(t, sp[n2 - 1]);
    }
    if (a.min_p > 0.f) {
      const int mk = a.min_keep < 1 ? 1 : (a.min_keep > V ? V : a.min_keep);
      t = fmaxf(t, fminf(a.min_p * sp[0], sp[mk - 1]));
    }
    s_thresh = t;
  }
  __syncthreads();
  const float thresh = s_thresh;
  const uint64_t draw = a.draw_base + (uint64_t)(row_pos ? row_pos[r] : 0) * a.draw_pos_mul;
  const uint32_t dlo = (uint32_t)draw, dhi = (uint32_t)(draw >> 32);
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < V; i += 1024) {
    const float e = expf(lg[i] - m);
    if (e >= thresh)
      argmax_combine(bv, bi, lg[i] * a.inv_temp + gumbel_for(i, dlo, dhi, (uint32_t)r, a.seed_lo, a.seed_hi), i);
  }
  warp_argmax(bv, bi);
  if (lane == 0) {
    red[warp] = bv;
    redi[warp] = bi;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w) argmax_combine(bv, bi, red[w], redi[w]);
    out[(size_t)r * out_stride] = bi == 0x7fffffff ? 0 : bi;
  }
}

int launch_sample(const float* logits, int ldl, int32_t* out, int out_stride, int R, int V,
                  const csmb_sampler& s, uint64_t draw_base, const int32_t* row_pos, uint32_t draw_pos_mul,
                  const int32_t* forced, int forced_stride, cudaStream_t st) {
  CSMB_REQUIRE(R >= 0 && V > 0 && s.temperature >= 0.f);
  if (R == 0) return CSMB_OK;
  SampleArgs a;
  a.inv_temp = s.temperature == 0.f ? 0.f : 1.f / s.temperature;
  a.top_k = s.top_k;
  a.top_p = s.top_p;
  a.min_p = s.min_p;
  a.min_keep = s.min_keep;
  a.seed_lo = (uint32_t)s.seed;
  a.seed_hi = (uint32_t)(s.seed >> 32);
  a.draw_base = draw_base;
  a.draw_pos_mul = draw_pos_mul;
  const bool filtered = !forced && s.temperature != 0.f &&
                        ((s.top_k > 0 && s.top_k < V) || (s.top_p > 0.f && s.top_p < 1.f) || s.min_p > 0.f);
  if (filtered) {
    if (V > 4096) return CSMB_ERR_UNSUPPORTED;
    k_sample_filtered<<<R, 1024, 0, st>>>(logits, ldl, out, out_stride, V, a, row_pos);
  } else {
    k_sample<<<R, 256, 0, st>>>(logits, ldl, out, out_stride, V, a, row_pos, forced, forced_stride);
  }
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

}  // namespace csmb

// ================================================================================================
// C ABI
using namespace csmb;

extern "C" {

int csmb_abi_version(void) { return CSMB_ABI_VERSION; }

const char* csmb_strerror(int status) {
  switch (status) {
    case CSMB_OK: return "ok";
    case CSMB_ERR_INVALID: return "invalid argument";
    case CSMB_ERR_CUDA: return "CUDA error (see csmb_last_cuda_error)";
    case CSMB_ERR_UNSUPPORTED: return "unsupported configuration";
    case CSMB_ERR_DEVICE: return "device is not sm_100";
    default: return "unknown status";
  }
}

const char* csmb_last_cuda_error(void) { return g_err; }

unsigned long long csmb_debug_launch_count(void) { return csmb::launch_count(); }

int csmb_check_device(int device) {
  cudaDeviceProp p;
  CSMB_CUDA(cudaGetDeviceProperties(&p, device));
  return (p.major == 10 && p.minor == 0) ? CSMB_OK : CSMB_ERR_DEVICE;
}

int csmb_embed_sum(const int32_t* tokens, const uint8_t* mask, const uint16_t* text_emb,
                   const uint16_t* audio_emb, float* out, int R, int d, int n_codebooks, int audio_vocab,
                   int device, void* stream) {
  CSMB_ENTER(device);
  return launch_embed_sum(tokens, mask, text_emb, audio_emb, out, R, d, n_codebooks, audio_vocab,
                          (cudaStream_t)stream);
}

int csmb_embed_audio(const int32_t* tokens, const uint16_t* audio_emb, float* out, int ldo, int R, int d,
                     int codebook, int audio_vocab, int device, void* stream) {
  CSMB_ENTER(device);
  return launch_embed_audio(tokens, 1, audio_emb, out, ldo, R, d, codebook, audio_vocab, (cudaStream_t)stream);
}

int csmb_rmsnorm(const float* x, int ldx, const float* w, float* y, int ldy, int R, int d, float eps, int device,
                 void* stream) {
  CSMB_ENTER(device);
  return launch_rmsnorm(x, ldx, w, y, ldy, R, d, eps, nullptr, (cudaStream_t)stream);
}

int csmb_linear(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
                int accumulate, int device, void* stream) {
  CSMB_ENTER(device);
  return launch_linear(x, ldx, W, y, ldy, R, N, K, accumulate, (cudaStream_t)stream);
}

size_t csmb_e4m3_blob_bytes(int N, int K) { return (N > 0 && K > 0) ? e4m3_blob_bytes(N, K) : 0; }

int csmb_linear_e4m3(const float* x, int ldx, const void* blob, float* y, int ldy, int R, int N, int K, int accumulate,
                     int device, void* stream) {
  CSMB_ENTER(device);
  return launch_linear_e4m3(x, ldx, blob, y, ldy, R, N, K, accumulate, (cudaStream_t)stream);
}

int csmb_swiglu(const float* gu, float* out, int R, int F, int device, void* stream) {
  CSMB_ENTER(device);
  return launch_swiglu(gu, out, R, F, (cudaStream_t)stream);
}

int csmb_rope_kv_append(float* qkv, const float* rope, float* kv_pool, const int32_t* block_table,
                        int max_pages, const int32_t* row_seq, const int32_t* row_pos, int R, int n_heads,
                        int n_kv_heads, int head_dim, int device, void* stream) {
  CSMB_ENTER(device);
  return launch_rope_kv_append(qkv, rope, kv_pool, block_table, max_pages, row_seq, row_pos, R, n_heads,
                               n_kv_heads, head_dim, (cudaStream_t)stream);
}

int csmb_attention(const float* qkv, int ldq, const float* kv_pool, const int32_t* block_table, int max_pages,
                   const int32_t* row_seq, const int32_t* row_pos, float* out, int R, int n_heads, int n_kv_heads,
                   int head_dim, int device, void* stream) {
  CSMB_ENTER(device);
  return launch_attention(qkv, ldq, kv_pool, block_table, max_pages, row_seq, row_pos, out, R, n_heads,
                          n_kv_heads, head_dim, max_pages * CSMB_PAGE, (cudaStream_t)stream);
}

int csmb_sample(const float* logits, int ldl, int32_t* out, int out_stride, int R, int V,
                const csmb_sampler* sampler, uint64_t draw, const int32_t* row_pos, uint32_t pos_mul, int device,
                void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(sampler != nullptr);
  return launch_sample(logits, ldl, out, out_stride, R, V, *sampler, draw, row_pos, pos_mul, nullptr, 0,
                       (cudaStream_t)stream);
}

}  // extern "C"
