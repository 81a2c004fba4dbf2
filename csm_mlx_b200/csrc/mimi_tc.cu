// Mimi codec at batch scale on the tensor cores (sm_100a): every SEANet Conv1d / ConvTranspose1d, every transformer
// Linear and the RVQ nearest-neighbour search as ONE persistent tcgen05 GEMM kernel with fused epilogues.
//
// The reference reaches the codec through moshi_mlx.models.mimi.Mimi (csm_mlx/tokenizers.py:61-85 encode, :148-150
// decode; generation.py:167-174).  BASELINE.json configs[4] runs it on 128 x 60 s of audio: ~85 TFLOP of dense
// contractions whose operands are fp32 in the reference.
//
//   Y[b][t][n] = epi( sum_{j < taps} sum_{c < C} A[b][t + j][c] * W[n][j*C + c] )            (time-major activations)
//
//   * Conv1d(k, stride 1): taps = k, C = Cin, A = the left-padded input rows.
//   * Conv1d(k = 2s, stride s): the input viewed as rows of s*Cin values is the same thing with taps = 2, C = s*Cin.
//   * ConvTranspose1d(k = 2s, stride s): taps = 2 over input rows (t-1, t), N = s*Cout phase-major outputs, which ARE the
//     time-major output rows t*s + r.
//   * Linear: taps = 1.
//   So a K block of 64 lies inside one tap: its A tile is a plain (non-overlapping) 2-D TMA box at row offset j.
//
// Precision: kind::f16 needs 16-bit operands.  Activations AND weights are kept as two bf16 planes, x = hi + lo up to
// 2^-17 relative, and three MMAs per K step accumulate Ahi.Whi + Alo.Whi + Ahi.Wlo into the fp32 TMEM accumulator (the
// dropped Alo.Wlo term is 2^-18): ~1e-5 of an fp32 contraction, which keeps the waveform > 80 dB SNR against the fp32
// oracle and the RVQ codes identical except at float-level near-ties.  The planes are written by the PRODUCING kernel's
// epilogue (with the consumer's ELU already applied), so no layer ever re-reads fp32 activations to convert them.
//
// Kernel: persistent, one CTA per SM, 18 warps: warp 0 = TMA producer (4 boxes per stage: Ahi, Alo 128 x 64, Whi,
// Wlo NT x 64, 128B swizzle), warp 1 = MMA issuer (tcgen05.mma kind::f16, M = 128 time rows, N = NT <= 128 output
// channels), warps 2-17 = epilogue (tcgen05.ld; bias, GELU, LayerScale, residual; fp32 and / or ELU'd bf16 hi+lo stores).
// TWO TMEM accumulators: the epilogue of tile i overlaps the loads and MMAs of tile i+1.
#include <math.h>

#include "tc.cuh"

namespace csmb {

constexpr int T3_MAX_STAGES = 6;
constexpr int T3_BM = 128;
// Epilogue warps: T3_EPI_WARPS / 4 per TMEM lane quarter, sharing a tile's 16-column chunks.  The epilogue is a chain of
// dependent ALU work (ELU, hi/lo split) and global loads / stores per row: one warp per scheduler cannot hide its own
// latencies (measured: 5.4 us per 128 x 32 tile with 4 warps), several can.
constexpr int T3_EPI_WARPS = 8;
constexpr int T3_THREADS = 64 + 32 * T3_EPI_WARPS;
// Per-warp staging buffer of the epilogue: a 32-row x 16-column chunk is transposed through shared memory so that every
// global store instruction writes whole 32-byte sectors (a row-per-lane store of 8 or 16 bytes per lane costs a full L2
// sector transaction each: measured 32 sectors per request and an L2-transaction-bound kernel).  fp32: [32][20] floats
// (row pitch 80 B, conflict-free for 16-byte accesses); planes: hi [32] x 48 B then lo [32] x 48 B.
constexpr int T3_STG_BYTES = 3072;
constexpr int T3_STG_F32_PITCH = 20;   // floats
constexpr int T3_STG_PL_PITCH = 48;    // bytes
constexpr size_t T3_SMEM_TOTAL = 225 * 1024;

struct T3Args {
  int T, rpb, B, N, C, taps, NT, colstride;
  int m_tiles, n_tiles, total_tiles, nstages, nk;
  float* y32; long long y_batch; int ldy;
  uint16_t *yhi, *ylo; long long p_batch; int ldp; int plane_act;
  const float *bias, *scale, *res; long long r_batch; int ldr;
  int act_out, vec;
  int* err;
};

__device__ __forceinline__ void t3_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
// ELU for the bf16 planes: exp(v) - 1 with the fast exponential (absolute error ~1e-7, far below the 2^-17 relative step of
// the hi/lo planes it is written to); the fp32 path keeps expm1f
__device__ __forceinline__ float t3_elu(float v) { return v > 0.f ? v : __expf(v) - 1.f; }
// 16-byte global load that does not allocate an L1 line (plain ld.global, coherent: the residual may alias the output)
__device__ __forceinline__ float4 ld_noalloc_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float t3_gelu(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(T3_THREADS, 1)
k_gemm_tc3(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
           const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo, const T3Args a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[T3_MAX_STAGES], empty[T3_MAX_STAGES], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = T3_BM * TC_BK * 2, w_bytes = (uint32_t)a.NT * TC_BK * 2;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * w_bytes;
  const int NS = a.nstages, nk = a.nk;
  uint32_t ncols = 32;
  while ((int)ncols < 2 * a.colstride) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < T3_MAX_STAGES; ++i) {
      tc_mbar_init(&full[i], 1);
      tc_mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(&tfull[i], 1);
      tc_mbar_init(&tempty[i], T3_EPI_WARPS);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ahi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_alo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
      uint32_t it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < a.total_tiles && ok; tile += gridDim.x) {
        const int nt = tile % a.n_tiles, rest = tile / a.n_tiles, mt = rest % a.m_tiles, b = rest / a.m_tiles;
        const int n0 = nt * a.NT, row0 = b * a.rpb + mt * T3_BM;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = (int)(it % (uint32_t)NS);
          const uint32_t par = (it / (uint32_t)NS) & 1u;
          if (!tc_mbar_wait(&empty[s], par ^ 1u, a.err)) { ok = false; break; }
          unsigned char* st = smem + (size_t)s * stage_bytes;
          const int k0 = kb * TC_BK;
          const int j = a.taps > 1 ? k0 / a.C : 0, c0 = a.taps > 1 ? k0 % a.C : k0;
          tc_mbar_expect_tx(&full[s], stage_bytes);
          tma_load_2d(st, &map_ahi, c0, row0 + j, &full[s]);
          tma_load_2d(st + a_bytes, &map_alo, c0, row0 + j, &full[s]);
          tma_load_2d(st + 2 * a_bytes, &map_whi, k0, n0, &full[s]);
          tma_load_2d(st + 2 * a_bytes + w_bytes, &map_wlo, k0, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(a.NT);
      uint32_t it = 0, tl = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < a.total_tiles && ok; tile += gridDim.x, ++tl) {
        const uint32_t acc = tl & 1u, u = tl >> 1;
        if (!tc_mbar_wait(&tempty[acc], (u & 1u) ^ 1u, a.err)) break;   // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)a.colstride;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = (int)(it % (uint32_t)NS);
          const uint32_t par = (it / (uint32_t)NS) & 1u;
          if (!tc_mbar_wait(&full[s], par, a.err)) { ok = false; break; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = s32(smem + (size_t)s * stage_bytes);
          const uint64_t dahi = umma_desc(sa), dalo = umma_desc(sa + a_bytes), dwhi = umma_desc(sa + 2 * a_bytes),
                         dwlo = umma_desc(sa + 2 * a_bytes + w_bytes);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            umma_f16(d_tmem, dahi + koff, dwhi + koff, idesc, (kb | k) != 0);
            umma_f16(d_tmem, dalo + koff, dwhi + koff, idesc, 1u);
            umma_f16(d_tmem, dahi + koff, dwlo + koff, idesc, 1u);
          }
          umma_commit(&empty[s]);
        }
        if (ok) umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ===== epilogue: warps 2.. -> TMEM lane quarter warp % 4; the T3_EPI_WARPS / 4 warps of a quarter share the tile's
    // 16-column chunks (chunk ci goes to sub-warp ci % NSUB)
    constexpr int NSUB = T3_EPI_WARPS / 4;
    const int q = warp & 3, sub = (warp - 2) >> 2;
    const int nchunks = a.NT / 16;
    unsigned char* stg = smem + (size_t)NS * stage_bytes + (size_t)(warp - 2) * T3_STG_BYTES;
    float* stg_f = reinterpret_cast<float*>(stg);
    // coalesced mappings of a 32 x 16 chunk: fp32 -> 4 instructions of (row = lane % 8 + 8 i, columns (lane / 8) * 4 ..+3);
    // planes -> 2 instructions of (row = lane % 16 + 16 i, columns (lane / 16) * 8 ..+7)
    const int fr = lane & 7, fc = (lane >> 3) * 4, pr = lane & 15, pc = (lane >> 4) * 8;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1u, u = tl >> 1;
      const int nt = tile % a.n_tiles, rest = tile / a.n_tiles, mt = rest % a.m_tiles, b = rest / a.m_tiles;
      const int n0 = nt * a.NT, tw = mt * T3_BM + q * 32;   // first row of this warp's 32 rows
      const int t = tw + lane;
      const bool rowok = t < a.T;
      if (!a.vec) {
        // ---- generic path (N not a multiple of 16: the 1-channel output conv): row per lane, scalar
        if (!tc_mbar_wait_warp(&tfull[acc], u & 1u, a.err)) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ci = sub; ci < nchunks; ci += NSUB) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)a.colstride + (uint32_t)(ci * 16), v);
          if (!rowok) continue;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + ci * 16 + e;
            if (n >= a.N) break;
            float xv = __uint_as_float(v[e]);
            if (a.bias) xv += __ldg(a.bias + n);
            if (a.act_out == 1) xv = t3_gelu(xv);
            if (a.scale) xv *= __ldg(a.scale + n);
            if (a.res) xv += a.res[(long long)b * a.r_batch + (long long)t * a.ldr + n];
            if (a.y32) a.y32[(long long)b * a.y_batch + (long long)t * a.ldy + n] = xv;
            if (a.yhi) {
              if (a.plane_act == 1) xv = t3_elu(xv);
              uint16_t hh, ll;
              split_bf16(xv, hh, ll);
              a.yhi[(long long)b * a.p_batch + (long long)t * a.ldp + n] = hh;
              a.ylo[(long long)b * a.p_batch + (long long)t * a.ldp + n] = ll;
            }
          }
        }
      } else {
        // ---- coalesced path.  The residual does not depend on the accumulator: its (coalesced) loads for the first chunk
        // go out before the wait, those of the next chunk before the current chunk's stores (in place: every element is
        // read before the same warp overwrites it, and no other warp touches it).
        // ALL of this warp's chunks of the tile (at most MAXC) have their residual in flight before the wait, each in its
        // own registers (a rotating buffer stalls on the copy of a load still in flight); streaming loads: with 225 KB of
        // shared memory the L1 has almost no lines left to allocate, and allocating loads serialise on them
        constexpr int MAXC = 8 / NSUB;   // NT <= 128: 8 chunks shared by NSUB sub-warps
        float4 rres_all[MAXC][4];
        if (a.res) {
#pragma unroll
          for (int k = 0; k < MAXC; ++k) {
            const int ci = sub + k * NSUB;
            if (ci < nchunks) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int tr = tw + fr + 8 * i, n = n0 + ci * 16 + fc;
                rres_all[k][i] = (tr < a.T && n < a.N) ? ld_noalloc_f4(a.res + (long long)b * a.r_batch + (long long)tr * a.ldr + n)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
        }
        if (!tc_mbar_wait_warp(&tfull[acc], u & 1u, a.err)) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
          const int ci = sub + k * NSUB;
          if (ci >= nchunks) break;
          const float4 (&rres)[4] = rres_all[k];
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)a.colstride + (uint32_t)(ci * 16), v);
          const int nc = n0 + ci * 16;   // first column of the chunk
          float x[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) x[e] = __uint_as_float(v[e]);
          if (a.bias) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 bv = (nc + j4 * 4 < a.N) ? __ldg(reinterpret_cast<const float4*>(a.bias + nc + j4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              x[j4 * 4] += bv.x; x[j4 * 4 + 1] += bv.y; x[j4 * 4 + 2] += bv.z; x[j4 * 4 + 3] += bv.w;
            }
          }
          if (a.act_out == 1) {
#pragma unroll
            for (int e = 0; e < 16; ++e) x[e] = t3_gelu(x[e]);
          }
          if (a.scale) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 sv = (nc + j4 * 4 < a.N) ? __ldg(reinterpret_cast<const float4*>(a.scale + nc + j4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              x[j4 * 4] *= sv.x; x[j4 * 4 + 1] *= sv.y; x[j4 * 4 + 2] *= sv.z; x[j4 * 4 + 3] *= sv.w;
            }
          }
          if (a.res) {
            // residual: coalesced registers -> staging -> this lane's row
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stg_f + (fr + 8 * i) * T3_STG_F32_PITCH + fc) = rres[i];
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 rv = *reinterpret_cast<const float4*>(stg_f + lane * T3_STG_F32_PITCH + j4 * 4);
              x[j4 * 4] += rv.x; x[j4 * 4 + 1] += rv.y; x[j4 * 4 + 2] += rv.z; x[j4 * 4 + 3] += rv.w;
            }
            __syncwarp();
          }
          if (a.y32) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              *reinterpret_cast<float4*>(stg_f + lane * T3_STG_F32_PITCH + j4 * 4) = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int tr = tw + fr + 8 * i, n = nc + fc;
              if (tr < a.T && n < a.N)
                *reinterpret_cast<float4*>(a.y32 + (long long)b * a.y_batch + (long long)tr * a.ldy + n) =
                    *reinterpret_cast<const float4*>(stg_f + (fr + 8 * i) * T3_STG_F32_PITCH + fc);
            }
            __syncwarp();
          }
          if (a.yhi) {
            if (a.plane_act == 1) {
#pragma unroll
              for (int e = 0; e < 16; ++e) x[e] = t3_elu(x[e]);
            }
            uint32_t h[8], l[8];
#pragma unroll
            for (int p2 = 0; p2 < 8; ++p2) {
              h[p2] = pack_bf16x2_rn(x[2 * p2], x[2 * p2 + 1]);
              l[p2] = pack_bf16x2_rn(x[2 * p2] - __uint_as_float(h[p2] << 16), x[2 * p2 + 1] - __uint_as_float(h[p2] & 0xffff0000u));
            }
            unsigned char* hs = stg + lane * T3_STG_PL_PITCH;
            unsigned char* ls = stg + 32 * T3_STG_PL_PITCH + lane * T3_STG_PL_PITCH;
            *reinterpret_cast<uint4*>(hs) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(hs + 16) = make_uint4(h[4], h[5], h[6], h[7]);
            *reinterpret_cast<uint4*>(ls) = make_uint4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<uint4*>(ls + 16) = make_uint4(l[4], l[5], l[6], l[7]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int row = pr + 16 * i, tr = tw + row, n = nc + pc;
              if (tr < a.T && n < a.N) {
                const long long o = (long long)b * a.p_batch + (long long)tr * a.ldp + n;
                *reinterpret_cast<uint4*>(a.yhi + o) = *reinterpret_cast<const uint4*>(stg + row * T3_STG_PL_PITCH + pc * 2);
                *reinterpret_cast<uint4*>(a.ylo + o) = *reinterpret_cast<const uint4*>(stg + 32 * T3_STG_PL_PITCH + row * T3_STG_PL_PITCH + pc * 2);
              }
            }
            __syncwarp();
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) t3_mbar_arrive(&tempty[acc]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ---- element-wise helpers of the plane pipeline -------------------------------------------------------------------------
// fp32 [B][T][C] (batch stride, row stride) -> bf16 hi/lo planes (their own strides), optional ELU first
__global__ void __launch_bounds__(256) k_split_planes(const float* __restrict__ x, long long x_batch, int ldx,
                                                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, long long p_batch,
                                                      int ldp, int T, int C, int act, size_t total4) {
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const int c4 = C / 4;
  const int c = (int)(i4 % c4) * 4;
  const size_t bt = i4 / c4;
  const int t = (int)(bt % T);
  const size_t b = bt / T;
  float4 v = *reinterpret_cast<const float4*>(x + b * x_batch + (size_t)t * ldx + c);
  if (act == 1) v = make_float4(t3_elu(v.x), t3_elu(v.y), t3_elu(v.z), t3_elu(v.w));
  const size_t o = b * p_batch + (size_t)t * ldp + c;
  store_split4(hi + o, lo + o, v.x, v.y, v.z, v.w);
}

// nn.LayerNorm (eps, affine) -> bf16 hi/lo planes, one warp per row (d % 128 == 0, d <= 1024); same arithmetic as k_layernorm
__global__ void __launch_bounds__(256) k_layernorm_planes(const float* __restrict__ x, long long x_batch,
                                                          const float* __restrict__ w, const float* __restrict__ bsh,
                                                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int R, int T,
                                                          int d, float eps) {
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
  float qq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float a0 = v[j].x - mean, a1 = v[j].y - mean, a2 = v[j].z - mean, a3 = v[j].w - mean;
      qq += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  const float rstd = rsqrtf(warp_sum(qq) / (float)d + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float4 g = *reinterpret_cast<const float4*>(w + j * 128 + lane * 4);
      const float4 bb = *reinterpret_cast<const float4*>(bsh + j * 128 + lane * 4);
      const size_t o = (size_t)r * d + j * 128 + lane * 4;
      store_split4(hi + o, lo + o, (v[j].x - mean) * rstd * g.x + bb.x, (v[j].y - mean) * rstd * g.y + bb.y,
                   (v[j].z - mean) * rstd * g.z + bb.z, (v[j].w - mean) * rstd * g.w + bb.w);
    }
}

// ---- whole-clip attention of the codec transformers ---------------------------------------------------------------------
// RoPE (adjacent pairs, angle = position * freqs[i]; position = row index: whole clips start at 0) on q and k, in place in
// qkv [B][T][3][H][64].
__global__ void __launch_bounds__(256) k_mimi_rope_inplace(float* __restrict__ qkv, const float* __restrict__ freqs, int T, int H) {
  const int bt = blockIdx.x, t = bt % T;
  float* row = qkv + (size_t)bt * 3 * H * 64;
  for (int i = threadIdx.x; i < H * 32; i += blockDim.x) {
    const int h = i >> 5, p = i & 31;
    float sn, cs;
    sincosf((float)t * freqs[p], &sn, &cs);
    const float2 qv = *reinterpret_cast<float2*>(row + h * 64 + 2 * p);
    *reinterpret_cast<float2*>(row + h * 64 + 2 * p) = make_float2(qv.x * cs - qv.y * sn, qv.x * sn + qv.y * cs);
    const float2 kv = *reinterpret_cast<float2*>(row + (H + h) * 64 + 2 * p);
    *reinterpret_cast<float2*>(row + (H + h) * 64 + 2 * p) = make_float2(kv.x * cs - kv.y * sn, kv.x * sn + kv.y * cs);
  }
}

// Causal attention over the last `ctx` positions, one block per (32 queries, head, clip): the <= ctx + 31 key and value rows
// the block's queries share are staged ONCE in shared memory (the per-query kernel of the streaming path re-reads them from
// L2 for every query: 128 KB per query, L2-bandwidth-bound at batch scale), and every warp works on FOUR consecutive queries
// at a time so that a key / value row read from shared memory is used four times (the kernel is shared-memory-bandwidth
// bound).  Per (query, key) the dot product runs over the 64 dims in ascending order and P.V over the keys in ascending
// order, as in k_mimi_attention; the output goes straight to the bf16 hi/lo planes of the out-projection's operand.
constexpr int AT_QT = 32, AT_PITCH = 68, AT_WARPS = 8, AT_QW = AT_QT / AT_WARPS;
static_assert(AT_QW == 4, "the warp body is written for four queries");
__global__ void __launch_bounds__(AT_WARPS * 32) k_mimi_attn_tile(const float* __restrict__ qkv, uint16_t* __restrict__ out_hi,
                                                                  uint16_t* __restrict__ out_lo, int T, int H, int ctx) {
  extern __shared__ float at_sm[];
  const int kt_max = AT_QT + ctx - 1, sc_len = ctx + AT_QW - 1;
  float* sK = at_sm;
  float* sV = sK + (size_t)kt_max * AT_PITCH;
  float* sQ = sV + (size_t)kt_max * AT_PITCH;          // [warp][4 queries][64]
  float* sS = sQ + AT_WARPS * AT_QW * 64;              // [warp][key][4 queries]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * AT_QT, h = blockIdx.y, b = blockIdx.z;
  const int k_first = t0 - ctx + 1 > 0 ? t0 - ctx + 1 : 0;
  const int k_last = t0 + AT_QT - 1 < T - 1 ? t0 + AT_QT - 1 : T - 1;
  const int nk = k_last - k_first + 1;
  const size_t rs = (size_t)3 * H * 64;
  const float* base = qkv + (size_t)b * T * rs;
  // four rows per thread and iteration: eight independent 16-byte loads in flight
  for (int idx0 = threadIdx.x; idx0 < nk * 16; idx0 += 4 * AT_WARPS * 32) {
    float4 kv[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + u * AT_WARPS * 32;
      if (idx < nk * 16) {
        const float* src = base + (size_t)(k_first + (idx >> 4)) * rs + (size_t)h * 64 + (idx & 15) * 4;
        kv[u] = *reinterpret_cast<const float4*>(src + (size_t)H * 64);
        vv[u] = *reinterpret_cast<const float4*>(src + (size_t)2 * H * 64);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = idx0 + u * AT_WARPS * 32;
      if (idx < nk * 16) {
        *reinterpret_cast<float4*>(sK + (size_t)(idx >> 4) * AT_PITCH + (idx & 15) * 4) = kv[u];
        *reinterpret_cast<float4*>(sV + (size_t)(idx >> 4) * AT_PITCH + (idx & 15) * 4) = vv[u];
      }
    }
  }
  const int tq0 = t0 + warp * AT_QW;
  float* sq = sQ + warp * AT_QW * 64;
  float* sc = sS + (size_t)warp * sc_len * AT_QW;
  const int nq = tq0 < T ? (T - tq0 < AT_QW ? T - tq0 : AT_QW) : 0;
  for (int i = 0; i < AT_QW; ++i) {   // this warp's query vectors (zeros past the end of the clip)
    const int t = tq0 + i;
    sq[i * 64 + lane] = t < T ? base[(size_t)t * rs + (size_t)h * 64 + lane] : 0.f;
    sq[i * 64 + lane + 32] = t < T ? base[(size_t)t * rs + (size_t)h * 64 + lane + 32] : 0.f;
  }
  __syncthreads();
  if (nq == 0) return;
  // union key window of the warp's queries: positions first0 .. tq0 + nq - 1
  const int first0 = tq0 - ctx + 1 > 0 ? tq0 - ctx + 1 : 0;
  const int Su = tq0 + nq - first0, off = first0 - k_first;
  float mx[AT_QW];
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) mx[i] = -INFINITY;
  for (int j = lane; j < Su; j += 64) {   // two keys per lane at a time x four queries: eight independent fma chains
    const float* kp[2];
    float dot[2][AT_QW];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      kp[u] = sK + (size_t)(off + (j + 32 * u < Su ? j + 32 * u : j)) * AT_PITCH;
#pragma unroll
      for (int i = 0; i < AT_QW; ++i) dot[u][i] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 64; c += 4) {
      float4 qv[AT_QW];
#pragma unroll
      for (int i = 0; i < AT_QW; ++i) qv[i] = *reinterpret_cast<const float4*>(sq + i * 64 + c);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 kv = *reinterpret_cast<const float4*>(kp[u] + c);
#pragma unroll
        for (int i = 0; i < AT_QW; ++i) {
          dot[u][i] = fmaf(kv.x, qv[i].x, dot[u][i]);
          dot[u][i] = fmaf(kv.y, qv[i].y, dot[u][i]);
          dot[u][i] = fmaf(kv.z, qv[i].z, dot[u][i]);
          dot[u][i] = fmaf(kv.w, qv[i].w, dot[u][i]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int jj = j + 32 * u;
      if (jj < Su) {
        const int pos = first0 + jj;
        float d[AT_QW];
#pragma unroll
        for (int i = 0; i < AT_QW; ++i) {
          const int t = tq0 + i;
          const bool valid = i < nq && pos <= t && pos > t - ctx;
          d[i] = valid ? dot[u][i] * 0.125f : -INFINITY;
          mx[i] = fmaxf(mx[i], d[i]);
        }
        *reinterpret_cast<float4*>(sc + (size_t)jj * AT_QW) = make_float4(d[0], d[1], d[2], d[3]);
      }
    }
  }
  float sum[AT_QW];
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) {
    mx[i] = warp_max(mx[i]);
    sum[i] = 0.f;
  }
  __syncwarp();
  for (int j = lane; j < Su; j += 32) {
    float4 e = *reinterpret_cast<const float4*>(sc + (size_t)j * AT_QW);
    e.x = expf(e.x - mx[0]);   // masked entries: exp(-inf) = 0
    e.y = nq > 1 ? expf(e.y - mx[1]) : 0.f;
    e.z = nq > 2 ? expf(e.z - mx[2]) : 0.f;
    e.w = nq > 3 ? expf(e.w - mx[3]) : 0.f;
    *reinterpret_cast<float4*>(sc + (size_t)j * AT_QW) = e;
    sum[0] += e.x; sum[1] += e.y; sum[2] += e.z; sum[3] += e.w;
  }
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) sum[i] = warp_sum(sum[i]);
  __syncwarp();
  float a0[AT_QW], a1[AT_QW];
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) a0[i] = a1[i] = 0.f;
  // keys in groups of 4: the loads of a group are independent; accumulation order stays j ascending per query
  for (int j0 = 0; j0 < Su; j0 += 4) {
    float4 pj[4];
    float v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j0 + u < Su ? j0 + u : Su - 1;
      const float* vp = sV + (size_t)(off + jj) * AT_PITCH;
      pj[u] = *reinterpret_cast<const float4*>(sc + (size_t)jj * AT_QW);
      v0[u] = vp[lane];
      v1[u] = vp[lane + 32];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (j0 + u < Su) {
        a0[0] = fmaf(pj[u].x, v0[u], a0[0]); a1[0] = fmaf(pj[u].x, v1[u], a1[0]);
        a0[1] = fmaf(pj[u].y, v0[u], a0[1]); a1[1] = fmaf(pj[u].y, v1[u], a1[1]);
        a0[2] = fmaf(pj[u].z, v0[u], a0[2]); a1[2] = fmaf(pj[u].z, v1[u], a1[2]);
        a0[3] = fmaf(pj[u].w, v0[u], a0[3]); a1[3] = fmaf(pj[u].w, v1[u], a1[3]);
      }
  }
#pragma unroll
  for (int i = 0; i < AT_QW; ++i) {
    if (i >= nq) break;
    const float inv = 1.f / sum[i];
    const size_t o = ((size_t)b * T + tq0 + i) * H * 64 + (size_t)h * 64;
    uint16_t hh, ll;
    split_bf16(a0[i] * inv, hh, ll);
    out_hi[o + lane] = hh;
    out_lo[o + lane] = ll;
    split_bf16(a1[i] * inv, hh, ll);
    out_hi[o + lane + 32] = hh;
    out_lo[o + lane + 32] = ll;
  }
}

// First SEANet encoder conv, Conv1d(1 -> C, k): x [B][(k-1) + N] fp32 (left-padded) -> y [B][ypad + N][C] fp32 and
// ELU(y) planes of the same layout.  One thread per (sample, 4 channels); weights [C][k].
__global__ void __launch_bounds__(256) k_conv_in(const float* __restrict__ x, long long x_batch, const float* __restrict__ w,
                                                 const float* __restrict__ bias, float* __restrict__ y, uint16_t* __restrict__ hi,
                                                 uint16_t* __restrict__ lo, long long y_batch, int N, int C, int k, size_t total) {
  extern __shared__ float sw[];  // [C][k] + [C]
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sw[C * k + i] = bias[i];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = C / 4;
  const int c = (int)(i % c4) * 4;
  const size_t bt = i / c4;
  const int t = (int)(bt % N);
  const size_t b = bt / N;
  const float* xp = x + b * x_batch + t;
  float acc[4] = {sw[C * k + c], sw[C * k + c + 1], sw[C * k + c + 2], sw[C * k + c + 3]};
  for (int j = 0; j < k; ++j) {
    const float xv = xp[j];
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = fmaf(xv, sw[(c + e) * k + j], acc[e]);
  }
  const size_t o = b * y_batch + (size_t)t * C + c;
  *reinterpret_cast<float4*>(y + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  store_split4(hi + o, lo + o, t3_elu(acc[0]), t3_elu(acc[1]), t3_elu(acc[2]), t3_elu(acc[3]));
}

// RVQ encode step on planes: like k_rvq_argmin_update, and the updated residual is also written as hi/lo planes (the next
// codebook's search reads them).  margin (optional) [M]: best-vs-second-best gap of the distance, for near-tie diagnostics.
__global__ void __launch_bounds__(256) k_rvq_argmin_update_planes(const float* __restrict__ dots, const float* __restrict__ c2,
                                                                  const float* __restrict__ codebook, float* __restrict__ r,
                                                                  uint16_t* __restrict__ rhi, uint16_t* __restrict__ rlo,
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
  for (int c = threadIdx.x; c < D; c += 256) {
    const float v = r[(size_t)m * D + c] - crow[c];
    r[(size_t)m * D + c] = v;
    uint16_t hh, ll;
    split_bf16(v, hh, ll);
    rhi[(size_t)m * D + c] = hh;
    rlo[(size_t)m * D + c] = ll;
  }
}

}  // namespace csmb

using namespace csmb;

extern "C" {

int csmb_gemm_tc3(const csmb_tc3* g, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(g && g->a_hi && g->a_lo && g->w_hi && g->w_lo && g->err_flag);
  CSMB_REQUIRE(g->B > 0 && g->T >= 0 && g->N > 0 && g->C > 0 && g->taps >= 1 && g->rpb > 0 && g->a_rows > 0);
  if (g->T == 0) return CSMB_OK;
  const int K = g->taps * g->C;
  CSMB_REQUIRE(g->lda % 8 == 0 && g->lda >= g->C && g->ldw % 8 == 0 && g->ldw >= K);
  CSMB_REQUIRE(g->taps == 1 ? (g->C % 8 == 0) : (g->C % TC_BK == 0));
  CSMB_REQUIRE(((reinterpret_cast<uintptr_t>(g->a_hi) | reinterpret_cast<uintptr_t>(g->a_lo) |
                 reinterpret_cast<uintptr_t>(g->w_hi) | reinterpret_cast<uintptr_t>(g->w_lo)) & 15) == 0);
  CSMB_REQUIRE(g->y32 || g->y_hi);
  CSMB_REQUIRE((g->y_hi == nullptr) == (g->y_lo == nullptr));
  const int N16 = ((g->N + 15) / 16) * 16;
  CSMB_REQUIRE(g->w_rows >= N16);
  T3Args a;
  a.T = g->T; a.rpb = g->rpb; a.B = g->B; a.N = g->N; a.C = g->C; a.taps = g->taps;
  a.NT = N16 <= 128 ? N16 : 128;
  a.colstride = a.NT < 32 ? 32 : ((a.NT + 31) / 32) * 32;
  a.m_tiles = cdiv(g->T, T3_BM);
  a.n_tiles = cdiv(N16, a.NT);
  const long long total = (long long)a.m_tiles * a.n_tiles * g->B;
  CSMB_REQUIRE(total < (1ll << 31));
  a.total_tiles = (int)total;
  a.nk = cdiv(K, TC_BK);
  const size_t stage = (size_t)2 * T3_BM * TC_BK * 2 + (size_t)2 * a.NT * TC_BK * 2;
  const size_t epi_smem = (size_t)T3_EPI_WARPS * T3_STG_BYTES;
  int nstages = (int)((T3_SMEM_TOTAL - 1024 - epi_smem) / stage);
  nstages = nstages > T3_MAX_STAGES ? T3_MAX_STAGES : nstages;
  CSMB_REQUIRE(nstages >= 2);
  a.nstages = nstages;
  a.y32 = g->y32; a.y_batch = g->y_batch; a.ldy = g->ldy;
  a.yhi = g->y_hi; a.ylo = g->y_lo; a.p_batch = g->p_batch; a.ldp = g->ldp; a.plane_act = g->plane_act;
  a.bias = g->bias; a.scale = g->scale; a.res = g->residual; a.r_batch = g->r_batch; a.ldr = g->ldr;
  a.act_out = g->act_out;
  a.err = g->err_flag;
  auto al = [](const void* p, uintptr_t m) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & m) == 0; };
  a.vec = (g->N % 16 == 0) && al(g->bias, 15) && al(g->scale, 15) &&
          (!g->y32 || (al(g->y32, 15) && g->ldy % 4 == 0 && g->y_batch % 4 == 0)) &&
          (!g->residual || (al(g->residual, 15) && g->ldr % 4 == 0 && g->r_batch % 4 == 0)) &&
          (!g->y_hi || (al(g->y_hi, 15) && al(g->y_lo, 15) && g->ldp % 8 == 0 && g->p_batch % 8 == 0));
  CUtensorMap mahi, malo, mwhi, mwlo;
  if (!tc_make_map_ld(&mahi, g->a_hi, g->a_rows, g->C, g->lda, T3_BM) || !tc_make_map_ld(&malo, g->a_lo, g->a_rows, g->C, g->lda, T3_BM) ||
      !tc_make_map_ld(&mwhi, g->w_hi, g->w_rows, K, g->ldw, a.NT) || !tc_make_map_ld(&mwlo, g->w_lo, g->w_rows, K, g->ldw, a.NT))
    return CSMB_ERR_UNSUPPORTED;
  const size_t smem = stage * nstages + epi_smem + 1024;
  CSMB_CUDA(cudaFuncSetAttribute(k_gemm_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3_SMEM_TOTAL));
  int sms = 0;
  CSMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int grid = a.total_tiles < sms ? a.total_tiles : sms;
  k_gemm_tc3<<<grid, T3_THREADS, smem, (cudaStream_t)stream>>>(mahi, malo, mwhi, mwlo, a);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_split_planes(const float* x, long long x_batch, int ldx, uint16_t* hi, uint16_t* lo, long long p_batch, int ldp,
                      int B, int T, int C, int act, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(x && hi && lo && C % 4 == 0 && ldx % 4 == 0 && x_batch % 4 == 0 && ldp % 4 == 0 && p_batch % 4 == 0);
  const size_t total4 = (size_t)B * T * (C / 4);
  if (total4 == 0) return CSMB_OK;
  k_split_planes<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, x_batch, ldx, hi, lo, p_batch, ldp, T, C,
                                                                                     act, total4);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_layernorm_planes(const float* x, long long x_batch, const float* w, const float* b, uint16_t* hi, uint16_t* lo, int B,
                          int T, int d, float eps, int device, void* stream) {
  CSMB_ENTER(device);
  const int R = B * T;
  if (R == 0) return CSMB_OK;
  CSMB_REQUIRE(d % 128 == 0 && d <= 1024 && x_batch % 4 == 0);
  k_layernorm_planes<<<cdiv(R, 8), 256, 0, (cudaStream_t)stream>>>(x, x_batch, w, b, hi, lo, R, T, d, eps);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_mimi_attention_planes(float* qkv, const float* freqs, uint16_t* out_hi, uint16_t* out_lo, int B, int T, int H, int ctx,
                               int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(qkv && freqs && out_hi && out_lo && B > 0 && T > 0 && H > 0 && ctx > 0);
  cudaStream_t st = (cudaStream_t)stream;
  k_mimi_rope_inplace<<<B * T, 256, 0, st>>>(qkv, freqs, T, H);
  CSMB_LAUNCH_CHECK();
  const size_t smem = ((size_t)2 * (AT_QT + ctx - 1) * AT_PITCH + AT_WARPS * AT_QW * 64 + (size_t)AT_WARPS * (ctx + AT_QW - 1) * AT_QW) * sizeof(float);
  CSMB_REQUIRE(smem <= 220 * 1024);
  CSMB_CUDA(cudaFuncSetAttribute(k_mimi_attn_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_mimi_attn_tile<<<dim3(cdiv(T, AT_QT), H, B), AT_WARPS * 32, smem, st>>>(qkv, out_hi, out_lo, T, H, ctx);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_conv_in_planes(const float* x, long long x_batch, const float* w, const float* bias, float* y, uint16_t* hi,
                        uint16_t* lo, long long y_batch, int B, int N, int C, int k, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(x && w && bias && y && hi && lo && C % 4 == 0 && y_batch % 4 == 0 && k >= 1 && C * (k + 1) * 4 <= 40000);
  const size_t total = (size_t)B * N * (C / 4);
  if (total == 0) return CSMB_OK;
  k_conv_in<<<(unsigned)((total + 255) / 256), 256, (size_t)C * (k + 1) * sizeof(float), (cudaStream_t)stream>>>(
      x, x_batch, w, bias, y, hi, lo, y_batch, N, C, k, total);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

int csmb_rvq_argmin_update_planes(const float* dots, const float* c2, const float* codebook, float* r, uint16_t* r_hi,
                                  uint16_t* r_lo, int32_t* codes, int M, int bins, int D, int K, int k, int F, int device,
                                  void* stream) {
  CSMB_ENTER(device);
  if (M == 0) return CSMB_OK;
  CSMB_REQUIRE(dots && c2 && codebook && r && r_hi && r_lo && codes);
  k_rvq_argmin_update_planes<<<M, 256, 0, (cudaStream_t)stream>>>(dots, c2, codebook, r, r_hi, r_lo, codes, bins, D, K, k, F);
  CSMB_LAUNCH_CHECK();
  return CSMB_OK;
}

}  // extern "C"
