// Persistent batched frame kernel (sm_100a): ONE cooperative launch computes one 80 ms frame of generate_frame
// (csm_mlx/generation.py:21-92 with T = 1, plus the input construction of :156-161) for B sequences in lock-step —
// backbone step over the paged KV cache, codebook-0 head + sampling, then the 31-step depth-decoder loop with heads,
// sampling and embedding gathers — without returning to the host.
//
// The dataflow is that of the fused kernel chain in batch_frame.cu (tcgen05 Linear -> fp32 split-K partials -> one
// fused "partial sum + op + bf16 hi/lo split" stage), but the ~1 250 kernel boundaries become grid barriers and the
// weight stream no longer stops at them:
//   * one CTA per SM, 8 warps.  Warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 4-7 = TMEM epilogue; in the
//     element-wise phases all 8 warps work.
//   * the shared-memory ring (W 128x64 | Xhi RNx64 | Xlo RNx64 per stage, 128B-swizzled) is persistent: as soon as the
//     MMAs of one Linear have drained a stage, the producer refills it with the NEXT Linear's weight tile — weights
//     depend on nothing — so the HBM stream runs through the element-wise phase and both grid barriers in between;
//     only the activation tiles are loaded after the barrier that publishes them.
//   * the TMEM accumulator (one allocation per launch) is reused by every Linear; all tensor maps (one per weight
//     matrix, six for the activation planes) live in the kernel's __grid_constant__ parameter block.
//   * data written by other CTAs during the launch is read with L2-only loads (ld.global.cg) or by TMA; generic-proxy
//     writes are ordered before async-proxy reads with fence.proxy.async on both sides of every barrier.
// All waits are bounded: on timeout a sticky error flag is raised and every later wait falls through.
#include <math.h>

#include "tc.cuh"

namespace csmb {

constexpr int BP_THREADS = 256;
constexpr int BP_MAX_STAGES = 8;
constexpr int BP_MAXMAPS = 124;
constexpr unsigned BP_SPIN = 1u << 22;

struct BpSample {
  float inv_temp;  // 0 => greedy
  uint32_t seed_lo, seed_hi;
  uint32_t draw_pos_mul;
  unsigned long long draw_base;
};

struct BpParams {
  CUtensorMap maps[BP_MAXMAPS];
  csmb_model m;
  int B, RN, max_pages, nstages, stage_stride, min_kblocks, max_ctas, attn_floats;
  float* kv_pool;
  unsigned long long kv_layer_stride;
  const int32_t* block_table;
  float* dec_kv_pool;
  unsigned long long dec_kv_layer_stride;
  const int32_t* prev_frame;
  const int32_t* pos;
  int32_t* frame;
  float *x, *dx, *h_last, *part;
  uint16_t *hi, *lo;
  unsigned long long* bar;       // [0] arrival counter (monotonic), [1] base of the next launch
  int* err;
  unsigned long long* prof;      // optional [grid][2][12] cycle counters of threads 0 and 128 (debug); null in production
  BpSample sa;
};

struct BpGemm {
  int wmap, xmap;  // xmap = hi plane, xmap + 1 = lo plane
  int R, N, K;
};

__host__ __device__ inline int bp_split(int tiles, int nk, int min_kblocks, int max_ctas) {
  int S = max_ctas / tiles;
  const int cap = nk / (min_kblocks > 0 ? min_kblocks : 1);
  S = S < cap ? S : cap;
  return S < 1 ? 1 : S;
}
__host__ __device__ inline int bp_tiles(int R, int N, int RN) { return ((N + TC_BM - 1) / TC_BM) * ((R + RN - 1) / RN); }

// map indices (shared by host and device)
__host__ __device__ inline int bp_wmap_backbone(int l, int j) { return l * 4 + j; }
__host__ __device__ inline int bp_wmap_decoder(const csmb_model& m, int l, int j) { return m.backbone.n_layers * 4 + l * 4 + j; }
__host__ __device__ inline int bp_wmap_c0(const csmb_model& m) { return (m.backbone.n_layers + m.decoder.n_layers) * 4; }
__host__ __device__ inline int bp_wmap_proj(const csmb_model& m) { return bp_wmap_c0(m) + 1; }
__host__ __device__ inline int bp_wmap_head(const csmb_model& m, int i) { return bp_wmap_c0(m) + 2 + i; }
__host__ __device__ inline int bp_xmap(const csmb_model& m, int kind) { return bp_wmap_c0(m) + 2 + (m.n_codebooks - 1) + kind * 2; }
enum { XK_DB = 0, XK_FB = 1, XK_DD = 2, XK_FD = 3 };

#ifdef __CUDACC__

struct BpCtx {
  const BpParams* p;
  unsigned char* ring;
  float* scratch;
  uint64_t *full, *empty, *acc_full;
  uint32_t tmem;
  uint32_t q_prod, q_mma, acc_par;
  int pre;                       // weight stages of the next Linear already issued by the producer
  unsigned long long bar_next;   // arrival count that completes the next barrier
  bool dead;
  int warp, lane;
  unsigned long long t_acc[12], t_last;
};
enum { BT_GEMM = 0, BT_PREFETCH = 1, BT_BARRIER = 2, BT_ATTN = 3, BT_NORM = 4, BT_SWIGLU = 5, BT_SAMPLE = 6, BT_EMBED = 7,
       BT_ACCWAIT = 8, BT_EPI = 9 };
__device__ __forceinline__ void bp_mark(BpCtx& c, int cat) {
  if (c.p->prof != nullptr && (threadIdx.x & 127) == 0) {
    const unsigned long long t = (unsigned long long)clock64();
    c.t_acc[cat] += t - c.t_last;
    c.t_last = t;
  }
}

__device__ __forceinline__ bool bp_wait(BpCtx& c, uint64_t* b, uint32_t parity) {
  if (c.dead) return false;
  if (!tc_mbar_wait(b, parity, c.p->err)) c.dead = true;
  return !c.dead;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// grid-wide barrier (all CTAs are co-resident: cooperative launch, one CTA per SM)
__device__ void bp_grid_sync(BpCtx& c) {
  asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of this phase before later TMA reads
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(c.p->bar, 1ull);
    unsigned spins = 0;
    while (ld_acquire_u64(c.p->bar) < c.bar_next) {
      if (++spins > BP_SPIN) {
        atomicExch(c.p->err, 2);
        break;
      }
      if ((spins & 1023) == 0 && *reinterpret_cast<volatile int*>(c.p->err) != 0) break;
    }
    __threadfence();
  }
  c.bar_next += gridDim.x;
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");
  bp_mark(c, BT_BARRIER);
}

// ---------------------------------------------------------------------------------------------- Linear phases
struct BpItem {
  int n0, r0, kb0, nk, z;
};
__device__ __forceinline__ int bp_items(const BpParams& p, const BpGemm& g, int& S) {
  const int tiles = bp_tiles(g.R, g.N, p.RN);
  S = bp_split(tiles, g.K / TC_BK, p.min_kblocks, p.max_ctas);
  return tiles * S;
}
__device__ __forceinline__ BpItem bp_item(const BpParams& p, const BpGemm& g, int item, int S) {
  const int tiles_n = (g.N + TC_BM - 1) / TC_BM, tiles = bp_tiles(g.R, g.N, p.RN);
  const int z = item / tiles, t = item % tiles, nk_total = g.K / TC_BK;
  BpItem it;
  it.z = z;
  it.n0 = (t % tiles_n) * TC_BM;
  it.r0 = (t / tiles_n) * p.RN;
  it.kb0 = (int)(((long long)nk_total * z) / S);
  it.nk = (int)(((long long)nk_total * (z + 1)) / S) - it.kb0;
  return it;
}

// producer thread: weight tiles of the first stages of this CTA's first item of Linear g (no activations yet)
__device__ void bp_prefetch(BpCtx& c, const BpGemm& g) {
  const BpParams& p = *c.p;
  c.pre = 0;
  if (c.warp != 0 || c.lane != 0 || g.wmap < 0) return;
  int S;
  const int items = bp_items(p, g, S);
  if ((int)blockIdx.x >= items) return;
  const BpItem it = bp_item(p, g, blockIdx.x, S);
  const uint32_t w_bytes = TC_BM * TC_BK * 2, x_bytes = (uint32_t)p.RN * TC_BK * 2;
  const int NS = p.nstages, pre = it.nk < NS ? it.nk : NS;
  for (int kb = 0; kb < pre; ++kb) {
    const uint32_t q = c.q_prod + kb;
    const int s = q % NS;
    if (!bp_wait(c, &c.empty[s], ((q / NS) & 1) ^ 1)) return;
    tc_mbar_expect_tx(&c.full[s], w_bytes + 2 * x_bytes);
    tma_load_2d(c.ring + (size_t)s * p.stage_stride, &p.maps[g.wmap], (it.kb0 + kb) * TC_BK, it.n0, &c.full[s]);
    c.pre = kb + 1;
  }
}

// Y partials of Linear g: part[z][r][n] = sum over this split's K range of X[r][k] W[n][k]
__device__ void bp_gemm(BpCtx& c, const BpGemm& g) {
  const BpParams& p = *c.p;
  int S;
  const int items = bp_items(p, g, S);
  const uint32_t w_bytes = TC_BM * TC_BK * 2, x_bytes = (uint32_t)p.RN * TC_BK * 2, x_off = w_bytes;
  const int NS = p.nstages, RN = p.RN;
  bool first = true;
  for (int item = blockIdx.x; item < items; item += gridDim.x, first = false) {
    const BpItem it = bp_item(p, g, item, S);
    if (c.warp == 0) {
      if (c.lane == 0) {
        const int pre = first ? c.pre : 0;
        for (int kb = 0; kb < it.nk; ++kb) {
          const uint32_t q = c.q_prod + kb;
          const int s = q % NS;
          unsigned char* st = c.ring + (size_t)s * p.stage_stride;
          if (kb >= pre) {
            if (!bp_wait(c, &c.empty[s], ((q / NS) & 1) ^ 1)) break;
            tc_mbar_expect_tx(&c.full[s], w_bytes + 2 * x_bytes);
            tma_load_2d(st, &p.maps[g.wmap], (it.kb0 + kb) * TC_BK, it.n0, &c.full[s]);
          }
          tma_load_2d(st + x_off, &p.maps[g.xmap], (it.kb0 + kb) * TC_BK, it.r0, &c.full[s]);
          tma_load_2d(st + x_off + x_bytes, &p.maps[g.xmap + 1], (it.kb0 + kb) * TC_BK, it.r0, &c.full[s]);
        }
        c.q_prod += it.nk;
        c.pre = 0;
      }
    } else if (c.warp == 1) {
      if (c.lane == 0) {
        const uint32_t idesc = umma_idesc(RN);
        for (int kb = 0; kb < it.nk; ++kb) {
          const uint32_t q = c.q_mma + kb;
          const int s = q % NS;
          if (!bp_wait(c, &c.full[s], (q / NS) & 1)) break;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = s32(c.ring + (size_t)s * p.stage_stride);
          const uint64_t da = umma_desc(sa), dhi = umma_desc(sa + x_off), dlo = umma_desc(sa + x_off + x_bytes);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t koff = (uint64_t)((k * 32) >> 4);
            umma_f16(c.tmem, da + koff, dhi + koff, idesc, (kb | k) != 0);
            umma_f16(c.tmem, da + koff, dlo + koff, idesc, 1u);
          }
          umma_commit(&c.empty[s]);
        }
        umma_commit(c.acc_full);
        c.q_mma += it.nk;
      }
    } else if (c.warp >= 4) {
      const int quarter = c.warp & 3;
      bp_mark(c, BT_GEMM);
      const bool ok = bp_wait(c, c.acc_full, c.acc_par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      bp_mark(c, BT_ACCWAIT);
      const int n = it.n0 + quarter * 32 + c.lane;
      if (ok) {
        float* dst0 = p.part + (size_t)it.z * g.R * g.N + n;
        for (int c0 = 0; c0 < RN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(c.tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
          if (n < g.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int r = it.r0 + c0 + j;
              if (c0 + j < RN && r < g.R) dst0[(size_t)r * g.N] = __uint_as_float(v[j]);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      bp_mark(c, BT_EPI);
    }
    c.acc_par ^= 1;
    if (item + (int)gridDim.x < items) {
      // a second item on this CTA: the accumulator must be drained before its MMAs overwrite it
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------- element-wise phases
struct BpPart {
  const float* p;
  int S;
  size_t stride;
  int ld;
};
__device__ __forceinline__ BpPart bp_part(const BpParams& p, const BpGemm& g) {
  int S;
  bp_items(p, g, S);
  return BpPart{p.part, S, (size_t)g.R * g.N, g.N};
}
__device__ __forceinline__ float bp_sum1(const BpPart& pi, size_t off) {
  float v = 0.f;
#pragma unroll 4
  for (int z = 0; z < pi.S; ++z) v += __ldcg(pi.p + (size_t)z * pi.stride + off);
  return v;
}
__device__ __forceinline__ float2 bp_sum2(const BpPart& pi, size_t off) {
  float2 v = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int z = 0; z < pi.S; ++z) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(pi.p + (size_t)z * pi.stride + off));
    v.x += t.x;
    v.y += t.y;
  }
  return v;
}
__device__ __forceinline__ float4 bp_sum4(const BpPart& pi, size_t off) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int z = 0; z < pi.S; ++z) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(pi.p + (size_t)z * pi.stride + off));
    v.x += t.x;
    v.y += t.y;
    v.z += t.z;
    v.w += t.w;
  }
  return v;
}
__device__ __forceinline__ float bp_block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  __syncthreads();
  return tot;
}

// x[b] = sum_k audio_emb[prev[b][k] + k*V]  (generation.py:156-161, models.py:82-92), RMSNorm(w) -> hi/lo.  d == 2048.
__device__ void bp_embed_norm(BpCtx& c, const float* w, float* red) {
  const BpParams& p = *c.p;
  const int d = p.m.backbone.d_model, V = p.m.audio_vocab, ncb = p.m.n_codebooks;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const int ch = threadIdx.x * 8;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int s = 0; s < ncb; ++s) {
      int t = p.prev_frame[(size_t)b * ncb + s];
      t = t < 0 ? 0 : (t >= V ? V - 1 : t);
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.m.audio_emb + ((size_t)t + (size_t)s * V) * d + ch));
      acc[0] += bf16lo(q.x); acc[1] += bf16hi(q.x); acc[2] += bf16lo(q.y); acc[3] += bf16hi(q.y);
      acc[4] += bf16lo(q.z); acc[5] += bf16hi(q.z); acc[6] += bf16lo(q.w); acc[7] += bf16hi(q.w);
    }
    float4* xo = reinterpret_cast<float4*>(p.x + (size_t)b * d + ch);
    xo[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    xo[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += acc[e] * acc[e];
    const float scale = rsqrtf(bp_block_sum(ss, red) / (float)d + p.m.backbone.eps);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(w + ch)), g1 = __ldg(reinterpret_cast<const float4*>(w + ch + 4));
    store_split4(p.hi + (size_t)b * d + ch, p.lo + (size_t)b * d + ch, acc[0] * scale * g0.x, acc[1] * scale * g0.y,
                 acc[2] * scale * g0.z, acc[3] * scale * g0.w);
    store_split4(p.hi + (size_t)b * d + ch + 4, p.lo + (size_t)b * d + ch + 4, acc[4] * scale * g1.x, acc[5] * scale * g1.y,
                 acc[6] * scale * g1.z, acc[7] * scale * g1.w);
  }
}

// residual update (mode 1: x = sum(part), 2: x += sum(part)) + RMSNorm -> hi/lo rows [rows][d]; input row = i*row_mul+row_add
template <int NV>
__device__ void bp_norm(BpCtx& c, float* x, const BpPart& part, int mode, const float* w, float eps, float* y32, int rows,
                        int row_mul, int row_add, float* red) {
  const BpParams& p = *c.p;
  constexpr int d = NV * 1024;
  for (int rout = blockIdx.x; rout < rows; rout += gridDim.x) {
    const int rin = rout * row_mul + row_add;
    float4 v[NV];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int ch = threadIdx.x * 4 + j * 1024;
      float* xp = x + (size_t)rin * d + ch;
      float4 s = bp_sum4(part, (size_t)rin * part.ld + ch);
      if (mode == 2) {
        const float4 o = __ldcg(reinterpret_cast<const float4*>(xp));
        s = make_float4(o.x + s.x, o.y + s.y, o.z + s.z, o.w + s.w);
      }
      *reinterpret_cast<float4*>(xp) = s;
      v[j] = s;
      ss += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;
    }
    const float scale = rsqrtf(bp_block_sum(ss, red) / (float)d + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int ch = threadIdx.x * 4 + j * 1024;
      const float4 g = __ldg(reinterpret_cast<const float4*>(w + ch));
      const float4 y = make_float4(v[j].x * scale * g.x, v[j].y * scale * g.y, v[j].z * scale * g.z, v[j].w * scale * g.w);
      store_split4(p.hi + (size_t)rout * d + ch, p.lo + (size_t)rout * d + ch, y.x, y.y, y.z, y.w);
      if (y32) *reinterpret_cast<float4*>(y32 + (size_t)rout * d + ch) = y;
    }
  }
}

// RoPE + KV append + GQA attention for item (sequence b, kv head): see k_attn_decode_fused in batch_frame.cu.  Each half
// of the CTA (4 warps = the 4 query heads of the group) takes one item at a time.
template <int HD>
__device__ void bp_attn(BpCtx& c, const csmb_llama& L, const BpPart& qkv, float* pool, const int32_t* block_table,
                        int max_pages, const int32_t* pos_arr, int pos0, int rps) {
  const BpParams& p = *c.p;
  const int H = L.n_heads, Hkv = L.n_kv_heads, G = H / Hkv;  // G == 4
  const int half_id = threadIdx.x >> 7, ht = threadIdx.x & 127, hw = ht >> 5, lane = c.lane;
  const int max_pos = max_pages * CSMB_PAGE;
  constexpr int half = HD / 2;
  float* sbase = c.scratch + (size_t)half_id * (p.attn_floats / 2);
  float* sq = sbase + (size_t)hw * (HD + max_pos);
  float* sc = sq + HD;
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD;
  const float scale = rsqrtf((float)HD);
  const float* rope = L.rope;
  for (int item = blockIdx.x * 2 + half_id; item < p.B * Hkv; item += gridDim.x * 2) {
    const int b = item / Hkv, kvh = item % Hkv, h = kvh * G + hw;
    const size_t head_off = (size_t)kvh * CSMB_PAGE * HD;
    for (int i = 0; i < rps; ++i) {
      const int r = b * rps + i;
      const int pos = (pos_arr ? pos_arr[b] : pos0) + i;
      const size_t roff = (size_t)r * qkv.ld;
      const float* rc = rope + (size_t)pos * half * 2;
      {
        const int lp = pos / CSMB_PAGE;
        const int page = block_table ? block_table[(size_t)b * max_pages + lp] : b * max_pages + lp;
        float* kdst = pool + (size_t)page * page_stride + head_off + (size_t)(pos % CSMB_PAGE) * HD;
        float* vdst = kdst + (size_t)Hkv * CSMB_PAGE * HD;
        for (int pr = ht; pr < half; pr += 128) {
          const float2 k = bp_sum2(qkv, roff + (size_t)(H + kvh) * HD + 2 * pr);
          const float2 cs = __ldg(reinterpret_cast<const float2*>(rc + 2 * pr));
          *reinterpret_cast<float2*>(kdst + 2 * pr) = make_float2(k.x * cs.x - k.y * cs.y, k.y * cs.x + k.x * cs.y);
        }
        for (int ch = ht * 2; ch < HD; ch += 256)
          *reinterpret_cast<float2*>(vdst + ch) = bp_sum2(qkv, roff + (size_t)(H + Hkv + kvh) * HD + ch);
      }
      for (int pr = lane; pr < half; pr += 32) {
        const float2 q = bp_sum2(qkv, roff + (size_t)h * HD + 2 * pr);
        const float2 cs = __ldg(reinterpret_cast<const float2*>(rc + 2 * pr));
        sq[2 * pr] = q.x * cs.x - q.y * cs.y;
        sq[2 * pr + 1] = q.y * cs.x + q.x * cs.y;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + half_id) : "memory");  // this item's k/v rows are stored, sq complete
      const int S = pos + 1;
      const int32_t* bt = block_table ? block_table + (size_t)b * max_pages : nullptr;
      float m = -INFINITY;
      for (int j = lane; j < S; j += 32) {
        const int page = bt ? bt[j / CSMB_PAGE] : b * max_pages + j / CSMB_PAGE;
        const float* kp = pool + (size_t)page * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD;
        float dot = 0.f;
#pragma unroll
        for (int ch = 0; ch < HD; ch += 4) {
          const float4 kv = __ldcg(reinterpret_cast<const float4*>(kp + ch));
          dot = fmaf(kv.x, sq[ch], dot);
          dot = fmaf(kv.y, sq[ch + 1], dot);
          dot = fmaf(kv.z, sq[ch + 2], dot);
          dot = fmaf(kv.w, sq[ch + 3], dot);
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
      for (int ii = 0; ii < PER; ++ii) acc[ii] = 0.f;
      // keys in groups of 8: the 8 x PER loads of a group are independent and in flight together
      for (int j0 = 0; j0 < S; j0 += 8) {
        float vv[8][PER];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + u < S ? j0 + u : S - 1;
          const int page = bt ? bt[j / CSMB_PAGE] : b * max_pages + j / CSMB_PAGE;
          const float* vp = pool + (size_t)page * page_stride + (size_t)Hkv * CSMB_PAGE * HD + head_off + (size_t)(j % CSMB_PAGE) * HD;
#pragma unroll
          for (int ii = 0; ii < PER; ++ii) vv[u][ii] = __ldcg(vp + lane + 32 * ii);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float pj = j0 + u < S ? sc[j0 + u] : 0.f;
#pragma unroll
          for (int ii = 0; ii < PER; ++ii) acc[ii] = fmaf(pj, vv[u][ii], acc[ii]);
        }
      }
      const size_t o = (size_t)r * H * HD + (size_t)h * HD;
#pragma unroll
      for (int ii = 0; ii < PER; ++ii) {
        uint16_t hh, ll;
        split_bf16(acc[ii] * inv, hh, ll);
        p.hi[o + lane + 32 * ii] = hh;
        p.lo[o + lane + 32 * ii] = ll;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + half_id) : "memory");  // sq / sc are reused
    }
  }
}

__device__ void bp_swiglu(BpCtx& c, const BpPart& gu, int R, int F) {
  const BpParams& p = *c.p;
  const size_t total4 = (size_t)R * F / 4;
  for (size_t i4 = (size_t)blockIdx.x * BP_THREADS + threadIdx.x; i4 < total4; i4 += (size_t)gridDim.x * BP_THREADS) {
    const size_t i = i4 * 4, r = i / F, f = i % F;
    const float4 g = bp_sum4(gu, r * gu.ld + f), u = bp_sum4(gu, r * gu.ld + F + f);
    store_split4(p.hi + i, p.lo + i, (g.x / (1.f + expf(-g.x))) * u.x, (g.y / (1.f + expf(-g.y))) * u.y,
                 (g.z / (1.f + expf(-g.z))) * u.z, (g.w / (1.f + expf(-g.w))) * u.w);
  }
}

// logits -> token -> frame[b][cb]; next decoder input row(s): see k_sample_embed in batch_frame.cu
__device__ void bp_sample_embed(BpCtx& c, const BpPart& lg, int cb, int embed, int out_mul, float* red_v, int* red_i) {
  const BpParams& p = *c.p;
  const int V = p.m.audio_vocab, d = p.m.backbone.d_model, ncb = p.m.n_codebooks;
  float* sl = c.scratch;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    for (int i = threadIdx.x; i < V; i += BP_THREADS) sl[i] = bp_sum1(lg, (size_t)b * lg.ld + i);
    __syncthreads();
    int tok;
    if (p.sa.inv_temp == 0.f) {
      tok = block_argmax(V, [&](int i) { return sl[i]; }, red_v, red_i);
    } else {
      const unsigned long long draw = p.sa.draw_base + (unsigned long long)cb + (unsigned long long)p.pos[b] * p.sa.draw_pos_mul;
      const uint32_t dlo = (uint32_t)draw, dhi = (uint32_t)(draw >> 32);
      tok = block_argmax(
          V, [&](int i) { return sl[i] * p.sa.inv_temp + gumbel_for(i, dlo, dhi, (uint32_t)b, p.sa.seed_lo, p.sa.seed_hi); },
          red_v, red_i);
    }
    if (threadIdx.x == 0) p.frame[(size_t)b * ncb + cb] = tok;
    if (embed) {
      const int t = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
      const uint16_t* src = p.m.audio_emb + ((size_t)t + (size_t)cb * V) * d;
      const size_t re = (size_t)(b * out_mul + out_mul - 1) * d;
      for (int ch = threadIdx.x * 8; ch < d; ch += BP_THREADS * 8) {
        *reinterpret_cast<uint4*>(p.hi + re + ch) = __ldg(reinterpret_cast<const uint4*>(src + ch));
        *reinterpret_cast<uint4*>(p.lo + re + ch) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (out_mul == 2) {
        const size_t rh = (size_t)(b * 2) * d;
        for (int ch = threadIdx.x * 4; ch < d; ch += BP_THREADS * 4) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(p.h_last + (size_t)b * d + ch));
          store_split4(p.hi + rh + ch, p.lo + rh + ch, v.x, v.y, v.z, v.w);
        }
      }
    }
    __syncthreads();  // sl is reused by the next row
  }
}

// ---------------------------------------------------------------------------------------------- the program
struct BpStack {
  const csmb_llama* L;
  int wbase;         // map index of layer 0's qkv
  int xk_d, xk_f;    // activation-plane kinds for K = d_model and K = d_ff
  float* x;
  float* pool;
  unsigned long long lstride;
  const int32_t* block_table;
  int max_pages;
  const int32_t* pos_arr;
  int pos0, rps;
  float* y32_final;
};

// one Llama stack; w.hi/lo already hold the first normalised input.  `after` = the Linear that follows the stack.
__device__ void bp_layers(BpCtx& c, const BpStack& s, const BpGemm& after, float* red) {
  const BpParams& p = *c.p;
  const csmb_llama& L = *s.L;
  const int d = L.d_model, F = L.d_ff, R = p.B * s.rps;
  const int nqkv = (L.n_heads + 2 * L.n_kv_heads) * L.head_dim;
  const int xd = bp_xmap(p.m, s.xk_d), xf = bp_xmap(p.m, s.xk_f);
  for (int l = 0; l < L.n_layers; ++l) {
    const BpGemm gq{s.wbase + l * 4 + 0, xd, R, nqkv, d}, go{s.wbase + l * 4 + 1, xd, R, d, d};
    const BpGemm gg{s.wbase + l * 4 + 2, xd, R, 2 * F, d}, gd{s.wbase + l * 4 + 3, xf, R, d, F};
    const BpGemm gn = (l + 1 < L.n_layers) ? BpGemm{s.wbase + (l + 1) * 4, xd, R, nqkv, d} : after;
    bp_gemm(c, gq);
    bp_mark(c, BT_GEMM);
    bp_prefetch(c, go);
    bp_mark(c, BT_PREFETCH);
    bp_grid_sync(c);
    if (L.head_dim == 64)
      bp_attn<64>(c, L, bp_part(p, gq), s.pool + (size_t)l * s.lstride, s.block_table, s.max_pages, s.pos_arr, s.pos0, s.rps);
    else
      bp_attn<128>(c, L, bp_part(p, gq), s.pool + (size_t)l * s.lstride, s.block_table, s.max_pages, s.pos_arr, s.pos0, s.rps);
    bp_mark(c, BT_ATTN);
    bp_grid_sync(c);
    bp_gemm(c, go);
    bp_mark(c, BT_GEMM);
    bp_prefetch(c, gg);
    bp_mark(c, BT_PREFETCH);
    bp_grid_sync(c);
    if (d == 1024) bp_norm<1>(c, s.x, bp_part(p, go), 2, L.norm_post[l], L.eps, nullptr, R, 1, 0, red);
    else bp_norm<2>(c, s.x, bp_part(p, go), 2, L.norm_post[l], L.eps, nullptr, R, 1, 0, red);
    bp_mark(c, BT_NORM);
    bp_grid_sync(c);
    bp_gemm(c, gg);
    bp_mark(c, BT_GEMM);
    bp_prefetch(c, gd);
    bp_mark(c, BT_PREFETCH);
    bp_grid_sync(c);
    bp_swiglu(c, bp_part(p, gg), R, F);
    bp_mark(c, BT_SWIGLU);
    bp_grid_sync(c);
    bp_gemm(c, gd);
    bp_mark(c, BT_GEMM);
    bp_prefetch(c, gn);
    bp_mark(c, BT_PREFETCH);
    bp_grid_sync(c);
    const bool last = l + 1 == L.n_layers;
    const float* nw = last ? L.norm_final : L.norm_in[l + 1];
    const int rows = last ? p.B : R, mul = last ? s.rps : 1, add = last ? s.rps - 1 : 0;
    if (d == 1024) bp_norm<1>(c, s.x, bp_part(p, gd), 2, nw, L.eps, last ? s.y32_final : nullptr, rows, mul, add, red);
    else bp_norm<2>(c, s.x, bp_part(p, gd), 2, nw, L.eps, last ? s.y32_final : nullptr, rows, mul, add, red);
    bp_mark(c, BT_NORM);
    bp_grid_sync(c);
  }
}

__global__ void __launch_bounds__(BP_THREADS, 1) k_frame_batch(const __grid_constant__ BpParams p) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * BP_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[8];
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  BpCtx c;
  c.p = &p;
  c.ring = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  c.scratch = reinterpret_cast<float*>(c.ring + (size_t)p.nstages * p.stage_stride);
  c.full = bars;
  c.empty = bars + BP_MAX_STAGES;
  c.acc_full = bars + 2 * BP_MAX_STAGES;
  c.q_prod = c.q_mma = 0;
  c.acc_par = 0;
  c.pre = 0;
  c.dead = false;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.bar_next = *reinterpret_cast<volatile unsigned long long*>(p.bar + 1) + gridDim.x;
  for (int i = 0; i < 12; ++i) c.t_acc[i] = 0;
  c.t_last = p.prof != nullptr ? (unsigned long long)clock64() : 0ull;
  uint32_t ncols = 32;
  while ((int)ncols < p.RN) ncols <<= 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < BP_MAX_STAGES; ++i) {
      tc_mbar_init(&c.full[i], 1);
      tc_mbar_init(&c.empty[i], 1);
    }
    tc_mbar_init(c.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (c.warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  c.tmem = tmem_base_s;

  const csmb_llama &Bk = p.m.backbone, &D = p.m.decoder;
  const int B = p.B, db = Bk.d_model, dd = D.d_model, V = p.m.audio_vocab, ncb = p.m.n_codebooks;
  const int nqkv_b = (Bk.n_heads + 2 * Bk.n_kv_heads) * Bk.head_dim, nqkv_d = (D.n_heads + 2 * D.n_kv_heads) * D.head_dim;
  const int x_db = bp_xmap(p.m, XK_DB), x_dd = bp_xmap(p.m, XK_DD);
  const int dec_pages = (ncb + CSMB_PAGE - 1) / CSMB_PAGE;
  const BpGemm g_c0{bp_wmap_c0(p.m), x_db, B, V, db};
  const BpGemm g_none{-1, 0, 0, 0, 0};

  // ---- backbone step (generation.py:34-42 with T = 1)
  bp_prefetch(c, BpGemm{bp_wmap_backbone(0, 0), x_db, B, nqkv_b, db});
  bp_embed_norm(c, Bk.norm_in[0], red);
  bp_mark(c, BT_EMBED);
  bp_grid_sync(c);
  {
    const BpStack s{&Bk, bp_wmap_backbone(0, 0), XK_DB, XK_FB, p.x, p.kv_pool, p.kv_layer_stride, p.block_table, p.max_pages,
                    p.pos, 0, 1, p.h_last};
    bp_layers(c, s, g_c0, red);
  }
  bp_gemm(c, g_c0);
  bp_mark(c, BT_GEMM);
  bp_prefetch(c, BpGemm{bp_wmap_proj(p.m), x_db, 2 * B, dd, db});
  bp_mark(c, BT_PREFETCH);
  bp_grid_sync(c);
  bp_sample_embed(c, bp_part(p, g_c0), 0, 1, 2, red_v, red_i);
  bp_mark(c, BT_SAMPLE);
  bp_grid_sync(c);
  // ---- depth decoder (generation.py:56-90)
  for (int i = 1; i < ncb; ++i) {
    const int rps = (i == 1) ? 2 : 1, R = B * rps;
    const BpGemm g_proj{bp_wmap_proj(p.m), x_db, R, dd, db};
    const BpGemm g_head{bp_wmap_head(p.m, i - 1), x_dd, B, V, dd};
    bp_gemm(c, g_proj);
    bp_mark(c, BT_GEMM);
    bp_prefetch(c, BpGemm{bp_wmap_decoder(p.m, 0, 0), x_dd, R, nqkv_d, dd});
    bp_mark(c, BT_PREFETCH);
    bp_grid_sync(c);
    if (dd == 1024) bp_norm<1>(c, p.dx, bp_part(p, g_proj), 1, D.norm_in[0], D.eps, nullptr, R, 1, 0, red);
    else bp_norm<2>(c, p.dx, bp_part(p, g_proj), 1, D.norm_in[0], D.eps, nullptr, R, 1, 0, red);
    bp_mark(c, BT_NORM);
    bp_grid_sync(c);
    {
      const BpStack s{&D, bp_wmap_decoder(p.m, 0, 0), XK_DD, XK_FD, p.dx, p.dec_kv_pool, p.dec_kv_layer_stride, nullptr,
                      dec_pages, nullptr, i == 1 ? 0 : i, rps, nullptr};
      bp_layers(c, s, g_head, red);
    }
    bp_gemm(c, g_head);
    bp_mark(c, BT_GEMM);
    bp_prefetch(c, i + 1 < ncb ? BpGemm{bp_wmap_proj(p.m), x_db, B, dd, db} : g_none);
    bp_mark(c, BT_PREFETCH);
    bp_grid_sync(c);
    bp_sample_embed(c, bp_part(p, g_head), i, i + 1 < ncb ? 1 : 0, 1, red_v, red_i);
    bp_mark(c, BT_SAMPLE);
    if (i + 1 < ncb) bp_grid_sync(c);
  }
  // every CTA has read bar[1] before the first barrier completed; CTA 0 publishes the next launch's base
  if (blockIdx.x == 0 && threadIdx.x == 0) p.bar[1] = c.bar_next - gridDim.x;
  if (p.prof != nullptr && (threadIdx.x & 127) == 0)
    for (int i = 0; i < 12; ++i) p.prof[((size_t)blockIdx.x * 2 + (threadIdx.x >> 7)) * 12 + i] = c.t_acc[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (c.warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(c.tmem), "r"(ncols) : "memory");
  }
}

#endif  // __CUDACC__

// ---------------------------------------------------------------------------------------------- host side
struct BpWs {
  unsigned long long* bar;
  int* err;
  float *x, *dx, *h_last, *part;
  uint16_t *hi, *lo;
  size_t bytes;
};

static int bp_rn(int B) { return ((B + 15) / 16) * 16; }

static BpWs bp_carve(const csmb_model& m, int B, int sms, void* base) {
  const csmb_llama &b = m.backbone, &d = m.decoder;
  const int qkv_b = (b.n_heads + 2 * b.n_kv_heads) * b.head_dim, qkv_d = (d.n_heads + 2 * d.n_kv_heads) * d.head_dim;
  const size_t R2 = (size_t)2 * B;
  size_t kmax = (size_t)(b.d_ff > d.d_ff ? b.d_ff : d.d_ff);
  kmax = kmax > (size_t)b.d_model ? kmax : (size_t)b.d_model;
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return r;
  };
  BpWs w;
  w.bar = (unsigned long long*)take(256);
  w.err = (int*)take(256);
  w.x = (float*)take((size_t)B * b.d_model * 4);
  w.dx = (float*)take(R2 * d.d_model * 4);
  w.h_last = (float*)take((size_t)B * b.d_model * 4);
  w.hi = (uint16_t*)take(R2 * kmax * 2);
  w.lo = (uint16_t*)take(R2 * kmax * 2);
  const int RN = bp_rn(B);
  size_t pf = 0;
  auto need = [&](int R, int N, int K) {
    // most generous split the knobs allow: 1 K block per CTA, every SM
    const size_t f = (size_t)bp_split(bp_tiles(R, N, RN), K / TC_BK, 1, sms) * R * N;
    pf = f > pf ? f : pf;
  };
  need(B, qkv_b, b.d_model); need(B, b.d_model, b.n_heads * b.head_dim); need(B, 2 * b.d_ff, b.d_model); need(B, b.d_model, b.d_ff);
  need(B, m.audio_vocab, b.d_model); need(B, m.audio_vocab, d.d_model);
  for (int R : {B, 2 * B}) {
    need(R, d.d_model, b.d_model); need(R, qkv_d, d.d_model); need(R, d.d_model, d.n_heads * d.head_dim);
    need(R, 2 * d.d_ff, d.d_model); need(R, d.d_model, d.d_ff);
  }
  w.part = (float*)take(pf * 4);
  w.bytes = off;
  return w;
}

static int g_bp_min_kblocks = 4;
static unsigned long long* g_bp_prof = nullptr;

static bool bp_supported(const csmb_model& m, const csmb_sampler& s, int B) {
  const csmb_llama &b = m.backbone, &d = m.decoder;
  auto llama_ok = [](const csmb_llama& L) {
    return (L.d_model == 1024 || L.d_model == 2048) && (L.head_dim == 64 || L.head_dim == 128) && L.n_kv_heads > 0 &&
           L.n_heads == 4 * L.n_kv_heads && L.d_ff % 64 == 0 && L.n_heads * L.head_dim == L.d_model && L.n_layers >= 1;
  };
  const bool filtered = s.temperature != 0.f && ((s.top_k > 0 && s.top_k < m.audio_vocab) || (s.top_p > 0.f && s.top_p < 1.f) || s.min_p > 0.f);
  const int nmaps = (b.n_layers + d.n_layers) * 4 + 2 + (m.n_codebooks - 1) + 8;
  return llama_ok(b) && llama_ok(d) && b.d_model == 2048 && m.audio_vocab <= 4096 && m.n_codebooks >= 2 && !filtered &&
         s.temperature >= 0.f && B >= 1 && B <= 256 && nmaps <= BP_MAXMAPS;
}

}  // namespace csmb

using namespace csmb;

extern "C" {

void csmb_debug_set_frame_batch(int min_kblocks) {
  if (min_kblocks >= 1) g_bp_min_kblocks = min_kblocks;
}
void csmb_debug_set_frame_batch_prof(unsigned long long* device_buf) { g_bp_prof = device_buf; }

size_t csmb_frame_batch_workspace_bytes(const csmb_model* m, int batch, int device) {
  if (!m || batch <= 0) return 0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 384;
  return bp_carve(*m, batch, sms, nullptr).bytes;
}

int csmb_frame_batch_supported(const csmb_model* m, const csmb_sampler* sampler, int batch) {
  return (m && sampler && bp_supported(*m, *sampler, batch)) ? 1 : 0;
}

int csmb_frame_batch(const csmb_model* m, const csmb_batch* bt, const int32_t* prev_frame, const int32_t* pos,
                     int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, void* workspace,
                     size_t workspace_bytes, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(m && bt && prev_frame && pos && frame && sampler && workspace && bt->batch > 0);
  CSMB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  const int B = bt->batch;
  if (!bp_supported(*m, *sampler, B)) return CSMB_ERR_UNSUPPORTED;
  int sms = 0;
  CSMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const BpWs w = bp_carve(*m, B, sms, workspace);
  CSMB_REQUIRE(workspace_bytes >= w.bytes);
  const csmb_llama &Bk = m->backbone, &D = m->decoder;

  static thread_local BpParams p;  // ~18 KB: keep it off the stack of foreign callers
  p.m = *m;
  p.B = B;
  p.RN = bp_rn(B);
  p.max_pages = bt->max_pages;
  p.min_kblocks = g_bp_min_kblocks;
  p.max_ctas = sms;
  p.kv_pool = bt->kv_pool;
  p.kv_layer_stride = bt->kv_layer_stride;
  p.block_table = bt->block_table;
  p.dec_kv_pool = bt->dec_kv_pool;
  p.dec_kv_layer_stride = bt->dec_kv_layer_stride;
  p.prev_frame = prev_frame;
  p.pos = pos;
  p.frame = frame;
  p.x = w.x; p.dx = w.dx; p.h_last = w.h_last; p.part = w.part; p.hi = w.hi; p.lo = w.lo;
  p.bar = w.bar;
  p.err = w.err;
  p.prof = g_bp_prof;
  p.sa.inv_temp = sampler->temperature == 0.f ? 0.f : 1.f / sampler->temperature;
  p.sa.seed_lo = (uint32_t)sampler->seed;
  p.sa.seed_hi = (uint32_t)(sampler->seed >> 32);
  p.sa.draw_pos_mul = (uint32_t)m->n_codebooks;
  p.sa.draw_base = draw_base;

  // shared memory: ring of [W 128x64 | Xhi RNx64 | Xlo RNx64] stages + scratch (attention scores / logits)
  const int dec_pages = cdiv(m->n_codebooks, CSMB_PAGE);
  const int pos_b = bt->max_pages * CSMB_PAGE, pos_d = dec_pages * CSMB_PAGE;
  const size_t attn_b = (size_t)2 * 4 * (Bk.head_dim + pos_b), attn_d = (size_t)2 * 4 * (D.head_dim + pos_d);
  size_t scratch_floats = attn_b > attn_d ? attn_b : attn_d;
  p.attn_floats = (int)scratch_floats;
  scratch_floats = scratch_floats > (size_t)m->audio_vocab ? scratch_floats : (size_t)m->audio_vocab;
  const size_t stage = ((size_t)TC_BM * TC_BK * 2 + 2 * (size_t)p.RN * TC_BK * 2 + 1023) & ~(size_t)1023;
  const size_t budget = 226 * 1024;  // dynamic + ~0.4 KB static per CTA (227 KB limit)
  CSMB_REQUIRE(scratch_floats * 4 + 1024 + 2 * stage <= budget);
  int nstages = (int)((budget - scratch_floats * 4 - 1024) / stage);
  nstages = nstages > BP_MAX_STAGES ? BP_MAX_STAGES : nstages;
  p.nstages = nstages;
  p.stage_stride = (int)stage;
  const size_t smem = stage * nstages + 1024 + scratch_floats * 4;

  // tensor maps
  bool ok = true;
  auto wmap = [&](int idx, const uint16_t* W, int N, int K) { ok = ok && tc_make_map(&p.maps[idx], W, N, K, TC_BM); };
  auto stack_maps = [&](const csmb_llama& L, int base) {
    const int nqkv = (L.n_heads + 2 * L.n_kv_heads) * L.head_dim;
    for (int l = 0; l < L.n_layers; ++l) {
      wmap(base + l * 4 + 0, L.wqkv[l], nqkv, L.d_model);
      wmap(base + l * 4 + 1, L.wo[l], L.d_model, L.n_heads * L.head_dim);
      wmap(base + l * 4 + 2, L.wgu[l], 2 * L.d_ff, L.d_model);
      wmap(base + l * 4 + 3, L.wdown[l], L.d_model, L.d_ff);
    }
  };
  stack_maps(Bk, bp_wmap_backbone(0, 0));
  stack_maps(D, bp_wmap_decoder(*m, 0, 0));
  wmap(bp_wmap_c0(*m), m->c0_head, m->audio_vocab, Bk.d_model);
  wmap(bp_wmap_proj(*m), m->projection, D.d_model, Bk.d_model);
  for (int i = 0; i + 1 < m->n_codebooks; ++i)
    wmap(bp_wmap_head(*m, i), m->audio_head_t + (size_t)i * m->audio_vocab * D.d_model, m->audio_vocab, D.d_model);
  const int xk[4] = {Bk.d_model, Bk.d_ff, D.d_model, D.d_ff};
  for (int k = 0; k < 4; ++k) {
    ok = ok && tc_make_map(&p.maps[bp_xmap(*m, k)], w.hi, 2 * B, xk[k], p.RN);
    ok = ok && tc_make_map(&p.maps[bp_xmap(*m, k) + 1], w.lo, 2 * B, xk[k], p.RN);
  }
  if (!ok) return CSMB_ERR_UNSUPPORTED;

  cudaStream_t st = (cudaStream_t)stream;
  CSMB_CUDA(cudaFuncSetAttribute(k_frame_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&p};
  CSMB_CUDA(cudaLaunchCooperativeKernel((void*)k_frame_batch, dim3(sms), dim3(BP_THREADS), args, smem, st));
  count_launch();
  return CSMB_OK;
}

}  // extern "C"
