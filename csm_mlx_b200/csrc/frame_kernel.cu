// Persistent whole-frame kernel for the batch-1 latency path (sm_100a).
//
// One cooperative launch computes one 80 ms frame of generate_frame (csm_mlx/generation.py:21-92, T=1) for ONE
// sequence: embedding sum of the previous frame, 16-layer backbone step over the paged KV cache, codebook0 head +
// sample, then the 31-step depth-decoder loop with per-step head, sampling and next-embedding gather — without
// returning to the host.  ~640 dependent GEMV phases per frame make per-kernel launches latency-bound; here
//   * one PRODUCER warp per CTA walks the statically known weight schedule (its row slice of every matrix, in
//     order) and streams it HBM -> shared memory with cp.async.bulk into a 12 x 16 KiB ring guarded by mbarriers;
//     it never waits for activations, so HBM stays busy across phase boundaries (ring = 28 MB chip-wide);
//   * eight CONSUMER warps per CTA wait at a grid-wide barrier only for the activation vector of the phase, keep
//     their K-slice of it in registers and reduce 1024-weight units (bf16 -> fp32 by bit shift, fp32 FMA) out of
//     the ring; rows are split evenly across the 148 CTAs so every matrix is read exactly once per use;
//   * activations (fp32) travel between CTAs through L2 (ld.global.cg), the depth decoder's per-frame KV state
//     (256 KiB) stays L2-resident, attention of the small decoder is recomputed by every CTA, the backbone's
//     attention is split over (kv-head, key-chunk) work items and merged by the consumers of the o-projection.
// Algorithmic bytes per launch: 9.107 GB of bf16 weights; nothing is read twice.
//
// All spin loops are bounded: on timeout a sticky abort flag is raised, every wait falls through and the host
// reports an error instead of hanging the GPU.
#include <cooperative_groups.h>
#include <math.h>

#include "ops.cuh"

namespace csmb {

constexpr int NCW = 8;                  // consumer warps
constexpr int NTHREADS = (NCW + 1) * 32;
constexpr int STAGE_BYTES = 16384;
constexpr int NSTAGES = 8;
constexpr int UNIT = 1024;              // weights per (warp, stage) unit
constexpr int MAXU = 128;               // max units per range per CTA (csm_1b: <= 112)
constexpr int MAX_SPLIT = 16;           // backbone attention chunks (128 keys each) per head
constexpr int PSTRIDE = 68;             // floats per attention partial: acc[64], m, l, pad (16-byte aligned rows)
constexpr int KROW = 512;               // floats per staged decoder-KV position: K (2 x 128) then V (2 x 128)
constexpr int KVS_BYTES = 32 * KROW * 4;  // shared-memory staging of one decoder layer's K/V
constexpr unsigned SPIN_LIMIT = 1u << 22;     // ~1-2 s of polling before a wait gives up

struct FrameParams {
  csmb_model m;
  // sequence state
  float* kv_pool;
  unsigned long long kv_layer_stride;
  const int32_t* block_table;  // this sequence's row
  float* dec_kv;               // [Ld][32 pos][2][Hkv_d*hd_d]
  const int32_t* prev_frame;   // [ncb]
  const int32_t* pos_ptr;      // position of this frame's backbone row
  int32_t* frame_out;          // [ncb]
  // scratch (global, L2-resident)
  float *xa, *xb;              // backbone residual stream ping-pong [d_b]
  float* qkv;                  // [2][max qkv]
  float* attn_part;            // [H_b][MAX_SPLIT][PSTRIDE]
  float* attn_out;             // [H_b*hd_b] merged attention output (single-chunk fast path)
  float* act;                  // [2][d_ff]
  float* h_last;               // [d_b]
  float* logits;               // [V]
  float *dxa, *dxb;            // decoder residual stream ping-pong [2][d_d]
  int pf_max;                  // producer: L2 prefetch distance in 16 KiB stages (0 = off)
  int pf_interval;             // producer: SM cycles between L2 prefetches (paces them at the HBM fair share)
  int dbg;                     // debug switches (0 in production): 1 = skip GEMV math, 2 = skip decoder attention math
  unsigned long long* prof;    // optional [gridDim][16] phase timers in ns (debug); null in production
  unsigned int* bar;           // grid barrier counter (zeroed by the host before launch)
  int* abort_flag;
  // sampling
  float inv_temp;
  uint32_t seed_lo, seed_hi;
  unsigned long long draw_base;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
// weights are read once per use: stream them through L2 with an evict-first policy so that they do not push the
// small hot set (activations, norm weights, RoPE rows, decoder KV) out of L2
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Activation loads.  Cross-CTA data is always consumed after a grid barrier whose acquire fence (fence.acq_rel.gpu
// by thread 0, then bar.sync) invalidates this SM's L1, so ordinary weak loads are both legal under the PTX memory
// model and fresh; ld.global.cg compiles to LDG.STRONG.GPU, which was measured to cost several L2 round trips per
// batch of eight (the "load" phase took 2.5-4.6k cycles).
#ifndef CSMB_FRAME_STRONG_LOADS
__device__ __forceinline__ float4 ldcg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
#else
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
#endif

struct Ring {
  unsigned char* data;
  uint64_t* full;
  uint64_t* empty;
};

struct Ctx {
  const FrameParams* p;
  Ring ring;
  float* part;     // smem [2 ranges][2 rows][MAXU][4]
  float* sred;     // smem scratch [NCW*4]
  float* sattn;    // smem scratch for attention
  float* kvs;      // smem staging of one decoder layer's cached K/V
  uint64_t* kvbar; // mbarrier of that staging copy
  uint32_t kv_phase;
  int G, cta, warp, lane, tid;
  uint32_t q;      // ring stage sequence number (same sequence in producer and consumers)
  uint32_t epoch;  // grid barrier epoch
  bool aborted;
  volatile int* s_abort;  // smem: thread 0 publishes the abort state to its CTA at every grid barrier
  uint64_t policy;  // L2 evict-first policy of the weight stream (producer)
  unsigned long long t_acc[12];
  unsigned long long t_last;
  int ntrace;
};

__device__ __forceinline__ unsigned long long gtime() { return (unsigned long long)clock64(); }  // SM cycles
// phase timers: thread 0 of every CTA accumulates the time since the previous mark into category `cat`
enum { T_SYNC = 0, T_LOAD = 1, T_GEMV = 2, T_FIN = 3, T_DATT = 4, T_BATT = 5, T_SAMPLE = 6, T_MERGE = 7, T_WAIT = 8, T_ARRIVE = 9 };
__device__ __forceinline__ void mark(Ctx& c, int cat) {
  if (c.p->prof != nullptr && c.tid == 0) {
    const unsigned long long t = gtime();
    c.t_acc[cat] += t - c.t_last;
    c.t_last = t;
  }
}

// raw event trace (debug bit 4): thread 0 of CTA 1 appends (id, clock) pairs after the per-CTA timer block
__device__ __forceinline__ void trace(Ctx& c, int id) {
  if ((c.p->dbg & 4) && c.p->prof != nullptr && c.tid == 0 && c.cta == 1 && c.ntrace < 6000) {
    unsigned long long* t = c.p->prof + 148 * 16 + 2 * (size_t)c.ntrace;
    t[0] = (unsigned long long)id;
    t[1] = (unsigned long long)clock64();
    c.ntrace++;
  }
}

__device__ __forceinline__ bool check_abort(Ctx& c) {
  if (!c.aborted && *reinterpret_cast<volatile int*>(c.p->abort_flag) != 0) c.aborted = true;
  return c.aborted;
}
__device__ __forceinline__ void raise_abort(Ctx& c, int code) {
  atomicCAS(c.p->abort_flag, 0, code);
  c.aborted = true;
}

__device__ __forceinline__ void mbar_wait(Ctx& c, uint64_t* b, uint32_t parity, int code) {
  if (c.aborted) return;
  unsigned spins = 0;
  while (!mbar_try_wait(b, parity)) {
    if (++spins > (SPIN_LIMIT >> 2)) {
      raise_abort(c, code);
      return;
    }
    if ((spins & 1023) == 0 && check_abort(c)) return;
  }
}

// Grid-wide barrier among the consumer warps of all CTAs: one monotonically increasing counter (zeroed by the host
// before the launch), one arriving/polling thread per CTA.  (A per-CTA flag array polled by every CTA was measured
// 3x slower: 148 x 160 loads per poll round serialise on a handful of L2 lines.)
__device__ __forceinline__ void grid_sync(Ctx& c) {
  trace(c, 10);
  named_bar_sync(1, NCW * 32);
  trace(c, 11);
  c.epoch++;
  if (c.tid == 0) {
    red_release_add(c.p->bar, 1u);
    trace(c, 12);
    mark(c, T_ARRIVE);
    const unsigned target = c.epoch * (unsigned)c.G;
    unsigned spins = 0;
    while (!c.aborted && ld_relaxed(c.p->bar) < target) {
      if (++spins > SPIN_LIMIT) raise_abort(c, 100 + (int)(c.epoch & 0xffff));
      if ((spins & 255) == 0) check_abort(c);
    }
    trace(c, 13);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    trace(c, 14);
    if (c.aborted) *c.s_abort = 1;
  }
  named_bar_sync(1, NCW * 32);
  trace(c, 15);
  if (*c.s_abort) c.aborted = true;
  mark(c, T_SYNC);
}

// ------------------------------------------------------------------------------------------------ row partition
struct Range {
  const uint16_t* p;
  int row0, rows, K;
};
__device__ __forceinline__ Range cta_range(const uint16_t* W, int N, int K, int cta, int G) {
  const int r0 = (int)(((unsigned)N * (unsigned)cta) / (unsigned)G), r1 = (int)(((unsigned)N * (unsigned)(cta + 1)) / (unsigned)G);
  return Range{W + (size_t)r0 * K, r0, r1 - r0, K};
}
__device__ __forceinline__ int n_stages(const Range& r) {
  return (int)(((size_t)r.rows * r.K * 2 + STAGE_BYTES - 1) / STAGE_BYTES);
}

// consumer: partial dot products of R activation rows with every unit of the range.  Two ring stages are processed
// per iteration so that the two units' load -> FMA -> shuffle chains overlap; the lane reduction stops after three
// shuffle steps and leaves 4 partials per unit in shared memory: part[(r*MAXU + u)*4 + 0..3], summed by the finaliser.
template <int R>
__device__ __forceinline__ void unit_dot(const unsigned char* base, const float (&xr)[R][32], float (&acc)[R]) {
#pragma unroll
  for (int i = 0; i < R; ++i) acc[i] = 0.f;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint4 w = *reinterpret_cast<const uint4*>(base + ch * 512);
    const float wf[8] = {bf16lo(w.x), bf16hi(w.x), bf16lo(w.y), bf16hi(w.y),
                         bf16lo(w.z), bf16hi(w.z), bf16lo(w.w), bf16hi(w.w)};
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i] = fmaf(wf[e], xr[i][ch * 8 + e], acc[i]);
  }
}
template <int R>
__device__ void consume(Ctx& c, const Range& r, const float (&xr)[R][32], float* part) {
  if (c.p->dbg & 8) return;  // timing experiment: no streaming at all (pure latency chain)
  const int KS = r.K / UNIT;
  const int units = r.rows * KS;
  const int nst = n_stages(r);
  for (int s = 0; s < nst; s += 2) {
    const bool two = s + 1 < nst;
    const int slot0 = c.q % NSTAGES, slot1 = (c.q + 1) % NSTAGES;
    const uint32_t par0 = (c.q / NSTAGES) & 1, par1 = ((c.q + 1) / NSTAGES) & 1;
    mark(c, T_GEMV);
    mbar_wait(c, &c.ring.full[slot0], par0, 2);
    if (two) mbar_wait(c, &c.ring.full[slot1], par1, 2);
    mark(c, T_WAIT);
    const int u0 = s * NCW + c.warp, u1 = u0 + NCW;
    const bool run = !c.aborted && !(c.p->dbg & 1);
    const bool do0 = u0 < units && run, do1 = two && u1 < units && run;
    float a0[R], a1[R];
#pragma unroll
    for (int i = 0; i < R; ++i) a0[i] = a1[i] = 0.f;
    const size_t woff = (size_t)c.warp * (UNIT * 2) + c.lane * 16;
    if (do0) unit_dot<R>(c.ring.data + (size_t)slot0 * STAGE_BYTES + woff, xr, a0);
    if (do1) unit_dot<R>(c.ring.data + (size_t)slot1 * STAGE_BYTES + woff, xr, a1);
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1)
#pragma unroll
      for (int i = 0; i < R; ++i) {
        if (do0) a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], o);
        if (do1) a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
      }
    if (c.lane < 4) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        if (do0) part[((size_t)i * MAXU + u0) * 4 + c.lane] = a0[i];
        if (do1) part[((size_t)i * MAXU + u1) * 4 + c.lane] = a1[i];
      }
    }
    __syncwarp();
    if (c.lane == 0) {
      mbar_arrive(&c.ring.empty[slot0]);
      if (two) mbar_arrive(&c.ring.empty[slot1]);
    }
    c.q += two ? 2 : 1;
  }
}

// consumer-side sync (shared-memory results visible to all consumer warps)
__device__ __forceinline__ void csync() { named_bar_sync(2, NCW * 32); }

// load the warp's K-slice (kseg = warp % KS) of R rows of a global fp32 vector into registers
template <int R>
__device__ __forceinline__ void load_slice(Ctx& c, const float* x, int ldx, int K, float (&xr)[R][32]) {
  const int kseg = c.warp % (K / UNIT);
  float4 v[R][8];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float* p = x + (size_t)i * ldx + kseg * UNIT + ch * 256 + c.lane * 8;
      v[i][ch * 2] = ldcg4(p);
      v[i][ch * 2 + 1] = ldcg4(p + 4);
    }
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xr[i][j * 4 + 0] = v[i][j].x; xr[i][j * 4 + 1] = v[i][j].y; xr[i][j * 4 + 2] = v[i][j].z; xr[i][j * 4 + 3] = v[i][j].w;
    }
}

// RMSNorm-ed K-slices of R rows: every load (norm weights, the slice, and for K = 2048 the other half that only
// feeds the statistics) is issued before the first dependent instruction, so the phase pays ONE L2 round trip.
struct NormW {
  float4 g[8];
};
// this warp's slice of a norm weight vector; issued BEFORE the grid barrier of the previous phase so that the (possibly
// HBM-cold) load is off the critical path
template <int KS>
__device__ __forceinline__ NormW prefetch_norm(Ctx& c, const float* w) {
  NormW n;
  const int kseg = c.warp % KS;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const float* wp = w + kseg * UNIT + ch * 256 + c.lane * 8;
    n.g[ch * 2] = __ldg(reinterpret_cast<const float4*>(wp));
    n.g[ch * 2 + 1] = __ldg(reinterpret_cast<const float4*>(wp + 4));
  }
  return n;
}
template <int R, int KS>
__device__ __forceinline__ void load_slice_norm(Ctx& c, const float* x, int ldx, int K, const NormW& nw, float eps,
                                                float (&xr)[R][32]) {
  const int kseg = c.warp % KS;
  float4 v[R][8], o[R][KS == 2 ? 8 : 1];
  const float4(&g)[8] = nw.g;
#pragma unroll
  for (int i = 0; i < R; ++i) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float* p = x + (size_t)i * ldx + kseg * UNIT + ch * 256 + c.lane * 8;
      v[i][ch * 2] = ldcg4(p);
      v[i][ch * 2 + 1] = ldcg4(p + 4);
    }
    if (KS == 2) {
      const float* q = x + (size_t)i * ldx + (kseg ^ 1) * UNIT + c.lane * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[i][j] = ldcg4(q + j * 128);
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += v[i][j].x * v[i][j].x + v[i][j].y * v[i][j].y + v[i][j].z * v[i][j].z + v[i][j].w * v[i][j].w;
    if (KS == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += o[i][j].x * o[i][j].x + o[i][j].y * o[i][j].y + o[i][j].z * o[i][j].z + o[i][j].w * o[i][j].w;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)K + eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xr[i][j * 4 + 0] = v[i][j].x * rstd * g[j].x; xr[i][j * 4 + 1] = v[i][j].y * rstd * g[j].y;
      xr[i][j * 4 + 2] = v[i][j].z * rstd * g[j].z; xr[i][j * 4 + 3] = v[i][j].w * rstd * g[j].w;
    }
  }
}

// sum the KS partials of local row j (fixed order -> deterministic)
__device__ __forceinline__ float row_total(const float* part, int j, int KS) {
  float s = 0.f;
  for (int k = 0; k < KS * 4; ++k) s += part[(size_t)j * KS * 4 + k];
  return s;
}

// ------------------------------------------------------------------------------------------------ sampling
__device__ __forceinline__ float gumbel_at(int idx, uint32_t dlo, uint32_t dhi, uint32_t k0, uint32_t k1) {
  uint32_t ctr[4] = {(uint32_t)(idx >> 2), dlo, dhi, 0u};
  philox4x32_10(ctr, k0, k1);
  return -logf(-logf(u01(ctr[idx & 3])));
}
// every CTA computes the same token from the logits in global memory (consumer warps only)
__device__ int sample_token(Ctx& c, const float* logits, int V, unsigned long long draw) {
  const FrameParams& p = *c.p;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const uint32_t dlo = (uint32_t)draw, dhi = (uint32_t)(draw >> 32);
  for (int i0 = c.tid; i0 < V; i0 += NCW * 32 * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * NCW * 32;
      v[j] = i < V ? __ldcg(logits + i) : -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * NCW * 32;
      if (i < V) {
        float t = v[j];
        if (p.inv_temp != 0.f) t = t * p.inv_temp + gumbel_at(i, dlo, dhi, p.seed_lo, p.seed_hi);
        argmax_combine(bv, bi, t, i);
      }
    }
  }
  warp_argmax(bv, bi);
  int* si = reinterpret_cast<int*>(c.sred + NCW);
  if (c.lane == 0) {
    c.sred[c.warp] = bv;
    si[c.warp] = bi;
  }
  csync();
  bv = c.sred[0];
  bi = si[0];
#pragma unroll
  for (int w = 1; w < NCW; ++w) argmax_combine(bv, bi, c.sred[w], si[w]);
  csync();
  mark(c, T_SAMPLE);
  return bi == 0x7fffffff ? 0 : bi;
}

// ------------------------------------------------------------------------------------------------ GEMV phases
// y[row] = (res ? res[row] : 0) + W[row,:] . x      for this CTA's rows; x given as register slices
template <int R>
__device__ void phase_linear(Ctx& c, const uint16_t* W, int N, int K, const float (&xr)[R][32], float* y, int ldy,
                             const float* res, int ldr) {
  const Range r = cta_range(W, N, K, c.cta, c.G);
  // this thread finalises output j = tid (rows*R <= 256 always: <= 21 rows per CTA x 2): fetch its residual early
  const int j = c.tid;
  const bool mine = j < r.rows * R;
  const int ri = mine ? j / r.rows : 0, rrow = mine ? j % r.rows : 0;
  float rv = 0.f;
  if (mine && res) rv = __ldcg(res + (size_t)ri * ldr + r.row0 + rrow);
  mark(c, T_LOAD);
  trace(c, 20);
  consume<R>(c, r, xr, c.part);
  trace(c, 21);
  mark(c, T_GEMV);
  csync();
  trace(c, 22);
  if (mine) y[(size_t)ri * ldy + r.row0 + rrow] = rv + row_total(c.part + (size_t)ri * MAXU * 4, rrow, K / UNIT);
  mark(c, T_FIN);
}

// SwiGLU MLP first half: act[f] = silu(Wg[f,:].x) * (Wu[f,:].x)
template <int R>
__device__ void phase_gate_up(Ctx& c, const uint16_t* Wgu, int F, int K, const float (&xr)[R][32], float* act) {
  const Range rg = cta_range(Wgu, F, K, c.cta, c.G);
  const Range ru = cta_range(Wgu + (size_t)F * K, F, K, c.cta, c.G);
  float* pg = c.part;
  float* pu = c.part + 2 * MAXU * 4;
  mark(c, T_LOAD);
  trace(c, 30);
  consume<R>(c, rg, xr, pg);
  consume<R>(c, ru, xr, pu);
  trace(c, 31);
  mark(c, T_GEMV);
  csync();
  trace(c, 32);
  const int KS = K / UNIT;
  for (int j = c.tid; j < rg.rows * R; j += NCW * 32) {
    const int i = j / rg.rows, row = j % rg.rows;
    const float g = row_total(pg + (size_t)i * MAXU * 4, row, KS), u = row_total(pu + (size_t)i * MAXU * 4, row, KS);
    act[(size_t)i * F + rg.row0 + row] = (g / (1.f + expf(-g))) * u;
  }
  mark(c, T_FIN);
}

// ------------------------------------------------------------------------------------------------ attention
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stage the cached K/V rows (positions < pos0) of one decoder layer into shared memory with ONE bulk copy (the rows
// are contiguous in dec_kv).  Issued right after a grid barrier, two phases before the attention that reads them.
__device__ __forceinline__ void bulk_g2s_plain(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void stage_decoder_kv(Ctx& c, int layer, int pos0) {
  if (pos0 == 0 || c.tid != 0 || c.aborted) return;
  const FrameParams& p = *c.p;
  const float* src = p.dec_kv + (size_t)layer * 32 * KROW;
  const uint32_t bytes = (uint32_t)pos0 * KROW * 4;
  asm volatile("fence.proxy.async;" ::: "memory");  // K/V rows were written with ordinary stores (by CTA 0)
  mbar_arrive_expect_tx(c.kvbar, bytes);
  bulk_g2s_plain(c.kvs, src, bytes, c.kvbar);
}

// Depth decoder attention, recomputed by every CTA: warp h = query head h (8 heads x 128), kv head h/4.
// qkv: this step's raw projections for R rows (positions pos0 .. pos0+R-1); cs: their RoPE (cos,sin) pairs for this
// lane (dims 4*lane .. 4*lane+3).  Keys < pos0 come from the staged shared-memory copy: QK is lane-per-key, PV is
// dims-over-lanes.  Result: sattn[R][1024] (every warp of the o-projection needs the whole vector).
template <int R>
__device__ void decoder_attention(Ctx& c, int layer, int pos0, const float* qkv, int ldq, const float4 (&cs)[R]) {
  const FrameParams& p = *c.p;
  constexpr int HD = 128, H = 8, HKV = 2, ROW = 2 * HKV * HD;
  const int h = c.warp, kvh = h / (H / HKV);
  float* kvl = p.dec_kv + (size_t)layer * 32 * ROW;
  float* sq = c.sattn + 2 * 1024 + c.warp * (2 * HD);  // this warp's rotated q rows, broadcast-read below
  const float scale = rsqrtf((float)HD);
  float4 q[R], k[R], v[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float* row = qkv + (size_t)i * ldq;
    q[i] = ldcg4(row + h * HD + c.lane * 4);
    k[i] = ldcg4(row + (H + kvh) * HD + c.lane * 4);
    v[i] = ldcg4(row + (H + HKV + kvh) * HD + c.lane * 4);
  }
  float qn[R][4], kn[R][4];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    qn[i][0] = q[i].x * cs[i].x - q[i].y * cs[i].y; qn[i][1] = q[i].y * cs[i].x + q[i].x * cs[i].y;
    qn[i][2] = q[i].z * cs[i].z - q[i].w * cs[i].w; qn[i][3] = q[i].w * cs[i].z + q[i].z * cs[i].w;
    kn[i][0] = k[i].x * cs[i].x - k[i].y * cs[i].y; kn[i][1] = k[i].y * cs[i].x + k[i].x * cs[i].y;
    kn[i][2] = k[i].z * cs[i].z - k[i].w * cs[i].w; kn[i][3] = k[i].w * cs[i].z + k[i].z * cs[i].w;
    *reinterpret_cast<float4*>(sq + i * HD + c.lane * 4) = make_float4(qn[i][0], qn[i][1], qn[i][2], qn[i][3]);
    if (c.cta == 0 && (h % (H / HKV)) == 0) {  // one writer per kv head persists K/V for later steps
      float* dst = kvl + (size_t)(pos0 + i) * ROW + kvh * HD + c.lane * 4;
      *reinterpret_cast<float4*>(dst) = make_float4(kn[i][0], kn[i][1], kn[i][2], kn[i][3]);
      *reinterpret_cast<float4*>(dst + HKV * HD) = v[i];
    }
  }
  if (pos0 > 0) {
    mbar_wait(c, c.kvbar, c.kv_phase, 4);
    c.kv_phase ^= 1;
  }
  csync();  // every warp's sq visible
  const float* krow = c.kvs + (size_t)c.lane * KROW + kvh * HD;   // lane-per-key
  const float* vcol = c.kvs + HKV * HD + kvh * HD + c.lane * 4;   // dims-over-lanes
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int pos = pos0 + i;
    float dot = 0.f;
    if (c.lane < pos0) {
      // lane j starts at 16-byte chunk j and wraps: rows are 2 KB apart (same bank), the rotation makes the
      // quarter-warps hit 8 distinct bank groups, and the per-lane q reads cover 512 contiguous bytes
#pragma unroll 8
      for (int t = 0; t < HD / 4; ++t) {
        const int d = ((t + c.lane) & (HD / 4 - 1)) * 4;
        const float4 kv = *reinterpret_cast<const float4*>(krow + d);
        const float4 qv = *reinterpret_cast<const float4*>(sq + i * HD + d);
        dot = fmaf(kv.x, qv.x, dot);
        dot = fmaf(kv.y, qv.y, dot);
        dot = fmaf(kv.z, qv.z, dot);
        dot = fmaf(kv.w, qv.w, dot);
      }
    }
    const float s_c = (c.lane < pos0) ? dot * scale : -INFINITY;
    float s_n[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float d = qn[i][0] * kn[j][0] + qn[i][1] * kn[j][1] + qn[i][2] * kn[j][2] + qn[i][3] * kn[j][3];
      d = warp_sum(d) * scale;
      s_n[j] = (pos0 + j <= pos) ? d : -INFINITY;
    }
    float m = warp_max(s_c);
#pragma unroll
    for (int j = 0; j < R; ++j) m = fmaxf(m, s_n[j]);
    const float e_c = (c.lane < pos0) ? expf(s_c - m) : 0.f;
    float sum = warp_sum(e_c);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int j = 0; j < pos0; ++j) {
      const float pj = __shfl_sync(0xffffffffu, e_c, j);
      const float4 vv = *reinterpret_cast<const float4*>(vcol + (size_t)j * KROW);
      o[0] = fmaf(pj, vv.x, o[0]); o[1] = fmaf(pj, vv.y, o[1]); o[2] = fmaf(pj, vv.z, o[2]); o[3] = fmaf(pj, vv.w, o[3]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const float e = (pos0 + j <= pos) ? expf(s_n[j] - m) : 0.f;
      sum += e;
      o[0] = fmaf(e, v[j].x, o[0]); o[1] = fmaf(e, v[j].y, o[1]); o[2] = fmaf(e, v[j].z, o[2]); o[3] = fmaf(e, v[j].w, o[3]);
    }
    const float inv = 1.f / sum;
    *reinterpret_cast<float4*>(c.sattn + i * 1024 + h * HD + c.lane * 4) =
        make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
  }
  csync();
}

// Backbone attention work item: kv head `kvh`, chunk `chunk` of 128 keys of S = pos+1.  8 warps = 4 query heads x
// 2 halves of 64 keys (2 keys per lane, lane-per-key for both QK and PV, reduce-scatter for the PV sum).  The two
// halves are merged in shared memory; the item writes either the final normalised head output (single chunk) or one
// partial (acc[64], m, l) per head to attn_part.
__device__ void backbone_attention_item(Ctx& c, int layer, int kvh, int chunk, int nchunks, int pos, const float* qkv) {
  const FrameParams& p = *c.p;
  constexpr int HD = 64, H = 32, HKV = 8;
  const int hq = kvh * (H / HKV) + (c.warp & 3), half = c.warp >> 2;
  const int S = pos + 1;
  const int kbeg = chunk * 128 + half * 64;
  float* pool = p.kv_pool + (size_t)layer * p.kv_layer_stride;
  const size_t page_stride = (size_t)2 * HKV * CSMB_PAGE * HD;
  const float* rope = p.m.backbone.rope + (size_t)pos * (HD / 2) * 2;
  float* sq = c.sattn + c.warp * HD;             // rotated q of this warp's head
  float* sm = c.sattn + NCW * HD + c.warp * PSTRIDE;  // this warp's partial for the intra-CTA merge
  {
    const float2 q = __ldcg(reinterpret_cast<const float2*>(qkv + hq * HD + c.lane * 2));
    const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + c.lane * 2));
    *reinterpret_cast<float2*>(sq + c.lane * 2) = make_float2(q.x * cs.x - q.y * cs.y, q.y * cs.x + q.x * cs.y);
  }
  const float2 kraw = __ldcg(reinterpret_cast<const float2*>(qkv + (H + kvh) * HD + c.lane * 2));
  const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + c.lane * 2));
  const float2 knew = make_float2(kraw.x * cs.x - kraw.y * cs.y, kraw.y * cs.x + kraw.x * cs.y);
  const float2 vnew = __ldcg(reinterpret_cast<const float2*>(qkv + (H + HKV + kvh) * HD + c.lane * 2));
  const bool has_new = (pos >= kbeg && pos < kbeg + 64);
  if (has_new && (c.warp & 3) == 0) {  // one writer per kv head: append to the paged cache
    const int page = p.block_table[pos / CSMB_PAGE];
    float* kd = pool + (size_t)page * page_stride + (size_t)kvh * CSMB_PAGE * HD + (size_t)(pos % CSMB_PAGE) * HD;
    *reinterpret_cast<float2*>(kd + c.lane * 2) = knew;
    *reinterpret_cast<float2*>(kd + (size_t)HKV * CSMB_PAGE * HD + c.lane * 2) = vnew;
  }
  __syncwarp();
  const float scale = rsqrtf((float)HD);
  // cached keys of this lane: j0 = kbeg + lane, j1 = kbeg + 32 + lane (valid if < pos)
  const float* kptr[2];
  bool valid[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = kbeg + t * 32 + c.lane;
    valid[t] = j < pos;
    const int jj = valid[t] ? j : 0;
    kptr[t] = pool + (size_t)p.block_table[jj / CSMB_PAGE] * page_stride + (size_t)kvh * CSMB_PAGE * HD +
              (size_t)(jj % CSMB_PAGE) * HD;
  }
  float sc[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    float4 kv[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) kv[d] = valid[t] ? ldcg4(kptr[t] + d * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float dot = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      dot = fmaf(kv[d].x, sq[d * 4], dot);
      dot = fmaf(kv[d].y, sq[d * 4 + 1], dot);
      dot = fmaf(kv[d].z, sq[d * 4 + 2], dot);
      dot = fmaf(kv[d].w, sq[d * 4 + 3], dot);
    }
    sc[t] = valid[t] ? dot * scale : -INFINITY;
  }
  float snew = -INFINITY;
  if (has_new) snew = warp_sum(sq[c.lane * 2] * knew.x + sq[c.lane * 2 + 1] * knew.y) * scale;
  const float m = fmaxf(warp_max(fmaxf(sc[0], sc[1])), snew);
  float e[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) e[t] = valid[t] ? expf(sc[t] - m) : 0.f;
  float sum = warp_sum(e[0] + e[1]);
  // PV, lane-per-key: acc[d] = e0 * V[j0][d] + e1 * V[j1][d], then reduce-scatter over the 32 lanes
  float acc[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) acc[d] = 0.f;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float* vp = kptr[t] + (size_t)HKV * CSMB_PAGE * HD;
    float4 vv[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) vv[d] = valid[t] ? ldcg4(vp + d * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      acc[d * 4] = fmaf(e[t], vv[d].x, acc[d * 4]);
      acc[d * 4 + 1] = fmaf(e[t], vv[d].y, acc[d * 4 + 1]);
      acc[d * 4 + 2] = fmaf(e[t], vv[d].z, acc[d * 4 + 2]);
      acc[d * 4 + 3] = fmaf(e[t], vv[d].w, acc[d * 4 + 3]);
    }
  }
  // reduce-scatter: after the five steps lane L holds the full sums of dims 2L and 2L+1 in acc[0], acc[1]
#pragma unroll
  for (int o = 16, n = 32; o > 0; o >>= 1, n >>= 1) {
    const bool up = (c.lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? acc[i] : acc[n + i];
      const float keep = up ? acc[n + i] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float2 out = make_float2(acc[0], acc[1]);
  if (has_new) {
    const float en = expf(snew - m);
    sum += en;
    out.x = fmaf(en, vnew.x, out.x);
    out.y = fmaf(en, vnew.y, out.y);
  }
  *reinterpret_cast<float2*>(sm + c.lane * 2) = out;
  if (c.lane == 0) {
    sm[HD] = m;      // -inf if this half saw no key
    sm[HD + 1] = sum;
  }
  csync();
  if (half == 0) {
    const float* s0 = sm;
    const float* s1 = sm + 4 * PSTRIDE;
    const float m0 = s0[HD], m1 = s1[HD];
    const float M = fmaxf(m0, m1);
    const float w0 = (m0 == -INFINITY) ? 0.f : expf(m0 - M), w1 = (m1 == -INFINITY) ? 0.f : expf(m1 - M);
    const float L = w0 * s0[HD + 1] + w1 * s1[HD + 1];
    const float2 a0 = *reinterpret_cast<const float2*>(s0 + c.lane * 2), a1 = *reinterpret_cast<const float2*>(s1 + c.lane * 2);
    float2 o2 = make_float2(w0 * a0.x + w1 * a1.x, w0 * a0.y + w1 * a1.y);
    if (nchunks == 1) {
      const float inv = 1.f / L;
      *reinterpret_cast<float2*>(p.attn_out + hq * HD + c.lane * 2) = make_float2(o2.x * inv, o2.y * inv);
    } else {
      float* dst = p.attn_part + ((size_t)hq * MAX_SPLIT + chunk) * PSTRIDE;
      *reinterpret_cast<float2*>(dst + c.lane * 2) = o2;
      if (c.lane == 0) {
        dst[HD] = M;
        dst[HD + 1] = L;
      }
    }
  }
  csync();
}

// merge of the per-chunk partials into the o-projection's register slice (K = 2048 -> 2 ksegs; a lane's 8
// consecutive k share one head).  Loads are issued in independent batches, two head-chunks at a time.
__device__ void merged_attention_slice(Ctx& c, int nchunks, float (&xr)[1][32]) {
  const FrameParams& p = *c.p;
  constexpr int HD = 64;
  const int kseg = c.warp % 2;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int k = kseg * UNIT + ch * 256 + c.lane * 8;
    const int h = k / HD, d = k % HD;
    const float* base = p.attn_part + (size_t)h * MAX_SPLIT * PSTRIDE;
    float2 ml[MAX_SPLIT];
#pragma unroll
    for (int s = 0; s < MAX_SPLIT; ++s)
      ml[s] = s < nchunks ? __ldcg(reinterpret_cast<const float2*>(base + s * PSTRIDE + HD)) : make_float2(-INFINITY, 0.f);
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < MAX_SPLIT; ++s) M = fmaxf(M, ml[s].x);
    float L = 0.f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
    for (int s0 = 0; s0 < MAX_SPLIT; s0 += 4) {
      if (s0 >= nchunks) break;
      float4 a[4][2];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const bool ok = s0 + s < nchunks;
        a[s][0] = ok ? ldcg4(base + (s0 + s) * PSTRIDE + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        a[s][1] = ok ? ldcg4(base + (s0 + s) * PSTRIDE + d + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float wgt = (ml[s0 + s].x == -INFINITY) ? 0.f : expf(ml[s0 + s].x - M);
        L = fmaf(wgt, ml[s0 + s].y, L);
        o[0] = fmaf(wgt, a[s][0].x, o[0]); o[1] = fmaf(wgt, a[s][0].y, o[1]);
        o[2] = fmaf(wgt, a[s][0].z, o[2]); o[3] = fmaf(wgt, a[s][0].w, o[3]);
        o[4] = fmaf(wgt, a[s][1].x, o[4]); o[5] = fmaf(wgt, a[s][1].y, o[5]);
        o[6] = fmaf(wgt, a[s][1].z, o[6]); o[7] = fmaf(wgt, a[s][1].w, o[7]);
      }
    }
    const float inv = 1.f / L;
#pragma unroll
    for (int e = 0; e < 8; ++e) xr[0][ch * 8 + e] = o[e] * inv;
  }
}

// ------------------------------------------------------------------------------------------------ schedule
// The weight schedule: range number idx of the frame, in exactly the order in which the consumers call consume().
//   backbone layer l: 5l + {0 qkv, 1 o, 2 gate, 3 up, 4 down};  then c0 head;  then per depth step i = 1..ncb-1:
//   projection, 4 x {qkv, o, gate, up, down}, audio_head[i-1].
__device__ bool sched_range(const FrameParams& p, int idx, int cta, int G, Range& out) {
  const csmb_llama &B = p.m.backbone, &D = p.m.decoder;
  const int db = B.d_model, dd = D.d_model, V = p.m.audio_vocab;
  auto layer_range = [&](const csmb_llama& L, int l, int k, int d) {
    const int nqkv = (L.n_heads + 2 * L.n_kv_heads) * L.head_dim;
    switch (k) {
      case 0: return cta_range(L.wqkv[l], nqkv, d, cta, G);
      case 1: return cta_range(L.wo[l], d, d, cta, G);
      case 2: return cta_range(L.wgu[l], L.d_ff, d, cta, G);
      case 3: return cta_range(L.wgu[l] + (size_t)L.d_ff * d, L.d_ff, d, cta, G);
      default: return cta_range(L.wdown[l], d, L.d_ff, cta, G);
    }
  };
  const int nb = B.n_layers * 5;
  if (idx < nb) {
    out = layer_range(B, idx / 5, idx % 5, db);
    return true;
  }
  if (idx == nb) {
    out = cta_range(p.m.c0_head, V, db, cta, G);
    return true;
  }
  const int per_step = 2 + D.n_layers * 5;
  const int j = idx - nb - 1, step = j / per_step, t = j % per_step;
  if (step >= p.m.n_codebooks - 1) return false;
  if (t == 0) out = cta_range(p.m.projection, dd, db, cta, G);
  else if (t == per_step - 1) out = cta_range(p.m.audio_head_t + (size_t)step * V * dd, V, dd, cta, G);
  else out = layer_range(D, (t - 1) / 5, (t - 1) % 5, dd);
  return true;
}

struct Cursor {
  int idx, stage, nst;
  Range r;
  bool valid;
};
__device__ __forceinline__ void cursor_load(const FrameParams& p, Cursor& k, int cta, int G) {
  k.valid = sched_range(p, k.idx, cta, G, k.r);
  k.stage = 0;
  k.nst = k.valid ? n_stages(k.r) : 0;
}
__device__ __forceinline__ void cursor_chunk(const Cursor& k, const char*& src, uint32_t& bytes) {
  const size_t total = (size_t)k.r.rows * k.r.K * 2, off = (size_t)k.stage * STAGE_BYTES;
  src = reinterpret_cast<const char*>(k.r.p) + off;
  bytes = (uint32_t)((total - off) < (size_t)STAGE_BYTES ? (total - off) : (size_t)STAGE_BYTES);
}
__device__ __forceinline__ void cursor_advance(const FrameParams& p, Cursor& k, int cta, int G) {
  if (++k.stage >= k.nst) {
    k.idx++;
    cursor_load(p, k, cta, G);
  }
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Producer (one thread per CTA).  Two cursors walk the schedule: the RING cursor copies 16 KiB stages into shared
// memory as slots free up; whenever the ring is full (the consumers are in a latency phase: barrier, activation
// load, attention) the PREFETCH cursor runs ahead, pulling future stages HBM -> L2 with cp.async.bulk.prefetch.L2 at
// this SM's fair share of the HBM rate (one stage per pf_interval cycles), at most pf_max stages ahead.  HBM thus
// keeps streaming through the latency phases; the ring later refills from L2.  Every byte still leaves HBM once.
__device__ void producer_main(Ctx& c) {
  const FrameParams& p = *c.p;
  if (p.dbg & 8) return;
  Cursor rc, pc;
  rc.idx = 0;
  cursor_load(p, rc, c.cta, c.G);
  pc = rc;
  int ahead = 0;  // stages the prefetch cursor is ahead of the ring cursor
  const int pf_max = p.pf_max, pf_interval = p.pf_interval;
  long long last_pf = clock64() - pf_interval;
  unsigned idle = 0;
  while (rc.valid && !c.aborted) {
    const int slot = c.q % NSTAGES;
    const uint32_t par = (c.q / NSTAGES) & 1;
    if (mbar_test(&c.ring.empty[slot], par ^ 1)) {
      const char* src;
      uint32_t n;
      cursor_chunk(rc, src, n);
      mbar_arrive_expect_tx(&c.ring.full[slot], n);
      bulk_g2s(c.ring.data + (size_t)slot * STAGE_BYTES, src, n, &c.ring.full[slot], c.policy);
      cursor_advance(p, rc, c.cta, c.G);
      ++c.q;
      if (ahead > 0) --ahead; else pc = rc;
      idle = 0;
    } else if (ahead < pf_max && pc.valid && clock64() - last_pf >= pf_interval) {
      const char* src;
      uint32_t n;
      cursor_chunk(pc, src, n);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(n) : "memory");
      cursor_advance(p, pc, c.cta, c.G);
      ++ahead;
      last_pf = clock64();
      idle = 0;
    } else {
      if (++idle > SPIN_LIMIT * 4u) raise_abort(c, 1);
      if ((idle & 1023) == 0) check_abort(c);
    }
  }
}

// one depth-decoder step with R rows (R = 2 for the first step: positions 0 and 1)
template <int R>
__device__ void decoder_step(Ctx& c, int step, int pos0, float (&xr)[R][32]) {
  const FrameParams& p = *c.p;
  const csmb_llama& D = p.m.decoder;
  const int db = p.m.backbone.d_model, dd = D.d_model, V = p.m.audio_vocab, F = D.d_ff;
  const int nqkv = (D.n_heads + 2 * D.n_kv_heads) * D.head_dim;
  // RoPE rows of this step's positions (same for all layers): loaded now, used two phases later
  float4 cs[R];
#pragma unroll
  for (int i = 0; i < R; ++i)
    cs[i] = __ldg(reinterpret_cast<const float4*>(D.rope + ((size_t)(pos0 + i) * (D.head_dim / 2) + c.lane * 2) * 2));
  // projection: xr already holds the input row slices (K = 2048)
  phase_linear<R>(c, p.m.projection, dd, db, xr, p.dxa, dd, nullptr, 0);
  NormW nw = prefetch_norm<1>(c, D.norm_in[0]);
  grid_sync(c);
  float* x = p.dxa;
  float* x1 = p.dxb;
  for (int l = 0; l < D.n_layers; ++l) {
    trace(c, 40);
    stage_decoder_kv(c, l, pos0);
    trace(c, 41);
    load_slice_norm<R, 1>(c, x, dd, dd, nw, D.eps, xr);
    phase_linear<R>(c, D.wqkv[l], nqkv, dd, xr, p.qkv, nqkv, nullptr, 0);
    grid_sync(c);
    trace(c, 50);
    decoder_attention<R>(c, l, pos0, p.qkv, nqkv, cs);
    trace(c, 51);
    mark(c, T_DATT);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const float* sp = c.sattn + i * 1024 + ch * 256 + c.lane * 8;
        const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
        xr[i][ch * 8 + 0] = a.x; xr[i][ch * 8 + 1] = a.y; xr[i][ch * 8 + 2] = a.z; xr[i][ch * 8 + 3] = a.w;
        xr[i][ch * 8 + 4] = b.x; xr[i][ch * 8 + 5] = b.y; xr[i][ch * 8 + 6] = b.z; xr[i][ch * 8 + 7] = b.w;
      }
    phase_linear<R>(c, D.wo[l], dd, dd, xr, x1, dd, x, dd);
    nw = prefetch_norm<1>(c, D.norm_post[l]);
    grid_sync(c);
    load_slice_norm<R, 1>(c, x1, dd, dd, nw, D.eps, xr);
    phase_gate_up<R>(c, D.wgu[l], F, dd, xr, p.act);
    grid_sync(c);
    load_slice<R>(c, p.act, F, F, xr);
    phase_linear<R>(c, D.wdown[l], dd, F, xr, x, dd, x1, dd);
    nw = prefetch_norm<1>(c, l + 1 < D.n_layers ? D.norm_in[l + 1] : D.norm_final);
    grid_sync(c);
  }
  // head on the last row
  float xh[1][32];
  const float* last = x + (size_t)(R - 1) * dd;
  load_slice_norm<1, 1>(c, last, dd, dd, nw, D.eps, xh);
  phase_linear<1>(c, p.m.audio_head_t + (size_t)(step - 1) * V * dd, V, dd, xh, p.logits, V, nullptr, 0);
  grid_sync(c);
}

__device__ __forceinline__ void embed_row_slice(Ctx& c, const uint16_t* row, int K, float* dst32) {
  const int kseg = c.warp % (K / UNIT);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(row + kseg * UNIT + ch * 256 + c.lane * 8));
    dst32[ch * 8 + 0] = bf16lo(w.x); dst32[ch * 8 + 1] = bf16hi(w.x);
    dst32[ch * 8 + 2] = bf16lo(w.y); dst32[ch * 8 + 3] = bf16hi(w.y);
    dst32[ch * 8 + 4] = bf16lo(w.z); dst32[ch * 8 + 5] = bf16hi(w.z);
    dst32[ch * 8 + 6] = bf16lo(w.w); dst32[ch * 8 + 7] = bf16hi(w.w);
  }
}

__device__ void consumer_main(Ctx& c) {
  const FrameParams& p = *c.p;
  const csmb_llama& B = p.m.backbone;
  const int db = B.d_model, V = p.m.audio_vocab, ncb = p.m.n_codebooks, F = B.d_ff;
  const int nqkv = (B.n_heads + 2 * B.n_kv_heads) * B.head_dim;
  const int pos = *p.pos_ptr;
  const int S = pos + 1;
  // ---- input embedding: x = sum_k audio_emb[prev[k] + k*V]   (generation.py:156-161 + models.py:82-92)
  // every CTA keeps its own copy in registers (slice) and CTA-distributed rows are not needed: the full vector is
  // written once to xa by CTA 0 for the residual path.
  {
    for (int k = c.tid * 8; k < db && c.cta == 0; k += NCW * 32 * 8) {
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      for (int cb = 0; cb < ncb; ++cb) {
        int t = p.prev_frame[cb];
        t = t < 0 ? 0 : (t >= V ? V - 1 : t);
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(p.m.audio_emb + ((size_t)t + (size_t)cb * V) * db + k));
        acc[0] += bf16lo(w.x); acc[1] += bf16hi(w.x); acc[2] += bf16lo(w.y); acc[3] += bf16hi(w.y);
        acc[4] += bf16lo(w.z); acc[5] += bf16hi(w.z); acc[6] += bf16lo(w.w); acc[7] += bf16hi(w.w);
      }
      if (c.cta == 0) {
        *reinterpret_cast<float4*>(p.xa + k) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(p.xa + k + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
    grid_sync(c);
  }
  float xr[2][32];
  float(&x1r)[1][32] = *reinterpret_cast<float(*)[1][32]>(&xr[0]);
  float* x = p.xa;
  float* x1 = p.xb;
  const int nchunks = (S + 127) / 128;  // <= MAX_SPLIT for S <= 2048
  NormW nw = prefetch_norm<2>(c, B.norm_in[0]);
  for (int l = 0; l < B.n_layers; ++l) {
    load_slice_norm<1, 2>(c, x, db, db, nw, B.eps, x1r);
    phase_linear<1>(c, B.wqkv[l], nqkv, db, x1r, p.qkv, nqkv, nullptr, 0);
    grid_sync(c);
    for (int item = c.cta; item < B.n_kv_heads * nchunks; item += c.G)
      backbone_attention_item(c, l, item % B.n_kv_heads, item / B.n_kv_heads, nchunks, pos, p.qkv);
    mark(c, T_BATT);
    grid_sync(c);
    if (nchunks == 1) load_slice<1>(c, p.attn_out, db, db, x1r);
    else merged_attention_slice(c, nchunks, x1r);
    mark(c, T_MERGE);
    phase_linear<1>(c, B.wo[l], db, db, x1r, x1, db, x, db);
    nw = prefetch_norm<2>(c, B.norm_post[l]);
    grid_sync(c);
    load_slice_norm<1, 2>(c, x1, db, db, nw, B.eps, x1r);
    phase_gate_up<1>(c, B.wgu[l], F, db, x1r, p.act);
    grid_sync(c);
    load_slice<1>(c, p.act, F, F, x1r);
    phase_linear<1>(c, B.wdown[l], db, F, x1r, x, db, x1, db);
    nw = prefetch_norm<2>(c, l + 1 < B.n_layers ? B.norm_in[l + 1] : B.norm_final);
    grid_sync(c);
  }
  // ---- final norm -> h_last (decoder input row 0) ; codebook0 head ; sample c0
  load_slice_norm<1, 2>(c, x, db, db, nw, B.eps, x1r);
  if (c.cta == 0 && c.warp < db / UNIT) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float* d = p.h_last + c.warp * UNIT + ch * 256 + c.lane * 8;
      *reinterpret_cast<float4*>(d) = make_float4(x1r[0][ch * 8], x1r[0][ch * 8 + 1], x1r[0][ch * 8 + 2], x1r[0][ch * 8 + 3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(x1r[0][ch * 8 + 4], x1r[0][ch * 8 + 5], x1r[0][ch * 8 + 6], x1r[0][ch * 8 + 7]);
    }
  }
  // keep the normalised slice: it is row 0 of the decoder's first input
  float hrow[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) hrow[e] = x1r[0][e];
  phase_linear<1>(c, p.m.c0_head, V, db, x1r, p.logits, V, nullptr, 0);
  grid_sync(c);
  const unsigned long long draw0 = p.draw_base + (unsigned long long)pos * (unsigned)ncb;
  int tok = sample_token(c, p.logits, V, draw0);
  if (c.cta == 0 && c.tid == 0) p.frame_out[0] = tok;
  // ---- depth decoder: step 1 has rows (h_last @ pos 0, embed(c0) @ pos 1)   (generation.py:56-90)
#pragma unroll
  for (int e = 0; e < 32; ++e) xr[0][e] = hrow[e];
  embed_row_slice(c, p.m.audio_emb + (size_t)tok * db, db, xr[1]);
  decoder_step<2>(c, 1, 0, xr);
  tok = sample_token(c, p.logits, V, draw0 + 1);
  if (c.cta == 0 && c.tid == 0) p.frame_out[1] = tok;
  for (int i = 2; i < ncb; ++i) {
    embed_row_slice(c, p.m.audio_emb + ((size_t)tok + (size_t)(i - 1) * V) * db, db, x1r[0]);
    decoder_step<1>(c, i, i, x1r);
    tok = sample_token(c, p.logits, V, draw0 + (unsigned)i);
    if (c.cta == 0 && c.tid == 0) p.frame_out[i] = tok;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_frame(const __grid_constant__ FrameParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2 * NSTAGES + 1];
  __shared__ float s_part[2 * 2 * MAXU * 4];  // [range][row][unit][4 partials]
  __shared__ float s_red[NCW * 4];
  __shared__ int s_abort;
  __shared__ __align__(16) float s_attn[2 * 1024 + NCW * 2 * 128];
  Ctx c;
  c.p = &p;
  c.ring.data = smem;
  c.ring.full = bars;
  c.ring.empty = bars + NSTAGES;
  c.kvbar = bars + 2 * NSTAGES;
  c.kv_phase = 0;
  c.part = s_part;
  c.sred = s_red;
  c.sattn = s_attn;
  c.kvs = reinterpret_cast<float*>(smem + (size_t)NSTAGES * STAGE_BYTES);
  c.G = gridDim.x;
  c.cta = blockIdx.x;
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.q = 0;
  c.epoch = 0;
  c.aborted = false;
  c.s_abort = &s_abort;
  c.ntrace = 0;
  for (int i = 0; i < 12; ++i) c.t_acc[i] = 0;
  c.t_last = (p.prof != nullptr) ? gtime() : 0ull;
  const unsigned long long t_begin = c.t_last;
  if (threadIdx.x == 0) {
    s_abort = 0;
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&c.ring.full[i], 1);
      mbar_init(&c.ring.empty[i], NCW);
    }
    mbar_init(c.kvbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (c.warp == NCW) {
    c.policy = make_evict_first_policy();
    if (c.lane == 0) producer_main(c);
  } else {
    consumer_main(c);
    if (p.prof != nullptr && c.tid == 0) {
      for (int i = 0; i < 12; ++i) p.prof[(size_t)blockIdx.x * 16 + i] = c.t_acc[i];
      p.prof[(size_t)blockIdx.x * 16 + 12] = gtime() - t_begin;
      p.prof[(size_t)blockIdx.x * 16 + 13] = c.epoch;
    }
  }
}

}  // namespace csmb

using namespace csmb;

static unsigned long long* g_prof_ptr = nullptr;  // debug only: csmb_debug_set_frame_prof
static int g_dbg_flags = 0;
static int g_pf_max = 0, g_pf_interval = 640;  // L2 prefetch measured slower than off (profiles/r01_frame_kernel_notes.md)

extern "C" {

/* debug: device buffer [n_sms][16] u64 receiving per-CTA phase timers (ns) of later csmb_frame_b1 launches; null = off */
void csmb_debug_set_frame_prof(unsigned long long* device_buf) { g_prof_ptr = device_buf; }
void csmb_debug_set_frame_flags(int flags) { g_dbg_flags = flags & 0xff; }
void csmb_debug_set_frame_prefetch(int max_stages, int interval_cycles) { g_pf_max = max_stages; g_pf_interval = interval_cycles; }

size_t csmb_frame_workspace_bytes(const csmb_model* m) {
  if (!m) return 0;
  const csmb_llama &B = m->backbone, &D = m->decoder;
  const size_t nqkv_b = (size_t)(B.n_heads + 2 * B.n_kv_heads) * B.head_dim, nqkv_d = (size_t)(D.n_heads + 2 * D.n_kv_heads) * D.head_dim;
  const size_t ff = B.d_ff > D.d_ff ? B.d_ff : D.d_ff;
  size_t f = 0;
  f += 2 * (size_t)B.d_model;                                   // xa, xb
  f += 2 * (nqkv_b > nqkv_d ? nqkv_b : nqkv_d);                  // qkv
  f += (size_t)B.n_heads * MAX_SPLIT * PSTRIDE + (size_t)B.d_model;  // attn_part, attn_out
  f += 2 * ff;                                                   // act
  f += (size_t)B.d_model;                                        // h_last
  f += (size_t)m->audio_vocab + 64;                              // logits
  f += 4 * (size_t)D.d_model;                                    // dxa, dxb
  f += (size_t)D.n_layers * 32 * 2 * D.n_kv_heads * D.head_dim;  // dec_kv
  return f * sizeof(float) + 2048 /* barrier flags + abort flag */ + 8192 /* alignment slack */;
}

/* Whole decode frame for ONE sequence in one persistent cooperative kernel (see top of file).
 * block_table: this sequence's row; pos: DEVICE int (position of this frame's backbone row);
 * workspace: csmb_frame_workspace_bytes() bytes.  Only temperature sampling without top-k/top-p/min-p (or greedy)
 * is fused; other sampler settings and other model shapes return CSMB_ERR_UNSUPPORTED (use csmb_decode_frame).
 * status (optional, DEVICE int[1]) receives 0 or a non-zero abort code if an internal wait timed out. */
int csmb_frame_b1(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table,
                  const int32_t* prev_frame, const int32_t* pos, int32_t* frame, const csmb_sampler* sampler,
                  uint64_t draw_base, void* workspace, size_t workspace_bytes, int32_t* status, int device,
                  void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(m && kv_pool && block_table && prev_frame && pos && frame && sampler && workspace);
  const csmb_llama &B = m->backbone, &D = m->decoder;
  const bool shape_ok = B.d_model == 2048 && B.n_heads == 32 && B.n_kv_heads == 8 && B.head_dim == 64 &&
                        B.d_ff % 1024 == 0 && D.d_model == 1024 && D.n_heads == 8 && D.n_kv_heads == 2 &&
                        D.head_dim == 128 && D.d_ff % 1024 == 0 && m->n_codebooks <= 32 && m->n_codebooks >= 3;
  if (!shape_ok) return CSMB_ERR_UNSUPPORTED;
  if (sampler->temperature != 0.f && ((sampler->top_k > 0 && sampler->top_k < m->audio_vocab) ||
                                     (sampler->top_p > 0.f && sampler->top_p < 1.f) || sampler->min_p > 0.f))
    return CSMB_ERR_UNSUPPORTED;
  CSMB_REQUIRE(workspace_bytes >= csmb_frame_workspace_bytes(m));
  cudaStream_t st = (cudaStream_t)stream;

  FrameParams p;
  p.m = *m;
  p.kv_pool = kv_pool;
  p.kv_layer_stride = kv_layer_stride;
  p.block_table = block_table;
  p.prev_frame = prev_frame;
  p.pos_ptr = pos;
  p.frame_out = frame;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.bar = reinterpret_cast<unsigned int*>(base);            // [<= 384] per-CTA epoch flags
  p.abort_flag = reinterpret_cast<int*>(base + 1536);
  float* f = reinterpret_cast<float*>(base + 2048);
  auto take = [&](size_t n) {
    float* r = f;
    f += (n + 63) & ~(size_t)63;
    return r;
  };
  const size_t nqkv_b = (size_t)(B.n_heads + 2 * B.n_kv_heads) * B.head_dim, nqkv_d = (size_t)(D.n_heads + 2 * D.n_kv_heads) * D.head_dim;
  const size_t ff = B.d_ff > D.d_ff ? B.d_ff : D.d_ff;
  p.xa = take(B.d_model);
  p.xb = take(B.d_model);
  p.qkv = take(2 * (nqkv_b > nqkv_d ? nqkv_b : nqkv_d));
  p.attn_part = take((size_t)B.n_heads * MAX_SPLIT * PSTRIDE);
  p.attn_out = take(B.d_model);
  p.act = take(2 * ff);
  p.h_last = take(B.d_model);
  p.logits = take(m->audio_vocab);
  p.dxa = take(2 * (size_t)D.d_model);
  p.dxb = take(2 * (size_t)D.d_model);
  p.dec_kv = take((size_t)D.n_layers * 32 * 2 * D.n_kv_heads * D.head_dim);
  CSMB_REQUIRE(reinterpret_cast<uintptr_t>(f) <= reinterpret_cast<uintptr_t>(workspace) + workspace_bytes);
  p.inv_temp = sampler->temperature == 0.f ? 0.f : 1.f / sampler->temperature;
  p.seed_lo = (uint32_t)sampler->seed;
  p.seed_hi = (uint32_t)(sampler->seed >> 32);
  p.draw_base = draw_base;
  p.prof = g_prof_ptr;
  p.dbg = g_dbg_flags;
  p.pf_max = g_pf_max;
  p.pf_interval = g_pf_interval;

  static const size_t dyn_smem = (size_t)NSTAGES * STAGE_BYTES + KVS_BYTES;  // ring + decoder KV staging
  int sms = 0;
  CSMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CSMB_CUDA(cudaFuncSetAttribute(k_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  CSMB_REQUIRE(sms <= 384);
  CSMB_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(base), 0, 2048, st));
  void* args[] = {&p};
  CSMB_CUDA(cudaLaunchCooperativeKernel((void*)k_frame, dim3(sms), dim3(NTHREADS), args, dyn_smem, st));
  count_launch();
  if (status) CSMB_CUDA(cudaMemcpyAsync(status, p.abort_flag, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return CSMB_OK;
}

}  // extern "C"
