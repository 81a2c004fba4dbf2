// Persistent whole-frame kernel for the batch-1 latency path (sm_100a).
//
// One cooperative launch computes one 80 ms frame of generate_frame (csm_mlx/generation.py:21-92, T=1) for ONE
// sequence: embedding sum of the previous frame, 16-layer backbone step over the paged KV cache, codebook0 head +
// sample, then the 31-step depth-decoder loop with per-step head, sampling and next-embedding gather — without
// returning to the host.  ~700 dependent GEMV phases per frame make per-kernel launches latency-bound; here
//   * one PRODUCER thread per CTA walks the statically known weight schedule (its row slice of every matrix, in
//     order) and streams it HBM -> shared memory with cp.async.bulk (L2 evict-first) into a 4 x 32 KiB ring guarded
//     by mbarriers; it never waits for activations;
//   * eight CONSUMER warps per CTA keep their K-slice of the phase's activation vector in registers and reduce
//     1024-weight units (bf16 -> fp32 by bit shift, fp32 FMA) out of the ring; rows are split evenly across the 148
//     CTAs so every matrix is read from HBM exactly once per use (9.107 GB per frame);
//   * there is NO grid barrier: every cross-CTA activation is stored as an 8-byte {fp32 value, tag} word (the tag
//     names the launch and the phase that produced it, like NCCL's LL protocol) and readers poll the data itself
//     with L1-bypassing loads, so a phase boundary costs one store flight + one L2 round trip instead of
//     release-fence + atomic + poll + acquire-fence (measured 2.1 us per barrier, x640 per frame).  A CTA can run at
//     most one phase ahead of the slowest one — to start phase q it needs every CTA's output of phase q-1 — which is
//     exactly what the buffer reuse distances tolerate (every buffer is rewritten only after its last reading phase);
//   * attention of the small decoder is recomputed by every CTA from its PRIVATE copy of the frame's decoder KV (no
//     cross-CTA dependency); the backbone's attention is split over (kv-head, 128-key chunk) work items.
//
// All spin loops are bounded: on timeout a sticky abort flag is raised, every wait falls through and the host
// reports an error instead of hanging the GPU.
#include <math.h>

#include <atomic>

#include "ops.cuh"

// This file is compiled twice: as is (CSMB_FK_FMT = 0: bf16 weights, namespace csmb::fk_bf16, plus the C entry points) and
// through frame_kernel_e4m3.cu (CSMB_FK_FMT = 1: weight-only FP8 blobs, namespace csmb::fk_e4m3, kernel and launcher only), so
// that the bf16 kernel's code is exactly what it was before the FP8 mode existed (a template parameter changed ptxas'
// inlining and cost the bf16 frame 3 %).
#ifndef CSMB_FK_FMT
#define CSMB_FK_FMT 0
#endif
#if CSMB_FK_FMT == 0
#define FK_NS fk_bf16
#else
#define FK_NS fk_e4m3
#endif

namespace csmb {
namespace FK_NS {

constexpr int FMT = CSMB_FK_FMT;        // CSMB_WEIGHTS_BF16 or CSMB_WEIGHTS_E4M3
constexpr int NCW = 8;                  // consumer warps
constexpr int NCT = NCW * 32;           // consumer threads
constexpr int NTHREADS = (NCW + 1) * 32;
#ifndef CSMB_FRAME_NAP
#define CSMB_FRAME_NAP 40   // ns between probes of a tagged word (0 … 200 ns: within 1 %; two probes in flight: 2-3 % slower)
#endif
#ifndef CSMB_FRAME_GS
#define CSMB_FRAME_GS 2
#endif
#ifndef CSMB_FRAME_STAGE_KB
#define CSMB_FRAME_STAGE_KB 32
#endif
// consume(): ring slots are handed back as soon as the stage's bytes are in registers, before the lane reduction
// (-1.5 % per frame), and the first butterfly step is a transposing exchange (half the shuffles, bit-identical sums;
// -0.3 %).  0 restores the older order for A/B runs (scripts/variant_ab.py).
#ifndef CSMB_FRAME_EARLY_ARRIVE
#define CSMB_FRAME_EARLY_ARRIVE 1
#endif
#ifndef CSMB_FRAME_XPOSE
#define CSMB_FRAME_XPOSE 1
#endif
// ring geometry: 4 x 32 KiB (default; measured 3.30 ms per frame) or 8 x 16 KiB (3.52 ms: twice the mbarrier handshakes
// and bulk copies per byte)
constexpr int STAGE_BYTES = CSMB_FRAME_STAGE_KB * 1024;
constexpr int NSTAGES = 128 * 1024 / STAGE_BYTES;
constexpr int SUB = STAGE_BYTES / 16384;                 // 16 KiB sub-stages (8 units of 1024 weights) per ring stage
constexpr int UNIT = 1024;              // weights per (warp, stage) unit
constexpr int MAXU = 192;               // max units per range per CTA (csm_1b: 112 on 148 CTAs, 176 on 96)
constexpr int MAX_SPLIT = 16;           // backbone attention chunks (128 keys each) per head
constexpr int PSTRIDE = 72;             // words per attention partial: acc[64], m, l, pad (32 heads x 72 = 9 x 256)
constexpr int KROW = 512;               // floats per decoder-KV position: K (2 x 128) then V (2 x 128)
constexpr int KVS_BYTES = 32 * KROW * 4;  // shared-memory staging of one decoder layer's K/V
constexpr unsigned SPIN_LIMIT = 1u << 22;  // ~1-2 s of polling before a wait gives up

struct FrameParams {
  csmb_model m;
  // sequence state
  float* kv_pool;
  unsigned long long kv_layer_stride;
  const int32_t* block_table;  // this sequence's row
  const int32_t* prev_frame;   // [ncb]
  const int32_t* pos_ptr;      // position of this frame's backbone row
  int32_t* frame_out;          // [ncb]
  float* h_last;               // [d_b] (plain; output only)
  const float* h_in;           // depth-only launch: the backbone's normalised last hidden row (prefill output); else null
  float* dec_kv;               // [CTA][Ld][32 pos][KROW]: every CTA keeps its own copy of the frame's decoder KV
  // cross-CTA activations: tagged words {value bits, tag}
  uint2 *xa, *xb;              // backbone residual stream ping-pong [d_b]
  uint2* qkv;                  // [2][max qkv]
  uint2* attn_part;            // [MAX_SPLIT chunks][H_b][PSTRIDE]
  uint2* attn_out;             // [H_b*hd_b] attention output (single-chunk fast path)
  uint2* act;                  // [2][d_ff]
  uint2* logits;               // [V]
  uint2 *dxa, *dxb;            // decoder residual stream ping-pong [2][d_d]
  unsigned* nonce;             // launch counter in the workspace: read by every CTA at start, bumped by CTA 0 at the
                               // end, so consecutive launches (also CUDA-graph replays) never share tags
  int* abort_flag;
  int32_t* status;             // caller's sticky status word (optional): receives the first abort code, never cleared here
  int pf_max, pf_interval;     // producer: L2 prefetch distance (16 KiB stages, 0 = off) and pacing (SM cycles)
  int dbg;                     // debug switches (0 in production): 1 = skip GEMV math, 8 = no weight streaming
  unsigned long long* prof;    // optional [gridDim][16] phase timers in SM cycles (debug); null in production
  // sampling
  float inv_temp;
  int top_k;                   // 0 = off; else keep the k most likely tokens (ties included), like k_sample_filtered
  float min_p;                 // 0 = off; else drop tokens whose probability is below min_p x the best one's
  float top_p;                 // 0 = off; else the nucleus: tokens whose strictly-more-likely mass is below top_p (k_sample_filtered)
  uint32_t seed_lo, seed_hi;
  unsigned long long draw_base;
  uint32_t seq;                // sequence word of the Philox counter (row index of this sequence in its batch; csmb_sample's r)
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
// e4m3 models: the depth decoder (4 layers + projection + one head = 115 MB at one byte per weight) is streamed 31 times per
// frame and nearly fits the 126 MB L2: its copies ask L2 to keep them (evict-last on a fraction of the lines, so that the
// tagged activation words and KV rows still find room), the backbone stays evict-first
__device__ __forceinline__ uint64_t make_keep_policy(float fraction) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(fraction));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- tagged ("LL") words: value and tag travel in one 8-byte (or two in one 16-byte) store/load
__device__ __forceinline__ void ll_st(uint2* p, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_st2(uint2* p, float a, float b, unsigned tag) {  // p 16-byte aligned
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag),
               "r"(__float_as_uint(b)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint2 ll_ld(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ll_ld2(const uint2* p) {  // two consecutive words, p 16-byte aligned
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

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
  float* sattn;    // smem: [0,2048) gathered activation vector / decoder attention output; then attention scratch
  float* kvs;      // smem staging of one decoder layer's cached K/V
  uint64_t* kvbar; // mbarrier of that staging copy
  uint32_t kv_phase;
  int G, cta, warp, lane, tid;
  uint32_t q;      // ring stage sequence number (same sequence in producer and consumers)
  uint32_t phase;  // activation phase counter (same sequence in every CTA); part of the tag
  uint32_t launch; // tag nonce of this launch (1 .. 2^20-1)
  bool aborted;
  bool prof_on;    // this thread records phase timers (debug)
  uint64_t policy;
  unsigned long long t_acc[12];
  unsigned long long t_last;
};

// phase timers (debug): thread 0 of every CTA accumulates the cycles since the previous mark into category `cat`
enum { T_POLL = 0, T_LOAD = 1, T_GEMV = 2, T_FIN = 3, T_DATT = 4, T_BATT = 5, T_SAMPLE = 6, T_MERGE = 7, T_WAIT = 8 };
__device__ __forceinline__ void mark(Ctx& c, int cat) {
  if (c.prof_on) {
    const unsigned long long t = (unsigned long long)clock64();
    c.t_acc[cat] += t - c.t_last;
    c.t_last = t;
  }
}

__device__ __forceinline__ bool check_abort(Ctx& c) {
  if (!c.aborted && *reinterpret_cast<volatile int*>(c.p->abort_flag) != 0) c.aborted = true;
  return c.aborted;
}
__device__ __forceinline__ void raise_abort(Ctx& c, int code) {
  atomicCAS(c.p->abort_flag, 0, code);
  if (c.p->status != nullptr) atomicCAS(c.p->status, 0, code);  // sticky across launches: the caller reads it whenever it likes
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
// bookkeeping of a tag-poll loop: returns false when the caller must give up
__device__ __forceinline__ bool poll_continue(Ctx& c, unsigned& spins, int code) {
  if (++spins > SPIN_LIMIT) {
    raise_abort(c, code);
    return false;
  }
  if ((spins & 255) == 0 && check_abort(c)) return false;
  return true;
}
__device__ __forceinline__ unsigned cur_tag(const Ctx& c) { return (c.launch << 12) | (c.phase & 0xfffu); }
// Stage 1 of every tagged read: lane 0 of the warp polls ONE word (with a short back-off) and the other 31 lanes sleep
// at the warp barrier; only then does the whole warp load and verify everything it needs.  Without this, 256 threads
// x several loads per poll round from each of the 148 CTAs hammer the very L2 lines the producers are storing to.
__device__ __forceinline__ void probe_wait(Ctx& c, const uint2* word, unsigned tag) {
  if (c.lane == 0 && !c.aborted) {
    unsigned spins = 0;
    while (ll_ld(word).y != tag) {
      __nanosleep(CSMB_FRAME_NAP);
      if (!poll_continue(c, spins, 150)) break;
    }
  }
  __syncwarp();
}
// consumer-side sync (shared-memory results visible to all consumer warps)
__device__ __forceinline__ void csync() { named_bar_sync(2, NCT); }

// two e4m3 -> two fp32, exact (hardware e4m3x2 -> f16x2, then f16 -> f32)
__device__ __forceinline__ float2 e4m3x2_to_f32(uint16_t v) {
  uint32_t h2;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"(v));
  float2 f;
  asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}" : "=f"(f.x), "=f"(f.y) : "r"(h2));
  return f;
}

// ------------------------------------------------------------------------------------------------ row partition
struct Range {
  const uint16_t* p;   // first weight of the range (bf16 elements; e4m3 instantiation: bytes)
  int row0, rows, K;
};
// this CTA's rows of the sub-matrix [rbase, rbase + N) of a matrix of Ntot rows: bf16 [Ntot][K], or (FMT = CSMB_WEIGHTS_E4M3) a
// blob of Ntot fp32 scales followed by the e4m3 bytes
__device__ __forceinline__ Range cta_range(const uint16_t* W, int Ntot, int rbase, int N, int K, int cta, int G) {
  const int r0 = (int)(((unsigned)N * (unsigned)cta) / (unsigned)G), r1 = (int)(((unsigned)N * (unsigned)(cta + 1)) / (unsigned)G);
  if (FMT == CSMB_WEIGHTS_E4M3) {
    const char* q = reinterpret_cast<const char*>(W) + e4m3_scale_bytes(Ntot) + (size_t)(rbase + r0) * K;
    return Range{reinterpret_cast<const uint16_t*>(q), r0, r1 - r0, K};
  }
  return Range{W + (size_t)(rbase + r0) * K, r0, r1 - r0, K};
}
// e4m3: the per-output-channel scales of that range's rows
__device__ __forceinline__ const float* range_scales(const uint16_t* W, int rbase, const Range& r) {
  return reinterpret_cast<const float*>(W) + rbase + r.row0;
}
// head i of audio_head_t: [V][dd] bf16, or one e4m3 blob per head
__device__ __forceinline__ const uint16_t* audio_head_ptr(const csmb_model& m, int i) {
  if (FMT == CSMB_WEIGHTS_E4M3)
    return reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(m.audio_head_t) +
                                             (size_t)i * e4m3_blob_bytes(m.audio_vocab, m.decoder.d_model));
  return m.audio_head_t + (size_t)i * m.audio_vocab * m.decoder.d_model;
}
// ring stages of a range: a stage holds SUB * 8 units of 1024 weights whatever the weight format (an e4m3 stage fills the
// first half of its 32 KiB slot), so the stage sequence — and every unit index — is the same for both formats
__device__ __forceinline__ int n_stages(const Range& r) {
  return (int)(((size_t)r.rows * r.K * 2 + STAGE_BYTES - 1) / STAGE_BYTES);
}

// ------------------------------------------------------------------------------------------------ consumer GEMV
// Partial dot products of R activation rows with every unit of the range.  GS ring stages are processed per iteration,
// all unconditionally (a missing stage just produces an unused value) so that their LDS -> FMA -> shuffle chains
// interleave and the per-iteration bookkeeping (mbarrier waits / arrives, loop control) is amortised; two accumulators
// per (unit, row) halve the dependent-FMA chain.  The lane reduction stops after three shuffle steps and leaves 4
// partials per unit: part[(r*MAXU + u)*4 + 0..3].
template <int R>
__device__ void consume(Ctx& c, const Range& r, const float (&xr)[R][32], float* part) {
  if (c.p->dbg & 8) return;  // timing experiment: no streaming at all (pure latency chain)
  constexpr int GS = CSMB_FRAME_GS;  // 16 KiB sub-stages per iteration; 4 was measured slower: longer waits, fewer free ring slots
  constexpr int SPI = GS / SUB;  // ring stages per iteration (2 with 16 KiB stages, 1 with 32 KiB stages)
  static_assert(GS % SUB == 0, "stage geometry");
  const int KS = r.K / UNIT;
  const int units = r.rows * KS;
  const int nst = n_stages(r);
  const bool math = !(c.p->dbg & 1);
  const size_t woff = (size_t)c.warp * (UNIT * 2) + c.lane * 16;
  for (int s = 0; s < nst; s += SPI) {
    const int ns = min(SPI, nst - s);
    mark(c, T_GEMV);
#pragma unroll
    for (int g = 0; g < SPI; ++g)
      if (g < ns) mbar_wait(c, &c.ring.full[(c.q + g) % NSTAGES], ((c.q + g) / NSTAGES) & 1, 2);
    mark(c, T_WAIT);
    float acc[GS][R][2];
#pragma unroll
    for (int g = 0; g < GS; ++g)
#pragma unroll
      for (int i = 0; i < R; ++i) acc[g][i][0] = acc[g][i][1] = 0.f;
    if (FMT == CSMB_WEIGHTS_E4M3 && math && !c.aborted) {
      // weight-only FP8: the same units at one byte per weight — 8 e4m3 per lane and chunk (one 8-byte load), widened to
      // fp32 with cvt.rn.f16x2.e4m3x2 + cvt.f32.f16 (exact); the per-channel scale is applied where the row is finalised
      const unsigned char* b[GS];
#pragma unroll
      for (int g = 0; g < GS; ++g)
        b[g] = c.ring.data + (size_t)((c.q + g / SUB) % NSTAGES) * STAGE_BYTES + (size_t)(g % SUB) * 8192 + (size_t)c.warp * UNIT +
               c.lane * 8;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint2 w[GS];
#pragma unroll
        for (int g = 0; g < GS; ++g) w[g] = *reinterpret_cast<const uint2*>(b[g] + ch * 256);
#pragma unroll
        for (int g = 0; g < GS; ++g) {
          const float2 f01 = e4m3x2_to_f32((uint16_t)(w[g].x & 0xffffu)), f23 = e4m3x2_to_f32((uint16_t)(w[g].x >> 16));
          const float2 f45 = e4m3x2_to_f32((uint16_t)(w[g].y & 0xffffu)), f67 = e4m3x2_to_f32((uint16_t)(w[g].y >> 16));
          const float f[8] = {f01.x, f01.y, f23.x, f23.y, f45.x, f45.y, f67.x, f67.y};
#pragma unroll
          for (int i = 0; i < R; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[g][i][e & 1] = fmaf(f[e], xr[i][ch * 8 + e], acc[g][i][e & 1]);
        }
      }
    } else if (FMT == CSMB_WEIGHTS_BF16 && math && !c.aborted) {
      const unsigned char* b[GS];
#pragma unroll
      for (int g = 0; g < GS; ++g)
        b[g] = c.ring.data + (size_t)((c.q + g / SUB) % NSTAGES) * STAGE_BYTES + (size_t)(g % SUB) * 16384 + woff;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint4 w[GS];
#pragma unroll
        for (int g = 0; g < GS; ++g) w[g] = *reinterpret_cast<const uint4*>(b[g] + ch * 512);
#pragma unroll
        for (int g = 0; g < GS; ++g) {
          const float f[8] = {bf16lo(w[g].x), bf16hi(w[g].x), bf16lo(w[g].y), bf16hi(w[g].y),
                              bf16lo(w[g].z), bf16hi(w[g].z), bf16lo(w[g].w), bf16hi(w[g].w)};
#pragma unroll
          for (int i = 0; i < R; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[g][i][e & 1] = fmaf(f[e], xr[i][ch * 8 + e], acc[g][i][e & 1]);
        }
      }
    }
    float sum[GS][R];
#pragma unroll
    for (int g = 0; g < GS; ++g)
#pragma unroll
      for (int i = 0; i < R; ++i) sum[g][i] = acc[g][i][0] + acc[g][i][1];
#if CSMB_FRAME_EARLY_ARRIVE
    // the stage's bytes are in registers (the sums above depend on every load): hand the ring slots back before the
    // lane reduction so that the producer's next bulk copy overlaps it
#pragma unroll
    for (int g = 0; g < GS; ++g)
#pragma unroll
      for (int i = 0; i < R; ++i) asm volatile("" ::"f"(sum[g][i]));
    __syncwarp();
    if (c.lane == 0) {
#pragma unroll
      for (int g = 0; g < SPI; ++g)
        if (g < ns) mbar_arrive(&c.ring.empty[(c.q + g) % NSTAGES]);
    }
#endif
#if CSMB_FRAME_XPOSE && CSMB_FRAME_GS == 2
    // transposing first step: the lower half-warp keeps unit 0, the upper half unit 1 (same additions, commuted, as the
    // plain butterfly: bit-identical partials), then two steps on one value instead of two
    {
      const bool up = (c.lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const float keep = up ? sum[1][i] : sum[0][i], give = up ? sum[0][i] : sum[1][i];
        float v = keep + __shfl_xor_sync(0xffffffffu, give, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        sum[0][i] = v;
      }
      const int g = c.lane >> 4;
      const int u = (s * SUB + g) * NCW + c.warp;
      if ((c.lane & 12) == 0 && g < ns * SUB && u < units) {
#pragma unroll
        for (int i = 0; i < R; ++i) part[((size_t)i * MAXU + u) * 4 + (c.lane & 3)] = sum[0][i];
      }
    }
#else
#pragma unroll
    for (int o = 16; o >= 4; o >>= 1)
#pragma unroll
      for (int g = 0; g < GS; ++g)
#pragma unroll
        for (int i = 0; i < R; ++i) sum[g][i] += __shfl_xor_sync(0xffffffffu, sum[g][i], o);
    if (c.lane < 4) {
#pragma unroll
      for (int g = 0; g < GS; ++g) {
        const int u = (s * SUB + g) * NCW + c.warp;
        if (g < ns * SUB && u < units) {
#pragma unroll
          for (int i = 0; i < R; ++i) part[((size_t)i * MAXU + u) * 4 + c.lane] = sum[g][i];
        }
      }
    }
#endif
#if !CSMB_FRAME_EARLY_ARRIVE
    __syncwarp();
    if (c.lane == 0) {
#pragma unroll
      for (int g = 0; g < SPI; ++g)
        if (g < ns) mbar_arrive(&c.ring.empty[(c.q + g) % NSTAGES]);
    }
#endif
    c.q += ns;
  }
}

// sum the 4*KS partials of local row j (fixed order -> deterministic)
__device__ __forceinline__ float row_total(const float* part, int j, int KS) {
  const float4* p4 = reinterpret_cast<const float4*>(part + (size_t)j * KS * 4);
  float4 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = k < KS ? p4[k] : make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  return s;
}

// ------------------------------------------------------------------------------------------------ activations
// Cooperative gather of NL*512 tagged words into shared memory: every thread polls its own 16-byte pairs until both
// carry `tag`.  The caller issues csync() before reading dst.
template <int NL>
__device__ __forceinline__ void gather_ll(Ctx& c, const uint2* src, unsigned tag, float* dst) {
  uint4 v[NL];
  unsigned spins = 0;
  probe_wait(c, src + (size_t)(NL - 1) * (NCT * 2) + c.warp * 64 + 63, tag);
  while (true) {
#pragma unroll
    for (int j = 0; j < NL; ++j) v[j] = ll_ld2(src + (size_t)j * (NCT * 2) + c.tid * 2);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NL; ++j) ok = ok && (v[j].y == tag) && (v[j].w == tag);
    if (ok || c.aborted) break;
    if (!poll_continue(c, spins, 200 + (int)(c.phase & 0x3ff))) break;
  }
#pragma unroll
  for (int j = 0; j < NL; ++j)
    *reinterpret_cast<float2*>(dst + (size_t)j * (NCT * 2) + c.tid * 2) = make_float2(__uint_as_float(v[j].x), __uint_as_float(v[j].z));
}

// this warp's K-slice (kseg = warp % KS) of R rows out of the gathered vector xs[R][K]
template <int R>
__device__ __forceinline__ void slice_from_smem(Ctx& c, const float* xs, int K, float (&xr)[R][32]) {
  const int kseg = c.warp % (K / UNIT);
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const float* p = xs + (size_t)i * K + kseg * UNIT + ch * 256 + c.lane * 8;
      const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      xr[i][ch * 8 + 0] = a.x; xr[i][ch * 8 + 1] = a.y; xr[i][ch * 8 + 2] = a.z; xr[i][ch * 8 + 3] = a.w;
      xr[i][ch * 8 + 4] = b.x; xr[i][ch * 8 + 5] = b.y; xr[i][ch * 8 + 6] = b.z; xr[i][ch * 8 + 7] = b.w;
    }
}

struct NormW {
  float4 g[8];
};
// this warp's slice of a norm weight vector; issued a phase early so that the (possibly HBM-cold) load is hidden
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
// RMSNorm-ed K-slices of R rows from the gathered vector xs[R][K] (every warp recomputes the row statistics)
template <int R, int KS>
__device__ __forceinline__ void norm_from_smem(Ctx& c, const float* xs, const NormW& nw, float eps, float (&xr)[R][32]) {
  constexpr int K = KS * UNIT;
  slice_from_smem<R>(c, xs, K, xr);
#pragma unroll
  for (int i = 0; i < R; ++i) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < K / 128; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(xs + (size_t)i * K + j * 128 + c.lane * 4);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)K + eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xr[i][j * 4 + 0] *= rstd * nw.g[j].x; xr[i][j * 4 + 1] *= rstd * nw.g[j].y;
      xr[i][j * 4 + 2] *= rstd * nw.g[j].z; xr[i][j * 4 + 3] *= rstd * nw.g[j].w;
    }
  }
}

// K = 8192 (MLP activation): every warp needs only its own 1024-slice, read straight from the tagged buffer
template <int R>
__device__ __forceinline__ void load_act_slice(Ctx& c, const uint2* act, int F, unsigned tag, float (&xr)[R][32]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const uint2* base = act + (size_t)i * F + c.warp * UNIT + c.lane * 8;
    uint4 v[16];
    unsigned spins = 0;
    probe_wait(c, act + (size_t)i * F + c.warp * UNIT + UNIT - 1, tag);
    while (true) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[ch * 4 + j] = ll_ld2(base + ch * 256 + j * 2);
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 16; ++j) ok = ok && (v[j].y == tag) && (v[j].w == tag);
      if (ok || c.aborted) break;
      if (!poll_continue(c, spins, 300)) break;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      xr[i][j * 2] = __uint_as_float(v[j].x);
      xr[i][j * 2 + 1] = __uint_as_float(v[j].z);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEMV phases
// y[row] = (res ? res[row] : 0) + W[row,:] . x  for this CTA's rows, written as tagged words with the current tag
template <int R>
__device__ void phase_linear(Ctx& c, const uint16_t* W, int N, int K, const float (&xr)[R][32], uint2* y, int ldy,
                             const uint2* res, int ldr) {
  const Range r = cta_range(W, N, 0, N, K, c.cta, c.G);
  // this thread finalises output j = tid (rows*R <= 44): fetch its residual early (its producer phase is long over)
  const int j = c.tid;
  const bool mine = j < r.rows * R;
  const int ri = mine ? j / r.rows : 0, rrow = mine ? j % r.rows : 0;
  float rv = 0.f;
  if (mine && res) rv = __uint_as_float(ll_ld(res + (size_t)ri * ldr + r.row0 + rrow).x);
  mark(c, T_LOAD);
  consume<R>(c, r, xr, c.part);
  mark(c, T_GEMV);
  csync();
  if (mine) {
    float t = row_total(c.part + (size_t)ri * MAXU * 4, rrow, K / UNIT);
    if (FMT == CSMB_WEIGHTS_E4M3) t *= __ldg(range_scales(W, 0, r) + rrow);   // weight-only FP8: per-output-channel scale on the finished dot product
    ll_st(y + (size_t)ri * ldy + r.row0 + rrow, rv + t, cur_tag(c));
  }
  mark(c, T_FIN);
}

// SwiGLU MLP first half: act[f] = silu(Wg[f,:].x) * (Wu[f,:].x)
template <int R>
__device__ void phase_gate_up(Ctx& c, const uint16_t* Wgu, int F, int K, const float (&xr)[R][32], uint2* act) {
  const Range rg = cta_range(Wgu, 2 * F, 0, F, K, c.cta, c.G);
  const Range ru = cta_range(Wgu, 2 * F, F, F, K, c.cta, c.G);
  float* pg = c.part;
  float* pu = c.part + 2 * MAXU * 4;
  mark(c, T_LOAD);
  consume<R>(c, rg, xr, pg);
  consume<R>(c, ru, xr, pu);
  mark(c, T_GEMV);
  csync();
  const int KS = K / UNIT;
  const unsigned tag = cur_tag(c);
  for (int j = c.tid; j < rg.rows * R; j += NCT) {
    const int i = j / rg.rows, row = j % rg.rows;
    float g = row_total(pg + (size_t)i * MAXU * 4, row, KS), u = row_total(pu + (size_t)i * MAXU * 4, row, KS);
    if (FMT == CSMB_WEIGHTS_E4M3) {
      g *= __ldg(range_scales(Wgu, 0, rg) + row);
      u *= __ldg(range_scales(Wgu, F, ru) + row);
    }
    ll_st(act + (size_t)i * F + rg.row0 + row, (g / (1.f + expf(-g))) * u, tag);
  }
  mark(c, T_FIN);
}

// ------------------------------------------------------------------------------------------------ sampling
__device__ __forceinline__ float gumbel_at(int idx, uint32_t dlo, uint32_t dhi, uint32_t seq, uint32_t k0, uint32_t k1) {
  uint32_t ctr[4] = {(uint32_t)(idx >> 2), dlo, dhi, seq};
  philox4x32_10(ctr, k0, k1);
  return -logf(-logf(u01(ctr[idx & 3])));
}
// every CTA computes the same token from the tagged logits (consumer warps only); V <= 9 * 256
__device__ int sample_token(Ctx& c, const uint2* logits, int V, unsigned tag, unsigned long long draw) {
  const FrameParams& p = *c.p;
  uint2 v[9];
  unsigned spins = 0;
  probe_wait(c, logits + min(V - 1, c.warp * (V / NCW) + V / NCW - 1), tag);
  while (true) {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int i = c.tid + j * NCT;
      v[j] = i < V ? ll_ld(logits + i) : make_uint2(0xff800000u, tag);
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 9; ++j) ok = ok && (v[j].y == tag);
    if (ok || c.aborted) break;
    if (!poll_continue(c, spins, 400)) break;
  }
  mark(c, T_POLL);
  // top-k / min-p (mlx_lm make_sampler as used by cli/generate.py:168-174): the same definition as k_sample_filtered —
  // e_i = exp(logit_i - max), threshold = max(k-th largest e, min_p), tokens with e_i >= threshold stay.  The k-th
  // largest e is found exactly by a 31-step bisection on its bit pattern (e >= 0 orders like an unsigned integer).
  const bool filt = p.inv_temp != 0.f && (p.top_k > 0 || p.min_p > 0.f || p.top_p > 0.f);
  float ev[9];
  float thresh = 0.f;
  if (filt) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 9; ++j)
      if (c.tid + j * NCT < V) m = fmaxf(m, __uint_as_float(v[j].x));
    m = warp_max(m);
    if (c.lane == 0) c.sred[c.warp] = m;
    csync();
    m = c.sred[0];
#pragma unroll
    for (int w = 1; w < NCW; ++w) m = fmaxf(m, c.sred[w]);
    csync();
#pragma unroll
    for (int j = 0; j < 9; ++j) ev[j] = (c.tid + j * NCT < V) ? expf(__uint_as_float(v[j].x) - m) : -1.f;
    unsigned T = 0u;
    if (p.top_k > 0) {
      int* cnt = reinterpret_cast<int*>(c.sred + 2 * NCW);
      for (int bit = 30; bit >= 0; --bit) {
        const unsigned cand = T | (1u << bit);
        int n = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j) n += (ev[j] >= 0.f && __float_as_uint(ev[j]) >= cand) ? 1 : 0;
        n = __reduce_add_sync(0xffffffffu, n);
        int* buf = cnt + (bit & 1) * NCW;
        if (c.lane == 0) buf[c.warp] = n;
        csync();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < NCW; ++w) tot += buf[w];
        if (tot >= p.top_k) T = cand;
      }
    }
    if (p.top_p > 0.f) {
      // nucleus: masses are exact integers q = floor(e * 2^32) (sums are order-independent, so this kernel, the chain's
      // sampler and the per-op sorter agree bit for bit); keep the tokens at or above the largest T whose mass
      // sum(q : e >= T) still reaches top_p * sum(q) — the same greedy bisection as for top-k, on mass instead of count
      unsigned long long* qs = reinterpret_cast<unsigned long long*>(c.sred + 4 * NCW);  // [2][NCW]
      auto block_mass = [&](unsigned cand, int slot) {
        unsigned long long n = 0;
#pragma unroll
        for (int j = 0; j < 9; ++j)
          if (ev[j] >= 0.f && __float_as_uint(ev[j]) >= cand) n += __float2ull_rz(ev[j] * 4294967296.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        unsigned long long* buf = qs + slot * NCW;
        if (c.lane == 0) buf[c.warp] = n;
        csync();
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < NCW; ++w) tot += buf[w];
        return tot;
      };
      const double need = (double)p.top_p * (double)block_mass(0u, 1);
      unsigned Tp = 0u;
      for (int bit = 30; bit >= 0; --bit) {
        const unsigned cand = Tp | (1u << bit);
        if ((double)block_mass(cand, bit & 1) >= need) Tp = cand;
      }
      T = T > Tp ? T : Tp;
    }
    thresh = fmaxf(__uint_as_float(T), p.min_p);
  }
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const uint32_t dlo = (uint32_t)draw, dhi = (uint32_t)(draw >> 32);
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    const int i = c.tid + j * NCT;
    if (i < V && !(filt && !(ev[j] >= thresh))) {
      float t = __uint_as_float(v[j].x);
      if (p.inv_temp != 0.f) t = t * p.inv_temp + gumbel_at(i, dlo, dhi, p.seq, p.seed_lo, p.seed_hi);
      argmax_combine(bv, bi, t, i);
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

// ------------------------------------------------------------------------------------------------ attention
// Stage this CTA's cached decoder K/V rows (positions < pos0) of one layer into shared memory with ONE bulk copy.
// The rows were written by this same CTA (ordinary stores, earlier steps): a proxy fence orders them before the copy.
__device__ __forceinline__ void stage_decoder_kv(Ctx& c, const float* kvl, int pos0) {
  if (pos0 == 0 || c.tid != 0 || c.aborted) return;
  const uint32_t bytes = (uint32_t)pos0 * KROW * 4;
  asm volatile("fence.proxy.async;" ::: "memory");
  mbar_arrive_expect_tx(c.kvbar, bytes);
  bulk_g2s_plain(c.kvs, kvl, bytes, c.kvbar);
}

// Depth decoder attention, recomputed by every CTA: warp h = query head h (8 heads x 128), kv head h/4.
// qkv: this step's raw tagged projections for R rows (positions pos0 .. pos0+R-1); cs: their RoPE (cos,sin) pairs for
// this lane (dims 4*lane .. 4*lane+3).  Keys < pos0 come from the staged shared-memory copy: QK is lane-per-key with a
// rotated chunk order (conflict-free although rows are 2 KiB apart), PV is dims-over-lanes.
// Result: sattn[R][1024] (every warp of the o-projection needs the whole vector).
template <int R>
__device__ void decoder_attention(Ctx& c, float* kvl, int pos0, const uint2* qkv, int ldq, unsigned tag, const float4 (&cs)[R]) {
  constexpr int HD = 128, H = 8, HKV = 2;
  const int h = c.warp, kvh = h / (H / HKV);
  float* sq = c.sattn + 2 * 1024 + c.warp * (2 * HD);  // this warp's rotated q rows, broadcast-read below
  const float scale = rsqrtf((float)HD);
  float4 q[R], k[R], v[R];
  {
    uint4 t[R][6];
    unsigned spins = 0;
    probe_wait(c, qkv + (size_t)(R - 1) * ldq + (H + HKV + kvh) * HD + HD - 1, tag);
    while (true) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const uint2* row = qkv + (size_t)i * ldq + c.lane * 4;
        t[i][0] = ll_ld2(row + h * HD);               t[i][1] = ll_ld2(row + h * HD + 2);
        t[i][2] = ll_ld2(row + (H + kvh) * HD);       t[i][3] = ll_ld2(row + (H + kvh) * HD + 2);
        t[i][4] = ll_ld2(row + (H + HKV + kvh) * HD); t[i][5] = ll_ld2(row + (H + HKV + kvh) * HD + 2);
      }
      bool ok = true;
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) ok = ok && (t[i][j].y == tag) && (t[i][j].w == tag);
      if (ok || c.aborted) break;
      if (!poll_continue(c, spins, 500)) break;
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      q[i] = make_float4(__uint_as_float(t[i][0].x), __uint_as_float(t[i][0].z), __uint_as_float(t[i][1].x), __uint_as_float(t[i][1].z));
      k[i] = make_float4(__uint_as_float(t[i][2].x), __uint_as_float(t[i][2].z), __uint_as_float(t[i][3].x), __uint_as_float(t[i][3].z));
      v[i] = make_float4(__uint_as_float(t[i][4].x), __uint_as_float(t[i][4].z), __uint_as_float(t[i][5].x), __uint_as_float(t[i][5].z));
    }
  }
  mark(c, T_POLL);
  float qn[R][4], kn[R][4];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    qn[i][0] = q[i].x * cs[i].x - q[i].y * cs[i].y; qn[i][1] = q[i].y * cs[i].x + q[i].x * cs[i].y;
    qn[i][2] = q[i].z * cs[i].z - q[i].w * cs[i].w; qn[i][3] = q[i].w * cs[i].z + q[i].z * cs[i].w;
    kn[i][0] = k[i].x * cs[i].x - k[i].y * cs[i].y; kn[i][1] = k[i].y * cs[i].x + k[i].x * cs[i].y;
    kn[i][2] = k[i].z * cs[i].z - k[i].w * cs[i].w; kn[i][3] = k[i].w * cs[i].z + k[i].z * cs[i].w;
    *reinterpret_cast<float4*>(sq + i * HD + c.lane * 4) = make_float4(qn[i][0], qn[i][1], qn[i][2], qn[i][3]);
    if ((h % (H / HKV)) == 0) {  // one writer per kv head keeps this CTA's private K/V copy for later steps
      float* dst = kvl + (size_t)(pos0 + i) * KROW + kvh * HD + c.lane * 4;
      *reinterpret_cast<float4*>(dst) = make_float4(kn[i][0], kn[i][1], kn[i][2], kn[i][3]);
      *reinterpret_cast<float4*>(dst + HKV * HD) = v[i];
    }
  }
  if (pos0 > 0) {
    mbar_wait(c, c.kvbar, c.kv_phase, 4);
    c.kv_phase ^= 1;
  }
  __syncwarp();  // sq is per-warp
  const float* krow = c.kvs + (size_t)c.lane * KROW + kvh * HD;   // lane-per-key
  const float* vcol = c.kvs + HKV * HD + kvh * HD + c.lane * 4;   // dims-over-lanes
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int pos = pos0 + i;
    float dot = 0.f;
    if (c.lane < pos0) {
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
// partial (acc[64], m, l) per head, as tagged words.
__device__ void backbone_attention_item(Ctx& c, int layer, int kvh, int chunk, int nchunks, int pos, const uint2* qkv,
                                        unsigned tag_in) {
  const FrameParams& p = *c.p;
  constexpr int HD = 64, H = 32, HKV = 8;
  const int hq = kvh * (H / HKV) + (c.warp & 3), half = c.warp >> 2;
  const int kbeg = chunk * 128 + half * 64;
  float* pool = p.kv_pool + (size_t)layer * p.kv_layer_stride;
  const size_t page_stride = (size_t)2 * HKV * CSMB_PAGE * HD;
  const float* rope = p.m.backbone.rope + (size_t)pos * (HD / 2) * 2;
  float* sq = c.sattn + c.warp * HD;                  // rotated q of this warp's head
  float* sm = c.sattn + NCW * HD + c.warp * PSTRIDE;  // this warp's partial for the intra-CTA merge
  float2 qraw, kraw, vnew;
  {
    uint4 t[3];
    unsigned spins = 0;
    probe_wait(c, qkv + (H + HKV + kvh) * HD + HD - 1, tag_in);
    while (true) {
      t[0] = ll_ld2(qkv + hq * HD + c.lane * 2);
      t[1] = ll_ld2(qkv + (H + kvh) * HD + c.lane * 2);
      t[2] = ll_ld2(qkv + (H + HKV + kvh) * HD + c.lane * 2);
      const bool ok = t[0].y == tag_in && t[0].w == tag_in && t[1].y == tag_in && t[1].w == tag_in && t[2].y == tag_in && t[2].w == tag_in;
      if (ok || c.aborted) break;
      if (!poll_continue(c, spins, 600)) break;
    }
    qraw = make_float2(__uint_as_float(t[0].x), __uint_as_float(t[0].z));
    kraw = make_float2(__uint_as_float(t[1].x), __uint_as_float(t[1].z));
    vnew = make_float2(__uint_as_float(t[2].x), __uint_as_float(t[2].z));
  }
  mark(c, T_POLL);
  const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + c.lane * 2));
  *reinterpret_cast<float2*>(sq + c.lane * 2) = make_float2(qraw.x * cs.x - qraw.y * cs.y, qraw.y * cs.x + qraw.x * cs.y);
  const float2 knew = make_float2(kraw.x * cs.x - kraw.y * cs.y, kraw.y * cs.x + kraw.x * cs.y);
  const bool has_new = (pos >= kbeg && pos < kbeg + 64);
  if (has_new && (c.warp & 3) == 0) {  // one writer per kv head: append to the paged cache (read by later launches)
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
    for (int d = 0; d < 16; ++d) kv[d] = valid[t] ? *reinterpret_cast<const float4*>(kptr[t] + d * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    for (int d = 0; d < 16; ++d) vv[d] = valid[t] ? *reinterpret_cast<const float4*>(vp + d * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    sm[HD] = m;  // -inf if this half saw no key
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
    const float2 o2 = make_float2(w0 * a0.x + w1 * a1.x, w0 * a0.y + w1 * a1.y);
    const unsigned tag = cur_tag(c);
    if (nchunks == 1) {
      const float inv = 1.f / L;
      ll_st2(p.attn_out + hq * HD + c.lane * 2, o2.x * inv, o2.y * inv, tag);
    } else {
      uint2* dst = p.attn_part + ((size_t)chunk * H + hq) * PSTRIDE;
      ll_st2(dst + c.lane * 2, o2.x, o2.y, tag);
      if (c.lane == 0) ll_st2(dst + HD, M, L, tag);
      if (c.lane >= 1 && c.lane <= 3) ll_st2(dst + HD + c.lane * 2, 0.f, 0.f, tag);  // pad words: the reader polls whole rows
    }
  }
  csync();
}

// merge of the per-chunk partials into the o-projection's register slice (K = 2048 -> 2 ksegs; a lane's 8
// consecutive k share one head).  Path for S > 128: chunk after chunk, the 32 x 72 tagged words of a chunk are
// gathered cooperatively into shared memory (one polled L2 round trip), then every lane folds its 4 head-slices into
// running (max, sum, acc) registers.
__device__ void merged_attention_slice(Ctx& c, int nchunks, unsigned tag, float (&xr)[1][32]) {
  const FrameParams& p = *c.p;
  constexpr int HD = 64, H = 32, WORDS = H * PSTRIDE;  // 2304 = 9 x 256
  const int kseg = c.warp % 2;
  float M[4], L[4], o[4][8];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    M[ch] = -INFINITY;
    L[ch] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[ch][e] = 0.f;
  }
  for (int s = 0; s < nchunks; ++s) {
    const uint2* src = p.attn_part + (size_t)s * WORDS;
    uint2 v[9];
    unsigned spins = 0;
    probe_wait(c, src + 8 * NCT + c.warp * 32 + 31, tag);
    while (true) {
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = ll_ld(src + j * NCT + c.tid);
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 9; ++j) ok = ok && (v[j].y == tag);
      if (ok || c.aborted) break;
      if (!poll_continue(c, spins, 700)) break;
    }
    csync();  // previous chunk fully consumed
#pragma unroll
    for (int j = 0; j < 9; ++j) c.sattn[j * NCT + c.tid] = __uint_as_float(v[j].x);
    csync();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int k = kseg * UNIT + ch * 256 + c.lane * 8;
      const int h = k / HD, d = k % HD;
      const float* base = c.sattn + h * PSTRIDE;
      const float ms = base[HD], ls = base[HD + 1];
      if (ms == -INFINITY) continue;
      const float Mn = fmaxf(M[ch], ms);
      const float wo = (M[ch] == -INFINITY) ? 0.f : expf(M[ch] - Mn), wn = expf(ms - Mn);
      L[ch] = L[ch] * wo + ls * wn;
      const float4 a = *reinterpret_cast<const float4*>(base + d), b = *reinterpret_cast<const float4*>(base + d + 4);
      o[ch][0] = o[ch][0] * wo + wn * a.x; o[ch][1] = o[ch][1] * wo + wn * a.y;
      o[ch][2] = o[ch][2] * wo + wn * a.z; o[ch][3] = o[ch][3] * wo + wn * a.w;
      o[ch][4] = o[ch][4] * wo + wn * b.x; o[ch][5] = o[ch][5] * wo + wn * b.y;
      o[ch][6] = o[ch][6] * wo + wn * b.z; o[ch][7] = o[ch][7] * wo + wn * b.w;
      M[ch] = Mn;
    }
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const float inv = 1.f / L[ch];
#pragma unroll
    for (int e = 0; e < 8; ++e) xr[0][ch * 8 + e] = o[ch][e] * inv;
  }
  csync();  // sattn is reused by the next gather
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
      case 0: return cta_range(L.wqkv[l], nqkv, 0, nqkv, d, cta, G);
      case 1: return cta_range(L.wo[l], d, 0, d, d, cta, G);
      case 2: return cta_range(L.wgu[l], 2 * L.d_ff, 0, L.d_ff, d, cta, G);
      case 3: return cta_range(L.wgu[l], 2 * L.d_ff, L.d_ff, L.d_ff, d, cta, G);
      default: return cta_range(L.wdown[l], d, 0, d, L.d_ff, cta, G);
    }
  };
  const int nb = B.n_layers * 5;
  if (idx < nb) {
    out = layer_range(B, idx / 5, idx % 5, db);
    return true;
  }
  if (idx == nb) {
    out = cta_range(p.m.c0_head, V, 0, V, db, cta, G);
    return true;
  }
  const int per_step = 2 + D.n_layers * 5;
  const int j = idx - nb - 1, step = j / per_step, t = j % per_step;
  if (step >= p.m.n_codebooks - 1) return false;
  if (t == 0) out = cta_range(p.m.projection, dd, 0, dd, db, cta, G);
  else if (t == per_step - 1) out = cta_range(audio_head_ptr(p.m, step), V, 0, V, dd, cta, G);
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
  constexpr int wb = FMT == CSMB_WEIGHTS_E4M3 ? 1 : 2;
  const size_t sb = (size_t)STAGE_BYTES / 2 * wb;   // bytes of a full stage in this range's weight format
  const size_t total = (size_t)k.r.rows * k.r.K * wb, off = (size_t)k.stage * sb;
  src = reinterpret_cast<const char*>(k.r.p) + off;
  bytes = (uint32_t)((total - off) < sb ? (total - off) : sb);
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

// Producer (one thread per CTA).  The RING cursor copies the schedule, stage by stage, into shared memory as slots
// free up.  While the ring is full (the consumers are in a latency phase) a second PREFETCH cursor runs ahead of it and
// pulls future stages HBM -> L2 (cp.async.bulk.prefetch.L2), one stage per pf_interval cycles (this SM's fair share
// of the HBM rate) and at most pf_max stages ahead, so HBM keeps streaming through the latency phases and the ring
// later refills from L2.  Every byte still leaves HBM once.  pf_max = 0 disables the prefetch cursor.
__device__ void producer_main(Ctx& c) {
  const FrameParams& p = *c.p;
  if (p.dbg & 8) return;
  Cursor rc, pc;
  rc.idx = p.h_in ? p.m.backbone.n_layers * 5 : 0;  // depth-only launches start at the c0 head
  cursor_load(p, rc, c.cta, c.G);
  pc = rc;
  uint32_t pq = 0;  // global stage number of the prefetch cursor (>= c.q)
  const int pf_max = p.pf_max, pf_interval = p.pf_interval;
  long long last_pf = clock64();
  unsigned n_pf = 0;
  const uint64_t policy_keep = make_keep_policy((float)((p.dbg >> 4) & 7) * 0.125f);
  while (rc.valid && !c.aborted) {
    const int slot = c.q % NSTAGES;
    const uint32_t par = (c.q / NSTAGES) & 1;
    unsigned idle = 0;
    while (!mbar_test(&c.ring.empty[slot], par ^ 1)) {
      if (pf_max > 0 && pc.valid && pq < c.q + (uint32_t)pf_max && clock64() - last_pf >= pf_interval) {
        if (pq >= c.q + NSTAGES) {  // stages inside the ring's own reach are fetched by the ring copies anyway
          const char* src;
          uint32_t n;
          cursor_chunk(pc, src, n);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(n) : "memory");
          last_pf = clock64();
          ++n_pf;
        }
        cursor_advance(p, pc, c.cta, c.G);
        ++pq;
      } else {
        if (++idle > SPIN_LIMIT * 8u) raise_abort(c, 1);
        if ((idle & 4095) == 0 && check_abort(c)) break;
      }
      if (c.aborted) break;
    }
    if (c.aborted) break;
    const char* src;
    uint32_t n;
    cursor_chunk(rc, src, n);
    mbar_arrive_expect_tx(&c.ring.full[slot], n);
    // dbg bits 4..6 (e4m3 only): keep k/8 of the decoder's lines in L2 across the 31 depth steps
    const bool keep = FMT == CSMB_WEIGHTS_E4M3 && (p.dbg & 0x70) != 0 && rc.idx > p.m.backbone.n_layers * 5;
    bulk_g2s(c.ring.data + (size_t)slot * STAGE_BYTES, src, n, &c.ring.full[slot], keep ? policy_keep : c.policy);
    cursor_advance(p, rc, c.cta, c.G);
    ++c.q;
    if (pq < c.q) {  // keep the prefetch cursor at or ahead of the ring cursor
      pc = rc;
      pq = c.q;
    }
  }
  if (p.prof != nullptr) p.prof[(size_t)c.cta * 16 + 14] = n_pf;
}

// one depth-decoder step with R rows (R = 2 for the first step: positions 0 and 1)
template <int R>
__device__ void decoder_step(Ctx& c, int step, int pos0, float (&xr)[R][32]) {
  const FrameParams& p = *c.p;
  const csmb_llama& D = p.m.decoder;
  const int db = p.m.backbone.d_model, dd = D.d_model, V = p.m.audio_vocab, F = D.d_ff;
  const int nqkv = (D.n_heads + 2 * D.n_kv_heads) * D.head_dim;
  constexpr int NLX = R * 1024 / (NCT * 2);  // gather loads per thread for an [R][1024] vector
  float* kv_mine = p.dec_kv + (size_t)c.cta * D.n_layers * 32 * KROW;
  // RoPE rows of this step's positions (same for all layers): loaded now, used two phases later
  float4 cs[R];
#pragma unroll
  for (int i = 0; i < R; ++i)
    cs[i] = __ldg(reinterpret_cast<const float4*>(D.rope + ((size_t)(pos0 + i) * (D.head_dim / 2) + c.lane * 2) * 2));
  // projection: xr already holds the input row slices (K = 2048)
  unsigned tag_x = cur_tag(c);
  phase_linear<R>(c, p.m.projection, dd, db, xr, p.dxa, dd, nullptr, 0);
  c.phase++;
  NormW nw = prefetch_norm<1>(c, D.norm_in[0]);
  uint2* x = p.dxa;
  uint2* x1 = p.dxb;
  for (int l = 0; l < D.n_layers; ++l) {
    float* kvl = kv_mine + (size_t)l * 32 * KROW;
    stage_decoder_kv(c, kvl, pos0);
    gather_ll<NLX>(c, x, tag_x, c.sattn);
    mark(c, T_POLL);
    csync();
    norm_from_smem<R, 1>(c, c.sattn, nw, D.eps, xr);
    const unsigned tag_qkv = cur_tag(c);
    phase_linear<R>(c, D.wqkv[l], nqkv, dd, xr, p.qkv, nqkv, nullptr, 0);
    c.phase++;
    decoder_attention<R>(c, kvl, pos0, p.qkv, nqkv, tag_qkv, cs);
    mark(c, T_DATT);
    slice_from_smem<R>(c, c.sattn, dd, xr);
    const unsigned tag_x1 = cur_tag(c);
    phase_linear<R>(c, D.wo[l], dd, dd, xr, x1, dd, x, dd);
    c.phase++;
    nw = prefetch_norm<1>(c, D.norm_post[l]);
    gather_ll<NLX>(c, x1, tag_x1, c.sattn);
    mark(c, T_POLL);
    csync();
    norm_from_smem<R, 1>(c, c.sattn, nw, D.eps, xr);
    const unsigned tag_act = cur_tag(c);
    phase_gate_up<R>(c, D.wgu[l], F, dd, xr, p.act);
    c.phase++;
    nw = prefetch_norm<1>(c, l + 1 < D.n_layers ? D.norm_in[l + 1] : D.norm_final);
    load_act_slice<R>(c, p.act, F, tag_act, xr);
    mark(c, T_POLL);
    tag_x = cur_tag(c);
    phase_linear<R>(c, D.wdown[l], dd, F, xr, x, dd, x1, dd);
    c.phase++;
  }
  // head on the last row
  gather_ll<NLX>(c, x, tag_x, c.sattn);
  mark(c, T_POLL);
  csync();
  float xh[1][32];
  norm_from_smem<1, 1>(c, c.sattn + (size_t)(R - 1) * dd, nw, D.eps, xh);
  phase_linear<1>(c, audio_head_ptr(p.m, step - 1), V, dd, xh, p.logits, V, nullptr, 0);
  c.phase++;
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
  // ---- phase 0: input embedding x = sum_k audio_emb[prev[k] + k*V]   (generation.py:156-161 + models.py:82-92),
  // computed by CTA 0 and published as tagged words
  unsigned tag_x = cur_tag(c);
  if (c.cta == 0 && p.h_in == nullptr) {
    for (int k = c.tid * 8; k < db; k += NCT * 8) {
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
#pragma unroll
      for (int e = 0; e < 8; e += 2) ll_st2(p.xa + k + e, acc[e], acc[e + 1], tag_x);
    }
  }
  c.phase++;
  float xr[2][32];
  float(&x1r)[1][32] = *reinterpret_cast<float(*)[1][32]>(&xr[0]);
  uint2* x = p.xa;
  uint2* x1 = p.xb;
  const int nchunks = (S + 127) / 128;  // <= MAX_SPLIT for S <= 2048
  constexpr int NLB = 2048 / (NCT * 2);
  NormW nw = prefetch_norm<2>(c, B.norm_in[0]);
  for (int l = 0; l < (p.h_in ? 0 : B.n_layers); ++l) {
    gather_ll<NLB>(c, x, tag_x, c.sattn);
    mark(c, T_POLL);
    csync();
    norm_from_smem<1, 2>(c, c.sattn, nw, B.eps, x1r);
    const unsigned tag_qkv = cur_tag(c);
    phase_linear<1>(c, B.wqkv[l], nqkv, db, x1r, p.qkv, nqkv, nullptr, 0);
    c.phase++;
    const unsigned tag_att = cur_tag(c);
    for (int item = c.cta; item < B.n_kv_heads * nchunks; item += c.G)
      backbone_attention_item(c, l, item % B.n_kv_heads, item / B.n_kv_heads, nchunks, pos, p.qkv, tag_qkv);
    mark(c, T_BATT);
    c.phase++;
    if (nchunks == 1) {
      gather_ll<NLB>(c, p.attn_out, tag_att, c.sattn);
      mark(c, T_POLL);
      csync();
      slice_from_smem<1>(c, c.sattn, db, x1r);
    } else {
      merged_attention_slice(c, nchunks, tag_att, x1r);
      mark(c, T_MERGE);
    }
    const unsigned tag_x1 = cur_tag(c);
    phase_linear<1>(c, B.wo[l], db, db, x1r, x1, db, x, db);
    c.phase++;
    nw = prefetch_norm<2>(c, B.norm_post[l]);
    gather_ll<NLB>(c, x1, tag_x1, c.sattn);
    mark(c, T_POLL);
    csync();
    norm_from_smem<1, 2>(c, c.sattn, nw, B.eps, x1r);
    const unsigned tag_act = cur_tag(c);
    phase_gate_up<1>(c, B.wgu[l], F, db, x1r, p.act);
    c.phase++;
    nw = prefetch_norm<2>(c, l + 1 < B.n_layers ? B.norm_in[l + 1] : B.norm_final);
    load_act_slice<1>(c, p.act, F, tag_act, x1r);
    mark(c, T_POLL);
    tag_x = cur_tag(c);
    phase_linear<1>(c, B.wdown[l], db, F, x1r, x, db, x1, db);
    c.phase++;
  }
  // ---- final norm -> h_last (decoder input row 0) ; codebook0 head ; sample c0
  if (p.h_in) {  // depth-only launch (first frame after a prefill): the row is given, already normalised
    for (int k = c.tid * 4; k < db; k += NCT * 4)
      *reinterpret_cast<float4*>(c.sattn + k) = __ldg(reinterpret_cast<const float4*>(p.h_in + k));
    csync();
    slice_from_smem<1>(c, c.sattn, db, x1r);
  } else {
    gather_ll<NLB>(c, x, tag_x, c.sattn);
    mark(c, T_POLL);
    csync();
    norm_from_smem<1, 2>(c, c.sattn, nw, B.eps, x1r);
  }
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
  unsigned tag_lg = cur_tag(c);
  phase_linear<1>(c, p.m.c0_head, V, db, x1r, p.logits, V, nullptr, 0);
  c.phase++;
  // draws are indexed by the position of the row whose hidden state is sampled from (the last prompt row when depth-only)
  const unsigned long long draw0 = p.draw_base + (unsigned long long)(p.h_in ? pos - 1 : pos) * (unsigned)ncb;
  int tok = sample_token(c, p.logits, V, tag_lg, draw0);
  if (c.cta == 0 && c.tid == 0) p.frame_out[0] = tok;
  // ---- depth decoder: step 1 has rows (h_last @ pos 0, embed(c0) @ pos 1)   (generation.py:56-90)
#pragma unroll
  for (int e = 0; e < 32; ++e) xr[0][e] = hrow[e];
  embed_row_slice(c, p.m.audio_emb + (size_t)tok * db, db, xr[1]);
  decoder_step<2>(c, 1, 0, xr);
  tag_lg = (c.launch << 12) | ((c.phase - 1) & 0xfffu);
  tok = sample_token(c, p.logits, V, tag_lg, draw0 + 1);
  if (c.cta == 0 && c.tid == 0) p.frame_out[1] = tok;
  for (int i = 2; i < ncb; ++i) {
    embed_row_slice(c, p.m.audio_emb + ((size_t)tok + (size_t)(i - 1) * V) * db, db, x1r[0]);
    decoder_step<1>(c, i, i, x1r);
    tag_lg = (c.launch << 12) | ((c.phase - 1) & 0xfffu);
    tok = sample_token(c, p.logits, V, tag_lg, draw0 + (unsigned)i);
    if (c.cta == 0 && c.tid == 0) p.frame_out[i] = tok;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_frame(const __grid_constant__ FrameParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2 * NSTAGES + 1];
  __shared__ __align__(16) float s_part[2 * 2 * MAXU * 4];  // [range][row][unit][4 partials]
  __shared__ __align__(16) float s_red[NCW * 8];
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
  c.phase = 0;
  const unsigned nonce0 = *reinterpret_cast<volatile unsigned*>(p.nonce);
  c.launch = (nonce0 % 0xfffffu) + 1u;
  c.aborted = false;
  c.prof_on = (p.prof != nullptr) && threadIdx.x == 0;
  for (int i = 0; i < 12; ++i) c.t_acc[i] = 0;
  c.t_last = (p.prof != nullptr) ? (unsigned long long)clock64() : 0ull;
  const unsigned long long t_begin = c.t_last;
  if (threadIdx.x == 0) {
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
    // CTA 0 can only get here after every CTA took part in every phase, i.e. after every CTA read the nonce
    if (c.cta == 0 && c.tid == 0) *p.nonce = nonce0 + 1u;
    if (p.prof != nullptr && c.tid == 0) {
      for (int i = 0; i < 12; ++i) p.prof[(size_t)blockIdx.x * 16 + i] = c.t_acc[i];
      p.prof[(size_t)blockIdx.x * 16 + 12] = (unsigned long long)clock64() - t_begin;
      p.prof[(size_t)blockIdx.x * 16 + 13] = c.phase;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side (per format)
static size_t frame_ws_words(const csmb_model* m) {
  const csmb_llama &B = m->backbone, &D = m->decoder;
  const size_t nqkv_b = (size_t)(B.n_heads + 2 * B.n_kv_heads) * B.head_dim, nqkv_d = (size_t)(D.n_heads + 2 * D.n_kv_heads) * D.head_dim;
  const size_t ff = B.d_ff > D.d_ff ? B.d_ff : D.d_ff;
  size_t w = 0;
  w += 2 * (size_t)B.d_model;                                        // xa, xb
  w += 2 * (nqkv_b > nqkv_d ? nqkv_b : nqkv_d);                       // qkv
  w += (size_t)B.n_heads * MAX_SPLIT * PSTRIDE + (size_t)B.d_model;   // attn_part, attn_out
  w += 2 * ff;                                                        // act
  w += (size_t)m->audio_vocab + 64;                                   // logits
  w += 4 * (size_t)D.d_model;                                         // dxa, dxb
  return w;
}

size_t frame_workspace_bytes(const csmb_model* m, int device) {
  if (!m) return 0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 384;
  const size_t dec_kv = (size_t)sms * m->decoder.n_layers * 32 * KROW * sizeof(float);  // private copy per CTA
  return frame_ws_words(m) * sizeof(uint2) + (size_t)m->backbone.d_model * sizeof(float) + dec_kv + 256 + 16384;
}

int launch_frame(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table,
                        const int32_t* prev_frame, const float* h_in, const int32_t* pos, int32_t* frame,
                        const csmb_sampler* sampler, uint64_t draw_base, uint32_t seq, const csmb_frame_opts* opts,
                        void* workspace, size_t workspace_bytes, int32_t* status, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(m && pos && frame && sampler && workspace);
  CSMB_REQUIRE(h_in || (kv_pool && block_table && prev_frame));
  const csmb_llama &B = m->backbone, &D = m->decoder;
  const bool shape_ok = B.d_model == 2048 && B.n_heads == 32 && B.n_kv_heads == 8 && B.head_dim == 64 &&
                        D.d_model == 1024 && D.n_heads == 8 && D.n_kv_heads == 2 && D.head_dim == 128 &&
                        D.d_ff == 8192 && B.d_ff == 8192 && m->n_codebooks <= 32 && m->n_codebooks >= 3 &&
                        m->audio_vocab <= 9 * NCT;
  if (!shape_ok) return CSMB_ERR_UNSUPPORTED;
  // fused samplers: greedy; temperature with optional top-k, top-p and / or min-p.  Only min-p with min_tokens_to_keep > 1
  // (it needs the sorted order) stays on the per-op path (csmb_decode_frame).
  if (sampler->temperature != 0.f && sampler->min_p > 0.f && sampler->min_keep > 1) return CSMB_ERR_UNSUPPORTED;
  CSMB_REQUIRE(workspace_bytes >= frame_workspace_bytes(m, device));
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  CSMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CSMB_REQUIRE(sms >= 128 && sms <= 384);

  FrameParams p;
  p.m = *m;
  p.kv_pool = kv_pool;
  p.kv_layer_stride = kv_layer_stride;
  p.block_table = block_table;
  p.prev_frame = prev_frame;
  p.h_in = h_in;
  p.pos_ptr = pos;
  p.frame_out = frame;
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255;
  p.abort_flag = reinterpret_cast<int*>(base);
  p.nonce = reinterpret_cast<unsigned*>(base + 128);
  uint2* w = reinterpret_cast<uint2*>(base + 256);
  auto take = [&](size_t n) {
    uint2* r = w;
    w += (n + 63) & ~(size_t)63;
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
  p.logits = take(m->audio_vocab);
  p.dxa = take(2 * (size_t)D.d_model);
  p.dxb = take(2 * (size_t)D.d_model);
  float* f = reinterpret_cast<float*>(w);
  p.h_last = f;
  f += B.d_model;
  p.dec_kv = f;
  f += (size_t)sms * D.n_layers * 32 * KROW;
  CSMB_REQUIRE(reinterpret_cast<uintptr_t>(f) <= reinterpret_cast<uintptr_t>(workspace) + workspace_bytes);
  // tags are 20-bit launch nonces + 12-bit phase numbers; the workspace must start zeroed (tag 0 is never produced)
  p.inv_temp = sampler->temperature == 0.f ? 0.f : 1.f / sampler->temperature;
  p.top_k = (sampler->top_k > 0 && sampler->top_k < m->audio_vocab) ? sampler->top_k : 0;
  p.min_p = sampler->min_p > 0.f ? sampler->min_p : 0.f;
  p.top_p = (sampler->top_p > 0.f && sampler->top_p < 1.f) ? sampler->top_p : 0.f;
  p.seed_lo = (uint32_t)sampler->seed;
  p.seed_hi = (uint32_t)(sampler->seed >> 32);
  p.draw_base = draw_base;
  p.seq = seq;
  p.status = status;
  p.prof = opts ? opts->prof : nullptr;
  p.dbg = opts ? (opts->flags & 0xff) : 0;
  p.pf_max = opts ? opts->prefetch_stages : 0;
  p.pf_interval = (opts && opts->prefetch_interval > 0) ? opts->prefetch_interval : 700;
  const int want_ctas = opts ? opts->ctas : 0;

  static const size_t dyn_smem = (size_t)NSTAGES * STAGE_BYTES + KVS_BYTES;  // ring + decoder KV staging
  CSMB_REQUIRE(m->weight_format == FMT);
  const void* kern = (const void*)k_frame;
  CSMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
  CSMB_CUDA(cudaMemsetAsync(reinterpret_cast<void*>(base), 0, 64, st));  // abort flag only; the nonce persists
  void* args[] = {&p};
  // fewest CTAs for which every linear's per-CTA slice fits the partial-sum buffer (MAXU units) and its rows fit one
  // finalising thread each (rows * 2 <= NCT)
  auto fits = [&](int G) {
    auto ok = [&](int N, int K) { const int rows = (N + G - 1) / G; return rows * (K / UNIT) <= MAXU && rows * 2 <= NCT; };
    return ok((int)nqkv_b, B.d_model) && ok(B.d_model, B.n_heads * B.head_dim) && ok(B.d_ff, B.d_model) &&
           ok(B.d_model, B.d_ff) && ok(m->audio_vocab, B.d_model) && ok(D.d_model, B.d_model) && ok((int)nqkv_d, D.d_model) &&
           ok(D.d_model, D.n_heads * D.head_dim) && ok(D.d_ff, D.d_model) && ok(D.d_model, D.d_ff) && ok(m->audio_vocab, D.d_model);
  };
  // default grid: the largest CTA count <= SMs that divides every matrix's row count, so that all slices are equal and
  // whole ring stages (csm_1b on 148 SMs: 128 CTAs, measured 3% faster than 148 uneven slices); the SMs left over
  // run the codec's streaming step of the previous frame concurrently (second stream).
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  int grid = sms;
  if (want_ctas > 0) {
    grid = want_ctas < sms ? want_ctas : sms;
  } else {
    const int g = gcd(gcd(gcd(B.d_model, B.d_ff), gcd(D.d_model, D.d_ff)), gcd((int)nqkv_b, (int)nqkv_d));
    for (int c = sms; c >= (sms * 3) / 4; --c)
      if (g % c == 0) { grid = c; break; }
  }
  while (grid < sms && !fits(grid)) ++grid;
  CSMB_REQUIRE(fits(grid));
  CSMB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(NTHREADS), args, dyn_smem, st));
  count_launch();
  return CSMB_OK;
}

}  // namespace FK_NS
}  // namespace csmb

#if CSMB_FK_FMT == 0
// the e4m3 instantiation of everything above lives in frame_kernel_e4m3.cu
namespace csmb {
namespace fk_e4m3 {
int launch_frame(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table, const int32_t* prev_frame,
                 const float* h_in, const int32_t* pos, int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, uint32_t seq,
                 const csmb_frame_opts* opts, void* workspace, size_t workspace_bytes, int32_t* status, int device, void* stream);
}
}  // namespace csmb

using namespace csmb;

// weight format of the model -> the kernel compiled for it
template <typename... A>
static int launch_frame(const csmb_model* m, A... a) {
  if (m && m->weight_format == CSMB_WEIGHTS_E4M3) return fk_e4m3::launch_frame(m, a...);
  return fk_bf16::launch_frame(m, a...);
}

extern "C" {

size_t csmb_frame_workspace_bytes(const csmb_model* m, int device) { return fk_bf16::frame_workspace_bytes(m, device); }

int csmb_frame_b1(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table,
                  const int32_t* prev_frame, const int32_t* pos, int32_t* frame, const csmb_sampler* sampler,
                  uint64_t draw_base, const csmb_frame_opts* opts, void* workspace, size_t workspace_bytes, int32_t* status,
                  int device, void* stream) {
  return launch_frame(m, kv_pool, kv_layer_stride, block_table, prev_frame, nullptr, pos, frame, sampler, draw_base, 0u, opts,
                      workspace, workspace_bytes, status, device, stream);
}

int csmb_frame_b1_slot(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table_row,
                       const int32_t* prev_frame_row, const int32_t* pos, int32_t* frame_row, const csmb_sampler* sampler,
                       uint64_t draw_base, int seq_index, const csmb_frame_opts* opts, void* workspace, size_t workspace_bytes,
                       int32_t* status, int device, void* stream) {
  if (seq_index < 0) return CSMB_ERR_INVALID;
  return launch_frame(m, kv_pool, kv_layer_stride, block_table_row, prev_frame_row, nullptr, pos, frame_row, sampler,
                      draw_base, (uint32_t)seq_index, opts, workspace, workspace_bytes, status, device, stream);
}

int csmb_frame_b1_depth(const csmb_model* m, const float* h_last, const int32_t* pos, int32_t* frame,
                        const csmb_sampler* sampler, uint64_t draw_base, const csmb_frame_opts* opts, void* workspace,
                        size_t workspace_bytes, int32_t* status, int device, void* stream) {
  if (!h_last) return CSMB_ERR_INVALID;
  return launch_frame(m, nullptr, 0, nullptr, nullptr, h_last, pos, frame, sampler, draw_base, 0u, opts, workspace,
                      workspace_bytes, status, device, stream);
}

}  // extern "C"
#endif  // CSMB_FK_FMT == 0
