// Fused batched decode frame (sm_100a): what csmb_decode_frame computes for B sequences in lock-step — generate_frame
// with T = 1 (csm_mlx/generation.py:21-92) plus the next-input construction (:156-161) — as a chain of ~1 100 kernels
// per frame instead of ~3 700:
//
//   * every Linear is ONE tcgen05 launch (k_gemm_part_t): 128 weight rows x all token rows per CTA tile, fp32 TMEM
//     accumulators, TMA-fed 128B-swizzled stages, split along K so that ~148 CTAs stream the matrix; it reads the
//     activations as bf16 hi + lo planes (x = hi + lo to 2^-17) that its PRODUCER kernel already wrote, and leaves
//     fp32 split-K partials [S][R][N];
//   * everything between two Linears is ONE kernel that starts by summing those partials in fixed order:
//       k_resid_norm_split    partial sum + residual add + RMSNorm + hi/lo split
//       k_attn_decode_fused   partial sum + RoPE + paged KV append + GQA attention + hi/lo split
//       (SwiGLU + hi/lo split run in the gate|up Linear's own epilogue, k_gemm_part_t<true>; k_swiglu_split is the
//        separate-launch form kept for A/B runs)
//       k_sample_embed        partial sum (logits) + sampling + next codebook embedding gather (+ hi/lo split)
//       k_frame_embed_norm    previous frame -> summed audio embeddings + first RMSNorm + hi/lo split
//   * all launches are chained with programmatic dependent launch: a GEMM CTA sets up its barriers / TMEM and already
//     streams its first weight stages (weights do not depend on the previous kernel) while its producer kernel is
//     still running; only the activation loads wait (griddepcontrol.wait).
//
// Arithmetic is that of the per-op path (fp32 accumulation, fixed summation orders): parity tests compare both with
// the batch-1 kernels and the oracle token for token.
#include <math.h>

#include <utility>

#include "tc.cuh"

namespace csmb {

constexpr int BF_MAX_STAGES = 10;
#ifndef CSMB_BF_NI
#define CSMB_BF_NI 2
#endif
constexpr int BF_NI = CSMB_BF_NI;                 // MMA-issuing warps (= fp32 accumulators) per Linear CTA: 1 or 2
constexpr int BF_THREADS = 256;                  // warp 0: TMA producer; warps 1 .. BF_NI: MMA issuers; warps 3..6: epilogue; all 8 warps share the epilogue (below)
static_assert(BF_NI <= 2, "warp roles below");
static_assert(BF_MAX_STAGES % BF_NI == 0, "a pipeline stage must always be consumed by the same issuer");
constexpr size_t BF_SMEM_BUDGET = 200 * 1024;

// Split-K geometry is a constant of the library (see bf_pick_split): at least BF_MIN_KBLOCKS 64-wide K blocks per CTA and
// about BF_SPLIT_CTAS CTAs per Linear.  No process-global tuning state: per-call switches travel in csmb_chain_opts.
constexpr int BF_MIN_KBLOCKS = 4;
constexpr int BF_SPLIT_CTAS = 148;

struct ChainCfg {
  int pdl;          // programmatic dependent launch
  int dbg;          // bits 0, 1: timing experiments (GpArgs::dbg); bit 2: separate k_swiglu_split launch (A/B)
  size_t smem;      // shared-memory budget of a Linear CTA (pipeline stages)
};
static ChainCfg chain_cfg(const csmb_chain_opts* o) {
  ChainCfg c{1, 0, BF_SMEM_BUDGET};
  if (o) {
    c.pdl = o->no_pdl ? 0 : 1;
    c.dbg = o->flags;
    if (o->smem_kb >= 48 && (size_t)o->smem_kb * 1024 <= BF_SMEM_BUDGET) c.smem = (size_t)o->smem_kb * 1024;
  }
  return c;
}

// ---------------------------------------------------------------------------------------------- timeline (debug builds)
// -DCSMB_TIMELINE (scripts/chain_timeline.py builds a second .so with it; never the shipped library): block (0,0,0) of every
// chain kernel records %globaltimer at entry, after griddepcontrol.wait, (Linears: accumulator complete) and at its end.
#ifdef CSMB_TIMELINE
constexpr int TL_SLOTS = 16384, TL_WORDS = 6;
__device__ unsigned long long g_tl[TL_SLOTS * TL_WORDS];
__device__ unsigned g_tl_n;
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool tl_block0() { return (blockIdx.x | blockIdx.y | blockIdx.z) == 0; }
__device__ __forceinline__ unsigned tl_enter(unsigned kind) {
  const unsigned slot = atomicAdd(&g_tl_n, 1u) % TL_SLOTS;
  g_tl[slot * TL_WORDS + 0] = kind | ((unsigned long long)(gridDim.x * gridDim.y * gridDim.z) << 8);
  g_tl[slot * TL_WORDS + 1] = tl_now();
  return slot;
}
__device__ __forceinline__ void tl_mark(unsigned slot, int i) { g_tl[slot * TL_WORDS + i] = tl_now(); }
#define TL_ENTER(kind) unsigned tl_slot = 0; if (tl_block0() && threadIdx.x == 0) tl_slot = tl_enter(kind);
#define TL_MARK(i) if (tl_block0() && threadIdx.x == 0) tl_mark(tl_slot, i);
#else
#define TL_ENTER(kind)
#define TL_MARK(i)
#endif

// ---------------------------------------------------------------------------------------------- GEMM
struct GpArgs {
  float* part;  // [S][R][N] fp32 partials
  int R, N, K, RN, nstages, S;
  int* err;
  int dbg;  // timing experiments only: bit 0 = no partial stores, bit 1 = no TMEM loads either (results are wrong)
  // GU variant only: the matrix is gate rows [0, F) then up rows [F, 2F); output planes [R][F]
  int F;
  uint16_t *out_hi, *out_lo;
  // L2 prefetch hint: bytes [pf, pf + pf_bytes) are weights a LATER Linear of the chain streams; CTA c of n asks L2 for slice c
  const char* pf;
  unsigned pf_bytes;
  int keep8;   // k: the weight stream asks L2 to keep k/8 of its lines (evict-last) instead of evict-first for all
};

// dynamic smem: [stage][ W 128x64 | Xhi RNx64 | Xlo RNx64 ] bf16, 1024-byte aligned tiles
//
// GU = true is the gate|up Linear of the MLP with SwiGLU fused into its epilogue (no split-K): a CTA owns 64 features,
// its weight tile is gate rows f0..f0+63 (UMMA rows 0..63) on top of up rows F+f0.. (rows 64..127; two 64-row TMA boxes
// per stage, the 128B swizzle only depends on the address inside a 1 KiB atom), the four epilogue warps park their TMEM
// quarters (gate: lanes 0..63, up: lanes 64..127) in the drained pipeline stages, and all 128 threads then write
// silu(gate) * up as bf16 hi + lo planes [R][F] for the down projection — the k_swiglu_split launch disappears.  Same
// sums and the same expression as k_swiglu_split over one partial: bit-identical planes.
//
// MMA issue: ONE thread issues one tcgen05.mma (M = 128, K = 16) per ~152 cycles whatever its N, a second issuing warp runs at
// the same rate beside it (scripts/micro/mma_issue.cu), and the in-kernel timeline (scripts/chain_timeline.py) showed the big
// Linears waiting for exactly that serial issue (gate|up, K = 1024: 64 instructions = 5 us after the token planes arrive).
// So the K blocks go alternately to BF_NI = 2 issuers, each with its own accumulator (TMEM columns [i * 2 RN, (i + 1) * 2 RN)),
// and the epilogue adds them in fixed order: out = (hi_0 + hi_1) + (lo_0 + lo_1).  BF_NI is a constant of the library and
// the number of stages is a multiple of it: a row's sums never depend on the row count, and a stage always belongs to the
// same issuer (an mbarrier parity wait cannot tell the second from the third use of a stage apart, so one thread must see
// every use of "its" stages in order).
//
// MODE 1 = PAIR (launched as clusters of two CTAs along the n-tile axis, csmb_chain_opts flags 65536 / 131072, bf_mode): the two
// CTAs share the token operand through ONE tcgen05.mma.cta_group::2 of M = 256 per K step, issued by the pair's leader: CTA r
// holds its own 128 weight rows and HALF of the stacked token operand (rank 0: the hi plane's RN rows, rank 1: the lo plane's)
// at the same stage offsets, so a CTA pulls W + RN rows per K block instead of W + 2 RN — with 64 sequences the token planes
// are half of what a CTA reads, every CTA re-reads all of them from L2, and the Linears of a 64-sequence step run at the L2 ->
// SM fill rate (128 CTAs x 422 KB in 5.4 us = 10 TB/s of the ~12 TB/s the chip delivers).  Both CTAs' TMA copies count their
// bytes on the LEADER's full barrier; the leader's commits are multicast to the empty / acc_full barriers of both CTAs; the
// accumulator rows of a CTA's weight tile land in its own TMEM, so the epilogue is unchanged.  Same products summed in the
// same order as MODE 0: bit-identical outputs.
//
// (A third mode — the last weight K blocks of a CTA parked in free tensor-memory columns before the previous kernel finishes and
// consumed as tcgen05.mma with A from TMEM — was built, bit-identical, and measured slower: profiles/r02_tmem_weights.md,
// r02_tmem_weights_experiment.patch.)
template <bool GU, int MODE>
__global__ void __launch_bounds__(BF_THREADS, 1)
k_gemm_part_t(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_hi,
              const __grid_constant__ CUtensorMap map_lo, const GpArgs a) {
  constexpr bool PAIR = MODE == 1;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[BF_MAX_STAGES], empty[BF_MAX_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ int gu_failed_s;
  pdl_launch_dependents();
  if (threadIdx.x == 0) gu_failed_s = 0;
#ifdef CSMB_TIMELINE
  __shared__ unsigned tl_slot_s;
  if (tl_block0() && threadIdx.x == 0) tl_slot_s = tl_enter(GU ? 2u : 1u);
#endif
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the tcgen05.mma operands in uniform
  // registers (otherwise every UTCHMMA sits in an ELECT / R2UR.BROADCAST loop: ~240 instead of ~150 cycles per instruction)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  // (timing experiment, flag 2048: token-plane boxes out of bounds -> TMA zero-fills them without reading L2; results are wrong)
  const int n0 = blockIdx.x * TC_BM, r0 = (a.dbg & 2048) ? (1 << 20) : blockIdx.y * a.RN;
  const int RN = a.RN;
  const uint32_t w_bytes = TC_BM * TC_BK * 2, x_bytes = (uint32_t)RN * TC_BK * 2;
  // a stage holds the weight tile and the token rows THIS CTA loads (PAIR: one plane); stage_tx = bytes counted on a full barrier
  const uint32_t x_off = w_bytes, stage_bytes = (w_bytes + (PAIR ? 1u : 2u) * x_bytes + 1023u) & ~1023u;
  const uint32_t stage_tx = PAIR ? 2u * (w_bytes + x_bytes) : w_bytes + 2u * x_bytes;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;          // rank 0 = the pair's leader (issues the MMAs)
  const uint32_t full0 = PAIR ? cluster_map_shared(s32(&full[0]), 0u) : 0u;   // the leader's full[0] (cluster address)
  const CUtensorMap* map_x = (PAIR && crank != 0u) ? &map_lo : &map_hi;       // PAIR: the plane this CTA loads
  const int nk_total = a.K / TC_BK, NS = a.nstages;
  const int kb0 = (int)(((long long)nk_total * blockIdx.z) / a.S), kb1 = (int)(((long long)nk_total * (blockIdx.z + 1)) / a.S);
  const int nk = kb1 - kb0;
  // The token rows' hi and lo planes sit one after the other in a stage and are ONE operand of 2 RN rows: a single
  // tcgen05.mma per K step leaves W.hi in accumulator columns [0, RN) and W.lo in [RN, 2 RN); the epilogue adds the halves.
  // (One thread issues one tcgen05.mma per ~150 cycles whatever its N — profiles/r02_mma_issue_microbench.log — and two
  // instructions per K step made these Linears issue-bound below the HBM stream rate.)
  const int ni = nk < BF_NI ? nk : BF_NI;   // issuers with work
  uint32_t ncols = 32;   // RN % 16 == 0 and the epilogue loads 16 columns at a time: no load reads past an accumulator
  while ((int)ncols < BF_NI * 2 * RN) ncols <<= 1;
  const uint64_t wpol = a.keep8 > 0 ? l2_policy_keep_fraction((float)a.keep8 * 0.125f) : l2_policy_evict_first();
  auto load_w = [&](unsigned char* dst, int kb, uint64_t* bar) {
    if (PAIR) {
      const uint32_t fb = full0 + (uint32_t)((bar - full) * sizeof(uint64_t));
      if (GU) {
        tma_load_2d_pair_hint(dst, &map_w, kb * TC_BK, n0 / 2, fb, wpol);
        tma_load_2d_pair_hint(dst + (TC_BM / 2) * TC_BK * 2, &map_w, kb * TC_BK, a.F + n0 / 2, fb, wpol);
      } else {
        tma_load_2d_pair_hint(dst, &map_w, kb * TC_BK, n0, fb, wpol);
      }
    } else if (a.dbg & 32) {   // A/B: default L2 policy
      if (GU) {
        tma_load_2d(dst, &map_w, kb * TC_BK, n0 / 2, bar);
        tma_load_2d(dst + (TC_BM / 2) * TC_BK * 2, &map_w, kb * TC_BK, a.F + n0 / 2, bar);
      } else {
        tma_load_2d(dst, &map_w, kb * TC_BK, n0, bar);
      }
    } else if (GU) {
      tma_load_2d_hint(dst, &map_w, kb * TC_BK, n0 / 2, bar, wpol);                                  // gate rows f0 .. f0+63
      tma_load_2d_hint(dst + (TC_BM / 2) * TC_BK * 2, &map_w, kb * TC_BK, a.F + n0 / 2, bar, wpol);  // up rows F+f0 ..
    } else {
      tma_load_2d_hint(dst, &map_w, kb * TC_BK, n0, bar, wpol);
    }
  };
  // the token rows of K block kb into a stage (after griddepcontrol.wait: the previous kernel writes them)
  auto load_x = [&](unsigned char* st, int kb, uint64_t* bar) {
    if (PAIR) {
      tma_load_2d_pair(st + x_off, map_x, kb * TC_BK, r0, full0 + (uint32_t)((bar - full) * sizeof(uint64_t)));
    } else {
      tma_load_2d(st + x_off, &map_hi, kb * TC_BK, r0, bar);
      tma_load_2d(st + x_off + x_bytes, &map_lo, kb * TC_BK, r0, bar);
    }
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < BF_MAX_STAGES; ++i) {
      tc_mbar_init(&full[i], 1);
      tc_mbar_init(&empty[i], 1);
    }
    tc_mbar_init(&acc_full, (uint32_t)ni);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {   // both CTAs of the pair allocate the same columns (nothing else holds TMEM on an SM a Linear CTA fits on)
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(ncols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_s)), "r"(ncols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before any copy or commit signals them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // Warp roles: warp 0 = TMA producer, warps 1 .. BF_NI = MMA issuers (one lane each); then ALL eight warps run the epilogue —
  // a warp reads the TMEM lane quarter warp % 4, so every quarter has two warps (3..6 take the even 16-column chunks of their
  // quarter, 0, 1, 2 and 7 the odd ones): with 64 token rows the epilogue is four chunks of TMEM loads, adds and 16 row
  // stores per thread, and four warps alone were its bottleneck (same sums, same stores: bit-identical).
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      if (PAIR) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(map_x) : "memory");
      } else {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
      }
      // weights do not depend on the previous kernel: fill the pipeline with them before waiting for it
      const int pre = nk < NS ? nk : NS;
      for (int kb = 0; kb < pre; ++kb) {
        if (crank == 0u) tc_mbar_expect_tx(&full[kb], stage_tx);   // (the peer's bytes may land first: the phase still needs this arrival)
        load_w(smem + (size_t)kb * stage_bytes, kb0 + kb, &full[kb]);
      }
      // L2 prefetch hints (no result depends on them).  (a) This CTA's slice of the NEXT big Linear's matrix, whose CTAs cannot
      // become resident before this kernel's CTAs leave: +1 % at B <= 16, +1.5 % at B = 64.  (b) The CTA's own weight blocks
      // beyond the ring (TMA tensor prefetch, same boxes): measured useless at small batches and -1.5 % at B = 64 on top of
      // (a) — off unless flag 64 asks for it (A/B).
      if (a.dbg & 64) {
        for (int kb = pre; kb < nk; ++kb) {
          if (GU) {
            tma_prefetch_2d(&map_w, (kb0 + kb) * TC_BK, n0 / 2);
            tma_prefetch_2d(&map_w, (kb0 + kb) * TC_BK, a.F + n0 / 2);
          } else {
            tma_prefetch_2d(&map_w, (kb0 + kb) * TC_BK, n0);
          }
        }
      }
      if (a.pf_bytes != 0 && !(a.dbg & 128)) {
        const unsigned nc = gridDim.x * gridDim.y * gridDim.z, c = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const unsigned chunk = (((a.pf_bytes + nc - 1) / nc) + 127u) & ~127u;
        const unsigned long long off = (unsigned long long)c * chunk;
        if (off < a.pf_bytes) {
          const unsigned end = a.pf_bytes - off < chunk ? a.pf_bytes : (unsigned)(off + chunk);
          for (unsigned o = (unsigned)off; o < end; o += 16384u) bulk_prefetch_l2(a.pf + o, (end - o < 16384u ? end - o : 16384u) & ~15u);
        }
      }
      pdl_wait();
#ifdef CSMB_TIMELINE
      if (tl_block0()) tl_mark(tl_slot_s, 2);
#endif
      for (int kb = 0; kb < pre; ++kb) load_x(smem + (size_t)kb * stage_bytes, kb0 + kb, &full[kb]);
      for (int kb = pre; kb < nk; ++kb) {
        const int s = kb % NS;
        const uint32_t par = (kb / NS) & 1;
        if (!tc_mbar_wait(&empty[s], par ^ 1, a.err)) break;
        unsigned char* st = smem + (size_t)s * stage_bytes;
        if (crank == 0u) tc_mbar_expect_tx(&full[s], stage_tx);
        load_w(st, kb0 + kb, &full[s]);
        load_x(st, kb0 + kb, &full[s]);
      }
    }
  } else if (warp <= BF_NI) {
    // ===== MMA issuers: issuer i = warp - 1 takes K blocks i, i + BF_NI, ... into accumulator i =====
    const int iss = warp - 1;
    if (lane == 0 && iss < ni && crank == 0u) {
      const uint32_t idesc = PAIR ? umma_idesc_pair(2 * RN) : umma_idesc(2 * RN);
      const uint32_t tacc = tmem_base + (uint32_t)(iss * 2 * RN);
      bool ok = true;
      for (int kb = iss; kb < nk && ok; kb += BF_NI) {
        const int s = kb % NS;
        const uint32_t par = (kb / NS) & 1;
        ok = tc_mbar_wait(&full[s], par, a.err);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = s32(smem + (size_t)s * stage_bytes);
        const uint64_t da = umma_desc(sa), dx = umma_desc(sa + x_off);   // hi rows then lo rows: 2 RN rows, 128 B apart
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t koff = (uint64_t)((k * 32) >> 4);
          if (PAIR) umma_f16_pair(tacc, da + koff, dx + koff, idesc, (kb != iss) || (k != 0));
          else umma_f16(tacc, da + koff, dx + koff, idesc, (kb != iss) || (k != 0));
        }
        if (PAIR) umma_commit_pair(&empty[s]);
        else umma_commit(&empty[s]);
      }
      if (PAIR) umma_commit_pair(&acc_full);
      else umma_commit(&acc_full);
    }
  }
  __syncwarp();
  {
    // ===== epilogue: TMEM lane quarter warp % 4; chunk parity by warp =====
    pdl_wait();  // `part` may still be read by the previous kernel of the stream
    const int quarter = warp & 3;
    const int cpar = (warp >= 3 && warp <= 6) ? 0 : 1;   // which 16-column chunks of the quarter this warp takes
    const bool ok = tc_mbar_wait(&acc_full, 0, a.err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef CSMB_TIMELINE
    if (tl_block0() && threadIdx.x == 96) tl_mark(tl_slot_s, 3);
#endif
    const uint32_t tq = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // 16 consecutive token columns c0 .. c0+15 of this thread's weight row: (sum of the issuers' W.hi) + (sum of their W.lo),
    // each in issuer order; all 2 * ni TMEM loads are in flight together (one tcgen05.wait::ld)
    auto load_cols = [&](int c0, float (&o)[16]) {
      uint32_t vh[BF_NI][16], vl[BF_NI][16];
#pragma unroll
      for (int i = 0; i < BF_NI; ++i)
        if (i < ni) {
          tmem_ld16_nowait(tq + (uint32_t)(i * 2 * RN + c0), vh[i]);
          tmem_ld16_nowait(tq + (uint32_t)(i * 2 * RN + RN + c0), vl[i]);
        }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float h = __uint_as_float(vh[0][j]), l = __uint_as_float(vl[0][j]);
#pragma unroll
        for (int i = 1; i < BF_NI; ++i)
          if (i < ni) {
            h += __uint_as_float(vh[i][j]);
            l += __uint_as_float(vl[i][j]);
          }
        o[j] = h + l;
      }
    };
    const int n = n0 + quarter * 32 + lane;
    if (GU) {
      // park gate (quarters 0, 1) and up (quarters 2, 3) values as [2][RN][64] fp32 in the drained stages
      float* ex = reinterpret_cast<float*>(smem);
      if (ok) {
        float* mine = ex + (size_t)(quarter >> 1) * RN * 64 + (quarter & 1) * 32 + lane;
        for (int c0 = cpar * 16; c0 < RN; c0 += 32) {
          float o[16];
          load_cols(c0, o);
#pragma unroll
          for (int j = 0; j < 16; ++j) mine[(size_t)(c0 + j) * 64] = o[j];
        }
      }
    } else if (ok) {
      float* dst0 = a.part + (size_t)blockIdx.z * a.R * a.N + n;
      for (int c0 = cpar * 16; c0 < RN; c0 += 32) {
        float o[16];
        if (a.dbg & 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = 0.f;
        } else {
          load_cols(c0, o);
        }
        if (n < a.N && !(a.dbg & 1)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = r0 + c0 + j;
            if (r < a.R) dst0[(size_t)r * a.N] = o[j];
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (GU && !ok) gu_failed_s = 1;   // a bounded wait timed out (a.err is set): the second half below stays off
  }
  __syncthreads();
  if (GU) {
    // Second half of the SwiGLU epilogue on ALL warps of the CTA (the producer and the MMA warps are done by now): the four
    // epilogue warps alone are ALU-bound on the exponentials, divisions and hi / lo splits of up to 64 rows x 64 features.
    const float* ex = reinterpret_cast<const float*>(smem);
    if (gu_failed_s == 0) {
      const int f0 = n0 / 2;
      const int rows = min(RN, a.R - r0);
      for (int idx = threadIdx.x; idx < rows * 16; idx += BF_THREADS) {
        const int t = idx >> 4, f4 = (idx & 15) * 4;
        const float4 g = *reinterpret_cast<const float4*>(ex + (size_t)t * 64 + f4);
        const float4 u = *reinterpret_cast<const float4*>(ex + (size_t)(RN + t) * 64 + f4);
        const size_t o = (size_t)(r0 + t) * a.F + f0 + f4;
        store_split4(a.out_hi + o, a.out_lo + o, (g.x / (1.f + expf(-g.x))) * u.x, (g.y / (1.f + expf(-g.y))) * u.y,
                     (g.z / (1.f + expf(-g.z))) * u.z, (g.w / (1.f + expf(-g.w))) * u.w);
      }
    }
  }
#ifdef CSMB_TIMELINE
  if (tl_block0() && threadIdx.x == 96) tl_mark(tl_slot_s, 4);
#endif
  if (PAIR) cluster_sync_all();   // both CTAs are done with their accumulators and with each other's shared memory
  if (warp == 1) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// Split-K factor of a Linear: a function of its shape (N, K) ONLY — never of the number of rows, of the device or of a
// tuning knob — so that a row's K blocks are summed in the same order in every batch: a sequence's tokens do not depend on
// how many other sequences share the step (generation.py:139-161 is a batch-1 loop) or on how a job is sharded over GPUs.
static int bf_pick_split(int N, int K) {
  const int tiles = cdiv(N, TC_BM);
  const int nk = K / TC_BK;
  int S = BF_SPLIT_CTAS / tiles;
  const int cap = nk / BF_MIN_KBLOCKS;
  S = S < cap ? S : cap;
  return S < 1 ? 1 : S;
}

// ---------------------------------------------------------------------------------------------- fused element kernels
struct PartIn {
  const float* p;  // partials [S][R][ld]
  int S;
  size_t stride;   // R * ld
  int ld;
};
// Asynchronous global -> shared copies (LDGSTS): no register destination, so a thread can have any number of them in flight —
// ptxas pipelines ordinary loads feeding a sum three to five at a time, which turned "sum S partials" into S / 4 dependent
// L2 round trips (profiles/r02_chain_timeline.md).  A thread that reads back only what it copied itself needs
// cp_async_wait_all() and no block barrier.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Fixed-order sums over the S split-K partials.  Loads go out in independent batches of 8 through the read-only path
// (one L2 round trip per batch instead of one per partial: these kernels are pure latency chains).
__device__ __forceinline__ float part_sum1(const PartIn& pi, size_t off) {
  float v = 0.f;
  for (int z0 = 0; z0 < pi.S; z0 += 8) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = (z0 + u < pi.S) ? __ldg(pi.p + (size_t)(z0 + u) * pi.stride + off) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (z0 + u < pi.S) v += t[u];
  }
  return v;
}
__device__ __forceinline__ float2 part_sum2(const PartIn& pi, size_t off) {
  float2 v = make_float2(0.f, 0.f);
  for (int z0 = 0; z0 < pi.S; z0 += 8) {
    float2 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      t[u] = (z0 + u < pi.S) ? __ldg(reinterpret_cast<const float2*>(pi.p + (size_t)(z0 + u) * pi.stride + off)) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (z0 + u < pi.S) {
        v.x += t[u].x;
        v.y += t[u].y;
      }
  }
  return v;
}
__device__ __forceinline__ float4 part_sum4(const PartIn& pi, size_t off) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z0 = 0; z0 < pi.S; z0 += 8) {
    float4 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      t[u] = (z0 + u < pi.S) ? __ldg(reinterpret_cast<const float4*>(pi.p + (size_t)(z0 + u) * pi.stride + off))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (z0 + u < pi.S) {
        v.x += t[u].x;
        v.y += t[u].y;
        v.z += t[u].z;
        v.w += t[u].w;
      }
  }
  return v;
}
// sum over a 256-thread block, broadcast to every thread (red: 8 floats of shared memory)
__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  __syncthreads();
  return tot;
}

// sum of squares of four values on top of ss, explicit fmas: every kernel that normalises a row rounds identically
__device__ __forceinline__ float sumsq4(const float4 s, float ss) {
  ss = __fmaf_rn(s.x, s.x, ss);
  ss = __fmaf_rn(s.y, s.y, ss);
  ss = __fmaf_rn(s.z, s.z, ss);
  return __fmaf_rn(s.w, s.w, ss);
}
// y = RMSNorm(v) * w for one row of d = NV * 1024 values held as v[j] = columns threadIdx.x*4 + j*1024 (256 threads),
// written as bf16 hi/lo planes (and fp32 y32 if given).  ss = this thread's sum of squares.
template <int NV>
__device__ __forceinline__ void norm_split_row(const float4 (&v)[NV], float ss, const float* __restrict__ w, float eps,
                                               uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, float* __restrict__ y32,
                                               float* red) {
  constexpr int d = NV * 1024;
  const float scale = rsqrtf(block_sum256(ss, red) / (float)d + eps);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = threadIdx.x * 4 + j * 1024;
    const float4 g = *reinterpret_cast<const float4*>(w + c);
    const float4 y = make_float4(v[j].x * scale * g.x, v[j].y * scale * g.y, v[j].z * scale * g.z, v[j].w * scale * g.w);
    store_split4(hi + c, lo + c, y.x, y.y, y.z, y.w);
    if (y32) *reinterpret_cast<float4*>(y32 + c) = y;
  }
}

// x[b][:] = sum_k audio_emb[prev[b][k] + k*V][:]   (generation.py:156-161 + models.py:82-92 + generation.py:32-36),
// then RMSNorm(w) -> hi/lo.  One block per sequence; d == 2048 (8 channels per thread).
__global__ void __launch_bounds__(256) k_frame_embed_norm(const int32_t* __restrict__ prev, const uint16_t* __restrict__ audio_emb,
                                                          int ncb, int V, int d, float* __restrict__ x,
                                                          const float* __restrict__ w, float eps,
                                                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                          const float* __restrict__ x_override,
                                                          const uint8_t* __restrict__ use_override) {
  __shared__ float red[8];
  pdl_launch_dependents();
  TL_ENTER(6u)
  pdl_wait();
  TL_MARK(2)
  const int b = blockIdx.x, c = threadIdx.x * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // a sequence admitted this step: its input row is the (already embedded) last row of its prompt, not a previous frame
  const bool over = use_override != nullptr && use_override[b] != 0;
  if (over) {
    const float4 a0 = *reinterpret_cast<const float4*>(x_override + (size_t)b * d + c);
    const float4 a1 = *reinterpret_cast<const float4*>(x_override + (size_t)b * d + c + 4);
    acc[0] = a0.x; acc[1] = a0.y; acc[2] = a0.z; acc[3] = a0.w; acc[4] = a1.x; acc[5] = a1.y; acc[6] = a1.z; acc[7] = a1.w;
  }
  for (int s = 0; s < (over ? 0 : ncb); ++s) {
    int t = prev[(size_t)b * ncb + s];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    const uint4 q = *reinterpret_cast<const uint4*>(audio_emb + ((size_t)t + (size_t)s * V) * d + c);
    acc[0] += bf16lo(q.x); acc[1] += bf16hi(q.x); acc[2] += bf16lo(q.y); acc[3] += bf16hi(q.y);
    acc[4] += bf16lo(q.z); acc[5] += bf16hi(q.z); acc[6] += bf16lo(q.w); acc[7] += bf16hi(q.w);
  }
  float4* xo = reinterpret_cast<float4*>(x + (size_t)b * d + c);
  xo[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  xo[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) ss += acc[e] * acc[e];
  const float scale = rsqrtf(block_sum256(ss, red) / (float)d + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(w + c), g1 = *reinterpret_cast<const float4*>(w + c + 4);
  store_split4(hi + (size_t)b * d + c, lo + (size_t)b * d + c, acc[0] * scale * g0.x, acc[1] * scale * g0.y,
               acc[2] * scale * g0.z, acc[3] * scale * g0.w);
  store_split4(hi + (size_t)b * d + c + 4, lo + (size_t)b * d + c + 4, acc[4] * scale * g1.x, acc[5] * scale * g1.y,
               acc[6] * scale * g1.z, acc[7] * scale * g1.w);
  TL_MARK(4)
}

// Row rin = blockIdx.x * row_mul + row_add (or row_idx[blockIdx.x] if given) of the residual stream:  mode 1: x = sum(part);  mode 2: x += sum(part);
// mode 0: x unchanged.  Then y = RMSNorm(x) * w -> hi/lo row blockIdx.x (and fp32 y32 if given).  d = NV * 1024.
template <int NV>
__global__ void __launch_bounds__(256) k_resid_norm_split(float* __restrict__ x, int ldx, PartIn part, int mode,
                                                          const float* __restrict__ w, float eps,
                                                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                          float* __restrict__ y32, int row_mul, int row_add,
                                                          const int32_t* __restrict__ row_idx, int staged) {
  extern __shared__ __align__(16) float pst[];   // staged: [S][d] partials of this row
  __shared__ float red[8];
  pdl_launch_dependents();
  TL_ENTER(3u)
  pdl_wait();
  TL_MARK(2)
  constexpr int d = NV * 1024;
  const int rin = row_idx ? row_idx[blockIdx.x] : blockIdx.x * row_mul + row_add, rout = blockIdx.x;
  if (staged && mode != 0) {
    // all S x NV 16-byte pieces of this thread's columns in flight at once
    for (int z = 0; z < part.S; ++z)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = threadIdx.x * 4 + j * 1024;
        cp_async16(pst + (size_t)z * d + c, part.p + (size_t)z * part.stride + (size_t)rin * part.ld + c);
      }
  }
  float4 v[NV];
  float ss = 0.f;
  float4 xo[NV];
  if (mode != 1) {
#pragma unroll
    for (int j = 0; j < NV; ++j) xo[j] = *reinterpret_cast<const float4*>(x + (size_t)rin * ldx + threadIdx.x * 4 + j * 1024);
  }
  if (staged && mode != 0) cp_async_wait_all();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = threadIdx.x * 4 + j * 1024;
    float* xp = x + (size_t)rin * ldx + c;
    float4 s;
    if (mode == 0) {
      s = xo[j];
    } else {
      if (staged) {
        s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < part.S; ++z) {
          const float4 t = *reinterpret_cast<const float4*>(pst + (size_t)z * d + c);
          s.x += t.x;
          s.y += t.y;
          s.z += t.z;
          s.w += t.w;
        }
      } else {
        s = part_sum4(part, (size_t)rin * part.ld + c);
      }
      if (mode == 2) s = make_float4(xo[j].x + s.x, xo[j].y + s.y, xo[j].z + s.z, xo[j].w + s.w);
      *reinterpret_cast<float4*>(xp) = s;
    }
    v[j] = s;
    ss = sumsq4(s, ss);
  }
  norm_split_row<NV>(v, ss, w, eps, hi + (size_t)rout * d, lo + (size_t)rout * d, y32 ? y32 + (size_t)rout * d : nullptr, red);
  TL_MARK(4)
}

// One block per (sequence b, kv head), warp g = query head kvh*G + g.  Stage 0 (all threads): sum the qkv partials of the
// sequence's rps new rows (positions pos .. pos+rps-1) for this group's G query heads, its k head and its v head, rotate
// q and k (attention.py:119-177), keep q in shared memory and append k/v to the paged cache (:236-237).  Then each
// warp attends over positions 0..pos of its head (attention.py:242-249).  Output rows [b*rps + i][H*HD] as hi/lo.
template <int HD>
__global__ void __launch_bounds__(256) k_attn_decode_fused(PartIn qkv, const float* __restrict__ rope, float* pool,
                                                           const int32_t* __restrict__ block_table, int max_pages,
                                                           const int32_t* __restrict__ pos_arr, int pos0, int rps, int H,
                                                           int Hkv, uint16_t* __restrict__ out_hi,
                                                           uint16_t* __restrict__ out_lo, int max_pos) {
  extern __shared__ float smem[];
  pdl_launch_dependents();
  TL_ENTER(4u)
  pdl_wait();
  TL_MARK(2)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / Hkv, kvh = blockIdx.x % Hkv;
  const int G = H / Hkv, h = kvh * G + warp;
  constexpr int half = HD / 2;
  float* sq_all = smem;                                    // [rps][G][HD] rotated queries
  float* sc = smem + (size_t)rps * G * HD + (size_t)warp * max_pos;  // this warp's scores
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD, head_off = (size_t)kvh * CSMB_PAGE * HD;
  const float scale = rsqrtf((float)HD);
  const int posb = pos_arr ? pos_arr[b] : pos0;
  {
    const int per_row = (G + 2) * half, total = rps * per_row;
#pragma unroll 2
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int i = idx / per_row, rem = idx % per_row, slot = rem / half, pr = rem % half;
      const int pos = posb + i;
      const int col = (slot < G ? (kvh * G + slot) : (slot == G ? H + kvh : H + Hkv + kvh)) * HD + 2 * pr;
      const float2 v = part_sum2(qkv, (size_t)(b * rps + i) * qkv.ld + col);
      const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + ((size_t)pos * half + pr) * 2));
      const float2 rot = make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
      if (slot < G) {
        *reinterpret_cast<float2*>(sq_all + ((size_t)i * G + slot) * HD + 2 * pr) = rot;
      } else {
        const int lp = pos / CSMB_PAGE;
        const int page = block_table ? block_table[(size_t)b * max_pages + lp] : b * max_pages + lp;
        float* kdst = pool + (size_t)page * page_stride + head_off + (size_t)(pos % CSMB_PAGE) * HD;
        if (slot == G) *reinterpret_cast<float2*>(kdst + 2 * pr) = rot;
        else *reinterpret_cast<float2*>(kdst + (size_t)Hkv * CSMB_PAGE * HD + 2 * pr) = v;
      }
    }
  }
  __syncthreads();  // this block's k/v rows are in the cache, all queries are in shared memory
  TL_MARK(3)
  const int32_t* bt = block_table ? block_table + (size_t)b * max_pages : nullptr;
  for (int i = 0; i < rps; ++i) {
    const int r = b * rps + i, pos = posb + i, S = pos + 1;
    const float* sq = sq_all + ((size_t)i * G + warp) * HD;
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      const int page = bt ? bt[j / CSMB_PAGE] : b * max_pages + j / CSMB_PAGE;
      const float* kp = pool + (size_t)page * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD;
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
    for (int ii = 0; ii < PER; ++ii) acc[ii] = 0.f;
    // keys in groups of 8: the loads of a group are independent and in flight together; accumulation order is j
    for (int j0 = 0; j0 < S; j0 += 8) {
      float vv[8][PER];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u < S ? j0 + u : S - 1;
        const int page = bt ? bt[j / CSMB_PAGE] : b * max_pages + j / CSMB_PAGE;
        const float* vp = pool + (size_t)page * page_stride + (size_t)Hkv * CSMB_PAGE * HD + head_off + (size_t)(j % CSMB_PAGE) * HD;
#pragma unroll
        for (int ii = 0; ii < PER; ++ii) vv[u][ii] = vp[lane + 32 * ii];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (j0 + u < S) {
          const float pj = sc[j0 + u];
#pragma unroll
          for (int ii = 0; ii < PER; ++ii) acc[ii] = fmaf(pj, vv[u][ii], acc[ii]);
        }
    }
    const size_t o = (size_t)r * H * HD + (size_t)h * HD;
#pragma unroll
    for (int ii = 0; ii < PER; ++ii) {
      uint16_t hh, ll;
      split_bf16(acc[ii] * inv, hh, ll);
      out_hi[o + lane + 32 * ii] = hh;
      out_lo[o + lane + 32 * ii] = ll;
    }
    __syncwarp();  // sc is reused by this warp's next row
  }
  TL_MARK(4)
}

// The same block for caches of at most AS_MAXPOS positions (the depth decoder: 32 codebook positions, re-created every
// frame, generation.py:70).  k_attn_decode_fused walks the cache through dependent L2 round trips (scores, then values in
// groups of 8 keys: 8.4 us per launch for 17 positions on average, the largest item of the decoder's frame-step,
// profiles/r02_chain_timeline.md); here every cached K / V row of the kv head is requested at once with cp.async straight
// into shared memory while the block sums the qkv partials of the new rows, and the attention itself runs out of shared
// memory.  Same sums in the same order as k_attn_decode_fused: bit-identical output planes.
constexpr int AS_MAXPOS = 32;
template <int HD>
__global__ void __launch_bounds__(256) k_attn_decode_small(PartIn qkv, const float* __restrict__ rope, float* pool,
                                                           const int32_t* __restrict__ block_table, int max_pages,
                                                           const int32_t* __restrict__ pos_arr, int pos0, int rps, int H,
                                                           int Hkv, uint16_t* __restrict__ out_hi,
                                                           uint16_t* __restrict__ out_lo) {
  constexpr int PITCH = HD + 4, half = HD / 2;
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  TL_ENTER(4u)
  pdl_wait();
  TL_MARK(2)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / Hkv, kvh = blockIdx.x % Hkv;
  const int G = H / Hkv, h = kvh * G + warp;
  float* sK = smem;                                  // [AS_MAXPOS][PITCH]
  float* sV = sK + AS_MAXPOS * PITCH;
  float* sq_all = sV + AS_MAXPOS * PITCH;            // [rps][G][HD] rotated queries
  float* sc = sq_all + (size_t)rps * G * HD + warp * AS_MAXPOS;   // this warp's scores
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD, head_off = (size_t)kvh * CSMB_PAGE * HD;
  const size_t v_off = (size_t)Hkv * CSMB_PAGE * HD;
  const float scale = rsqrtf((float)HD);
  const int posb = pos_arr ? pos_arr[b] : pos0;
  // positions 0 .. posb-1 are in the cache (written by earlier launches): all of them in flight at once
  for (int idx = threadIdx.x; idx < posb * (HD / 4); idx += blockDim.x) {
    const int j = idx / (HD / 4), c4 = (idx % (HD / 4)) * 4;
    const int page = block_table ? block_table[(size_t)b * max_pages + j / CSMB_PAGE] : b * max_pages + j / CSMB_PAGE;
    const float* src = pool + (size_t)page * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD + c4;
    cp_async16(sK + j * PITCH + c4, src);
    cp_async16(sV + j * PITCH + c4, src + v_off);
  }
  // the new rows' qkv partials (this block's G query heads, its k head, its v head) and their RoPE rows: also asynchronous
  const int per_row = (G + 2) * half, total = rps * per_row;
  float2* stg = reinterpret_cast<float2*>(sq_all + (size_t)rps * G * HD + (size_t)G * AS_MAXPOS);   // [S][total]
  float2* srope = stg + (size_t)qkv.S * total;                                                      // [rps][half]
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = idx / per_row, rem = idx % per_row, slot = rem / half, pr = rem % half;
    const int col = (slot < G ? (kvh * G + slot) : (slot == G ? H + kvh : H + Hkv + kvh)) * HD + 2 * pr;
    const float* src = qkv.p + (size_t)(b * rps + i) * qkv.ld + col;
    for (int z = 0; z < qkv.S; ++z) cp_async8(stg + (size_t)z * total + idx, src + (size_t)z * qkv.stride);
  }
  for (int idx = threadIdx.x; idx < rps * half; idx += blockDim.x)
    cp_async8(srope + idx, rope + ((size_t)(posb + idx / half) * half + idx % half) * 2);
  cp_async_wait_all();
  __syncthreads();
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int i = idx / per_row, rem = idx % per_row, slot = rem / half, pr = rem % half;
    const int pos = posb + i;
    float2 v = make_float2(0.f, 0.f);
    for (int z = 0; z < qkv.S; ++z) {   // part_sum2's order
      const float2 t = stg[(size_t)z * total + idx];
      v.x += t.x;
      v.y += t.y;
    }
    const float2 cs = srope[i * half + pr];
    const float2 rot = make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
    if (slot < G) {
      *reinterpret_cast<float2*>(sq_all + ((size_t)i * G + slot) * HD + 2 * pr) = rot;
    } else {
      const int lp = pos / CSMB_PAGE;
      const int page = block_table ? block_table[(size_t)b * max_pages + lp] : b * max_pages + lp;
      float* kdst = pool + (size_t)page * page_stride + head_off + (size_t)(pos % CSMB_PAGE) * HD;
      if (slot == G) {
        *reinterpret_cast<float2*>(kdst + 2 * pr) = rot;
        *reinterpret_cast<float2*>(sK + pos * PITCH + 2 * pr) = rot;
      } else {
        *reinterpret_cast<float2*>(kdst + v_off + 2 * pr) = v;
        *reinterpret_cast<float2*>(sV + pos * PITCH + 2 * pr) = v;
      }
    }
  }
  __syncthreads();  // the kv head's whole cache and all queries are in shared memory
  TL_MARK(3)
  for (int i = 0; i < rps; ++i) {
    const int r = b * rps + i, S = posb + i + 1;
    const float* sq = sq_all + ((size_t)i * G + warp) * HD;
    float m = -INFINITY;
    for (int j = lane; j < S; j += 32) {
      const float* kp = sK + j * PITCH;
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
    for (int ii = 0; ii < PER; ++ii) acc[ii] = 0.f;
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
      const float pj = sc[j];
#pragma unroll
      for (int ii = 0; ii < PER; ++ii) acc[ii] = fmaf(pj, sV[j * PITCH + lane + 32 * ii], acc[ii]);
    }
    const size_t o = (size_t)r * H * HD + (size_t)h * HD;
#pragma unroll
    for (int ii = 0; ii < PER; ++ii) {
      uint16_t hh, ll;
      split_bf16(acc[ii] * inv, hh, ll);
      out_hi[o + lane + 32 * ii] = hh;
      out_lo[o + lane + 32 * ii] = ll;
    }
    __syncwarp();  // sc is reused by this warp's next row
  }
  TL_MARK(4)
}

// The same block for long caches (the backbone: up to max_pos positions), one new row per sequence.  The cache is walked in
// chunks of AC_CHUNK positions staged in shared memory with cp.async (every row of a chunk in flight at once; chunk 0 —
// keys AND values — is requested before the new row's qkv partials are summed): a cache of S positions costs about
// 2 ceil(S / AC_CHUNK) - 1 L2 / DRAM round trips instead of ceil(S / 32) + ceil(S / 8).  Scores for all positions first,
// softmax, then the values in position order: the sums and their order are those of k_attn_decode_fused (bit-identical).
constexpr int AC_CHUNK = 64;
template <int HD>
__global__ void __launch_bounds__(256) k_attn_decode_chunked(PartIn qkv, const float* __restrict__ rope, float* pool,
                                                             const int32_t* __restrict__ block_table, int max_pages,
                                                             const int32_t* __restrict__ pos_arr, int pos0, int H, int Hkv,
                                                             uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                                                             int max_pos) {
  constexpr int PITCH = HD + 4, half = HD / 2;
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  TL_ENTER(4u)
  pdl_wait();
  TL_MARK(2)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / Hkv, kvh = blockIdx.x % Hkv;
  const int G = H / Hkv, h = kvh * G + warp;
  float* sK = smem;                                   // [AC_CHUNK][PITCH]
  float* sV = sK + AC_CHUNK * PITCH;                  // [AC_CHUNK][PITCH]
  float* sq_all = sV + AC_CHUNK * PITCH;              // [G][HD] rotated queries
  float* knew = sq_all + (size_t)G * HD;              // [HD] rotated key of the new row
  float* vnew = knew + HD;                            // [HD]
  float* sc = vnew + HD + (size_t)warp * max_pos;     // this warp's scores [max_pos]
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD, head_off = (size_t)kvh * CSMB_PAGE * HD;
  const size_t v_off = (size_t)Hkv * CSMB_PAGE * HD;
  const float scale = rsqrtf((float)HD);
  const int posb = pos_arr ? pos_arr[b] : pos0, S = posb + 1;
  const int32_t* bt = block_table ? block_table + (size_t)b * max_pages : nullptr;
  // cached rows j0 .. min(j0 + AC_CHUNK, posb) - 1 of K (what & 1) and / or V (what & 2) -> shared memory, asynchronously
  auto stage_chunk = [&](int j0, int what) {
    const int n = (posb - j0 < AC_CHUNK ? posb - j0 : AC_CHUNK) * (HD / 4);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
      const int jl = idx / (HD / 4), c4 = (idx % (HD / 4)) * 4, j = j0 + jl;
      const int page = bt ? bt[j / CSMB_PAGE] : b * max_pages + j / CSMB_PAGE;
      const float* src = pool + (size_t)page * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD + c4;
      if (what & 1) cp_async16(sK + jl * PITCH + c4, src);
      if (what & 2) cp_async16(sV + jl * PITCH + c4, src + v_off);
    }
  };
  stage_chunk(0, 3);
  // the new row's qkv partials and its RoPE row
  const int per_row = (G + 2) * half;
  float2* stg = reinterpret_cast<float2*>(sc - (size_t)warp * max_pos + (size_t)G * max_pos);   // [S_split][per_row]
  float2* srope = stg + (size_t)qkv.S * per_row;                                                // [half]
  for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
    const int slot = idx / half, pr = idx % half;
    const int col = (slot < G ? (kvh * G + slot) : (slot == G ? H + kvh : H + Hkv + kvh)) * HD + 2 * pr;
    const float* src = qkv.p + (size_t)b * qkv.ld + col;
    for (int z = 0; z < qkv.S; ++z) cp_async8(stg + (size_t)z * per_row + idx, src + (size_t)z * qkv.stride);
  }
  for (int idx = threadIdx.x; idx < half; idx += blockDim.x) cp_async8(srope + idx, rope + ((size_t)posb * half + idx) * 2);
  cp_async_wait_all();
  __syncthreads();
  for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
    const int slot = idx / half, pr = idx % half;
    float2 v = make_float2(0.f, 0.f);
    for (int z = 0; z < qkv.S; ++z) {   // part_sum2's order
      const float2 t = stg[(size_t)z * per_row + idx];
      v.x += t.x;
      v.y += t.y;
    }
    const float2 cs = srope[pr];
    const float2 rot = make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
    if (slot < G) {
      *reinterpret_cast<float2*>(sq_all + (size_t)slot * HD + 2 * pr) = rot;
    } else {
      const int lp = posb / CSMB_PAGE;
      const int page = bt ? bt[lp] : b * max_pages + lp;
      float* kdst = pool + (size_t)page * page_stride + head_off + (size_t)(posb % CSMB_PAGE) * HD;
      if (slot == G) {
        *reinterpret_cast<float2*>(kdst + 2 * pr) = rot;
        *reinterpret_cast<float2*>(knew + 2 * pr) = rot;
      } else {
        *reinterpret_cast<float2*>(kdst + v_off + 2 * pr) = v;
        *reinterpret_cast<float2*>(vnew + 2 * pr) = v;
      }
    }
  }
  __syncthreads();
  TL_MARK(3)
  const float* sq = sq_all + (size_t)warp * HD;
  // ---- scores, chunk by chunk (chunk 0 is already in shared memory)
  float m = -INFINITY;
  for (int j0 = 0; j0 < S; j0 += AC_CHUNK) {
    if (j0 > 0) {
      __syncthreads();   // everybody is done with the previous K chunk
      stage_chunk(j0, 1);
      cp_async_wait_all();
      __syncthreads();
    }
    const int j1 = j0 + AC_CHUNK < S ? j0 + AC_CHUNK : S;
    for (int j = j0 + lane; j < j1; j += 32) {
      const float* kp = j == posb ? knew : sK + (j - j0) * PITCH;
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
  // ---- values in position order (chunk 0 of V was staged with chunk 0 of K)
  for (int j0 = 0; j0 < S; j0 += AC_CHUNK) {
    if (j0 > 0) {
      __syncthreads();
      stage_chunk(j0, 2);
      cp_async_wait_all();
      __syncthreads();
    }
    const int j1 = j0 + AC_CHUNK < S ? j0 + AC_CHUNK : S;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const float pj = sc[j];
      const float* vp = j == posb ? vnew : sV + (j - j0) * PITCH;
#pragma unroll
      for (int ii = 0; ii < PER; ++ii) acc[ii] = fmaf(pj, vp[lane + 32 * ii], acc[ii]);
    }
  }
  const size_t o = (size_t)b * H * HD + (size_t)h * HD;
#pragma unroll
  for (int ii = 0; ii < PER; ++ii) {
    uint16_t hh, ll;
    split_bf16(acc[ii] * inv, hh, ll);
    out_hi[o + lane + 32 * ii] = hh;
    out_lo[o + lane + 32 * ii] = ll;
  }
  TL_MARK(4)
}

// ---- prompt rows (T > 1) through the chain's kernels: generation.py:34-42 with many rows per sequence --------------------
// x[r][:] = sum over the masked slots of the row's embeddings (models.py:82-92 + generation.py:32-36; text id in the last
// column), then RMSNorm(w) -> hi/lo.  One block per row; d == 2048 (8 channels per thread).
__global__ void __launch_bounds__(256) k_prefill_embed_norm(const int32_t* __restrict__ tokens, const uint8_t* __restrict__ mask,
                                                            const uint16_t* __restrict__ text_emb,
                                                            const uint16_t* __restrict__ audio_emb, int ncb, int V, int d,
                                                            float* __restrict__ x, const float* __restrict__ w, float eps,
                                                            uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, c = threadIdx.x * 8;
  const int32_t* tk = tokens + (size_t)r * (ncb + 1);
  const uint8_t* mk = mask + (size_t)r * (ncb + 1);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int s = 0; s <= ncb; ++s) {
    if (!mk[s]) continue;
    const uint16_t* row = (s < ncb) ? audio_emb + ((size_t)tk[s] + (size_t)s * V) * d : text_emb + (size_t)tk[s] * d;
    const uint4 q = *reinterpret_cast<const uint4*>(row + c);
    acc[0] += bf16lo(q.x); acc[1] += bf16hi(q.x); acc[2] += bf16lo(q.y); acc[3] += bf16hi(q.y);
    acc[4] += bf16lo(q.z); acc[5] += bf16hi(q.z); acc[6] += bf16lo(q.w); acc[7] += bf16hi(q.w);
  }
  float4* xo = reinterpret_cast<float4*>(x + (size_t)r * d + c);
  xo[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  xo[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) ss += acc[e] * acc[e];
  const float scale = rsqrtf(block_sum256(ss, red) / (float)d + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(w + c), g1 = *reinterpret_cast<const float4*>(w + c + 4);
  store_split4(hi + (size_t)r * d + c, lo + (size_t)r * d + c, acc[0] * scale * g0.x, acc[1] * scale * g0.y,
               acc[2] * scale * g0.z, acc[3] * scale * g0.w);
  store_split4(hi + (size_t)r * d + c + 4, lo + (size_t)r * d + c + 4, acc[4] * scale * g1.x, acc[5] * scale * g1.y,
               acc[6] * scale * g1.z, acc[7] * scale * g1.w);
}

// One block per prompt row: sum the qkv partials, rotate q and k (attention.py:119-177) at the row's position, keep q
// (fp32, [R][H*HD]) and append k / v to the paged cache of the row's sequence (attention.py:236-237).
__global__ void __launch_bounds__(512) k_prefill_rope_append(PartIn qkv, const float* __restrict__ rope, float* __restrict__ pool,
                                                             const int32_t* __restrict__ block_table, int max_pages,
                                                             const int32_t* __restrict__ row_seq,
                                                             const int32_t* __restrict__ row_pos, int H, int Hkv, int HD,
                                                             float* __restrict__ qout) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, pos = row_pos[r], seq = row_seq[r];
  const int half = HD / 2;
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD;
  const int page = block_table[(size_t)seq * max_pages + pos / CSMB_PAGE];
  float* kbase = pool + (size_t)page * page_stride + (size_t)(pos % CSMB_PAGE) * HD;
  const int total = (H + 2 * Hkv) * half;
  // one (head, pair) per thread (gridDim.y blocks per row): a row's partial loads are all in flight at once
  {
    const int idx = blockIdx.y * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int head = idx / half, pr = idx % half;
    const float2 v = part_sum2(qkv, (size_t)r * qkv.ld + (size_t)head * HD + 2 * pr);
    if (head < H + Hkv) {
      const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + ((size_t)pos * half + pr) * 2));
      const float2 rot = make_float2(v.x * cs.x - v.y * cs.y, v.y * cs.x + v.x * cs.y);
      if (head < H)
        *reinterpret_cast<float2*>(qout + (size_t)r * H * HD + (size_t)head * HD + 2 * pr) = rot;
      else
        *reinterpret_cast<float2*>(kbase + (size_t)(head - H) * CSMB_PAGE * HD + 2 * pr) = rot;
    } else {
      *reinterpret_cast<float2*>(kbase + (size_t)Hkv * CSMB_PAGE * HD + (size_t)(head - H - Hkv) * CSMB_PAGE * HD + 2 * pr) = v;
    }
  }
}

// One block per (prompt row, kv head), warp g = query head kvh*G + g: causal attention over positions 0 .. pos of the row's
// sequence (attention.py:242-249; the cache already holds the whole prompt).  Same arithmetic as k_attn_decode_fused.
template <int HD>
__global__ void __launch_bounds__(256) k_prefill_attn(const float* __restrict__ q, const float* __restrict__ pool,
                                                      const int32_t* __restrict__ block_table, int max_pages,
                                                      const int32_t* __restrict__ row_seq, const int32_t* __restrict__ row_pos,
                                                      int H, int Hkv, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                                                      int max_pos) {
  extern __shared__ float smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x / Hkv, kvh = blockIdx.x % Hkv;
  const int G = H / Hkv, h = kvh * G + warp;
  float* sq = smem + (size_t)warp * HD;
  float* sc = smem + (size_t)G * HD + (size_t)warp * max_pos;
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD, head_off = (size_t)kvh * CSMB_PAGE * HD;
  const float scale = rsqrtf((float)HD);
  const int S = row_pos[r] + 1;
  const int32_t* bt = block_table + (size_t)row_seq[r] * max_pages;
  for (int c = lane; c < HD; c += 32) sq[c] = q[(size_t)r * H * HD + (size_t)h * HD + c];
  __syncwarp();
  float m = -INFINITY;
  for (int j = lane; j < S; j += 32) {
    const float* kp = pool + (size_t)bt[j / CSMB_PAGE] * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD;
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
  for (int ii = 0; ii < PER; ++ii) acc[ii] = 0.f;
  for (int j0 = 0; j0 < S; j0 += 8) {
    float vv[8][PER];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + u < S ? j0 + u : S - 1;
      const float* vp = pool + (size_t)bt[j / CSMB_PAGE] * page_stride + (size_t)Hkv * CSMB_PAGE * HD + head_off + (size_t)(j % CSMB_PAGE) * HD;
#pragma unroll
      for (int ii = 0; ii < PER; ++ii) vv[u][ii] = vp[lane + 32 * ii];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (j0 + u < S) {
        const float pj = sc[j0 + u];
#pragma unroll
        for (int ii = 0; ii < PER; ++ii) acc[ii] = fmaf(pj, vv[u][ii], acc[ii]);
      }
  }
  const size_t o = (size_t)r * H * HD + (size_t)h * HD;
#pragma unroll
  for (int ii = 0; ii < PER; ++ii) {
    uint16_t hh, ll;
    split_bf16(acc[ii] * inv, hh, ll);
    out_hi[o + lane + 32 * ii] = hh;
    out_lo[o + lane + 32 * ii] = ll;
  }
}

// Prompt attention (attention.py:242-249 with T > 1), one block per (tile of PA_QT consecutive prompt rows, kv head).  The G
// query heads that share the kv head are handled together (GQA: mx.repeat, attention.py:242-245): a warp takes a row and
// computes its G heads at once, so a staged key / value row is used G times.  Keys and values of the tile's sequence are staged
// in shared memory in chunks of PA_KC positions (the per-row kernel re-read them from L2 for every row: O(T^2) traffic) with
// an online softmax across chunks (running max / sum / accumulator per (row, head)).  A tile that spans two sequences
// processes them one after the other.  Output: bf16 hi/lo planes of the o-projection's operand.
constexpr int PA_QT = 32, PA_KC = 256, PA_WARPS = 8, PA_G = 4;
template <int HD>
__global__ void __launch_bounds__(PA_WARPS * 32) k_prefill_attn_tile(const float* __restrict__ q, const float* __restrict__ pool,
                                                                     const int32_t* __restrict__ block_table, int max_pages,
                                                                     const int32_t* __restrict__ row_seq,
                                                                     const int32_t* __restrict__ row_pos, int R, int H, int Hkv,
                                                                     uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo) {
  static_assert(HD == 64, "two output dims per lane");
  constexpr int PITCH = HD + 4;
  extern __shared__ float pa_sm[];
  float* sK = pa_sm;
  float* sV = sK + (size_t)PA_KC * PITCH;
  float* sQ = sV + (size_t)PA_KC * PITCH;           // [warp][G][HD]
  float* sS = sQ + PA_WARPS * PA_G * HD;            // [warp][PA_KC][G]
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * PA_QT, kvh = blockIdx.y;
  const int r1 = r0 + PA_QT < R ? r0 + PA_QT : R;
  const size_t page_stride = (size_t)2 * Hkv * CSMB_PAGE * HD, head_off = (size_t)kvh * CSMB_PAGE * HD;
  const float scale = rsqrtf((float)HD);
  float* sq = sQ + warp * PA_G * HD;
  float* sc = sS + (size_t)warp * PA_KC * PA_G;
  for (int seg0 = r0; seg0 < r1;) {
    // segment: consecutive rows of the tile that belong to one sequence
    const int seq = row_seq[seg0];
    int seg1 = seg0 + 1, pmax = row_pos[seg0];
    while (seg1 < r1 && row_seq[seg1] == seq) {
      pmax = row_pos[seg1] > pmax ? row_pos[seg1] : pmax;
      ++seg1;
    }
    const int32_t* bt = block_table + (size_t)seq * max_pages;
    // this warp's rows of the segment: seg0 + warp, + PA_WARPS, ... (at most PA_QT / PA_WARPS = 4); running softmax state in registers
    constexpr int RPW = PA_QT / PA_WARPS;
    float m_run[RPW][PA_G], l_run[RPW][PA_G], a0[RPW][PA_G], a1[RPW][PA_G];
#pragma unroll
    for (int i = 0; i < RPW; ++i)
#pragma unroll
      for (int g = 0; g < PA_G; ++g) {
        m_run[i][g] = -INFINITY;
        l_run[i][g] = a0[i][g] = a1[i][g] = 0.f;
      }
    for (int k0 = 0; k0 <= pmax; k0 += PA_KC) {
      const int nk = pmax + 1 - k0 < PA_KC ? pmax + 1 - k0 : PA_KC;
      __syncthreads();   // the previous chunk (or segment) is fully consumed
      for (int idx0 = threadIdx.x; idx0 < nk * (HD / 4); idx0 += 4 * PA_WARPS * 32) {
        float4 kv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + u * PA_WARPS * 32;
          if (idx < nk * (HD / 4)) {
            const int j = k0 + idx / (HD / 4), c4 = (idx % (HD / 4)) * 4;
            const float* src = pool + (size_t)bt[j / CSMB_PAGE] * page_stride + head_off + (size_t)(j % CSMB_PAGE) * HD + c4;
            kv[u] = *reinterpret_cast<const float4*>(src);
            vv[u] = *reinterpret_cast<const float4*>(src + (size_t)Hkv * CSMB_PAGE * HD);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + u * PA_WARPS * 32;
          if (idx < nk * (HD / 4)) {
            *reinterpret_cast<float4*>(sK + (size_t)(idx / (HD / 4)) * PITCH + (idx % (HD / 4)) * 4) = kv[u];
            *reinterpret_cast<float4*>(sV + (size_t)(idx / (HD / 4)) * PITCH + (idx % (HD / 4)) * 4) = vv[u];
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = seg0 + warp + i * PA_WARPS;
        if (r >= seg1) break;
        const int pos = row_pos[r];
        if (pos < k0) continue;                                   // causal: nothing of this chunk is visible to the row
        const int S = pos - k0 + 1 < nk ? pos - k0 + 1 : nk;      // visible keys of the chunk
        __syncwarp();
        for (int g = 0; g < PA_G; ++g) {
          const float* qp = q + (size_t)r * H * HD + (size_t)(kvh * PA_G + g) * HD;
          sq[g * HD + lane] = qp[lane];
          sq[g * HD + lane + 32] = qp[lane + 32];
        }
        __syncwarp();
        float mx[PA_G];
#pragma unroll
        for (int g = 0; g < PA_G; ++g) mx[g] = -INFINITY;
        for (int j = lane; j < S; j += 64) {   // two keys per lane x G heads: 2 G independent fma chains
          const float* kp[2];
          float dot[2][PA_G];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            kp[u] = sK + (size_t)(j + 32 * u < S ? j + 32 * u : j) * PITCH;
#pragma unroll
            for (int g = 0; g < PA_G; ++g) dot[u][g] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < HD; c += 4) {
            float4 qv[PA_G];
#pragma unroll
            for (int g = 0; g < PA_G; ++g) qv[g] = *reinterpret_cast<const float4*>(sq + g * HD + c);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float4 kv = *reinterpret_cast<const float4*>(kp[u] + c);
#pragma unroll
              for (int g = 0; g < PA_G; ++g) {
                dot[u][g] = fmaf(kv.x, qv[g].x, dot[u][g]);
                dot[u][g] = fmaf(kv.y, qv[g].y, dot[u][g]);
                dot[u][g] = fmaf(kv.z, qv[g].z, dot[u][g]);
                dot[u][g] = fmaf(kv.w, qv[g].w, dot[u][g]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (j + 32 * u < S) {
              float d[PA_G];
#pragma unroll
              for (int g = 0; g < PA_G; ++g) {
                d[g] = dot[u][g] * scale;
                mx[g] = fmaxf(mx[g], d[g]);
              }
              *reinterpret_cast<float4*>(sc + (size_t)(j + 32 * u) * PA_G) = make_float4(d[0], d[1], d[2], d[3]);
            }
        }
        float corr[PA_G], sum[PA_G];
#pragma unroll
        for (int g = 0; g < PA_G; ++g) {
          const float mnew = fmaxf(m_run[i][g], warp_max(mx[g]));
          corr[g] = expf(m_run[i][g] - mnew);   // exp(-inf) = 0 on the first chunk
          m_run[i][g] = mnew;
          sum[g] = 0.f;
        }
        __syncwarp();
        for (int j = lane; j < S; j += 32) {
          float4 e = *reinterpret_cast<const float4*>(sc + (size_t)j * PA_G);
          e.x = expf(e.x - m_run[i][0]); e.y = expf(e.y - m_run[i][1]); e.z = expf(e.z - m_run[i][2]); e.w = expf(e.w - m_run[i][3]);
          *reinterpret_cast<float4*>(sc + (size_t)j * PA_G) = e;
          sum[0] += e.x; sum[1] += e.y; sum[2] += e.z; sum[3] += e.w;
        }
#pragma unroll
        for (int g = 0; g < PA_G; ++g) {
          l_run[i][g] = l_run[i][g] * corr[g] + warp_sum(sum[g]);
          a0[i][g] *= corr[g];
          a1[i][g] *= corr[g];
        }
        __syncwarp();
        for (int j0 = 0; j0 < S; j0 += 4) {
          float4 pj[4];
          float v0[4], v1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int jj = j0 + u < S ? j0 + u : S - 1;
            const float* vp = sV + (size_t)jj * PITCH;
            pj[u] = *reinterpret_cast<const float4*>(sc + (size_t)jj * PA_G);
            v0[u] = vp[lane];
            v1[u] = vp[lane + 32];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (j0 + u < S) {
              a0[i][0] = fmaf(pj[u].x, v0[u], a0[i][0]); a1[i][0] = fmaf(pj[u].x, v1[u], a1[i][0]);
              a0[i][1] = fmaf(pj[u].y, v0[u], a0[i][1]); a1[i][1] = fmaf(pj[u].y, v1[u], a1[i][1]);
              a0[i][2] = fmaf(pj[u].z, v0[u], a0[i][2]); a1[i][2] = fmaf(pj[u].z, v1[u], a1[i][2]);
              a0[i][3] = fmaf(pj[u].w, v0[u], a0[i][3]); a1[i][3] = fmaf(pj[u].w, v1[u], a1[i][3]);
            }
        }
      }
    }
    // write this segment's rows
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = seg0 + warp + i * PA_WARPS;
      if (r >= seg1) break;
#pragma unroll
      for (int g = 0; g < PA_G; ++g) {
        const float inv = 1.f / l_run[i][g];
        const size_t o = (size_t)r * H * HD + (size_t)(kvh * PA_G + g) * HD;
        uint16_t hh, ll;
        split_bf16(a0[i][g] * inv, hh, ll);
        out_hi[o + lane] = hh;
        out_lo[o + lane] = ll;
        split_bf16(a1[i][g] * inv, hh, ll);
        out_hi[o + lane + 32] = hh;
        out_lo[o + lane + 32] = ll;
      }
    }
    seg0 = seg1;
  }
}

// act[r][f] = silu(gate) * up from the fused gate|up partials (mlx_lm MLP) -> hi/lo [R][F]
__global__ void __launch_bounds__(256) k_swiglu_split(PartIn gu, int F, size_t total4, uint16_t* __restrict__ hi,
                                                      uint16_t* __restrict__ lo) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const size_t i = i4 * 4, r = i / F, f = i % F;
  const float4 g = part_sum4(gu, r * gu.ld + f), u = part_sum4(gu, r * gu.ld + F + f);
  store_split4(hi + i, lo + i, (g.x / (1.f + expf(-g.x))) * u.x, (g.y / (1.f + expf(-g.y))) * u.y,
               (g.z / (1.f + expf(-g.z))) * u.z, (g.w / (1.f + expf(-g.w))) * u.w);
}

struct BfSample {
  int top_k;       // 0 = off
  float min_p;     // 0 = off
  float top_p;     // 0 = off
  float inv_temp;  // 0 => greedy
  uint32_t seed_lo, seed_hi;
  uint64_t draw_base;
  uint32_t draw_pos_mul;
};

// One block per sequence: logits = sum of the head's partials; token = argmax / Gumbel-max (generation.py:51-54,
// 81-84; same draw indexing as csmb_sample); frame[b][cb] = token; then the next depth step's input row
// embed_audio(cb, token) (generation.py:86-89) as hi (the embedding is bf16: lo = 0) at row b*out_mul + out_mul-1,
// and, for the first depth step (out_mul == 2), row 2b = h_last[b] (generation.py:56-64).
// With a projected-embedding table (ptab = rows of codebook cb of csmb_build_proj_table, out_mul == 1) the next depth
// step's projection Linear and first RMSNorm are done here instead: dx[b] = ptab[token] (= projection . embed_audio(cb,
// token), the very fp32 values that Linear would produce) and hi/lo row b = split(RMSNorm(dx[b]) * nw), d_d = NVD * 1024.
struct ProjIn {
  const float* ptab;  // [V][dd] or null
  float* dx;          // decoder residual stream [B][dd]
  const float* nw;    // decoder layer 0 input-norm weight
  float eps;
  int dd;
};
__global__ void __launch_bounds__(256) k_sample_embed(PartIn lg, int V, BfSample a, const int32_t* __restrict__ row_pos,
                                                      int cb, int32_t* __restrict__ frame, int ncb,
                                                      const uint16_t* __restrict__ audio_emb, int d, int embed,
                                                      const float* __restrict__ h_last, uint16_t* __restrict__ hi,
                                                      uint16_t* __restrict__ lo, int out_mul, ProjIn pj, int staged) {
  extern __shared__ float sl[];
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ unsigned long long red_q[8];
  pdl_launch_dependents();
  TL_ENTER(5u)
  pdl_wait();
  TL_MARK(2)
  const int b = blockIdx.x;
  if (staged) {
    // logits = sum of the head's split-K partials in part_sum1's order; every partial of this thread's vocabulary entries is
    // requested at once (cp.async into sl[V ..]), one L2 round trip instead of one per entry
    float* stg = sl + V;
    const float* lp = lg.p + (size_t)b * lg.ld;
    for (int z = 0; z < lg.S; ++z)
      for (int i = threadIdx.x; i < V; i += 256) cp_async4(stg + (size_t)z * V + i, lp + (size_t)z * lg.stride + i);
    cp_async_wait_all();
    for (int i = threadIdx.x; i < V; i += 256) {
      float v = 0.f;
      for (int z = 0; z < lg.S; ++z) v += stg[(size_t)z * V + i];
      sl[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < V; i += 256) sl[i] = part_sum1(lg, (size_t)b * lg.ld + i);
  }
  __syncthreads();
  TL_MARK(3)
  int tok;
  if (a.inv_temp == 0.f) {
    tok = block_argmax(V, [&](int i) { return sl[i]; }, red_v, red_i);
  } else {
    // top-k / min-p exactly as k_sample_filtered defines them: e_i = exp(logit_i - max), keep e_i >= max(k-th largest e,
    // min_p); the k-th largest e by a 31-step bisection on its bit pattern (see sample_token in frame_kernel.cu)
    float thresh = -1.f, m = 0.f;
    if (a.top_k > 0 || a.min_p > 0.f || a.top_p > 0.f) {
      float mx = -INFINITY;
      for (int i = threadIdx.x; i < V; i += 256) mx = fmaxf(mx, sl[i]);
      mx = warp_max(mx);
      if ((threadIdx.x & 31) == 0) red_v[threadIdx.x >> 5] = mx;
      __syncthreads();
      m = red_v[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = fmaxf(m, red_v[w]);
      __syncthreads();
      unsigned T = 0u;
      if (a.top_k > 0) {
        for (int bit = 30; bit >= 0; --bit) {
          const unsigned cand = T | (1u << bit);
          int n = 0;
          for (int i = threadIdx.x; i < V; i += 256) n += (__float_as_uint(expf(sl[i] - m)) >= cand) ? 1 : 0;
          n = __reduce_add_sync(0xffffffffu, n);
          if ((threadIdx.x & 31) == 0) red_i[threadIdx.x >> 5] = n;
          __syncthreads();
          int tot = 0;
#pragma unroll
          for (int w = 0; w < 8; ++w) tot += red_i[w];
          __syncthreads();
          if (tot >= a.top_k) T = cand;
        }
      }
      if (a.top_p > 0.f) {
        // nucleus on exact integer masses q = floor(e * 2^32): largest T with sum(q : e >= T) >= top_p * sum(q)
        // (identical to sample_token of frame_kernel.cu and to the sorter of k_sample_filtered)
        auto block_mass = [&](unsigned cand) {
          unsigned long long n = 0;
          for (int i = threadIdx.x; i < V; i += 256) {
            const float e = expf(sl[i] - m);
            if (__float_as_uint(e) >= cand) n += __float2ull_rz(e * 4294967296.f);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
          if ((threadIdx.x & 31) == 0) red_q[threadIdx.x >> 5] = n;
          __syncthreads();
          unsigned long long tot = 0;
#pragma unroll
          for (int w = 0; w < 8; ++w) tot += red_q[w];
          __syncthreads();
          return tot;
        };
        const double need = (double)a.top_p * (double)block_mass(0u);
        unsigned Tp = 0u;
        for (int bit = 30; bit >= 0; --bit) {
          const unsigned cand = Tp | (1u << bit);
          if ((double)block_mass(cand) >= need) Tp = cand;
        }
        T = T > Tp ? T : Tp;
      }
      thresh = fmaxf(__uint_as_float(T), a.min_p);
    }
    const uint64_t draw = a.draw_base + (uint64_t)(row_pos ? row_pos[b] : 0) * a.draw_pos_mul;
    const uint32_t dlo = (uint32_t)draw, dhi = (uint32_t)(draw >> 32);
    tok = block_argmax(
        V,
        [&](int i) {
          if (thresh >= 0.f && !(expf(sl[i] - m) >= thresh)) return -INFINITY;
          return sl[i] * a.inv_temp + gumbel_for(i, dlo, dhi, (uint32_t)b, a.seed_lo, a.seed_hi);
        },
        red_v, red_i);
  }
  if (threadIdx.x == 0) frame[(size_t)b * ncb + cb] = tok;
  TL_MARK(4)
  if (!embed) return;
  const int t = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
  if (pj.ptab != nullptr) {
    const float* src32 = pj.ptab + (size_t)t * pj.dd;
    float* xr = pj.dx + (size_t)b * pj.dd;
    if (pj.dd == 1024) {
      float4 v[1];
      v[0] = __ldg(reinterpret_cast<const float4*>(src32 + threadIdx.x * 4));
      *reinterpret_cast<float4*>(xr + threadIdx.x * 4) = v[0];
      norm_split_row<1>(v, sumsq4(v[0], 0.f), pj.nw, pj.eps, hi + (size_t)b * 1024, lo + (size_t)b * 1024, nullptr, red_v);
    } else {
      float4 v[2];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        v[j] = __ldg(reinterpret_cast<const float4*>(src32 + threadIdx.x * 4 + j * 1024));
        *reinterpret_cast<float4*>(xr + threadIdx.x * 4 + j * 1024) = v[j];
        ss = sumsq4(v[j], ss);
      }
      norm_split_row<2>(v, ss, pj.nw, pj.eps, hi + (size_t)b * 2048, lo + (size_t)b * 2048, nullptr, red_v);
    }
    return;
  }
  const uint16_t* src = audio_emb + ((size_t)t + (size_t)cb * V) * d;
  const size_t re = (size_t)(b * out_mul + out_mul - 1) * d;
  for (int c = threadIdx.x * 8; c < d; c += 256 * 8) {
    *reinterpret_cast<uint4*>(hi + re + c) = *reinterpret_cast<const uint4*>(src + c);
    *reinterpret_cast<uint4*>(lo + re + c) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (h_last != nullptr && out_mul == 2) {
    const size_t rh = (size_t)(b * 2) * d;
    for (int c = threadIdx.x * 4; c < d; c += 256 * 4) {
      const float4 v = *reinterpret_cast<const float4*>(h_last + (size_t)b * d + c);
      store_split4(hi + rh + c, lo + rh + c, v.x, v.y, v.z, v.w);
    }
  }
}

// out[i] = sum of the S partials in the chain's fixed order (part_sum4): builds the projected-embedding table
__global__ void __launch_bounds__(256) k_part_reduce(PartIn pi, size_t total4, float* __restrict__ out) {
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  *reinterpret_cast<float4*>(out + i4 * 4) = part_sum4(pi, i4 * 4);
}

__global__ void __launch_bounds__(256) k_part_reduce1(PartIn pi, size_t total, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = part_sum1(pi, i);
}

// ---------------------------------------------------------------------------------------------- host side
template <typename... KArgs, typename... Args>
static cudaError_t bf_launch(const ChainCfg& cc, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cc.pdl ? 1 : 0;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
// the same as clusters of two CTAs along x (CTA pairs of one TPC: k_gemm_part_t<., true>)
template <typename... KArgs, typename... Args>
static cudaError_t bf_launch_pair(const ChainCfg& cc, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cc.pdl ? 2 : 1;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

struct FastWs {
  int* err;
  float *x, *dx, *h_last, *part;
  uint16_t *hi, *lo;    // activation planes every Linear reads ([rows][K])
  uint16_t *hi2, *lo2;  // SwiGLU output planes [rows][d_ff]: written by the gate|up Linear's epilogue while other CTAs of that
                        // launch still read hi / lo, read by the down projection
  size_t part_floats, bytes;
  ChainCfg cc;
};

static size_t bf_part_floats(int R, int N, int K) { return (size_t)bf_pick_split(N, K) * R * N; }

static FastWs bf_carve(const csmb_model& m, int B, void* base, const ChainCfg& cc = ChainCfg{1, 0, BF_SMEM_BUDGET}) {
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
  FastWs w;
  w.cc = cc;
  w.err = (int*)take(256);
  w.x = (float*)take((size_t)B * b.d_model * 4);
  w.dx = (float*)take(R2 * d.d_model * 4);
  w.h_last = (float*)take((size_t)B * b.d_model * 4);
  w.hi = (uint16_t*)take(R2 * kmax * 2);
  w.lo = (uint16_t*)take(R2 * kmax * 2);
  w.hi2 = (uint16_t*)take(R2 * kmax * 2);
  w.lo2 = (uint16_t*)take(R2 * kmax * 2);
  size_t pf = 0;
  auto need = [&](int R, int N, int K) {
    const size_t f = bf_part_floats(R, N, K);
    pf = f > pf ? f : pf;
  };
  need(B, qkv_b, b.d_model); need(B, b.d_model, b.n_heads * b.head_dim); need(B, 2 * b.d_ff, b.d_model); need(B, b.d_model, b.d_ff);
  need(B, m.audio_vocab, b.d_model); need(B, m.audio_vocab, d.d_model);
  for (int R : {B, 2 * B}) {
    need(R, d.d_model, b.d_model); need(R, qkv_d, d.d_model); need(R, d.d_model, d.n_heads * d.head_dim);
    need(R, 2 * d.d_ff, d.d_model); need(R, d.d_model, d.d_ff);
  }
  w.part_floats = pf;
  w.part = (float*)take(w.part_floats * 4);
  w.bytes = off;
  return w;
}

// Pipeline depth of a Linear CTA: what the shared-memory budget allows, at most BF_MAX_STAGES, a multiple of BF_NI.  (Measured
// on the 2-issuer kernel, B = 8: 6 stages 4.40 ms per frame-step, 8 stages 4.30, 10 stages 4.25; shrinking the ring of a
// Linear with few K blocks to make room for other kernels' CTAs on its SM: 4.37.  B = 64 has 32 KiB stages: 6 either way.
// profiles/r02_ring_depth.log)
static int bf_ring_depth(int fit, int dbg) {
  int n = fit > BF_MAX_STAGES ? BF_MAX_STAGES : fit;
  if ((dbg & 256) && n > 8) n = 8;   // A/B: at most 8 stages
  n -= n % BF_NI;                    // a stage always belongs to the same issuer
  return n;
}

// CTA pairs (k_gemm_part_t<., true>) for a Linear of `tiles` n-tiles: opt-in per call (csmb_chain_opts flag 4096) while it is
// being measured; needs an even number of n-tiles (a pair = two neighbouring tiles, same token rows, same K range).  Outputs
// are bit-identical either way, so the choice may depend on anything.
#ifndef CSMB_PAIR_MIN_ROWS
#define CSMB_PAIR_MIN_ROWS 0   // > 0: every Linear over at least this many token rows runs as CTA pairs (prompt passes)
#endif
// -> MODE of k_gemm_part_t: 0 = single CTAs (shipped), 1 = CTA pairs
static int bf_mode(const ChainCfg& cc, int tiles, bool gu, int R) {
  if (tiles % 2 == 0 && ((CSMB_PAIR_MIN_ROWS > 0 && R >= CSMB_PAIR_MIN_ROWS) || (cc.dbg & (gu ? 65536 : 131072)) != 0)) return 1;
  return 0;
}
// shared memory of a Linear CTA: ring stage size, ring depth and the dynamic allocation for a MODE
struct BfSmem { size_t stage; int nstages; size_t bytes; };
static BfSmem bf_smem(const ChainCfg& cc, int RN, int mode) {
  BfSmem r;
  r.stage = ((size_t)TC_BM * TC_BK * 2 + (mode == 1 ? 1 : 2) * (size_t)RN * TC_BK * 2 + 1023) & ~(size_t)1023;
  r.nstages = bf_ring_depth((int)(cc.smem / r.stage), cc.dbg);
  r.bytes = r.stage * r.nstages + 1024;
  return r;
}
constexpr int BF_SMEM_ATTR = (int)(BF_SMEM_BUDGET + 1024);   // cudaFuncAttributeMaxDynamicSharedMemorySize of the Linears

// y = x W^T for the R rows whose planes are xhi / xlo -> split-K partials in w.part
static int bf_gemm(const FastWs& w, const uint16_t* W, int R, int N, int K, PartIn* out, cudaStream_t st,
                   const uint16_t* xhi = nullptr, const uint16_t* xlo = nullptr, const void* pf = nullptr, size_t pf_bytes = 0,
                   int keep8 = 0) {
  CSMB_REQUIRE(R > 0 && N > 0 && K % TC_BK == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0);
  const int S = bf_pick_split(N, K);
  CSMB_REQUIRE((size_t)S * R * N <= w.part_floats);
  const int RN = R <= 128 ? ((R + 15) / 16) * 16 : 128;   // token rows per tile: 2 RN <= 256 accumulator columns (hi | lo)
  CUtensorMap mw, mhi, mlo;
  if (!tc_make_map(&mw, W, N, K, TC_BM) || !tc_make_map(&mhi, xhi ? xhi : w.hi, R, K, RN) ||
      !tc_make_map(&mlo, xlo ? xlo : w.lo, R, K, RN))
    return CSMB_ERR_UNSUPPORTED;
  const int mode = bf_mode(w.cc, cdiv(N, TC_BM), false, R);
  const BfSmem sm = bf_smem(w.cc, RN, mode);
  const int nstages = sm.nstages;
  CSMB_REQUIRE(nstages >= 2);
  GpArgs a{w.part, R, N, K, RN, nstages, S, w.err, w.cc.dbg & (3 | 32 | 64 | 128 | 2048), 0, nullptr, nullptr, static_cast<const char*>(pf), (unsigned)pf_bytes, keep8};
  const size_t smem = sm.bytes;
  dim3 grid(cdiv(N, TC_BM), cdiv(R, RN), S);
  if (mode == 1) {
    CSMB_CUDA(cudaFuncSetAttribute(k_gemm_part_t<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_ATTR));
    CSMB_CUDA(bf_launch_pair(w.cc, k_gemm_part_t<false, 1>, grid, dim3(BF_THREADS), smem, st, mw, mhi, mlo, a));
  } else {
    CSMB_CUDA(cudaFuncSetAttribute(k_gemm_part_t<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_ATTR));
    CSMB_CUDA(bf_launch(w.cc, k_gemm_part_t<false, 0>, grid, dim3(BF_THREADS), smem, st, mw, mhi, mlo, a));
  }
  *out = PartIn{w.part, S, (size_t)R * N, N};
  return CSMB_OK;
}

// SwiGLU MLP first half in one launch: w.hi2 / w.lo2 [R][F] = split(silu(x Wg^T) * (x Wu^T)), Wgu = gate rows then up rows
static int bf_gemm_gu(const FastWs& w, const uint16_t* Wgu, int R, int F, int K, cudaStream_t st, const void* pf = nullptr,
                      size_t pf_bytes = 0, int keep8 = 0) {
  CSMB_REQUIRE(R > 0 && F % (TC_BM / 2) == 0 && K % TC_BK == 0 && (reinterpret_cast<uintptr_t>(Wgu) & 15) == 0);
  const int RN = R <= 128 ? ((R + 15) / 16) * 16 : 128;   // token rows per tile: 2 RN <= 256 accumulator columns (hi | lo)
  CUtensorMap mw, mhi, mlo;
  if (!tc_make_map(&mw, Wgu, 2 * F, K, TC_BM / 2) || !tc_make_map(&mhi, w.hi, R, K, RN) || !tc_make_map(&mlo, w.lo, R, K, RN))
    return CSMB_ERR_UNSUPPORTED;
  const int mode = bf_mode(w.cc, F / (TC_BM / 2), true, R);
  const BfSmem sm = bf_smem(w.cc, RN, mode);
  const int nstages = sm.nstages;
  CSMB_REQUIRE(nstages >= 2 && (size_t)nstages * sm.stage >= (size_t)2 * RN * 64 * sizeof(float));
  GpArgs a{nullptr, R, 2 * F, K, RN, nstages, 1, w.err, w.cc.dbg & (32 | 64 | 128 | 2048), F, w.hi2, w.lo2, static_cast<const char*>(pf), (unsigned)pf_bytes, keep8};
  const size_t smem = sm.bytes;
  dim3 grid(F / (TC_BM / 2), cdiv(R, RN), 1);
  if (mode == 1) {
    CSMB_CUDA(cudaFuncSetAttribute(k_gemm_part_t<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_ATTR));
    CSMB_CUDA(bf_launch_pair(w.cc, k_gemm_part_t<true, 1>, grid, dim3(BF_THREADS), smem, st, mw, mhi, mlo, a));
  } else {
    CSMB_CUDA(cudaFuncSetAttribute(k_gemm_part_t<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_ATTR));
    CSMB_CUDA(bf_launch(w.cc, k_gemm_part_t<true, 0>, grid, dim3(BF_THREADS), smem, st, mw, mhi, mlo, a));
  }
  return CSMB_OK;
}

static int bf_norm(const FastWs& w, float* x, int d, PartIn part, int mode, const float* nw, float eps, float* y32, int rows,
                   int row_mul, int row_add, cudaStream_t st, const int32_t* row_idx = nullptr) {
  // the row's S partials staged in shared memory with cp.async when they fit (S * d * 4 bytes)
  size_t smem = mode != 0 ? (size_t)part.S * d * sizeof(float) : 0;
  const int staged = (smem > 0 && smem <= 96 * 1024 && !(w.cc.dbg & 16)) ? 1 : 0;
  if (!staged) smem = 0;
  if (d == 1024) {
    if (smem > 48 * 1024) CSMB_CUDA(cudaFuncSetAttribute(k_resid_norm_split<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CSMB_CUDA(bf_launch(w.cc, k_resid_norm_split<1>, dim3(rows), dim3(256), smem, st, x, d, part, mode, nw, eps, w.hi, w.lo, y32, row_mul, row_add, row_idx, staged));
  } else if (d == 2048) {
    if (smem > 48 * 1024) CSMB_CUDA(cudaFuncSetAttribute(k_resid_norm_split<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    CSMB_CUDA(bf_launch(w.cc, k_resid_norm_split<2>, dim3(rows), dim3(256), smem, st, x, d, part, mode, nw, eps, w.hi, w.lo, y32, row_mul, row_add, row_idx, staged));
  } else
    return CSMB_ERR_UNSUPPORTED;
  return CSMB_OK;
}

static int bf_attn(const FastWs& w, const csmb_llama& L, PartIn qkv, float* pool, const int32_t* block_table, int max_pages,
                   const int32_t* pos_arr, int pos0, int rps, int B, cudaStream_t st) {
  const int G = L.n_heads / L.n_kv_heads, max_pos = max_pages * CSMB_PAGE;
  CSMB_REQUIRE(G >= 1 && G <= 8);
  const dim3 grid(B * L.n_kv_heads), block(32 * G);
  if (max_pos <= AS_MAXPOS && !(w.cc.dbg & 8)) {
    // short caches (the depth decoder): everything staged in shared memory at once
    const size_t smem = ((size_t)2 * AS_MAXPOS * (L.head_dim + 4) + (size_t)rps * G * L.head_dim + (size_t)G * AS_MAXPOS) * sizeof(float) +
                        ((size_t)qkv.S * rps * (G + 2) * (L.head_dim / 2) + (size_t)rps * (L.head_dim / 2)) * sizeof(float2);
    CSMB_REQUIRE(smem <= 160 * 1024);
    if (L.head_dim == 64) CSMB_CUDA(cudaFuncSetAttribute(k_attn_decode_small<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (L.head_dim == 128) CSMB_CUDA(cudaFuncSetAttribute(k_attn_decode_small<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (L.head_dim == 64)
      CSMB_CUDA(bf_launch(w.cc, k_attn_decode_small<64>, grid, block, smem, st, qkv, L.rope, pool, block_table, max_pages, pos_arr, pos0,
                          rps, L.n_heads, L.n_kv_heads, w.hi, w.lo));
    else if (L.head_dim == 128)
      CSMB_CUDA(bf_launch(w.cc, k_attn_decode_small<128>, grid, block, smem, st, qkv, L.rope, pool, block_table, max_pages, pos_arr, pos0,
                          rps, L.n_heads, L.n_kv_heads, w.hi, w.lo));
    else
      return CSMB_ERR_UNSUPPORTED;
    return CSMB_OK;
  }
  if (rps == 1 && !(w.cc.dbg & 8)) {
    // long caches, one new row per sequence (the backbone): chunks of the cache staged in shared memory
    const size_t smem = ((size_t)2 * AC_CHUNK * (L.head_dim + 4) + (size_t)(G + 2) * L.head_dim + (size_t)G * max_pos) * sizeof(float) +
                        ((size_t)qkv.S * (G + 2) * (L.head_dim / 2) + (size_t)(L.head_dim / 2)) * sizeof(float2);
    if (smem <= 160 * 1024) {
      if (L.head_dim == 64) {
        CSMB_CUDA(cudaFuncSetAttribute(k_attn_decode_chunked<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CSMB_CUDA(bf_launch(w.cc, k_attn_decode_chunked<64>, grid, block, smem, st, qkv, L.rope, pool, block_table, max_pages, pos_arr, pos0,
                            L.n_heads, L.n_kv_heads, w.hi, w.lo, max_pos));
        return CSMB_OK;
      }
      if (L.head_dim == 128) {
        CSMB_CUDA(cudaFuncSetAttribute(k_attn_decode_chunked<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CSMB_CUDA(bf_launch(w.cc, k_attn_decode_chunked<128>, grid, block, smem, st, qkv, L.rope, pool, block_table, max_pages, pos_arr, pos0,
                            L.n_heads, L.n_kv_heads, w.hi, w.lo, max_pos));
        return CSMB_OK;
      }
    }
  }
  const size_t smem = ((size_t)rps * G * L.head_dim + (size_t)G * max_pos) * sizeof(float);
  CSMB_REQUIRE(smem <= 48 * 1024);
  if (L.head_dim == 64)
    CSMB_CUDA(bf_launch(w.cc, k_attn_decode_fused<64>, grid, block, smem, st, qkv, L.rope, pool, block_table, max_pages, pos_arr, pos0,
                        rps, L.n_heads, L.n_kv_heads, w.hi, w.lo, max_pos));
  else if (L.head_dim == 128)
    CSMB_CUDA(bf_launch(w.cc, k_attn_decode_fused<128>, grid, block, smem, st, qkv, L.rope, pool, block_table, max_pages, pos_arr, pos0,
                        rps, L.n_heads, L.n_kv_heads, w.hi, w.lo, max_pos));
  else
    return CSMB_ERR_UNSUPPORTED;
  return CSMB_OK;
}

// one Llama stack over R = B*rps rows whose first normalised input is already in w.hi / w.lo.  `final_norm`, the
// weight of the norm after the last layer, is applied to rows b*rps + rps-1 only (R -> B rows).
static int bf_layers(const FastWs& w, const csmb_llama& L, float* x, float* pool, size_t layer_stride,
                     const int32_t* block_table, int max_pages, const int32_t* pos_arr, int pos0, int rps, int B,
                     float* y32_final, cudaStream_t st, int keep8 = 0) {
  const int d = L.d_model, F = L.d_ff, R = B * rps;
  const int nqkv = (L.n_heads + 2 * L.n_kv_heads) * L.head_dim;
  int rc;
  PartIn part;
  for (int l = 0; l < L.n_layers; ++l) {
    if ((rc = bf_gemm(w, L.wqkv[l], R, nqkv, d, &part, st, nullptr, nullptr, nullptr, 0, keep8))) return rc;
    if ((rc = bf_attn(w, L, part, pool + (size_t)l * layer_stride, block_table, max_pages, pos_arr, pos0, rps, B, st))) return rc;
    if ((rc = bf_gemm(w, L.wo[l], R, d, L.n_heads * L.head_dim, &part, st, nullptr, nullptr, nullptr, 0, keep8))) return rc;
    if ((rc = bf_norm(w, x, d, part, 2, L.norm_post[l], L.eps, nullptr, R, 1, 0, st))) return rc;
    if (w.cc.dbg & 4) {
      if ((rc = bf_gemm(w, L.wgu[l], R, 2 * F, d, &part, st))) return rc;
      const size_t total4 = (size_t)R * F / 4;
      CSMB_CUDA(bf_launch(w.cc, k_swiglu_split, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, st, part, F, total4, w.hi2, w.lo2));
    } else {
      // its CTAs also pull the down matrix towards L2: the down Linear's CTAs only become resident as these leave
      if ((rc = bf_gemm_gu(w, L.wgu[l], R, F, d, st, L.wdown[l], (size_t)d * F * 2, keep8))) return rc;
    }
    if ((rc = bf_gemm(w, L.wdown[l], R, d, F, &part, st, w.hi2, w.lo2, nullptr, 0, keep8))) return rc;
    if (l + 1 < L.n_layers) {
      if ((rc = bf_norm(w, x, d, part, 2, L.norm_in[l + 1], L.eps, nullptr, R, 1, 0, st))) return rc;
    } else {
      if ((rc = bf_norm(w, x, d, part, 2, L.norm_final, L.eps, y32_final, B, rps, rps - 1, st))) return rc;
    }
  }
  return CSMB_OK;
}

static bool bf_supported(const csmb_model& m, const csmb_sampler& s) {
  const csmb_llama &b = m.backbone, &d = m.decoder;
  auto llama_ok = [](const csmb_llama& L) {
    return (L.d_model == 1024 || L.d_model == 2048) && (L.head_dim == 64 || L.head_dim == 128) && L.n_kv_heads > 0 &&
           L.n_heads % L.n_kv_heads == 0 && L.n_heads / L.n_kv_heads <= 8 && L.d_ff % 64 == 0 && L.d_ff % 4 == 0 &&
           L.n_heads * L.head_dim == L.d_model;
  };
  // fused samplers: greedy; temperature with optional top-k, top-p and / or min-p (min_keep 1)
  const bool filtered = s.temperature != 0.f && s.min_p > 0.f && s.min_keep > 1;
  if (m.weight_format != CSMB_WEIGHTS_BF16) return false;   // weight-only FP8 models run on the row-based GEMV path
  return llama_ok(b) && llama_ok(d) && b.d_model == 2048 && m.audio_vocab <= 8192 && m.n_codebooks >= 2 && !filtered &&
         s.temperature >= 0.f;
}

}  // namespace csmb

using namespace csmb;

extern "C" {

size_t csmb_decode_frame_fast_workspace_bytes(const csmb_model* m, int batch) {
  if (!m || batch <= 0) return 0;
  return bf_carve(*m, batch, nullptr).bytes;
}

int csmb_decode_frame_fast_supported(const csmb_model* m, const csmb_sampler* sampler) {
  return (m && sampler && bf_supported(*m, *sampler)) ? 1 : 0;
}

int csmb_decode_frame_fast(const csmb_model* m, const csmb_batch* bt, const int32_t* prev_frame, const int32_t* pos,
                           int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, void* workspace,
                           size_t workspace_bytes, int device, void* stream) {
  return csmb_decode_frame_fast_admit(m, bt, prev_frame, pos, frame, sampler, draw_base, nullptr, nullptr, nullptr, workspace,
                                      workspace_bytes, device, stream);
}

int csmb_decode_frame_fast_admit(const csmb_model* m, const csmb_batch* bt, const int32_t* prev_frame, const int32_t* pos,
                                 int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, const float* x_override,
                                 const uint8_t* use_override, const csmb_chain_opts* opts, void* workspace,
                                 size_t workspace_bytes, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE((x_override == nullptr) == (use_override == nullptr));
  CSMB_REQUIRE(m && bt && prev_frame && pos && frame && sampler && workspace && bt->batch > 0);
  CSMB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (!bf_supported(*m, *sampler)) return CSMB_ERR_UNSUPPORTED;
  const csmb_llama &Bk = m->backbone, &D = m->decoder;
  const int B = bt->batch, db = Bk.d_model, dd = D.d_model, V = m->audio_vocab, ncb = m->n_codebooks;
  FastWs w = bf_carve(*m, B, workspace, chain_cfg(opts));
  CSMB_REQUIRE(workspace_bytes >= w.bytes);
  const float* ptab = opts ? opts->proj_table : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const int dec_pages = cdiv(ncb, CSMB_PAGE);
  BfSample sa;
  sa.inv_temp = sampler->temperature == 0.f ? 0.f : 1.f / sampler->temperature;
  sa.top_k = (sampler->top_k > 0 && sampler->top_k < V) ? sampler->top_k : 0;
  sa.min_p = sampler->min_p > 0.f ? sampler->min_p : 0.f;
  sa.top_p = (sampler->top_p > 0.f && sampler->top_p < 1.f) ? sampler->top_p : 0.f;
  sa.seed_lo = (uint32_t)sampler->seed;
  sa.seed_hi = (uint32_t)(sampler->seed >> 32);
  sa.draw_pos_mul = (uint32_t)ncb;
  int rc;
  PartIn part;
  // logits row + (staged) the head's split-K partials of the row: (1 + S) * V floats
  auto lg_smem_of = [&](int K, int* staged) {
    const size_t full = (size_t)(1 + bf_pick_split(V, K)) * V * sizeof(float);
    *staged = (full <= 96 * 1024 && !(w.cc.dbg & 16)) ? 1 : 0;
    return *staged ? full : (size_t)V * sizeof(float);
  };
  CSMB_CUDA(cudaFuncSetAttribute(k_sample_embed, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  int lg_staged = 0;
  size_t lg_smem = lg_smem_of(db, &lg_staged);
  const ProjIn no_proj{nullptr, nullptr, nullptr, 0.f, 0};

  // ---- backbone step (generation.py:34-42 with T = 1)
  CSMB_CUDA(bf_launch(w.cc, k_frame_embed_norm, dim3(B), dim3(256), 0, st, prev_frame, m->audio_emb, ncb, V, db, w.x, Bk.norm_in[0],
                      Bk.eps, w.hi, w.lo, x_override, use_override));
  if ((rc = bf_layers(w, Bk, w.x, bt->kv_pool, bt->kv_layer_stride, bt->block_table, bt->max_pages, pos, 0, 1, B, w.h_last, st)))
    return rc;
  if ((rc = bf_gemm(w, m->c0_head, B, V, db, &part, st))) return rc;
  sa.draw_base = draw_base;
  CSMB_CUDA(bf_launch(w.cc, k_sample_embed, dim3(B), dim3(256), lg_smem, st, part, V, sa, pos, 0, frame, ncb, m->audio_emb, db, 1,
                      (const float*)w.h_last, w.hi, w.lo, 2, no_proj, lg_staged));
  lg_smem = lg_smem_of(dd, &lg_staged);
  // ---- depth decoder (generation.py:56-90).  Its 222 MB of layer weights are streamed 31 times per frame-step: the Linears ask
  // L2 to keep dec_keep8 / 8 of those lines (evict-last), so that that share of every later step comes from L2
  const int dec_keep8 = (w.cc.dbg >> 12) & 7;
  for (int i = 1; i < ncb; ++i) {
    const int rps = (i == 1) ? 2 : 1, R = B * rps;
    if (i == 1 || ptab == nullptr) {
      // projection Linear + first RMSNorm; with the table, k_sample_embed of step i-1 already left dx and its planes
      if ((rc = bf_gemm(w, m->projection, R, dd, db, &part, st))) return rc;
      if ((rc = bf_norm(w, w.dx, dd, part, 1, D.norm_in[0], D.eps, nullptr, R, 1, 0, st))) return rc;
    }
    if ((rc = bf_layers(w, D, w.dx, bt->dec_kv_pool, bt->dec_kv_layer_stride, nullptr, dec_pages, nullptr, i == 1 ? 0 : i, rps, B,
                        nullptr, st, dec_keep8)))
      return rc;
    if ((rc = bf_gemm(w, m->audio_head_t + (size_t)(i - 1) * V * dd, B, V, dd, &part, st))) return rc;
    sa.draw_base = draw_base + (uint64_t)i;
    const int embed = i + 1 < ncb ? 1 : 0;
    const ProjIn pj = (ptab && embed) ? ProjIn{ptab + (size_t)i * V * dd, w.dx, D.norm_in[0], D.eps, dd} : no_proj;
    CSMB_CUDA(bf_launch(w.cc, k_sample_embed, dim3(B), dim3(256), lg_smem, st, part, V, sa, pos, i, frame, ncb, m->audio_emb, db,
                        embed, (const float*)nullptr, w.hi, w.lo, 1, pj, lg_staged));
  }
  return CSMB_OK;
}

// ---- prompt rows through the chain's kernels ------------------------------------------------------------------------------
static size_t prefill_q_bytes(const csmb_model& m, int R) {
  return ((size_t)R * m.backbone.n_heads * m.backbone.head_dim * sizeof(float) + 255) & ~(size_t)255;
}
size_t csmb_prefill_fast_workspace_bytes(const csmb_model* m, int rows) {
  if (!m || rows <= 0) return 0;
  return bf_carve(*m, rows, nullptr).bytes + prefill_q_bytes(*m, rows);
}

int csmb_prefill_fast(const csmb_model* m, const csmb_batch* bt, const int32_t* tokens, const uint8_t* mask,
                      const int32_t* row_seq, const int32_t* row_pos, int R, const int32_t* last_rows, int n_last,
                      float* h_last, float* c0_logits, void* workspace, size_t workspace_bytes, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(m && bt && tokens && mask && row_seq && row_pos && workspace && R > 0);
  CSMB_REQUIRE((n_last == 0) == (last_rows == nullptr) && n_last >= 0 && n_last <= R && (n_last == 0 || h_last));
  CSMB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  csmb_sampler greedy = {0.f, 0, 0.f, 0.f, 1, 0};
  if (!bf_supported(*m, greedy)) return CSMB_ERR_UNSUPPORTED;
  const csmb_llama& L = m->backbone;
  const int d = L.d_model, F = L.d_ff, H = L.n_heads, Hkv = L.n_kv_heads, HD = L.head_dim, V = m->audio_vocab;
  const int nqkv = (H + 2 * Hkv) * HD, G = H / Hkv, max_pos = bt->max_pages * CSMB_PAGE;
  FastWs w = bf_carve(*m, R, workspace, ChainCfg{1, 0, BF_SMEM_BUDGET});
  CSMB_REQUIRE(workspace_bytes >= w.bytes + prefill_q_bytes(*m, R));
  float* qbuf = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.bytes);
  const size_t at_smem = ((size_t)G * HD + (size_t)G * max_pos) * sizeof(float);
  CSMB_REQUIRE(at_smem <= 48 * 1024 && G >= 1 && G <= 8 && (HD == 64 || HD == 128));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  PartIn part;
  CSMB_CUDA(bf_launch(w.cc, k_prefill_embed_norm, dim3(R), dim3(256), 0, st, tokens, mask, m->text_emb, m->audio_emb, m->n_codebooks,
                      V, d, w.x, L.norm_in[0], L.eps, w.hi, w.lo));
  for (int l = 0; l < L.n_layers; ++l) {
    float* pool = bt->kv_pool + (size_t)l * bt->kv_layer_stride;
    if ((rc = bf_gemm(w, L.wqkv[l], R, nqkv, d, &part, st))) return rc;
    CSMB_CUDA(bf_launch(w.cc, k_prefill_rope_append, dim3(R, cdiv(nqkv / 2, 512)), dim3(512), 0, st, part, L.rope, pool,
                        bt->block_table, bt->max_pages, row_seq, row_pos, H, Hkv, HD, qbuf));
    if (HD == 64 && G == PA_G) {
      const size_t tile_smem = ((size_t)2 * PA_KC * (HD + 4) + PA_WARPS * PA_G * HD + (size_t)PA_WARPS * PA_KC * PA_G) * sizeof(float);
      CSMB_CUDA(cudaFuncSetAttribute(k_prefill_attn_tile<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
      CSMB_CUDA(bf_launch(w.cc, k_prefill_attn_tile<64>, dim3(cdiv(R, PA_QT), Hkv), dim3(PA_WARPS * 32), tile_smem, st, (const float*)qbuf,
                          (const float*)pool, bt->block_table, bt->max_pages, row_seq, row_pos, R, H, Hkv, w.hi, w.lo));
    } else if (HD == 64)
      CSMB_CUDA(bf_launch(w.cc, k_prefill_attn<64>, dim3(R * Hkv), dim3(32 * G), at_smem, st, (const float*)qbuf, (const float*)pool,
                          bt->block_table, bt->max_pages, row_seq, row_pos, H, Hkv, w.hi, w.lo, max_pos));
    else
      CSMB_CUDA(bf_launch(w.cc, k_prefill_attn<128>, dim3(R * Hkv), dim3(32 * G), at_smem, st, (const float*)qbuf, (const float*)pool,
                          bt->block_table, bt->max_pages, row_seq, row_pos, H, Hkv, w.hi, w.lo, max_pos));
    if ((rc = bf_gemm(w, L.wo[l], R, d, H * HD, &part, st))) return rc;
    if ((rc = bf_norm(w, w.x, d, part, 2, L.norm_post[l], L.eps, nullptr, R, 1, 0, st))) return rc;
    if ((rc = bf_gemm_gu(w, L.wgu[l], R, F, d, st))) return rc;
    if ((rc = bf_gemm(w, L.wdown[l], R, d, F, &part, st, w.hi2, w.lo2))) return rc;
    if (l + 1 < L.n_layers) {
      if ((rc = bf_norm(w, w.x, d, part, 2, L.norm_in[l + 1], L.eps, nullptr, R, 1, 0, st))) return rc;
    } else if (n_last > 0) {
      // final norm on the listed rows only (generation.py:40: the last position of every sequence)
      if ((rc = bf_norm(w, w.x, d, part, 2, L.norm_final, L.eps, h_last, n_last, 1, 0, st, last_rows))) return rc;
    }
  }
  if (n_last > 0 && c0_logits) {
    if ((rc = bf_gemm(w, m->c0_head, n_last, V, d, &part, st))) return rc;
    const size_t total = (size_t)n_last * V;
    k_part_reduce1<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, total, c0_logits);
    CSMB_LAUNCH_CHECK();
  }
  return CSMB_OK;
}

#ifdef CSMB_TIMELINE
// debug builds only: copies the recorded slots ([n][TL_WORDS] words) to `out`, returns the number recorded, optionally resets
int csmb_debug_timeline_read(unsigned long long* out, int max_slots, int reset) {
  unsigned n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_tl_n, sizeof(n));
  const int m = (int)(n < (unsigned)max_slots ? n : (unsigned)max_slots);
  if (out && m > 0) cudaMemcpyFromSymbol(out, g_tl, (size_t)m * TL_WORDS * sizeof(unsigned long long));
  if (reset) {
    n = 0;
    cudaMemcpyToSymbol(g_tl_n, &n, sizeof(n));
  }
  return m;
}
#endif

// ---- projected-embedding table -------------------------------------------------------------------------------------
size_t csmb_proj_table_bytes(const csmb_model* m) {
  if (!m) return 0;
  return (size_t)m->n_codebooks * m->audio_vocab * m->decoder.d_model * sizeof(float);
}
size_t csmb_proj_table_workspace_bytes(const csmb_model* m) {
  if (!m) return 0;
  const size_t V = m->audio_vocab, db = m->backbone.d_model, dd = m->decoder.d_model;
  return 256 + ((V * db * 2 + 255) & ~(size_t)255) + bf_part_floats((int)V, (int)dd, (int)db) * sizeof(float);
}

int csmb_build_proj_table(const csmb_model* m, float* table, void* workspace, size_t workspace_bytes, int device, void* stream) {
  CSMB_ENTER(device);
  CSMB_REQUIRE(m && table && workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  if (m->weight_format != CSMB_WEIGHTS_BF16) return CSMB_ERR_UNSUPPORTED;
  CSMB_REQUIRE(workspace_bytes >= csmb_proj_table_workspace_bytes(m));
  const int V = m->audio_vocab, db = m->backbone.d_model, dd = m->decoder.d_model, ncb = m->n_codebooks;
  CSMB_REQUIRE(db % TC_BK == 0 && dd % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  char* p = static_cast<char*>(workspace);
  FastWs w = {};
  w.cc = ChainCfg{0, 0, BF_SMEM_BUDGET};
  w.err = reinterpret_cast<int*>(p);
  uint16_t* zero = reinterpret_cast<uint16_t*>(p + 256);
  const size_t zbytes = ((size_t)V * db * 2 + 255) & ~(size_t)255;
  w.part = reinterpret_cast<float*>(p + 256 + zbytes);
  w.part_floats = bf_part_floats(V, dd, db);
  CSMB_CUDA(cudaMemsetAsync(p, 0, 256 + zbytes, st));
  int rc;
  for (int cb = 0; cb < ncb; ++cb) {
    // the rows of codebook cb are exact bf16 values: hi plane = the embedding table itself, lo plane = 0, which is what
    // k_sample_embed hands to the projection Linear of the chain; same kernel, same split, same summation order
    PartIn part;
    if ((rc = bf_gemm(w, m->projection, V, dd, db, &part, st, m->audio_emb + (size_t)cb * V * db, zero))) return rc;
    const size_t total4 = (size_t)V * dd / 4;
    k_part_reduce<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(part, total4, table + (size_t)cb * V * dd);
    CSMB_LAUNCH_CHECK();
  }
  return CSMB_OK;
}

}  // extern "C"
