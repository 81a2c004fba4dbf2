// tcgen05 / TMEM / TMA / mbarrier / programmatic-dependent-launch helpers shared by the tensor-core linears
// (gemm_tc.cu: prefill and the per-op API; batch_frame.cu: the fused batched decode frame).  sm_100a only.
#pragma once
#include <cuda.h>

#include "ops.cuh"

namespace csmb {

constexpr int TC_BM = 128;      // weight rows per tile (UMMA M)
constexpr int TC_BK = 64;       // K per stage = one 128-byte swizzle row of bf16
constexpr int TC_THREADS = 192; // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr unsigned TC_SPIN = 1u << 24;

// bf16 [rows][K] row-major tensor map, box = box_rows x 64 elements, 128-byte swizzle, zero fill out of bounds
bool tc_make_map(CUtensorMap* m, const void* base, int rows, int K, int box_rows);
// the same with a leading dimension: row r starts at element r * ld (ld >= K, ld % 8 == 0)
bool tc_make_map_ld(CUtensorMap* m, const void* base, long long rows, int K, int ld, int box_rows);

#ifdef __CUDACC__

// ---- bf16 hi/lo split of fp32 activations: x = hi + lo up to 2^-17 relative ---------------------------------------
// Round to nearest even with the hardware conversion (cvt.rn.bf16[x2].f32: one instruction instead of ~5 integer ops per
// value; for finite inputs the same bits as the integer form u += 0x7fff + ((u >> 16) & 1), which this replaced).
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
  uint16_t h;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(f));
  return h;
}
// two fp32 -> packed bf16 pair (element 0 in the low half)
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float e0, float e1) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
__device__ __forceinline__ void split_bf16(float f, uint16_t& h, uint16_t& l) {
  h = f32_to_bf16_rn(f);
  l = f32_to_bf16_rn(f - __uint_as_float((uint32_t)h << 16));
}
// four values -> 4 hi + 4 lo, 8-byte stores (dst 8-byte aligned)
__device__ __forceinline__ void store_split4(uint16_t* hi, uint16_t* lo, float a, float b, float c, float d) {
  const uint32_t h01 = pack_bf16x2_rn(a, b), h23 = pack_bf16x2_rn(c, d);
  const uint32_t l01 = pack_bf16x2_rn(a - __uint_as_float(h01 << 16), b - __uint_as_float(h01 & 0xffff0000u));
  const uint32_t l23 = pack_bf16x2_rn(c - __uint_as_float(h23 << 16), d - __uint_as_float(h23 & 0xffff0000u));
  *reinterpret_cast<uint2*>(hi) = make_uint2(h01, h23);
  *reinterpret_cast<uint2*>(lo) = make_uint2(l01, l23);
}

// ---- programmatic dependent launch (griddepcontrol): both are no-ops for a kernel launched without the attribute ----
// launch_dependents: the next kernel of the stream may start its prologue now; wait: everything the previous kernel
// wrote is visible from here on.  EVERY kernel of a PDL chain must execute pdl_wait(), otherwise completion is no
// longer transitive along the stream.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier / TMA / UMMA PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: returns false on timeout (the kernel then drains without touching memory it does not own)
__device__ __forceinline__ bool tc_mbar_wait(uint64_t* b, uint32_t parity, int* err) {
  unsigned spins = 0;
  while (!tc_mbar_try(b, parity)) {
    if (++spins > TC_SPIN) {
      atomicExch(err, 1);
      return false;
    }
    if ((spins & 4095) == 0 && *reinterpret_cast<volatile int*>(err) != 0) return false;
  }
  return true;
}
// Warp-collective form: lane 0 polls, the result is broadcast (a whole warp spinning on mbarrier.try_wait floods the shared
// memory pipe that the TMA writes and the epilogue's staging traffic also need).  All 32 lanes must call it.
__device__ __forceinline__ bool tc_mbar_wait_warp(uint64_t* b, uint32_t parity, int* err) {
  int ok = 1;
  if ((threadIdx.x & 31) == 0) ok = tc_mbar_wait(b, parity, err) ? 1 : 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(x), "r"(y)
      : "memory");
}
// the same with an L2 eviction policy (createpolicy): weights are read once per frame-step — evict-first keeps them from
// pushing the step's small, re-read data (KV rows, partials, planes, norm weights, RoPE rows) out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// evict-last for `fraction` of the lines, evict-first for the rest: a matrix set that is streamed again and again (the depth
// decoder: 31 times per frame) and is larger than L2 keeps that share of itself resident
__device__ __forceinline__ uint64_t l2_policy_keep_fraction(float fraction) {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(p) : "f"(fraction));
  return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(x), "r"(y), "l"(policy)
      : "memory");
}
// L2 prefetch of a tensor-map box / of a byte range (fire and forget: no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {   // p 16-byte aligned, bytes % 16 == 0
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle, rows of 64 bf16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);        // [0,14)  start address >> 4
  d |= (uint64_t)1 << 16;                              // [16,30) leading byte offset >> 4 (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // [32,46) stride byte offset >> 4: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                              // [46,48) descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // [61,64) layout: SWIZZLE_128B
  return d;
}
// instruction descriptor: kind::f16, A = B = bf16, D = fp32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 1u << 7;                      // a_format = BF16
  d |= 1u << 10;                     // b_format = BF16
  d |= (uint32_t)(n >> 3) << 17;     // n_dim
  d |= (uint32_t)(TC_BM >> 4) << 24; // m_dim
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a 2-CTA cluster (one TPC) run ONE tcgen05.mma of M = 256 ---------------------
// CTA r holds weight rows [128 r, 128 r + 128) of the A tile and half of the B operand's N rows at the SAME shared-memory
// offsets; accumulator rows 128 r .. land in CTA r's own TMEM.  Only the leader (rank 0) issues the instruction; both CTAs load
// their halves with the cta_group::2 form of the TMA copy, whose complete_tx goes to the LEADER's mbarrier; the leader's
// commit is multicast to the barrier at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// every thread of every CTA of the cluster (a superset of __syncthreads)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the executing CTA's shared-memory address `saddr` as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// box -> the executing CTA's shared memory; bytes are counted on the mbarrier at cluster address `bar_cluster`
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int x, int y, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s32(dst)),
      "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* map, int x, int y, uint32_t bar_cluster,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(s32(dst)),
      "l"(map), "r"(bar_cluster), "r"(x), "r"(y), "l"(policy)
      : "memory");
}
// instruction descriptor as umma_idesc with M = 256 (pair)
__device__ __forceinline__ uint32_t umma_idesc_pair(int n) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)((2 * TC_BM) >> 4) << 24;
  return d;
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once all earlier tcgen05.mma of this thread are complete) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of the accumulator -> registers (waits for the load)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 columns without the wait: the caller issues several loads and then one tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of the accumulator -> registers (waits for the load)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// (the cvt form is now store_split4 itself; the name is kept for the codec's callers)
__device__ __forceinline__ void store_split4_cvt(uint16_t* hi, uint16_t* lo, float a, float b, float c, float d) {
  store_split4(hi, lo, a, b, c, d);
}

#endif  // __CUDACC__

}  // namespace csmb
