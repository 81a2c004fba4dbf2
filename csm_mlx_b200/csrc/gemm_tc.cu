// Tensor-core linear for many rows (prefill, batched decode): tcgen05.mma with TMEM accumulators fed by TMA.
//
//   Y[R][N] (+)= X[R][K] . W[N][K]^T        W bf16 (HBM, read once), X fp32, Y fp32
//
// "Swap-AB": the weight matrix is the UMMA A operand (M = 128 weight rows per CTA tile, K-major, straight from the
// checkpoint layout) and the token rows are the B operand (N = up to 256 rows, K-major), so a decode batch of 16..256
// sequences or a prompt of that many rows fills the N dimension of ONE instruction shape and the weights are streamed
// exactly once.  kind::f16 needs 16-bit operands: X is split on the fly into bf16 hi + bf16 lo (x = hi + lo up to
// 2^-17 relative), and two MMAs per K-step accumulate W.hi + W.lo into the same fp32 TMEM accumulator — the weights
// are exactly bf16, so the result matches an fp32 reference to ~1e-5 relative instead of bf16's 4e-3.
//
// Pipeline per CTA (192 threads): warp 0 = TMA producer (one elected lane; 128B-swizzled boxes W 128x64, Xhi/Xlo
// RNx64 per stage, 4 stages, mbarrier complete_tx), warp 1 = MMA issuer (one lane; 4 x 2 tcgen05.mma per stage,
// tcgen05.commit frees the stage / publishes the accumulator), warps 2-5 = epilogue (tcgen05.ld 32 lanes x 32 columns,
// residual add, coalesced stores: for a fixed token row the 32 lanes of a warp write 32 consecutive features).
// Reference ops replaced: every nn.Linear of the path at T > 1 / B > 8 (attention.py:216-218,253; mlx_lm MLP;
// generation.py:42,75,79).
#include "tc.cuh"

namespace csmb {

constexpr int TC_STAGES = 4;

// ---- bf16 hi/lo split of the activations -----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split_bf16(const float* __restrict__ x, int ldx, uint16_t* __restrict__ hi,
                                                    uint16_t* __restrict__ lo, int K, size_t total) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  const size_t r = i / K, k = i % K;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + k);
  const float f[4] = {v.x, v.y, v.z, v.w};
  uint16_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = f32_to_bf16_rn(f[j]);
    l[j] = f32_to_bf16_rn(f[j] - __uint_as_float((uint32_t)h[j] << 16));
  }
  *reinterpret_cast<uint2*>(hi + i) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
  *reinterpret_cast<uint2*>(lo + i) = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
}

struct TcArgs {
  float* y;
  int ldy, R, N, K, accumulate, RN;  // RN: token rows per CTA tile (multiple of 16, <= 256)
  int nstages;                       // smem pipeline depth (<= TC_STAGES)
  int S;                             // split-K factor = gridDim.z; S > 1: fp32 partials [S][R][N] instead of y
  float* part;
  int* err;
};

// dynamic smem: [stage][ W 128x64 | Xhi RNx64 | Xlo RNx64 ] bf16, 1024-byte aligned tiles
__global__ void __launch_bounds__(TC_THREADS, 1)
k_linear_tc(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_hi,
            const __grid_constant__ CUtensorMap map_lo, const TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  // 128B-swizzled tiles must start on 1024-byte boundaries; the dynamic segment only guarantees 16
  unsigned char* smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TC_BM, r0 = blockIdx.y * a.RN;
  const int RN = a.RN;
  const uint32_t w_bytes = TC_BM * TC_BK * 2, x_bytes = (uint32_t)RN * TC_BK * 2;
  const uint32_t x_off = w_bytes, stage_bytes = (w_bytes + 2 * x_bytes + 1023u) & ~1023u;
  const int nk_total = a.K / TC_BK, NS = a.nstages;
  const int kb0 = (int)(((long long)nk_total * blockIdx.z) / a.S), kb1 = (int)(((long long)nk_total * (blockIdx.z + 1)) / a.S);
  const int nk = kb1 - kb0;
  uint32_t ncols = 32;
  while ((int)ncols < RN) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) {
      tc_mbar_init(&full[i], 1);
      tc_mbar_init(&empty[i], 1);
    }
    tc_mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM allocation (one warp), address lands in shared memory
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % NS;
        const uint32_t par = (kb / NS) & 1;
        if (!tc_mbar_wait(&empty[s], par ^ 1, a.err)) break;
        unsigned char* st = smem + (size_t)s * stage_bytes;
        tc_mbar_expect_tx(&full[s], w_bytes + 2 * x_bytes);
        tma_load_2d(st, &map_w, (kb0 + kb) * TC_BK, n0, &full[s]);
        tma_load_2d(st + x_off, &map_hi, (kb0 + kb) * TC_BK, r0, &full[s]);
        tma_load_2d(st + x_off + x_bytes, &map_lo, (kb0 + kb) * TC_BK, r0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(RN);
      bool ok = true;
      for (int kb = 0; kb < nk && ok; ++kb) {
        const int s = kb % NS;
        const uint32_t par = (kb / NS) & 1;
        ok = tc_mbar_wait(&full[s], par, a.err);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = s32(smem + (size_t)s * stage_bytes);
        const uint64_t da = umma_desc(sa), dhi = umma_desc(sa + x_off), dlo = umma_desc(sa + x_off + x_bytes);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint64_t koff = (uint64_t)((k * 32) >> 4);  // 16 bf16 = 32 bytes along K inside the swizzled row
          umma_f16(tmem_base, da + koff, dhi + koff, idesc, (kb | k) != 0);
          umma_f16(tmem_base, da + koff, dlo + koff, idesc, 1u);
        }
        umma_commit(&empty[s]);  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(&acc_full);    // accumulator complete (tcgen05.commit tracks all prior MMAs of this thread)
    }
  } else {
    // ===== epilogue: warps 2..5 -> TMEM lane quarters (warp % 4) =====
    const int quarter = warp & 3;
    const bool ok = tc_mbar_wait(&acc_full, 0, a.err);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int n = n0 + quarter * 32 + lane;
    if (ok) {
      for (int c0 = 0; c0 < RN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        tmem_ld32(taddr, v);
        if (n < a.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int r = r0 + c0 + j;
            if (c0 + j < RN && r < a.R) {
              const float val = __uint_as_float(v[j]);
              if (a.S > 1) {
                a.part[((size_t)blockIdx.z * a.R + r) * a.N + n] = val;
              } else {
                float* dst = a.y + (size_t)r * a.ldy + n;
                *dst = a.accumulate ? *dst + val : val;
              }
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// split-K: y[r][n] (+)= sum_z part[z][r][n], fixed order (deterministic)
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, int S, float* __restrict__ y, int ldy,
                                                       int R, int N, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * N) return;
  const size_t r = i / N, n = i % N;
  float v = 0.f;
  for (int z = 0; z < S; ++z) v += part[(size_t)z * R * N + i];
  float* dst = y + r * ldy + n;
  *dst = accumulate ? *dst + v : v;
}

// ---- host side ---------------------------------------------------------------------------------------------------
// Split-K factor: a function of the Linear's shape (N, K) ONLY.  A row's K blocks are therefore summed in the same
// order whatever the number of rows in the call (and whichever other sequences' rows they are): a sequence's tokens do
// not depend on its neighbours in a batch, like the reference's batch-1 loop (generation.py:139-161).
static int pick_split(int N, int K) {
  const int tiles = cdiv(N, TC_BM);
  const int nk = K / TC_BK;
  int S = 1;
  while (S < 16 && tiles * S * 2 <= 160 && nk / (S * 2) >= 4) S *= 2;
  return S;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// bf16 [rows][K] row-major, box = box_rows x 64 elements, 128-byte swizzle, zero fill out of bounds
bool tc_make_map(CUtensorMap* m, const void* base, int rows, int K, int box_rows) {
  return tc_make_map_ld(m, base, rows, K, K, box_rows);
}
// the same with a leading dimension: row r starts at element r * ld (ld >= K, ld % 8 == 0)
bool tc_make_map_ld(CUtensorMap* m, const void* base, long long rows, int K, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t linear_tc_workspace_bytes(int R, int N, int K) {
  const int S = pick_split(N, K);
  return (size_t)2 * R * K * sizeof(uint16_t) + (S > 1 ? (size_t)S * R * N * sizeof(float) : 0) + 1024;
}

// workspace: 2*R*K bf16 (hi, lo) + an int error flag; all on `st`
int launch_linear_tc(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K, int accumulate,
                     void* workspace, size_t workspace_bytes, cudaStream_t st) {
  CSMB_REQUIRE(R > 0 && N > 0 && K > 0 && K % TC_BK == 0 && ldx % 4 == 0 && workspace);
  CSMB_REQUIRE(workspace_bytes >= linear_tc_workspace_bytes(R, N, K));
  CSMB_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
  int* err = reinterpret_cast<int*>(workspace);
  uint16_t* hi = reinterpret_cast<uint16_t*>(reinterpret_cast<char*>(workspace) + 256);
  uint16_t* lo = hi + (size_t)R * K;
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(lo + (size_t)R * K) + ((256 - (((size_t)4 * R * K) & 255)) & 255));
  const size_t total = (size_t)R * K;
  const int S = pick_split(N, K);
  // the error flag (first int of the workspace) is sticky: the owner zeroes the workspace once
  k_split_bf16<<<(unsigned)((total / 4 + 255) / 256), 256, 0, st>>>(x, ldx, hi, lo, K, total);
  CSMB_LAUNCH_CHECK();
  // token rows per tile: all of them if <= 256, else tiles of 128
  int RN = R <= 256 ? ((R + 15) / 16) * 16 : 128;
  CUtensorMap mw, mhi, mlo;
  if (!tc_make_map(&mw, W, N, K, TC_BM) || !tc_make_map(&mhi, hi, R, K, RN) || !tc_make_map(&mlo, lo, R, K, RN)) return CSMB_ERR_UNSUPPORTED;
  const size_t stage = ((size_t)TC_BM * TC_BK * 2 + 2 * (size_t)RN * TC_BK * 2 + 1023) & ~(size_t)1023;
  int nstages = (int)((200 * 1024) / stage);
  nstages = nstages > TC_STAGES ? TC_STAGES : nstages;
  CSMB_REQUIRE(nstages >= 2);
  TcArgs a{y, ldy, R, N, K, accumulate, RN, nstages, S, part, err};
  const size_t smem = stage * nstages + 1024;
  CSMB_CUDA(cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(N, TC_BM), cdiv(R, RN), S);
  k_linear_tc<<<grid, TC_THREADS, smem, st>>>(mw, mhi, mlo, a);
  CSMB_LAUNCH_CHECK();
  if (S > 1) {
    const size_t tot = (size_t)R * N;
    k_splitk_reduce<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(part, S, y, ldy, R, N, accumulate);
    CSMB_LAUNCH_CHECK();
  }
  return CSMB_OK;
}

}  // namespace csmb

using namespace csmb;

extern "C" {

size_t csmb_linear_tc_workspace_bytes(int R, int N, int K) { return linear_tc_workspace_bytes(R, N, K) + 256; }

/* Tensor-core (tcgen05 / TMEM / TMA) variant of csmb_linear for R >= 9 rows and K % 64 == 0.
 * workspace: csmb_linear_tc_workspace_bytes(R, N, K) bytes, 256-byte aligned, zeroed once by its owner. */
int csmb_linear_tc(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K, int accumulate,
                   void* workspace, size_t workspace_bytes, int device, void* stream) {
  CSMB_ENTER(device);
  return launch_linear_tc(x, ldx, W, y, ldy, R, N, K, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
