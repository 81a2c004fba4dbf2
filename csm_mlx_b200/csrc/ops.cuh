// Internal launchers for the elementary LM kernels (no device guard; caller has made the device current).
#pragma once
#include "common.cuh"

namespace csmb {

int launch_embed_sum(const int32_t* tokens, const uint8_t* mask, const uint16_t* text_emb,
                     const uint16_t* audio_emb, float* out, int R, int d, int ncb, int audio_vocab,
                     cudaStream_t st);
// tokens read with stride tok_stride; token < 0 is clamped to 0.  If `src` != nullptr and sel_col >= 0 the
// token is instead taken from forced when forced != nullptr (teacher forcing).
int launch_embed_audio(const int32_t* tokens, int tok_stride, const uint16_t* audio_emb, float* out, int ldo,
                       int R, int d, int codebook, int audio_vocab, cudaStream_t st);
int launch_rmsnorm(const float* x, int ldx, const float* w, float* y, int ldy, int R, int d, float eps,
                   const int32_t* row_idx, cudaStream_t st);
int launch_linear(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
                  int accumulate, cudaStream_t st);
// e4m3 blob (include/csm_b200.h, CSMB_WEIGHTS_E4M3): scales then bytes
#ifdef __CUDACC__
#define CSMB_HD __host__ __device__
#else
#define CSMB_HD
#endif
CSMB_HD static inline size_t e4m3_scale_bytes(int N) { return ((size_t)N * 4 + 255) & ~(size_t)255; }
CSMB_HD static inline size_t e4m3_blob_bytes(int N, int K) { return e4m3_scale_bytes(N) + (((size_t)N * K + 255) & ~(size_t)255); }
int launch_linear_e4m3(const float* x, int ldx, const void* blob, float* y, int ldy, int R, int N, int K, int accumulate,
                       cudaStream_t st);
int launch_swiglu(const float* gu, float* out, int R, int F, cudaStream_t st);
int launch_rope_kv_append(float* qkv, const float* rope, float* kv_pool, const int32_t* block_table,
                          int max_pages, const int32_t* row_seq, const int32_t* row_pos, int R, int H, int Hkv,
                          int hd, cudaStream_t st);
int launch_attention(const float* qkv, int ldq, const float* kv_pool, const int32_t* block_table, int max_pages,
                     const int32_t* row_seq, const int32_t* row_pos, float* out, int R, int H, int Hkv, int hd,
                     int max_pos, cudaStream_t st);
// draw index for row r: draw_base + (row_pos ? row_pos[r] : 0) * draw_pos_mul; sequence word = r.
int launch_sample(const float* logits, int ldl, int32_t* out, int out_stride, int R, int V,
                  const csmb_sampler& s, uint64_t draw_base, const int32_t* row_pos, uint32_t draw_pos_mul,
                  const int32_t* forced, int forced_stride, cudaStream_t st);

// tcgen05 / TMEM / TMA linear for R >= 9 rows, K % 64 == 0 (gemm_tc.cu).  workspace: linear_tc_workspace_bytes(R, N, K), zeroed once.
size_t linear_tc_workspace_bytes(int R, int N, int K);
int launch_linear_tc(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K, int accumulate,
                     void* workspace, size_t workspace_bytes, cudaStream_t st);

}  // namespace csmb
