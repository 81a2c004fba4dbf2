/* csm_b200.h — C ABI of libcsm_b200.so: the sm_100a kernels behind the csm_mlx generation hot path.
 *
 * The reference (sethdford/csm-mlx) has no FFI of its own: every tensor op is dispatched into
 * mlx / mlx_lm / moshi_mlx from Python.  This header is the boundary a binding for that path would
 * target; each entry point cites the reference call it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - plain pointers + sizes only; every pointer is a DEVICE pointer unless the name says host.
 *  - activations fp32, Linear/embedding weights bf16 (uint16_t bit patterns), norm weights fp32.
 *  - every function takes the CUDA device ordinal and a cudaStream_t (as void*), is asynchronous on
 *    that stream, allocates nothing and may be called from any host thread (the reference's demo
 *    hops threads between frames, run_streaming_csm_mlx.py:984-1000).  The library keeps NO mutable
 *    state that changes what a later call computes: tuning and debug switches travel per call in
 *    caller-owned structs (csmb_chain_opts, csmb_frame_opts, NULL = defaults).  The only process-wide
 *    data are diagnostics (the launch counter and the text of the last CUDA error).
 *  - batch invariance: what a call computes for one row / one sequence never depends on how many other
 *    rows or sequences share the call (split-K factors and tilings are functions of the Linear's shape
 *    only), matching the reference's batch-1 loop (generation.py:139-161).
 *  - return value: 0 = ok, negative = csmb_status; csmb_strerror() gives text.
 *  - "rows": all LM ops work on R flattened token rows; row r belongs to sequence row_seq[r] at
 *    position row_pos[r] (prefill: many rows per sequence; decode: one row per sequence).
 *  - KV caches are paged: pool layout [n_pages][2 (K,V)][n_kv_heads][CSMB_PAGE][head_dim] fp32 per
 *    layer, block_table[seq][logical_page] -> physical page.
 */
#ifndef CSM_B200_H
#define CSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSMB_ABI_VERSION 3
#define CSMB_PAGE 16          /* tokens per KV page */
#define CSMB_MAX_LAYERS 16
#define CSMB_MAX_CODEBOOKS 32

typedef enum {
  CSMB_OK = 0,
  CSMB_ERR_INVALID = -1,      /* bad argument (shape, alignment, null pointer) */
  CSMB_ERR_CUDA = -2,         /* a CUDA runtime call failed; see csmb_last_cuda_error */
  CSMB_ERR_UNSUPPORTED = -3,  /* configuration not compiled in */
  CSMB_ERR_DEVICE = -4        /* device is not sm_100 */
} csmb_status;

int csmb_abi_version(void);
const char* csmb_strerror(int status);
/* text of the last CUDA error seen by this library on any thread (diagnostic only; racy by nature). */
const char* csmb_last_cuda_error(void);
/* number of kernels this library has launched (enqueued or captured) so far in this process; diagnostic. */
unsigned long long csmb_debug_launch_count(void);
/* 0 if `device` is an sm_100 part this library was compiled for. */
int csmb_check_device(int device);

/* ---------------------------------------------------------------- elementary ops (LM) ---------- */

/* CSM.embed_tokens + mask-multiply + sum(-2)  (csm_mlx/models.py:82-92, generation.py:32-36).
 * tokens/mask [R][n_codebooks+1] (audio codebooks first, text id last); out [R][d] fp32. */
int csmb_embed_sum(const int32_t* tokens, const uint8_t* mask, const uint16_t* text_emb,
                   const uint16_t* audio_emb, float* out, int R, int d, int n_codebooks, int audio_vocab,
                   int device, void* stream);

/* CSM.embed_audio (models.py:79-80): out[r] = audio_emb[tokens[r] + codebook*audio_vocab], fp32. */
int csmb_embed_audio(const int32_t* tokens, const uint16_t* audio_emb, float* out, int ldo, int R, int d,
                     int codebook, int audio_vocab, int device, void* stream);

/* nn.RMSNorm of mlx_lm's TransformerBlock (constructed models.py:50-51): y = x*rsqrt(mean(x^2)+eps)*w. */
int csmb_rmsnorm(const float* x, int ldx, const float* w, float* y, int ldy, int R, int d, float eps,
                 int device, void* stream);

/* nn.Linear without bias (attention.py:216-218,253; mlx_lm MLP; codebook0_head generation.py:42;
 * projection generation.py:75):  y[R][N] (+)= x[R][K] . W[N][K]^T.  accumulate!=0 adds into y (residual). */
int csmb_linear(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
                int accumulate, int device, void* stream);

/* Tensor-core variant of csmb_linear for many rows (prefill, batched decode): tcgen05.mma (kind::f16, fp32 TMEM
 * accumulators) fed by TMA; the weight tile is the UMMA A operand (M = 128 rows), the token rows are N (<= 256 per
 * tile); X is split into bf16 hi + lo on the fly and both halves are accumulated, so results match fp32 math on the
 * bf16 weights to ~1e-5.  Small-N shapes are split along K over up to 16 CTAs per tile (fp32 partials, fixed-order
 * reduction; the split factor depends on (N, K) only, never on R).  Requires K % 64 == 0; workspace = csmb_linear_tc_workspace_bytes(R, N, K) bytes, 256-byte aligned, zeroed
 * once by its owner (first int = sticky error flag set if an internal bounded wait timed out). */
size_t csmb_linear_tc_workspace_bytes(int R, int N, int K);
int csmb_linear_tc(const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K, int accumulate,
                   void* workspace, size_t workspace_bytes, int device, void* stream);

/* SwiGLU of mlx_lm MLP: out[r][f] = silu(gu[r][f]) * gu[r][F+f]   (gu = fused gate|up output). */
int csmb_swiglu(const float* gu, float* out, int R, int F, int device, void* stream);

/* Llama3ScaledRoPE.__call__ on q and k (attention.py:119-177, 226-228) + KVCache.update_and_fetch
 * (attention.py:236-237).  qkv [R][(H+2*Hkv)*hd] (q heads, then k heads, then v heads): q is rotated in
 * place; rotated k and v are written to the paged pool at (row_seq, row_pos).
 * rope [max_pos][hd/2][2] fp32 (cos, sin); adjacent-pair convention. */
int csmb_rope_kv_append(float* qkv, const float* rope, float* kv_pool, const int32_t* block_table,
                        int max_pages, const int32_t* row_seq, const int32_t* row_pos, int R, int n_heads,
                        int n_kv_heads, int head_dim, int device, void* stream);

/* mx.repeat + scaled_dot_product_attention (attention.py:242-249): for row r, head h, softmax over
 * positions 0..row_pos[r] of sequence row_seq[r] (causal), GQA head h -> kv head h/(H/Hkv).
 * q is read from qkv (leading dim ldq); out [R][H*hd]. */
int csmb_attention(const float* qkv, int ldq, const float* kv_pool, const int32_t* block_table,
                   int max_pages, const int32_t* row_seq, const int32_t* row_pos, float* out, int R,
                   int n_heads, int n_kv_heads, int head_dim, int device, void* stream);

/* Sampling (generation.py:51-54, 81-84 and mlx_lm.sample_utils.make_sampler as used by
 * cli/generate.py:168-174).  logits [R][V] fp32 -> out[r*out_stride] int32.
 * temperature==0: argmax (lowest index on ties).  Otherwise top-k (k>0), top-p (0<p<1), min-p (>0,
 * keeping at least min_keep) filters on softmax(logits) — top-p on exact integer masses floor(exp(l - max) * 2^32): a
 * token stays while the mass of the strictly more likely tokens is below top_p * total, independent of any summation
 * order, so every sampler of the library picks the same nucleus —, then categorical(logits/temperature) by the
 * Gumbel-max trick: argmax_i(logits[i]/temperature - log(-log(u_i))), u_i from Philox4x32-10 with
 * key = seed and counter = (i/4, d_lo, d_hi, r), word i%4 of the output block, where the draw index
 * d = draw + (row_pos ? row_pos[r] : 0) * pos_mul  (row_pos: optional DEVICE array [R]). */
typedef struct {
  float temperature;
  int top_k;
  float top_p;
  float min_p;
  int min_keep;
  uint64_t seed;
} csmb_sampler;

int csmb_sample(const float* logits, int ldl, int32_t* out, int out_stride, int R, int V,
                const csmb_sampler* sampler /*host*/, uint64_t draw, const int32_t* row_pos, uint32_t pos_mul,
                int device, void* stream);

/* ---------------------------------------------------------------- fused LM path ---------------- */

typedef struct {
  int n_layers, d_model, n_heads, n_kv_heads, head_dim, d_ff;
  float eps;
  const uint16_t* wqkv[CSMB_MAX_LAYERS];   /* [(H+2Hkv)*hd][d]   q rows, k rows, v rows */
  const uint16_t* wo[CSMB_MAX_LAYERS];     /* [d][H*hd] */
  const uint16_t* wgu[CSMB_MAX_LAYERS];    /* [2*d_ff][d]        gate rows then up rows */
  const uint16_t* wdown[CSMB_MAX_LAYERS];  /* [d][d_ff] */
  const float* norm_in[CSMB_MAX_LAYERS];
  const float* norm_post[CSMB_MAX_LAYERS];
  const float* norm_final;
  const float* rope;                       /* [max_pos][hd/2][2] */
} csmb_llama;

typedef struct {
  csmb_llama backbone, decoder;
  const uint16_t* text_emb;     /* [n_text_vocab][d_b] */
  const uint16_t* audio_emb;    /* [n_codebooks*audio_vocab][d_b] */
  const uint16_t* projection;   /* [d_d][d_b] */
  const uint16_t* c0_head;      /* [audio_vocab][d_b] */
  const uint16_t* audio_head_t; /* [n_codebooks-1][audio_vocab][d_d]: checkpoint's (in,out) transposed */
  int n_text_vocab, audio_vocab, n_codebooks, max_pos;
  int weight_format;            /* CSMB_WEIGHTS_BF16 (0) or CSMB_WEIGHTS_E4M3: every nn.Linear matrix above (wqkv, wo, wgu, wdown,
                                   projection, c0_head, audio_head_t — not the embedding tables) is an e4m3 blob, see below */
  int reserved;
} csmb_model;

/* Weight-only FP8 (the analogue of mlx.nn.quantize on the reference's Linear layers, /root/reference README.md:92-128).
 * A matrix [N][K] is stored as one blob: N fp32 per-output-channel scales, padded to 256 bytes, then N*K e4m3 bytes
 * (OCP FP8 E4M3, finite-only "fn" variant) row-major, the whole blob padded to 256 bytes:  W[n][k] = scale[n] * e4m3[n][k].
 * audio_head_t is n_codebooks-1 such blobs of [audio_vocab][d_d] back to back.  csmb_e4m3_blob_bytes gives the size.
 * Served by the batch-1 frame kernel (csmb_frame_b1*: a second instantiation of the persistent kernel streams the blobs at
 * one byte per weight) and by the row-based entry points (csmb_backbone_forward, csmb_depth_decode, csmb_decode_frame: GEMV
 * kernels); both widen e4m3 -> fp32 in registers and apply the scale to the finished dot product.  The tensor-core paths
 * (csmb_decode_frame_fast*, csmb_prefill_fast, csmb_build_proj_table) return CSMB_ERR_UNSUPPORTED for such a model. */
#define CSMB_WEIGHTS_BF16 0
#define CSMB_WEIGHTS_E4M3 1
size_t csmb_e4m3_blob_bytes(int N, int K);
/* csmb_linear on an e4m3 blob: y[R][N] (+)= scale[n] * sum_k x[r][k] * e4m3[n][k]   (K % 16 == 0, blob 16-byte aligned) */
int csmb_linear_e4m3(const float* x, int ldx, const void* blob, float* y, int ldy, int R, int N, int K, int accumulate,
                     int device, void* stream);

/* Per-call view of a batch of sequences being generated. */
typedef struct {
  int batch;                    /* sequences */
  int max_pages;                /* block_table row length */
  float* kv_pool;               /* backbone: [n_layers][n_pages_total] pages (see top of file) */
  size_t kv_layer_stride;       /* floats between consecutive layers in kv_pool */
  const int32_t* block_table;   /* [batch][max_pages] */
  float* dec_kv_pool;           /* decoder: [n_layers][batch*ceil(n_codebooks/PAGE)] pages */
  size_t dec_kv_layer_stride;
  void* workspace;              /* csmb_lm_workspace_bytes() bytes */
  size_t workspace_bytes;
  int flags;                    /* CSMB_BATCH_* */
} csmb_batch;

/* csmb_batch.flags bit: the row-based entry points (csmb_backbone_forward, csmb_depth_decode, csmb_decode_frame) run every
 * Linear on the tensor-core path whatever the row count (instead of GEMV kernels up to 8 rows), so that a row's result
 * is independent of the rows it is batched with.  A serving loop that admits prompts into a running batch sets it. */
#define CSMB_BATCH_ROW_INVARIANT 1

/* bytes of scratch for up to max_rows rows (prefill) / batch sequences (decode). */
size_t csmb_lm_workspace_bytes(const csmb_model* m /*host*/, int max_rows);

/* model.backbone(...) over R rows, then codebook0_head on the rows listed in last_rows
 * (generation.py:34-42).  Writes h_last [n_last][d_b] and c0_logits [n_last][audio_vocab]. */
int csmb_backbone_forward(const csmb_model* m, const csmb_batch* b, const int32_t* tokens,
                          const uint8_t* mask, const int32_t* row_seq, const int32_t* row_pos, int R,
                          const int32_t* last_rows, int n_last, float* h_last, float* c0_logits, int device,
                          void* stream);

/* The 31-step depth-decoder loop of generate_frame (generation.py:56-90) for `batch` sequences:
 * given h_last and the already-sampled c0 (frame[:,0]), fills frame[:,i] for i in [step_begin, step_end)
 * (1 <= step_begin <= step_end <= n_codebooks; the whole loop is [1, n_codebooks)); all sampling on device.
 * Step i reads its input token from frame[:,i-1] (or forced[:,i-1]), so a host that wants to post-process
 * samples can run one step per call and overwrite frame[:,i] in between.
 * logits_out (optional) [batch][n_codebooks][audio_vocab] receives every step's logits (slot 0 unused).
 * forced (optional) [batch][n_codebooks]: teacher forcing — propagate these tokens instead of the samples.
 * RNG draw index of codebook i for sequence b: draw_base + i + (pos ? pos[b] : 0) * n_codebooks. */
int csmb_depth_decode(const csmb_model* m, const csmb_batch* b, const float* h_last, int32_t* frame,
                      const csmb_sampler* sampler, uint64_t draw_base, const int32_t* pos, float* logits_out,
                      const int32_t* forced, int step_begin, int step_end, int device, void* stream);

/* One whole decode frame for batch sequences whose previous frame is prev_frame [batch][n_codebooks]
 * at positions pos[batch] (generation.py:21-92 with T=1, plus :156-161 for the input construction):
 * embed -> backbone -> c0 -> depth loop.  Writes frame [batch][n_codebooks]. */
int csmb_decode_frame(const csmb_model* m, const csmb_batch* b, const int32_t* prev_frame,
                      const int32_t* pos, int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base,
                      int device, void* stream);

/* Per-call switches of the fused chain; caller-owned, read during the call only; NULL = all defaults. */
typedef struct {
  int no_pdl;               /* 1 = plain stream order instead of programmatic dependent launch (A/B timing) */
  int flags;                /* debug / A-B switches, 0 = the shipped chain.  Timing experiments whose results are WRONG: 1 = Linears
                               skip their partial stores, 2 = and their TMEM loads, 2048 = token planes not read (zero fill).
                               Same tokens, other kernels / hints: 4 = SwiGLU as a separate launch, 8 = the round-1 attention
                               kernel (dependent L2 round trips) instead of the shared-memory staged ones, 16 = partial sums
                               with plain loads instead of cp.async staging, 32 = default L2 policy for the weight stream,
                               64 = also L2-prefetch a Linear CTA's own weight blocks beyond its ring, 128 = no L2 prefetch of
                               the next big Linear's weights,
                               256 = at most 8 pipeline stages, bits 12..14 = k: the depth decoder's Linears ask L2 to keep
                               k/8 of their weight lines (evict-last) across the 31 depth steps (measured slower; default 0),
                               65536 = the gate|up Linears, 131072 = the other Linears with an even number of n-tiles run
                               as CTA pairs (2-CTA clusters, one tcgen05.mma.cta_group::2 of M = 256 per K step, each CTA loads
                               half of the token operand; bit-identical, measured no faster: profiles/r02_cta_pair.md) */
  int smem_kb;              /* shared-memory budget of a Linear CTA in KiB (48..200, 0 = 200): <= 100 lets two Linear CTAs
                               (of this or of another stream's chain) share an SM */
  const float* proj_table;  /* optional DEVICE table of csmb_build_proj_table: depth steps >= 2 read projection(embedding)
                               rows from it instead of running the projection Linear (same fp32 values, 60 launches less) */
} csmb_chain_opts;

/* Throughput path: the same frame as csmb_decode_frame for `batch` sequences in lock-step, as a fused kernel chain
 * (csrc/batch_frame.cu): every nn.Linear of the frame (attention.py:216-218,253; mlx_lm MLP; generation.py:42,75,79)
 * is one tcgen05/TMEM/TMA launch that reads bf16 hi+lo activation planes written by its producer kernel and leaves
 * fp32 split-K partials; everything between two Linears (partial reduction + residual + RMSNorm, RoPE + KV append +
 * attention, SwiGLU, sampling + next embedding: generation.py:51-56,81-89, attention.py:226-249) is one kernel.
 * Launches are chained with programmatic dependent launch, so weight streaming continues across kernel boundaries.
 * Same argument meaning as csmb_decode_frame (b->workspace is not used); workspace = csmb_decode_frame_fast_workspace_bytes
 * bytes, 256-byte aligned, zeroed once by its owner (first int = sticky error flag of the bounded waits).
 * Fused samplers: greedy, or temperature with optional top-k, top-p and / or min-p; min-p with min_keep > 1 (or an unsupported
 * model shape) returns CSMB_ERR_UNSUPPORTED — use csmb_decode_frame.  csmb_decode_frame_fast_supported returns 1/0 up front. */
size_t csmb_decode_frame_fast_workspace_bytes(const csmb_model* m /*host*/, int batch);
int csmb_decode_frame_fast_supported(const csmb_model* m /*host*/, const csmb_sampler* sampler /*host*/);
int csmb_decode_frame_fast(const csmb_model* m, const csmb_batch* b, const int32_t* prev_frame, const int32_t* pos,
                           int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, void* workspace,
                           size_t workspace_bytes, int device, void* stream);
/* csmb_decode_frame_fast with admission: for sequences b with use_override[b] != 0 (DEVICE uint8 [batch]) the backbone input
 * row is x_override[b] (DEVICE fp32 [batch][d_backbone]) — the already embedded LAST row of a prompt whose earlier rows
 * were run through csmb_backbone_forward — instead of the embedding of prev_frame[b].  This is how a serving loop admits a
 * new request into a free slot in the same step in which the running sequences decode (generation.py:34-42 for T > 1 is
 * split into rows 0..T-2 and the causal last row).  x_override / use_override both null and opts null =
 * csmb_decode_frame_fast. */
int csmb_decode_frame_fast_admit(const csmb_model* m, const csmb_batch* b, const int32_t* prev_frame, const int32_t* pos,
                                 int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, const float* x_override,
                                 const uint8_t* use_override, const csmb_chain_opts* opts /*host*/, void* workspace,
                                 size_t workspace_bytes, int device, void* stream);

/* csmb_backbone_forward on the chain's kernels: the prompt rows of any number of sequences (generation.py:34-42 with T > 1)
 * with one tcgen05 launch per Linear, RMSNorm / residual / SwiGLU fused as in csmb_decode_frame_fast, and a per-row attention
 * kernel: 8 launches per layer instead of 14.  Same arguments as csmb_backbone_forward (b->workspace is not used; last_rows /
 * h_last / c0_logits may be null with n_last = 0: a serving loop that only fills the KV cache); workspace =
 * csmb_prefill_fast_workspace_bytes(m, R) bytes, 256-byte aligned, first int zeroed once by its owner (sticky error flag).
 * A row's result does not depend on the other rows of the call. */
size_t csmb_prefill_fast_workspace_bytes(const csmb_model* m /*host*/, int rows);
int csmb_prefill_fast(const csmb_model* m, const csmb_batch* b, const int32_t* tokens, const uint8_t* mask,
                      const int32_t* row_seq, const int32_t* row_pos, int R, const int32_t* last_rows, int n_last,
                      float* h_last, float* c0_logits, void* workspace, size_t workspace_bytes, int device, void* stream);

/* Projected-embedding table for csmb_chain_opts.proj_table: table[cb][token][:] = projection . embed_audio(cb, token)
 * (generation.py:75 applied to models.py:79-80 rows) for every codebook and token, [n_codebooks][audio_vocab][d_decoder]
 * fp32, computed with the chain's own projection Linear (same tcgen05 tiles, same split-K, same summation order), so
 * reading a row is bit-identical to running that Linear.  csmb_proj_table_bytes = size of the table;
 * workspace = csmb_proj_table_workspace_bytes bytes, 256-byte aligned (scratch, free afterwards). */
size_t csmb_proj_table_bytes(const csmb_model* m /*host*/);
size_t csmb_proj_table_workspace_bytes(const csmb_model* m /*host*/);
int csmb_build_proj_table(const csmb_model* m, float* table, void* workspace, size_t workspace_bytes, int device,
                          void* stream);

/* Batch-1 latency path: ONE persistent cooperative kernel per frame (csrc/frame_kernel.cu) doing what
 * csmb_decode_frame does for a single sequence — generate_frame with T=1 (generation.py:21-92) plus the input
 * construction of :156-161 — with a producer warp per CTA streaming every weight matrix exactly once through a
 * shared-memory ring; the ~640 dependent GEMV phases synchronise through tagged activation words, not barriers.
 * workspace must be zero-initialised once (csmb_frame_workspace_bytes; ~40 MB on B200: one private decoder-KV
 * copy per SM) and belongs to one sequence.
 * block_table: this sequence's row of the paged-KV table; pos: DEVICE int, position of this frame's backbone row.
 * Fused samplers: greedy, or temperature with optional top-k, top-p and / or min-p; min-p with min_keep > 1 (and model shapes
 * other than csm_1b) returns CSMB_ERR_UNSUPPORTED — use csmb_decode_frame. */
size_t csmb_frame_workspace_bytes(const csmb_model* m /*host*/, int device);
/* Per-call switches of the frame kernel; caller-owned, read during the call only; NULL = all defaults. */
typedef struct {
  int ctas;                  /* CTAs of the launch; 0 = automatic: the largest count <= SMs that splits every weight matrix
                                into equal row slices (128 for csm_1b on a 148-SM B200).  The SMs left over let the codec's
                                streaming step of the previous frame run beside the frame kernel on a second stream
                                (generation.py:251 moved off the critical path). */
  int flags;                 /* debug: bit 0 = skip the GEMV arithmetic (timing experiments only; results are wrong); bits 4..6 =
                                k (e4m3 models): the decoder's ring copies ask L2 to keep k/8 of their lines (A/B, same tokens) */
  int prefetch_stages;       /* L2 prefetch distance in 16 KiB stages per SM (0 = off) */
  int prefetch_interval;     /* SM cycles between prefetches (0 = 700) */
  unsigned long long* prof;  /* debug: DEVICE buffer [n_sms][16] u64 filled with per-CTA phase timers; null = off */
} csmb_frame_opts;
/* status (optional, DEVICE int, zeroed once by its owner): receives the abort code of the FIRST internal wait that ever timed
 * out on this state (all waits are bounded) and is never cleared by the library, so the owner may look at it at any later
 * point — e.g. once per utterance — without losing an earlier frame's failure. */
int csmb_frame_b1(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table,
                  const int32_t* prev_frame, const int32_t* pos, int32_t* frame, const csmb_sampler* sampler,
                  uint64_t draw_base, const csmb_frame_opts* opts /*host*/, void* workspace, size_t workspace_bytes,
                  int32_t* status, int device, void* stream);
/* csmb_frame_b1 for ONE sequence of a batched state (serving with a single active slot): the pointers are that sequence's
 * rows (block_table + slot * max_pages, prev_frame + slot * n_codebooks, pos + slot, frame + slot * n_codebooks) and
 * seq_index is its row in the batch, so that sampling draws the same Philox noise as csmb_decode_frame(_fast) would for
 * that row (csmb_sample's r).  Everything else as csmb_frame_b1 (which is seq_index 0). */
int csmb_frame_b1_slot(const csmb_model* m, float* kv_pool, size_t kv_layer_stride, const int32_t* block_table_row,
                       const int32_t* prev_frame_row, const int32_t* pos, int32_t* frame_row, const csmb_sampler* sampler,
                       uint64_t draw_base, int seq_index, const csmb_frame_opts* opts /*host*/, void* workspace,
                       size_t workspace_bytes, int32_t* status, int device, void* stream);
/* The same kernel without its backbone part, for the first frame after a prompt (generation.py:139-146 with the
 * whole prompt as `tokens`): h_last = csmb_backbone_forward's normalised last hidden row [d_backbone] of this sequence,
 * pos = DEVICE int holding the sequence length (the sampled row is pos-1, which also indexes the random draws exactly
 * like csmb_sample + csmb_depth_decode do).  Codebook-0 head, sampling and the 31 depth steps run in one launch. */
int csmb_frame_b1_depth(const csmb_model* m, const float* h_last, const int32_t* pos, int32_t* frame,
                        const csmb_sampler* sampler, uint64_t draw_base, const csmb_frame_opts* opts /*host*/, void* workspace,
                        size_t workspace_bytes, int32_t* status, int device, void* stream);

/* ---------------------------------------------------------------- Mimi codec ------------------- */
/* moshi_mlx Mimi.encode / decode / decode_step (csm_mlx/tokenizers.py:14-21,70,150; generation.py:224-225,
 * 251,258).  Activations are time-major fp32 [batch][time][channels]; see csrc/mimi.cu for how Conv1d and
 * ConvTranspose1d map onto the strided-row GEMM. */

/* Y[b][t][n] = epi( sum_{j<K} actA(A[b*a_batch + t*lda + j]) * W[n][j] ), epi = (+bias[n]) -> act_out ->
 * (*scale[n]) -> (+residual[b*r_batch + t*ldr + n]).  act_in: 0 none, 1 ELU.  act_out: 0 none, 1 GELU(erf). */
int csmb_gemm_f32(const float* A, long long a_batch, int lda, const float* W, float* Y, long long y_batch, int ldy,
                  const float* bias, const float* scale, const float* residual, long long r_batch, int ldr, int B,
                  int T, int N, int K, int act_in, int act_out, int device, void* stream);

int csmb_layernorm(const float* x, long long x_batch, const float* w, const float* b, float* y, int B, int T, int d,
                   float eps, int device, void* stream);

/* qkv [B][T][3][H][64]: RoPE (adjacent pairs, angle = pos*freqs[i]) on q,k in place; rotated k and v written to
 * the ring cache [B][cap][2][H][64] at slot pos%cap; out [B][T][H*64] = causal attention over the last `ctx`
 * positions.  *pos0 (device int) = absolute position of step 0 of this call; cap >= ctx + T - 1. */
int csmb_mimi_attention(float* qkv, float* cache, const float* freqs, const int* pos0, float* out, int B, int T,
                        int H, int cap, int ctx, int device, void* stream);

/* RVQ dequantise gather: codes [B][K][F] -> sem [B][F][D] = C_0[c_0], ac [B][F][D] = sum_{k>=1} C_k[c_k];
 * codebooks [K][bins][D].  Ids outside [0,bins) are clamped. */
int csmb_rvq_gather(const int32_t* codes, const float* codebooks, float* sem, float* ac, int B, int K, int F,
                    int bins, int D, int device, void* stream);

/* depthwise ConvTranspose1d(k=4,s=2) x2 upsampler: x [B][T][C], xprev [B][C] (row before t=0), w [C][4] ->
 * y [B][2T][C]. */
int csmb_upsample_dw(const float* x, const float* xprev, const float* w, float* y, int B, int T, int C, int device,
                     void* stream);

/* One residual-VQ encode step: idx = argmin_j (c2[j] - 2*dots[m][j]); codes[b][k][f] = idx; r[m] -= C[idx]. */
int csmb_rvq_argmin_update(const float* dots, const float* c2, const float* codebook, float* r, int32_t* codes,
                           int M, int bins, int D, int K, int k, int F, int device, void* stream);

/* dst[b][dst_t0+t][:] = src[b][src_t0 + (replicate ? 0 : t)][:] for t < T (padding / context fill). */
int csmb_copy_rows(const float* src, long long s_batch, int src_t0, float* dst, long long d_batch, int dst_t0, int B,
                   int T, int C, int replicate, int device, void* stream);
/* streaming context carry: buf[b][0:pad] = buf[b][T:T+pad]. */
int csmb_shift_rows(float* buf, long long batch, int B, int T, int pad, int C, int device, void* stream);
/* *p += v on the device (position counters of captured graphs). */
int csmb_add_int(int* p, int v, int device, void* stream);

/* ---------------------------------------------------------------- Mimi codec on the tensor cores ---------- */
/* Batch-scale codec path (csrc/mimi_tc.cu): Mimi.encode / Mimi.decode of many clips (csm_mlx/tokenizers.py:61-85, 148-150;
 * BASELINE.json configs[4]).  Activations and weights travel as two bf16 planes, x = hi + lo up to 2^-17, and
 *
 *   Y[b][t][n] = epi( sum_{j < taps} sum_{c < C} A[b][t + j][c] * W[n][j*C + c] ),   A = a_hi + a_lo, W = w_hi + w_lo
 *
 * runs as ONE persistent tcgen05 kernel (three MMAs per K step: Ahi.Whi + Alo.Whi + Ahi.Wlo, fp32 TMEM accumulators, TMA
 * 128B-swizzled boxes, two accumulators so the epilogue overlaps the next tile).  Conv1d(k, stride 1): taps = k, C = Cin on
 * the left-padded input; Conv1d(k = 2s, stride s): the input viewed as rows of s*Cin values, taps = 2; ConvTranspose1d
 * (k = 2s, stride s): taps = 2 over rows (t-1, t), N = s*Cout phase-major = the time-major output; Linear: taps = 1.
 * epi = (+bias[n]) -> GELU(act_out = 1) -> (*scale[n]) -> (+residual[b][t][n]); the result goes to y32 (fp32) and / or, after
 * an optional ELU (plane_act = 1: the activation the CONSUMING conv applies), to the bf16 planes y_hi / y_lo.
 * a_*: [a_rows][lda] bf16 (row b*rpb + t + j of the plane is A[b][t + j]); w_*: [w_rows >= N rounded up to 16][ldw], rows
 * beyond N zero.  taps > 1 needs C % 64 == 0; lda, ldw % 8 == 0; all plane bases 16-byte aligned.
 * err_flag: DEVICE int zeroed once by its owner; set (sticky) if a bounded wait inside the kernel timed out. */
typedef struct {
  const uint16_t *a_hi, *a_lo;
  long long a_rows;
  int lda, rpb, C, taps;
  const uint16_t *w_hi, *w_lo;
  int w_rows, ldw;
  float* y32;
  long long y_batch;
  int ldy;
  uint16_t *y_hi, *y_lo;
  long long p_batch;
  int ldp, plane_act;
  const float *bias, *scale, *residual;
  long long r_batch;
  int ldr;
  int B, T, N, act_out;
  int* err_flag;
} csmb_tc3;
int csmb_gemm_tc3(const csmb_tc3* g /*host*/, int device, void* stream);

/* fp32 [B][T][C] (batch stride x_batch, row stride ldx) -> bf16 hi/lo planes (p_batch, ldp); act = 1 applies ELU first. */
int csmb_split_planes(const float* x, long long x_batch, int ldx, uint16_t* hi, uint16_t* lo, long long p_batch, int ldp,
                      int B, int T, int C, int act, int device, void* stream);
/* csmb_layernorm writing dense [B*T][d] bf16 hi/lo planes (the operand of the next tensor-core Linear). */
int csmb_layernorm_planes(const float* x, long long x_batch, const float* w, const float* b, uint16_t* hi, uint16_t* lo, int B,
                          int T, int d, float eps, int device, void* stream);
/* Whole-clip attention of the codec transformers (moshi StreamingMultiheadAttention with context 250, RoPE base 10 000 on
 * adjacent pairs): qkv [B][T][3][H][64] fp32 from the in-projection; q and k are rotated IN PLACE (position = row index), then
 * every position attends causally to the last `ctx` positions; out_hi / out_lo [B][T][H*64] are the bf16 planes the
 * out-projection reads.  One block per (32 queries, head, clip) with the shared keys / values staged in shared memory. */
int csmb_mimi_attention_planes(float* qkv, const float* freqs, uint16_t* out_hi, uint16_t* out_lo, int B, int T, int H, int ctx,
                               int device, void* stream);
/* First SEANet encoder layer, Conv1d(1 -> C, k) + bias: x [B][(k-1) + N] fp32 (left-padded, batch stride x_batch) ->
 * y fp32 and ELU(y) planes, both [B][..][C] with batch stride y_batch (pointers already at the first output row). */
int csmb_conv_in_planes(const float* x, long long x_batch, const float* w, const float* bias, float* y, uint16_t* hi,
                        uint16_t* lo, long long y_batch, int B, int N, int C, int k, int device, void* stream);
/* csmb_rvq_argmin_update that also writes the updated residual as planes (operand of the next codebook's search). */
int csmb_rvq_argmin_update_planes(const float* dots, const float* c2, const float* codebook, float* r, uint16_t* r_hi,
                                  uint16_t* r_lo, int32_t* codes, int M, int bins, int D, int K, int k, int F, int device,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSM_B200_H */
