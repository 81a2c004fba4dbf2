// Host-side orchestration of the row-based LM path: backbone forward over R rows, the 31-step depth
// loop and a whole decode frame, as sequences of the kernels in ops.cu on one stream.  No host sync, no
// allocation: capturable into a CUDA graph by the caller.
//
// Restates (paths relative to /root/reference): csm_mlx/generation.py:21-92 (generate_frame),
// :156-161 (next-input construction); mlx_lm LlamaModel block structure (models.py:50-51,70-77).
#include "ops.cuh"

namespace csmb {

struct Workspace {
  float *x, *xn, *qkv, *attn, *gu, *act;  // [rows][...]
  float *dx;                               // decoder residual stream [rows][d_d]
  float *din;                              // decoder input embeddings [rows][d_b]
  float *hn;                               // normalised last hidden [batch][max(d_b,d_d)]
  float *logits;                           // [batch][V]
  float *h_last;                           // [batch][d_b]
  int32_t *tokens;                         // [batch][ncb+1]
  uint8_t *mask;                           // [batch][ncb+1]
  int32_t *iota;                           // [rows] 0..rows-1
  int32_t *dec_seq, *dec_pos;              // [rows] decoder row maps (rebuilt per step)
  int32_t *dec_bt;                         // [batch][dec_pages]
  int32_t *c0_tmp;                         // [batch]
  void* tc;                                // scratch of the tensor-core linear (bf16 hi/lo split, split-K partials)
  size_t tc_bytes;
  size_t bytes;
  bool row_invariant;                      // CSMB_BATCH_ROW_INVARIANT: every Linear on the tensor-core path, whatever R
  bool e4m3;                               // CSMB_WEIGHTS_E4M3: the Linear matrices are weight-only FP8 blobs (GEMV kernels)
};

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
// From this many rows on, linears run on the tensor cores (gemm_tc.cu); <= 8: GEMV kernels.  A batch flagged
// CSMB_BATCH_ROW_INVARIANT takes the tensor-core path for ANY row count, so that what a row computes never depends on
// how many other rows (other sequences' prompts, other requests of a serving step) share the call.
constexpr int TC_MIN_ROWS = 9;

struct Workspace;
static int lin(const Workspace& w, const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
               int accumulate, cudaStream_t st);

static Workspace carve(const csmb_model& m, int max_rows, void* base) {
  const csmb_llama &b = m.backbone, &d = m.decoder;
  const int qkv_b = (b.n_heads + 2 * b.n_kv_heads) * b.head_dim, qkv_d = (d.n_heads + 2 * d.n_kv_heads) * d.head_dim;
  const size_t dmax = (size_t)(b.d_model > d.d_model ? b.d_model : d.d_model);
  const size_t qkvmax = (size_t)(qkv_b > qkv_d ? qkv_b : qkv_d);
  const size_t ffmax = (size_t)(b.d_ff > d.d_ff ? b.d_ff : d.d_ff);
  const size_t attnmax = (size_t)((b.n_heads * b.head_dim > d.n_heads * d.head_dim) ? b.n_heads * b.head_dim
                                                                                   : d.n_heads * d.head_dim);
  const size_t R = (size_t)max_rows;
  const int dec_pages = cdiv(m.n_codebooks, CSMB_PAGE);
  size_t off = 0;
  char* p = static_cast<char*>(base);
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align256(bytes);
    return r;
  };
  Workspace w;
  w.x = (float*)take(R * dmax * 4);
  w.xn = (float*)take(R * dmax * 4);
  w.qkv = (float*)take(R * qkvmax * 4);
  w.attn = (float*)take(R * attnmax * 4);
  w.gu = (float*)take(R * 2 * ffmax * 4);
  w.act = (float*)take(R * ffmax * 4);
  w.dx = (float*)take(R * (size_t)d.d_model * 4);
  w.din = (float*)take(R * (size_t)b.d_model * 4);
  w.hn = (float*)take(R * dmax * 4);
  w.logits = (float*)take(R * (size_t)m.audio_vocab * 4);
  w.h_last = (float*)take(R * (size_t)b.d_model * 4);
  w.tokens = (int32_t*)take(R * (size_t)(m.n_codebooks + 1) * 4);
  w.mask = (uint8_t*)take(R * (size_t)(m.n_codebooks + 1));
  w.iota = (int32_t*)take(R * 4);
  w.dec_seq = (int32_t*)take(R * 4);
  w.dec_pos = (int32_t*)take(R * 4);
  w.dec_bt = (int32_t*)take(R * (size_t)dec_pages * 4);
  w.c0_tmp = (int32_t*)take(R * 4);
  // tensor-core linear scratch: the largest requirement over this model's (N, K) shapes at max_rows rows
  size_t tcb = 0;
  {
    const int shapes[][2] = {{qkv_b, b.d_model}, {b.d_model, b.n_heads * b.head_dim}, {2 * b.d_ff, b.d_model}, {b.d_model, b.d_ff},
                             {qkv_d, d.d_model}, {d.d_model, d.n_heads * d.head_dim}, {2 * d.d_ff, d.d_model}, {d.d_model, d.d_ff},
                             {m.audio_vocab, b.d_model}, {m.audio_vocab, d.d_model}, {d.d_model, b.d_model}};
    for (auto& sh : shapes)
      if (sh[1] % 64 == 0) {
        const size_t need = linear_tc_workspace_bytes(max_rows, sh[0], sh[1]);
        tcb = need > tcb ? need : tcb;
      }
  }
  w.tc = take(tcb);
  w.tc_bytes = tcb;
  w.bytes = off;
  w.row_invariant = false;
  w.e4m3 = m.weight_format == CSMB_WEIGHTS_E4M3;
  return w;
}

static int lin(const Workspace& w, const float* x, int ldx, const uint16_t* W, float* y, int ldy, int R, int N, int K,
               int accumulate, cudaStream_t st) {
  // weight-only FP8: e4m3 rows widened in registers (a row's sums do not depend on the other rows of the call either)
  if (w.e4m3) return launch_linear_e4m3(x, ldx, W, y, ldy, R, N, K, accumulate, st);
  if ((R >= TC_MIN_ROWS || w.row_invariant) && K % 64 == 0 && w.tc != nullptr && w.tc_bytes >= linear_tc_workspace_bytes(R, N, K))
    return launch_linear_tc(x, ldx, W, y, ldy, R, N, K, accumulate, w.tc, w.tc_bytes, st);
  return launch_linear(x, ldx, W, y, ldy, R, N, K, accumulate, st);
}

// ---- small index kernels ------------------------------------------------------------------------
__global__ void k_iota(int32_t* a, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
// decoder row maps: first step has 2 rows per sequence (positions 0,1), later steps 1 row at `pos`.
__global__ void k_dec_rows(int32_t* seq, int32_t* posv, int32_t* bt, int batch, int rows_per_seq, int pos0,
                           int dec_pages) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < batch * rows_per_seq) {
    seq[i] = i / rows_per_seq;
    posv[i] = pos0 + i % rows_per_seq;
  }
  if (i < batch * dec_pages) bt[i] = i;
}
// next backbone input from the previous frame: tokens = [frame, 0], mask = [1.., 0]  (generation.py:156-161)
__global__ void k_frame_to_input(const int32_t* frame, int32_t* tokens, uint8_t* mask, int batch, int ncb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * (ncb + 1)) return;
  int b = i / (ncb + 1), c = i % (ncb + 1);
  tokens[i] = c < ncb ? frame[b * ncb + c] : 0;
  mask[i] = c < ncb ? 1 : 0;
}
// din[(b*2+0)] = h_last[b]
__global__ void k_copy_rows(const float* src, int lds, float* dst, int ldd, int d, int R) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * d) return;
  size_t r = i / d, c = i % d;
  dst[r * ldd + c] = src[r * lds + c];
}

// ---- one Llama stack over R rows ----------------------------------------------------------------
static int llama_layers(const csmb_llama& L, float* x, const Workspace& w, float* kv_pool, size_t kv_layer_stride,
                        const int32_t* block_table, int max_pages, const int32_t* row_seq, const int32_t* row_pos,
                        int R, int max_pos, cudaStream_t st) {
  const int d = L.d_model, H = L.n_heads, Hkv = L.n_kv_heads, hd = L.head_dim, F = L.d_ff;
  const int nqkv = (H + 2 * Hkv) * hd;
  int rc;
  for (int l = 0; l < L.n_layers; ++l) {
    float* pool = kv_pool + (size_t)l * kv_layer_stride;
    if ((rc = launch_rmsnorm(x, d, L.norm_in[l], w.xn, d, R, d, L.eps, nullptr, st))) return rc;
    if ((rc = lin(w, w.xn, d, L.wqkv[l], w.qkv, nqkv, R, nqkv, d, 0, st))) return rc;
    if ((rc = launch_rope_kv_append(w.qkv, L.rope, pool, block_table, max_pages, row_seq, row_pos, R, H, Hkv, hd, st)))
      return rc;
    if ((rc = launch_attention(w.qkv, nqkv, pool, block_table, max_pages, row_seq, row_pos, w.attn, R, H, Hkv, hd,
                               max_pos, st)))
      return rc;
    if ((rc = lin(w, w.attn, H * hd, L.wo[l], x, d, R, d, H * hd, 1, st))) return rc;
    if ((rc = launch_rmsnorm(x, d, L.norm_post[l], w.xn, d, R, d, L.eps, nullptr, st))) return rc;
    if ((rc = lin(w, w.xn, d, L.wgu[l], w.gu, 2 * F, R, 2 * F, d, 0, st))) return rc;
    if ((rc = launch_swiglu(w.gu, w.act, R, F, st))) return rc;
    if ((rc = lin(w, w.act, F, L.wdown[l], x, d, R, d, F, 1, st))) return rc;
  }
  return CSMB_OK;
}

static int check_model(const csmb_model* m) {
  CSMB_REQUIRE(m != nullptr);
  CSMB_REQUIRE(m->backbone.n_layers > 0 && m->backbone.n_layers <= CSMB_MAX_LAYERS);
  CSMB_REQUIRE(m->decoder.n_layers > 0 && m->decoder.n_layers <= CSMB_MAX_LAYERS);
  CSMB_REQUIRE(m->n_codebooks >= 2 && m->n_codebooks <= CSMB_MAX_CODEBOOKS);
  CSMB_REQUIRE(m->backbone.n_heads * m->backbone.head_dim == m->backbone.d_model);
  CSMB_REQUIRE(m->decoder.n_heads * m->decoder.head_dim == m->decoder.d_model);
  CSMB_REQUIRE(m->weight_format == CSMB_WEIGHTS_BF16 || m->weight_format == CSMB_WEIGHTS_E4M3);
  if (m->weight_format == CSMB_WEIGHTS_E4M3)
    CSMB_REQUIRE(m->backbone.d_model % 16 == 0 && m->decoder.d_model % 16 == 0 && m->backbone.d_ff % 16 == 0 &&
                 m->decoder.d_ff % 16 == 0);
  return CSMB_OK;
}

static int backbone_forward(const csmb_model& m, const csmb_batch& b, const Workspace& w, const int32_t* tokens,
                            const uint8_t* mask, const int32_t* row_seq, const int32_t* row_pos, int R,
                            const int32_t* last_rows, int n_last, float* h_last, float* c0_logits,
                            cudaStream_t st) {
  const csmb_llama& L = m.backbone;
  int rc;
  if ((rc = launch_embed_sum(tokens, mask, m.text_emb, m.audio_emb, w.x, R, L.d_model, m.n_codebooks,
                             m.audio_vocab, st)))
    return rc;
  if ((rc = llama_layers(L, w.x, w, b.kv_pool, b.kv_layer_stride, b.block_table, b.max_pages, row_seq, row_pos, R,
                         m.max_pos, st)))
    return rc;
  if ((rc = launch_rmsnorm(w.x, L.d_model, L.norm_final, h_last, L.d_model, n_last, L.d_model, L.eps, last_rows, st)))
    return rc;
  if (c0_logits)
    if ((rc = lin(w, h_last, L.d_model, m.c0_head, c0_logits, m.audio_vocab, n_last, m.audio_vocab,
                            L.d_model, 0, st)))
      return rc;
  return CSMB_OK;
}

static int depth_decode(const csmb_model& m, const csmb_batch& b, const Workspace& w, const float* h_last,
                        int32_t* frame, const csmb_sampler& sampler, uint64_t draw_base, const int32_t* pos,
                        float* logits_out, const int32_t* forced, int step_begin, int step_end, cudaStream_t st) {
  const csmb_llama& L = m.decoder;
  const int B = b.batch, db = m.backbone.d_model, dd = L.d_model, V = m.audio_vocab, ncb = m.n_codebooks;
  const int dec_pages = cdiv(ncb, CSMB_PAGE);
  const int32_t* toks = forced ? forced : frame;  // tokens that are propagated (teacher forcing or own samples)
  int rc;
  for (int i = step_begin; i < step_end; ++i) {
    const int rps = (i == 1) ? 2 : 1;
    const int R = B * rps;
    if (i == 1) {
      // rows (b,0) = h_last[b], rows (b,1) = embed_audio(0, c0[b])   (generation.py:56-64)
      size_t tot = (size_t)B * db;
      k_copy_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h_last, db, w.din, 2 * db, db, B);
      CSMB_LAUNCH_CHECK();
      if ((rc = launch_embed_audio(toks, ncb, m.audio_emb, w.din + db, 2 * db, B, db, 0, V, st))) return rc;
    } else {
      // decoder_inputs = embed_audio(i-1, c_{i-1})   (generation.py:86-89)
      if ((rc = launch_embed_audio(toks + (i - 1), ncb, m.audio_emb, w.din, db, B, db, i - 1, V, st))) return rc;
    }
    k_dec_rows<<<cdiv(B * 2 > B * dec_pages ? B * 2 : B * dec_pages, 128), 128, 0, st>>>(
        w.dec_seq, w.dec_pos, w.dec_bt, B, rps, i == 1 ? 0 : i, dec_pages);
    CSMB_LAUNCH_CHECK();
    if ((rc = lin(w, w.din, db, m.projection, w.dx, dd, R, dd, db, 0, st))) return rc;
    if ((rc = llama_layers(L, w.dx, w, b.dec_kv_pool, b.dec_kv_layer_stride, w.dec_bt, dec_pages, w.dec_seq,
                           w.dec_pos, R, dec_pages * CSMB_PAGE, st)))
      return rc;
    // last position of every sequence -> final norm -> audio_head[i-1]   (generation.py:79)
    const float* last = (i == 1) ? w.dx + dd : w.dx;
    const int ldlast = (i == 1) ? 2 * dd : dd;
    if ((rc = launch_rmsnorm(last, ldlast, L.norm_final, w.hn, dd, B, dd, L.eps, nullptr, st))) return rc;
    float* lg = logits_out ? logits_out + (size_t)i * V : w.logits;
    const int ldl = logits_out ? ncb * V : V;
    const uint16_t* head = w.e4m3 ? reinterpret_cast<const uint16_t*>(reinterpret_cast<const char*>(m.audio_head_t) +
                                                                      (size_t)(i - 1) * e4m3_blob_bytes(V, dd))
                                  : m.audio_head_t + (size_t)(i - 1) * V * dd;
    if ((rc = lin(w, w.hn, dd, head, lg, ldl, B, V, dd, 0, st))) return rc;
    if ((rc = launch_sample(lg, ldl, frame + i, ncb, B, V, sampler, draw_base + (uint64_t)i, pos, (uint32_t)ncb,
                            forced ? forced + i : nullptr, ncb, st)))
      return rc;
  }
  return CSMB_OK;
}

}  // namespace csmb

using namespace csmb;

extern "C" {

size_t csmb_lm_workspace_bytes(const csmb_model* m, int max_rows) {
  if (!m || max_rows <= 0) return 0;
  return carve(*m, max_rows, nullptr).bytes;
}

int csmb_backbone_forward(const csmb_model* m, const csmb_batch* b, const int32_t* tokens, const uint8_t* mask,
                          const int32_t* row_seq, const int32_t* row_pos, int R, const int32_t* last_rows,
                          int n_last, float* h_last, float* c0_logits, int device, void* stream) {
  CSMB_ENTER(device);
  int rc = check_model(m);
  if (rc) return rc;
  CSMB_REQUIRE(b && b->workspace && R > 0 && n_last > 0 && n_last <= R);
  CSMB_REQUIRE(b->workspace_bytes >= carve(*m, R, nullptr).bytes);
  Workspace w = carve(*m, R, b->workspace);
  w.row_invariant = (b->flags & CSMB_BATCH_ROW_INVARIANT) != 0;
  return backbone_forward(*m, *b, w, tokens, mask, row_seq, row_pos, R, last_rows, n_last, h_last, c0_logits,
                          (cudaStream_t)stream);
}

int csmb_depth_decode(const csmb_model* m, const csmb_batch* b, const float* h_last, int32_t* frame,
                      const csmb_sampler* sampler, uint64_t draw_base, const int32_t* pos, float* logits_out,
                      const int32_t* forced, int step_begin, int step_end, int device, void* stream) {
  CSMB_ENTER(device);
  int rc = check_model(m);
  if (rc) return rc;
  CSMB_REQUIRE(b && b->workspace && b->batch > 0 && sampler);
  CSMB_REQUIRE(step_begin >= 1 && step_begin <= step_end && step_end <= m->n_codebooks);
  CSMB_REQUIRE(b->workspace_bytes >= carve(*m, 2 * b->batch, nullptr).bytes);
  Workspace w = carve(*m, 2 * b->batch, b->workspace);
  w.row_invariant = (b->flags & CSMB_BATCH_ROW_INVARIANT) != 0;
  return depth_decode(*m, *b, w, h_last, frame, *sampler, draw_base, pos, logits_out, forced, step_begin, step_end,
                      (cudaStream_t)stream);
}

int csmb_decode_frame(const csmb_model* m, const csmb_batch* b, const int32_t* prev_frame, const int32_t* pos,
                      int32_t* frame, const csmb_sampler* sampler, uint64_t draw_base, int device, void* stream) {
  CSMB_ENTER(device);
  int rc = check_model(m);
  if (rc) return rc;
  CSMB_REQUIRE(b && b->workspace && b->batch > 0 && sampler && prev_frame && pos && frame);
  const int B = b->batch, ncb = m->n_codebooks;
  CSMB_REQUIRE(b->workspace_bytes >= carve(*m, 2 * B, nullptr).bytes);
  Workspace w = carve(*m, 2 * B, b->workspace);
  w.row_invariant = (b->flags & CSMB_BATCH_ROW_INVARIANT) != 0;
  cudaStream_t st = (cudaStream_t)stream;
  k_frame_to_input<<<cdiv(B * (ncb + 1), 128), 128, 0, st>>>(prev_frame, w.tokens, w.mask, B, ncb);
  CSMB_LAUNCH_CHECK();
  k_iota<<<cdiv(B, 128), 128, 0, st>>>(w.iota, B);
  CSMB_LAUNCH_CHECK();
  if ((rc = backbone_forward(*m, *b, w, w.tokens, w.mask, w.iota, pos, B, w.iota, B, w.h_last, w.logits, st)))
    return rc;
  if ((rc = launch_sample(w.logits, m->audio_vocab, frame, ncb, B, m->audio_vocab, *sampler, draw_base, pos,
                          (uint32_t)ncb, nullptr, 0, st)))
    return rc;
  return depth_decode(*m, *b, w, w.h_last, frame, *sampler, draw_base, pos, nullptr, nullptr, 1, ncb, st);
}

}  // extern "C"
