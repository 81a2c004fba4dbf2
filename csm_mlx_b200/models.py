"""CSM model object: parameter tree, weight loading, device-side layout.

Mirrors ``/root/reference/csm_mlx/models.py:12-92`` (``ModelArgs``, ``csm_1b``, ``CSM``) and the inherited
``mlx.nn.Module.load_weights`` (README.md:40,140; run_streaming_csm_mlx.py:738).  The arithmetic of the
reference's ``LlamaModel`` / ``Attention`` modules lives in libcsm_b200.so; this class only owns the weights
(bf16 in HBM, fused per layer the way the kernels stream them) and the descriptor handed to the C ABI.

HBM layout (csm_1b, bf16): per backbone layer qkv [3072][2048], o [2048][2048], gate|up [16384][2048],
down [2048][8192]; per decoder layer qkv [1536][1024], o [1024][1024], gate|up [16384][1024],
down [1024][8192]; text_emb [128256][2048]; audio_emb [65632][2048]; projection [1024][2048];
c0_head [2051][2048]; audio_head stored TRANSPOSED as [31][2051][1024] so every head row is a contiguous
K-vector like any other Linear.  ≈3.1 GB total; norm weights and RoPE tables fp32.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple, Union

import torch

from . import _lib
from .attention import llama3_rope_table
from .config import BACKBONE_CONFIGURATION, DECODER_CONFIGURATION, MAX_SEQ_LEN, LlamaArgs


@dataclass
class ModelArgs:
    backbone_name: str
    decoder_name: str
    n_text_vocab: int
    n_audio_vocab: int
    n_audio_codebooks: int


def csm_1b() -> ModelArgs:
    """models.py:21-28."""
    return ModelArgs(backbone_name="1b", decoder_name="100m", n_text_vocab=128256, n_audio_vocab=2051,
                     n_audio_codebooks=32)


def csm_tiny() -> ModelArgs:
    """Structure-preserving miniature for fast tests (not a reference configuration)."""
    return ModelArgs(backbone_name="tiny", decoder_name="tiny", n_text_vocab=512, n_audio_vocab=67,
                     n_audio_codebooks=4)


class _Param:
    """Stand-in for an mlx module that only has a ``weight``."""

    def __init__(self) -> None:
        self.weight: Optional[torch.Tensor] = None


class _LlamaStack:
    """What callers read off ``model.backbone`` / ``model.decoder``: ``.layers`` and ``.args``
    (generation.py:70,127,132)."""

    def __init__(self, args: LlamaArgs):
        self.args = args
        self.layers: List[dict] = [dict() for _ in range(args.num_hidden_layers)]
        # device tensors, filled by CSM._finalize
        self.wqkv: List[torch.Tensor] = []
        self.wo: List[torch.Tensor] = []
        self.wgu: List[torch.Tensor] = []
        self.wdown: List[torch.Tensor] = []
        self.norm_in: List[torch.Tensor] = []
        self.norm_post: List[torch.Tensor] = []
        self.norm_final: Optional[torch.Tensor] = None
        self.rope: Optional[torch.Tensor] = None


def _expected_shapes(args: ModelArgs) -> Dict[str, Tuple[int, ...]]:
    from .random_init import csm_param_shapes

    return {n: tuple(s) for n, s, _ in csm_param_shapes(args.backbone_name, args.decoder_name, args.n_text_vocab,
                                                       args.n_audio_vocab, args.n_audio_codebooks)}


class CSM:
    def __init__(self, args: ModelArgs, device: Union[str, torch.device, None] = None):
        self.args = args
        self.n_text_vocab = args.n_text_vocab
        self.n_audio_vocab = args.n_audio_vocab
        self.n_audio_codebooks = args.n_audio_codebooks
        b, d = BACKBONE_CONFIGURATION[args.backbone_name], DECODER_CONFIGURATION[args.decoder_name]
        self.n_backbone_embedding = b.num_attention_heads * b.head_dim
        self.n_decoder_embedding = d.num_attention_heads * d.head_dim
        self.backbone = _LlamaStack(b)
        self.decoder = _LlamaStack(d)
        self.text_embeddings = _Param()
        self.audio_embeddings = _Param()
        self.projection = _Param()
        self.codebook0_head = _Param()
        self._audio_head_t: Optional[torch.Tensor] = None  # [ncb-1][V][d_d]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else torch.device("cuda")
        self._desc: Optional[_lib.Model] = None
        self._proj_table: Optional[torch.Tensor] = None
        self.weights_version = 0   # bumped whenever the weights change: caches of derived values (KV prefixes) key on it
        self._loaded = False
        self.quantized = False   # quantize_weights(): the Linear matrices are weight-only FP8 blobs

    # ------------------------------------------------------------------ reference-visible views
    @property
    def audio_head(self) -> Optional[torch.Tensor]:
        """(n_codebooks-1, d_decoder, n_audio_vocab) view, the reference's layout (models.py:65-67)."""
        return None if self._audio_head_t is None else self._audio_head_t.transpose(1, 2)

    def parameters(self) -> Dict[str, torch.Tensor]:
        """Flat dict keyed like the reference parameter tree; views into the fused storage (a quantised
        model: the dequantised fp32 matrices, i.e. exactly the numbers its kernels multiply with)."""
        self._require_loaded()
        if self.quantized:
            return self._dequantized_parameters()
        out = {
            "text_embeddings.weight": self.text_embeddings.weight,
            "audio_embeddings.weight": self.audio_embeddings.weight,
            "projection.weight": self.projection.weight,
            "codebook0_head.weight": self.codebook0_head.weight,
            "audio_head": self.audio_head,
        }
        for name, st in (("backbone", self.backbone), ("decoder", self.decoder)):
            a = st.args
            nq, nkv = a.num_attention_heads * a.head_dim, a.num_key_value_heads * a.head_dim
            for l in range(a.num_hidden_layers):
                p = f"{name}.layers.{l}."
                out[p + "self_attn.q_proj.weight"] = st.wqkv[l][:nq]
                out[p + "self_attn.k_proj.weight"] = st.wqkv[l][nq:nq + nkv]
                out[p + "self_attn.v_proj.weight"] = st.wqkv[l][nq + nkv:]
                out[p + "self_attn.o_proj.weight"] = st.wo[l]
                out[p + "mlp.gate_proj.weight"] = st.wgu[l][:a.intermediate_size]
                out[p + "mlp.up_proj.weight"] = st.wgu[l][a.intermediate_size:]
                out[p + "mlp.down_proj.weight"] = st.wdown[l]
                out[p + "input_layernorm.weight"] = st.norm_in[l]
                out[p + "post_attention_layernorm.weight"] = st.norm_post[l]
            out[f"{name}.norm.weight"] = st.norm_final
        return out

    # ------------------------------------------------------------------ loading
    def load_weights(self, file_or_weights: Union[str, Iterable[Tuple[str, object]], Dict[str, object]],
                     strict: bool = True) -> "CSM":
        """``mlx.nn.Module.load_weights`` semantics: a ``.safetensors``/``.npz`` path or (name, array) pairs;
        ``strict`` ⇒ ValueError on missing / unexpected names or shape mismatch.  Returns self."""
        if isinstance(file_or_weights, (str, bytes)) or hasattr(file_or_weights, "__fspath__"):
            path = str(file_or_weights)
            if path.endswith(".safetensors"):
                from safetensors.torch import load_file

                weights = load_file(path)
            elif path.endswith(".npz"):
                import numpy as np

                weights = {k: torch.from_numpy(v) for k, v in np.load(path).items()}
            else:
                raise ValueError(f"Unsupported weights file format: {path}")
        elif isinstance(file_or_weights, dict):
            weights = dict(file_or_weights)
        else:
            weights = dict(list(file_or_weights))
        weights = {k: (v if isinstance(v, torch.Tensor) else torch.as_tensor(_to_numpy(v))) for k, v in weights.items()}
        weights = normalize_checkpoint_keys(weights)

        expected = _expected_shapes(self.args)
        if strict:
            extra = sorted(set(weights) - set(expected))
            if extra:
                raise ValueError(f"Received parameters not in model: {' '.join(extra[:8])}{' …' if len(extra) > 8 else ''}.")
            missing = sorted(set(expected) - set(weights))
            if missing:
                raise ValueError(f"Missing parameters: {' '.join(missing[:8])}{' …' if len(missing) > 8 else ''}.")
        for k, v in weights.items():
            if k in expected and tuple(v.shape) != expected[k]:
                raise ValueError(f"Expected shape {expected[k]} but received shape {tuple(v.shape)} for parameter {k}")
        known = {k: v for k, v in weights.items() if k in expected}
        if not self._loaded:
            missing = sorted(set(expected) - set(known))
            if missing:
                raise ValueError(
                    "first load_weights() must provide every parameter (the model holds no random init): missing "
                    + " ".join(missing[:8]))
            self._finalize(known)
        else:
            if self.quantized:
                raise RuntimeError("load_weights() on a quantised model: load into a fresh CSM and quantise again")
            cur = self.parameters()
            self._proj_table = None  # derived from projection / audio_embeddings: rebuilt on next use
            self.weights_version += 1
            for k, v in known.items():
                if k == "audio_head":
                    self._audio_head_t.copy_(v.transpose(1, 2).to(self._audio_head_t.dtype))
                else:
                    cur[k].copy_(v.to(cur[k].dtype))
        return self

    def _finalize(self, W: Dict[str, torch.Tensor]) -> None:
        _lib.require_device(self.device)
        dev = self.device
        bf = lambda t: t.to(device=dev, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
        self.text_embeddings.weight = bf(W["text_embeddings.weight"])
        self.audio_embeddings.weight = bf(W["audio_embeddings.weight"])
        self.projection.weight = bf(W["projection.weight"])
        self.codebook0_head.weight = bf(W["codebook0_head.weight"])
        self._audio_head_t = bf(W["audio_head"].transpose(1, 2))
        for name, st in (("backbone", self.backbone), ("decoder", self.decoder)):
            a = st.args
            for l in range(a.num_hidden_layers):
                p = f"{name}.layers.{l}."
                st.wqkv.append(bf(torch.cat([W[p + "self_attn.q_proj.weight"], W[p + "self_attn.k_proj.weight"],
                                             W[p + "self_attn.v_proj.weight"]], dim=0)))
                st.wo.append(bf(W[p + "self_attn.o_proj.weight"]))
                st.wgu.append(bf(torch.cat([W[p + "mlp.gate_proj.weight"], W[p + "mlp.up_proj.weight"]], dim=0)))
                st.wdown.append(bf(W[p + "mlp.down_proj.weight"]))
                st.norm_in.append(f32(W[p + "input_layernorm.weight"]))
                st.norm_post.append(f32(W[p + "post_attention_layernorm.weight"]))
            st.norm_final = f32(W[f"{name}.norm.weight"])
            st.rope = f32(llama3_rope_table(a.head_dim, a.rope_theta, float(a.rope_scaling.get("factor", 1.0)),
                                            MAX_SEQ_LEN))
        self._loaded = True
        self._desc = None
        self.weights_version += 1

    # ------------------------------------------------------------------ weight-only FP8 (nn.quantize analogue)
    def quantize_weights(self) -> "CSM":
        """In place: every Linear matrix (q|k|v, o, gate|up, down of both stacks, projection, codebook0_head, audio_head)
        becomes an E4M3 blob with one fp32 scale per output channel (``quantization.py``; /root/reference README.md:92-128 calls
        ``nn.quantize(csm)`` at this point).  Embeddings and norm weights are untouched.  Halves the streamed bytes of a frame;
        served by the batch-1 frame kernel and the row-based GEMV path (the tensor-core chain declines a quantised model)."""
        from . import quantization as qz

        self._require_loaded()
        if self.quantized:
            return self
        self._qshapes: Dict[str, Tuple[int, int]] = {}

        def blob(t: torch.Tensor, name: str) -> torch.Tensor:
            q, s = qz.quantize_rows_e4m3(t)
            self._qshapes[name] = (int(t.shape[0]), int(t.shape[1]))
            return qz.pack_blob(q, s)

        for name, st in (("backbone", self.backbone), ("decoder", self.decoder)):
            for l in range(st.args.num_hidden_layers):
                st.wqkv[l] = blob(st.wqkv[l], f"{name}.{l}.wqkv")
                st.wo[l] = blob(st.wo[l], f"{name}.{l}.wo")
                st.wgu[l] = blob(st.wgu[l], f"{name}.{l}.wgu")
                st.wdown[l] = blob(st.wdown[l], f"{name}.{l}.wdown")
        self.projection.weight = blob(self.projection.weight, "projection")
        self.codebook0_head.weight = blob(self.codebook0_head.weight, "codebook0_head")
        heads = [blob(self._audio_head_t[i], "audio_head") for i in range(self._audio_head_t.shape[0])]
        self._audio_head_t = torch.cat(heads)      # n_codebooks-1 blobs back to back
        self.quantized = True
        self._proj_table = None
        self._desc = None
        self.weights_version += 1
        return self

    def _dequantized_parameters(self) -> Dict[str, torch.Tensor]:
        from . import quantization as qz

        def deq(blob: torch.Tensor, name: str) -> torch.Tensor:
            n, k = self._qshapes[name]
            return qz.dequantize_rows_e4m3(*qz.unpack_blob(blob, n, k))

        out = {"text_embeddings.weight": self.text_embeddings.weight, "audio_embeddings.weight": self.audio_embeddings.weight,
               "projection.weight": deq(self.projection.weight, "projection"),
               "codebook0_head.weight": deq(self.codebook0_head.weight, "codebook0_head")}
        v, dd = self._qshapes["audio_head"]
        nb = qz.blob_bytes(v, dd)
        out["audio_head"] = torch.stack([deq(self._audio_head_t[i * nb:(i + 1) * nb], "audio_head").t()
                                         for i in range(self.n_audio_codebooks - 1)])
        for name, st in (("backbone", self.backbone), ("decoder", self.decoder)):
            a = st.args
            nq, nkv = a.num_attention_heads * a.head_dim, a.num_key_value_heads * a.head_dim
            for l in range(a.num_hidden_layers):
                p = f"{name}.layers.{l}."
                qkv, gu = deq(st.wqkv[l], f"{name}.{l}.wqkv"), deq(st.wgu[l], f"{name}.{l}.wgu")
                out[p + "self_attn.q_proj.weight"], out[p + "self_attn.k_proj.weight"] = qkv[:nq], qkv[nq:nq + nkv]
                out[p + "self_attn.v_proj.weight"] = qkv[nq + nkv:]
                out[p + "self_attn.o_proj.weight"] = deq(st.wo[l], f"{name}.{l}.wo")
                out[p + "mlp.gate_proj.weight"], out[p + "mlp.up_proj.weight"] = gu[:a.intermediate_size], gu[a.intermediate_size:]
                out[p + "mlp.down_proj.weight"] = deq(st.wdown[l], f"{name}.{l}.wdown")
                out[p + "input_layernorm.weight"] = st.norm_in[l]
                out[p + "post_attention_layernorm.weight"] = st.norm_post[l]
            out[f"{name}.norm.weight"] = st.norm_final
        return out

    def _require_loaded(self) -> None:
        if not self._loaded:
            raise RuntimeError("CSM weights are not loaded: call model.load_weights(...) first")

    # ------------------------------------------------------------------ C descriptor
    def desc(self) -> _lib.Model:
        self._require_loaded()
        if self._desc is None:
            m = _lib.Model()
            for st, dst in ((self.backbone, m.backbone), (self.decoder, m.decoder)):
                a = st.args
                dst.n_layers, dst.d_model = a.num_hidden_layers, a.hidden_size
                dst.n_heads, dst.n_kv_heads, dst.head_dim = a.num_attention_heads, a.num_key_value_heads, a.head_dim
                dst.d_ff, dst.eps = a.intermediate_size, a.rms_norm_eps
                for l in range(a.num_hidden_layers):
                    dst.wqkv[l] = st.wqkv[l].data_ptr()
                    dst.wo[l] = st.wo[l].data_ptr()
                    dst.wgu[l] = st.wgu[l].data_ptr()
                    dst.wdown[l] = st.wdown[l].data_ptr()
                    dst.norm_in[l] = st.norm_in[l].data_ptr()
                    dst.norm_post[l] = st.norm_post[l].data_ptr()
                dst.norm_final = st.norm_final.data_ptr()
                dst.rope = st.rope.data_ptr()
            m.text_emb = self.text_embeddings.weight.data_ptr()
            m.audio_emb = self.audio_embeddings.weight.data_ptr()
            m.projection = self.projection.weight.data_ptr()
            m.c0_head = self.codebook0_head.weight.data_ptr()
            m.audio_head_t = self._audio_head_t.data_ptr()
            m.n_text_vocab, m.audio_vocab = self.n_text_vocab, self.n_audio_vocab
            m.n_codebooks, m.max_pos = self.n_audio_codebooks, MAX_SEQ_LEN
            m.weight_format = _lib.WEIGHTS_E4M3 if self.quantized else _lib.WEIGHTS_BF16
            self._desc = m
        return self._desc

    def proj_table(self) -> Optional[torch.Tensor]:
        """(n_codebooks, n_audio_vocab, d_decoder) fp32 table of ``projection(embed_audio(cb, token))`` (generation.py:75 over
        models.py:79-80 rows) for the batched chain's depth steps, built once on first use with the chain's own projection
        Linear (``csmb_build_proj_table``: bit-identical to running it).  None if the shapes are not the chain's."""
        self._require_loaded()
        if self.quantized:
            return None   # the chain declines weight-only FP8 models
        if self._proj_table is None:
            if self.n_backbone_embedding % 64 != 0 or self.n_decoder_embedding not in (1024, 2048):
                return None
            import ctypes as C

            dev_idx = _lib.require_device(self.device)
            l, m = _lib.lib(), self.desc()
            tab = torch.empty((self.n_audio_codebooks, self.n_audio_vocab, self.n_decoder_embedding), device=self.device,
                              dtype=torch.float32)
            ws = torch.empty((l.csmb_proj_table_workspace_bytes(C.byref(m)),), device=self.device, dtype=torch.uint8)
            _lib.check(l.csmb_build_proj_table(C.byref(m), tab.data_ptr(), ws.data_ptr(), ws.numel(), dev_idx,
                                               _lib.stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()  # the scratch is freed on return
            if int(ws[:4].view(torch.int32).item()) != 0:
                raise _lib.CsmbError("csmb_build_proj_table: a bounded wait of the tensor-core linear timed out")
            self._proj_table = tab
        return self._proj_table

    # ------------------------------------------------------------------ reference methods (models.py:79-92)
    def embed_audio(self, codebook: int, tokens: torch.Tensor) -> torch.Tensor:
        """(B,1) or (B,) int → (B,1,D) fp32 rows of ``audio_embeddings`` at ``tokens + codebook*n_audio_vocab``."""
        self._require_loaded()
        dev_idx = _lib.require_device(self.device)
        flat = tokens.reshape(-1).to(device=self.device, dtype=torch.int32).contiguous()
        d = self.n_backbone_embedding
        out = torch.empty((flat.numel(), d), device=self.device, dtype=torch.float32)
        _lib.check(_lib.lib().csmb_embed_audio(_lib.ptr(flat), _lib.ptr(self.audio_embeddings.weight), _lib.ptr(out),
                                               d, flat.numel(), d, int(codebook), self.n_audio_vocab, dev_idx,
                                               _lib.stream_ptr(self.device)))
        return out.reshape(*tokens.shape, d)

    def embed_frames(self, tokens: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """``embed_tokens`` × mask, summed over the 33 slots (models.py:82-92 + generation.py:32-36):
        (B,T,33) → (B,T,D) fp32."""
        self._require_loaded()
        dev_idx = _lib.require_device(self.device)
        tk = tokens.to(device=self.device, dtype=torch.int32).contiguous()
        mk = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        R = tk.numel() // tk.shape[-1]
        d = self.n_backbone_embedding
        out = torch.empty((R, d), device=self.device, dtype=torch.float32)
        _lib.check(_lib.lib().csmb_embed_sum(_lib.ptr(tk), _lib.ptr(mk), _lib.ptr(self.text_embeddings.weight),
                                             _lib.ptr(self.audio_embeddings.weight), _lib.ptr(out), R, d,
                                             self.n_audio_codebooks, self.n_audio_vocab, dev_idx,
                                             _lib.stream_ptr(self.device)))
        return out.reshape(*tk.shape[:-1], d)


_SESAME_RENAMES = (
    (".attn.q_proj.", ".self_attn.q_proj."), (".attn.k_proj.", ".self_attn.k_proj."), (".attn.v_proj.", ".self_attn.v_proj."),
    (".attn.output_proj.", ".self_attn.o_proj."), (".mlp.w1.", ".mlp.gate_proj."), (".mlp.w3.", ".mlp.up_proj."),
    (".mlp.w2.", ".mlp.down_proj."), (".sa_norm.scale", ".input_layernorm.weight"),
    (".mlp_norm.scale", ".post_attention_layernorm.weight"), (".norm.scale", ".norm.weight"),
)


def normalize_checkpoint_keys(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accepts the original ``sesame/csm-1b`` (torchtune) parameter names next to the mlx ones of
    ``senstella/csm-1b-mlx``: same tensors, same adjacent-pair RoPE row order, different
    names (``attn.output_proj`` / ``mlp.w1,w3,w2`` / ``sa_norm.scale`` …; a leading ``model.`` is dropped)."""
    if not any((".attn." in k) or k.endswith(".scale") or k.startswith("model.") or ".mlp.w" in k for k in weights):
        return weights
    out = {}
    for k, v in weights.items():
        k = k[len("model."):] if k.startswith("model.") else k
        for a, b in _SESAME_RENAMES:
            k = k.replace(a, b)
        out[k] = v
    return out


def _to_numpy(v):
    import numpy as np

    return np.asarray(v)
