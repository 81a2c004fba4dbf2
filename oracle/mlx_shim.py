"""A minimal stand-in for ``mlx`` / ``mlx_lm`` / ``moshi_mlx`` so that the REFERENCE'S OWN Python modules
(``/root/reference/csm_mlx/{attention,models,generation,tokenizers,segment,config}.py``) can be imported and run,
unmodified, in this container — where none of those packages can be installed.

TEST INFRASTRUCTURE (see oracle/__init__.py): used only by ``scripts/make_reference_golden.py`` to generate
``tests/golden/reference_*.npz``; the product never imports it.

What runs for real through this shim is every line the reference owns on the hot path: ``Llama3ScaledRoPE`` and
``Attention`` (attention.py), the ``CSM`` parameter tree and ``embed_tokens`` / ``embed_audio`` (models.py),
``generate_frame`` and the ``generate`` driver (generation.py), the frame assembly of tokenizers.py.  What the shim
has to supply is the third-party part, restated from the published mlx / mlx-lm sources (pyproject pins
``mlx>=0.22.1``, ``mlx-lm>=0.22.0``, ``moshi-mlx>=0.2.3``; no lock file, exact versions unknown):

* ``mlx.core``: ``array`` is a ``torch.Tensor`` subclass with the handful of mlx methods the reference uses
  (``astype``, permutation-style ``transpose``); the ~20 functions used (``concat``, ``expand_dims``, ``stack``,
  ``repeat`` …) map onto torch CPU fp32 ops.
* ``mlx.nn``: ``Module`` (attribute tree + ``load_weights``), ``Linear`` (x Wᵀ + b), ``Embedding``, ``RMSNorm``
  (fp32, x·rsqrt(mean x² + eps)·w), ``Identity``.
* ``mlx_lm.models.llama``: ``ModelArgs``, ``LlamaModel`` = embed → N × [h += attn(norm(h)); h += down(silu(gate(n))·up(n))]
  → norm, with a causal additive mask iff T > 1; ``mlx_lm.models.base.scaled_dot_product_attention`` =
  softmax(q·kᵀ·scale + mask)·v; ``mlx_lm.models.cache.KVCache`` = growing (B, H, S, hd) K/V with ``offset``.
* ``moshi_mlx``, ``audiofile``, ``audresample``: import stubs only (the codec is pinned against HF elsewhere).
"""

from __future__ import annotations

import contextlib
import sys
import types
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch


# --------------------------------------------------------------------------------------------- mlx.core
class array(torch.Tensor):
    """mx.array over torch: results of torch ops stay ``array`` through the default ``__torch_function__``."""

    @staticmethod
    def __new__(cls, data=None, dtype=None):
        if isinstance(data, (list, tuple)) and len(data) and isinstance(data[0], torch.Tensor):
            t = torch.stack([torch.as_tensor(d) for d in data])
            t = t.to(dtype) if dtype is not None else t
        else:
            t = torch.as_tensor(data, dtype=dtype)
        if t.dtype == torch.float64:
            t = t.to(torch.float32)          # mlx has no float64 on the default path
        return t.as_subclass(cls)

    def astype(self, dtype):
        return self.to(dtype)

    def transpose(self, *axes):              # mlx: a full permutation (or none = reverse)
        if len(axes) == 1 and isinstance(axes[0], (list, tuple)):
            axes = tuple(axes[0])
        if not axes:
            axes = tuple(reversed(range(self.dim())))
        if len(axes) == 2 and self.dim() != 2:       # torch-style swap used inside this stand-in (invalid in mlx)
            return torch.Tensor.transpose(self, axes[0], axes[1])
        return self.permute(*axes)


def _wrap(t):
    return t.as_subclass(array) if isinstance(t, torch.Tensor) and not isinstance(t, array) else t


def _make_core() -> types.ModuleType:
    mx = types.ModuleType("mlx.core")
    mx.array = array
    mx.float32, mx.float16, mx.bfloat16 = torch.float32, torch.float16, torch.bfloat16
    mx.int32, mx.int64, mx.bool_, mx.uint32 = torch.int32, torch.int64, torch.bool, torch.int64

    def _shape(shape):
        return (shape,) if isinstance(shape, int) else tuple(shape)

    mx.zeros = lambda shape, dtype=torch.float32: _wrap(torch.zeros(_shape(shape), dtype=dtype))
    mx.ones = lambda shape, dtype=torch.float32: _wrap(torch.ones(_shape(shape), dtype=dtype))
    mx.ones_like = lambda a: _wrap(torch.ones_like(a))
    mx.zeros_like = lambda a: _wrap(torch.zeros_like(a))

    def arange(*args, dtype=None):
        t = torch.arange(*args)
        if dtype is not None:
            t = t.to(dtype)
        elif not t.is_floating_point():
            t = t.to(torch.int32)
        return _wrap(t)

    mx.arange = arange
    mx.concat = mx.concatenate = lambda arrs, axis=0: _wrap(torch.cat([torch.as_tensor(a) for a in arrs], dim=axis))
    mx.stack = lambda arrs, axis=0: _wrap(torch.stack(list(arrs), dim=axis))
    mx.expand_dims = lambda a, axis: _wrap(torch.unsqueeze(a, axis))
    mx.repeat = lambda a, repeats, axis=None: _wrap(torch.repeat_interleave(a, repeats, dim=axis))
    mx.argmax = lambda a, axis=None, keepdims=False: _wrap(torch.argmax(a, dim=axis, keepdim=keepdims).to(torch.int32))
    mx.matmul = lambda a, b: _wrap(torch.matmul(a, b))
    mx.einsum = lambda eq, *ops: _wrap(torch.einsum(eq.replace(" ", ""), *ops))
    mx.cos = lambda a: _wrap(torch.cos(a))
    mx.sin = lambda a: _wrap(torch.sin(a))
    mx.power = lambda a, b: _wrap(torch.pow(a, b))
    mx.where = lambda c, a, b: _wrap(torch.where(c, a, b))
    mx.softmax = lambda a, axis=-1: _wrap(torch.softmax(a, dim=axis))
    mx.eval = lambda *a, **k: None

    class Stream:  # noqa: D401
        pass

    mx.Stream = Stream
    mx.new_stream = lambda device=None: Stream()
    mx.default_device = lambda: "cpu"
    mx.stream = lambda s=None: contextlib.nullcontext()
    rnd = types.ModuleType("mlx.core.random")

    def categorical(logits, axis=-1):
        raise NotImplementedError("mx.random.categorical: the golden vectors are generated with temperature 0")

    rnd.categorical = categorical
    rnd.seed = lambda s: None
    mx.random = rnd
    return mx


# --------------------------------------------------------------------------------------------- mlx.nn
class Module:
    def __init__(self):
        pass

    def __call__(self, *a, **k):
        raise NotImplementedError

    def load_weights(self, file_or_weights, strict: bool = True):
        """(name, array) pairs keyed by dotted attribute paths, list indices included (mlx.nn.Module.load_weights)."""
        pairs = list(file_or_weights.items()) if isinstance(file_or_weights, dict) else list(file_or_weights)
        for name, value in pairs:
            obj: Any = self
            parts = name.split(".")
            for p in parts[:-1]:
                obj = obj[int(p)] if isinstance(obj, (list, tuple)) else getattr(obj, p)
            if strict and not hasattr(obj, parts[-1]):
                raise ValueError(f"Received parameters not in model: {name}.")
            setattr(obj, parts[-1], array(torch.as_tensor(value, dtype=torch.float32)))
        return self

    def eval(self):
        return self


class Identity(Module):
    def __call__(self, x, *a, **k):
        return x


class Linear(Module):
    def __init__(self, input_dims: int, output_dims: int, bias: bool = True):
        super().__init__()
        self.weight = None      # (out, in); filled by load_weights (the reference's random init is never used)
        self.bias = None
        self._shape = (output_dims, input_dims)
        self._has_bias = bias

    def __call__(self, x):
        y = _wrap(torch.matmul(x, self.weight.t()))
        return y + self.bias if self.bias is not None else y


class Embedding(Module):
    def __init__(self, num_embeddings: int, dims: int):
        super().__init__()
        self.weight = None
        self._shape = (num_embeddings, dims)

    def __call__(self, x):
        return _wrap(self.weight[torch.as_tensor(x).long()])


class RMSNorm(Module):
    def __init__(self, dims: int, eps: float = 1e-5):
        super().__init__()
        self.weight = array(torch.ones(dims))
        self.eps = eps

    def __call__(self, x):
        xf = x.to(torch.float32)
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        return _wrap((y * self.weight).to(x.dtype))


def _silu(x):
    return _wrap(x * torch.sigmoid(x))


# --------------------------------------------------------------------------------------------- mlx_lm
@dataclass
class BaseModelArgs:
    @classmethod
    def from_dict(cls, params):
        import inspect

        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})


@dataclass
class LlamaArgs(BaseModelArgs):
    model_type: str = "llama"
    hidden_size: int = 0
    num_hidden_layers: int = 0
    intermediate_size: int = 0
    num_attention_heads: int = 0
    rms_norm_eps: float = 1e-5
    vocab_size: int = 0
    head_dim: Optional[int] = None
    max_position_embeddings: Optional[int] = None
    num_key_value_heads: Optional[int] = None
    attention_bias: bool = False
    mlp_bias: bool = False
    rope_theta: float = 10000
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Any]] = None
    tie_word_embeddings: bool = True


def create_attention_mask(h, cache=None):
    """mlx_lm.models.base: a causal mask only when more than one query row is processed."""
    T = h.shape[1]
    if T <= 1:
        return None
    offset = 0
    if cache is not None and cache[0] is not None:
        offset = cache[0].offset
    rows = torch.arange(offset, offset + T)[:, None]
    cols = torch.arange(offset + T)[None, :]
    return _wrap(torch.where(cols <= rows, 0.0, float("-inf")).to(torch.float32))


def scaled_dot_product_attention(queries, keys, values, cache=None, scale: float = 1.0, mask=None):
    scores = torch.matmul(queries.to(torch.float32) * scale, keys.to(torch.float32).transpose(-1, -2))
    if mask is not None:
        scores = scores + mask
    return _wrap(torch.matmul(torch.softmax(scores, dim=-1), values.to(torch.float32)).to(queries.dtype))


class KVCache:
    def __init__(self):
        self.keys = None
        self.values = None
        self.offset = 0

    def update_and_fetch(self, keys, values):
        self.keys = keys if self.keys is None else _wrap(torch.cat([self.keys, keys], dim=2))
        self.values = values if self.values is None else _wrap(torch.cat([self.values, values], dim=2))
        self.offset += keys.shape[2]
        return self.keys, self.values


class _MLP(Module):
    def __init__(self, args: LlamaArgs):
        super().__init__()
        self.gate_proj = Linear(args.hidden_size, args.intermediate_size, bias=args.mlp_bias)
        self.down_proj = Linear(args.intermediate_size, args.hidden_size, bias=args.mlp_bias)
        self.up_proj = Linear(args.hidden_size, args.intermediate_size, bias=args.mlp_bias)

    def __call__(self, x):
        return self.down_proj(_silu(self.gate_proj(x)) * self.up_proj(x))


class _TransformerBlock(Module):
    def __init__(self, args: LlamaArgs):
        super().__init__()
        self.self_attn = None                      # the reference installs its own Attention (models.py:70-77)
        self.mlp = _MLP(args)
        self.input_layernorm = RMSNorm(args.hidden_size, eps=args.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(args.hidden_size, eps=args.rms_norm_eps)

    def __call__(self, x, mask=None, cache=None):
        h = x + self.self_attn(self.input_layernorm(x), mask, cache)
        return h + self.mlp(self.post_attention_layernorm(h))


class LlamaModel(Module):
    def __init__(self, args: LlamaArgs):
        super().__init__()
        self.args = args
        self.vocab_size = args.vocab_size
        self.num_hidden_layers = args.num_hidden_layers
        self.embed_tokens = Embedding(args.vocab_size, args.hidden_size)   # replaced by Identity in the reference
        self.layers = [_TransformerBlock(args) for _ in range(args.num_hidden_layers)]
        self.norm = RMSNorm(args.hidden_size, eps=args.rms_norm_eps)

    def __call__(self, inputs, mask=None, cache=None):
        h = self.embed_tokens(inputs)
        if mask is None:
            mask = create_attention_mask(h, cache)
        if cache is None:
            cache = [None] * len(self.layers)
        for layer, c in zip(self.layers, cache):
            h = layer(h, mask, cache=c)
        return self.norm(h)


# --------------------------------------------------------------------------------------------- installation
def install(reference_root: str = "/root/reference") -> types.ModuleType:
    """Registers the stand-in modules and a ``csm_mlx`` package whose sub-modules load from the reference tree WITHOUT
    running its ``__init__`` (which imports the fine-tuning stack).  Returns the ``mlx.core`` stand-in."""
    import os

    mx = _make_core()
    mlx = types.ModuleType("mlx")
    nn = types.ModuleType("mlx.nn")
    for k, v in dict(Module=Module, Linear=Linear, Embedding=Embedding, RMSNorm=RMSNorm, Identity=Identity, silu=_silu).items():
        setattr(nn, k, v)
    mlx.core, mlx.nn = mx, nn
    mods = {"mlx": mlx, "mlx.core": mx, "mlx.nn": nn, "mlx.core.random": mx.random}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        mods[name] = m
        return m

    mod("mlx_lm")
    mod("mlx_lm.models")
    mod("mlx_lm.models.base", BaseModelArgs=BaseModelArgs, scaled_dot_product_attention=scaled_dot_product_attention,
        create_attention_mask=create_attention_mask)
    mod("mlx_lm.models.llama", ModelArgs=LlamaArgs, LlamaModel=LlamaModel)
    mod("mlx_lm.models.cache", KVCache=KVCache)

    class _NoCodec:
        def __init__(self, *a, **k):
            raise RuntimeError("moshi_mlx is not available: the caller must patch get_audio_tokenizer")

    mod("moshi_mlx")
    mod("moshi_mlx.models")
    mod("moshi_mlx.models.mimi", Mimi=_NoCodec, mimi_202407=lambda n: None)
    mod("audiofile", read=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("audiofile is not available")))
    mod("audresample", resample=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("audresample is not available")))
    pkg = types.ModuleType("csm_mlx")
    pkg.__path__ = [os.path.join(reference_root, "csm_mlx")]
    mods["csm_mlx"] = pkg
    for k in list(sys.modules):
        if k == "csm_mlx" or k.startswith("csm_mlx."):
            del sys.modules[k]
    import importlib.machinery

    for name, m in mods.items():
        if getattr(m, "__spec__", None) is None:
            m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=hasattr(m, "__path__"))
    sys.modules.update(mods)
    return mx
