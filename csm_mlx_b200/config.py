"""Hyper-parameters of the CSM hot path.

Mirrors ``/root/reference/csm_mlx/config.py:3-53`` (BACKBONE_CONFIGURATION["1b"],
DECODER_CONFIGURATION["100m"], TOKENIZERS) with a plain dataclass instead of
``mlx_lm.models.llama.ModelArgs``; field names are kept so callers that read
``model.backbone.args.<field>`` (generation.py:132) keep working.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional


@dataclass
class LlamaArgs:
    model_type: str = "llama"
    vocab_size: int = 128_256
    num_hidden_layers: int = 16
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 64
    intermediate_size: int = 8192
    hidden_size: int = 2048
    rms_norm_eps: float = 1e-5
    rope_scaling: Dict[str, object] = field(default_factory=lambda: {
        "factor": 32.0,
        "high_freq_factor": 4.0,
        "low_freq_factor": 1.0,
        "original_max_position_embeddings": 8192,
        "rope_type": "llama3",
    })
    rope_theta: float = 500_000.0
    # Unset in the reference's config ⇒ generate() falls back to 2048 (generation.py:132).
    max_position_embeddings: Optional[int] = None


BACKBONE_CONFIGURATION = {
    "1b": LlamaArgs(),
    # Structure-preserving miniature used by CPU-side host-logic tests (not a reference config).
    "tiny": LlamaArgs(vocab_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                      head_dim=32, intermediate_size=256, hidden_size=128),
}

DECODER_CONFIGURATION = {
    "100m": LlamaArgs(num_hidden_layers=4, num_attention_heads=8, num_key_value_heads=2, head_dim=128,
                      intermediate_size=8192, hidden_size=1024),
    "tiny": LlamaArgs(vocab_size=512, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                      head_dim=32, intermediate_size=128, hidden_size=64),
}

TOKENIZERS = {
    "audio": {
        "repo_id": "kyutai/moshiko-pytorch-bf16",
        "filename": "tokenizer-e351c8d8-checkpoint125.safetensors",
    },
    "text": {"repo_id": "unsloth/Llama-3.2-1B"},
}

# RoPE table length: Llama3ScaledRoPE(max_seq_len=2048) (attention.py:38); also the context cap.
MAX_SEQ_LEN = 2048
