"""csm_mlx_b200 — B200-native drop-in for the ``csm_mlx`` generation hot path.

Same export list as ``/root/reference/csm_mlx/__init__.py:1-16``.  The training exports (``CSMDataset``,
``CSMTrainer``, ``TrainArgs``) are outside the generation hot path and raise ``NotImplementedError``
when used; ``load_adapters`` folds a LoRA / full adapter into the dense weights (``adapters.py``).  ``import csm_mlx`` resolves to a thin alias package of this one.
"""

from .generation import generate, generate_batch, generate_frame, make_cache, stream_generate
from .models import CSM, ModelArgs, csm_1b, csm_tiny
from .sample_utils import make_logits_processors, make_sampler
from .segment import Segment
from .serving import ContextCache, Engine, KVPrefixCache


def _out_of_scope(name):
    class _Stub:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} belongs to the fine-tuning stack, which is outside the generation hot path")

    _Stub.__name__ = name
    return _Stub


CSMDataset = _out_of_scope("CSMDataset")
CSMTrainer = _out_of_scope("CSMTrainer")
TrainArgs = _out_of_scope("TrainArgs")


from .adapters import load_adapters  # noqa: E402  (inference-side half of finetune/utils.py:84-108: adapters are merged)
from .quantization import quantize  # noqa: E402  (what README.md:92-128 does with mlx.nn.quantize(csm): weight-only FP8 here)


__all__ = [
    "generate",
    "stream_generate",
    "CSM",
    "csm_1b",
    "Segment",
    "CSMDataset",
    "CSMTrainer",
    "TrainArgs",
    "load_adapters",
]
