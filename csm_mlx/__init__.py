"""Alias so that ``import csm_mlx`` / ``from csm_mlx import CSM, csm_1b, generate`` (README.md:29-55 of the
reference) resolve to the B200 implementation in ``csm_mlx_b200``."""

import sys as _sys

import csm_mlx_b200 as _impl
from csm_mlx_b200 import *  # noqa: F401,F403
from csm_mlx_b200 import (attention, config, generation, models, sample_utils, segment, serving, tokenizers, utils)  # noqa: F401
from csm_mlx_b200 import __all__  # noqa: F401
from csm_mlx_b200 import generate_batch, generate_frame, make_cache, make_logits_processors, make_sampler  # noqa: F401
from csm_mlx_b200 import ContextCache, Engine, quantize  # noqa: F401  (throughput form of the path; nn.quantize analogue)

for _name in ("attention", "config", "generation", "models", "sample_utils", "segment", "serving", "tokenizers", "utils"):
    _sys.modules[f"csm_mlx.{_name}"] = getattr(_impl, _name)
