import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` under gpurun)")


@pytest.fixture(scope="session")
def csm_weights():
    from csm_mlx_b200.random_init import random_csm_weights

    return random_csm_weights()


@pytest.fixture(scope="session")
def mimi_weights():
    from csm_mlx_b200.random_init import random_mimi_weights

    return random_mimi_weights()


@pytest.fixture(scope="session")
def oracle_1b(csm_weights):
    from oracle import lm

    return lm.OracleCSM(lm.CSM_1B, csm_weights)


@pytest.fixture(scope="session")
def device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def model_1b(csm_weights, device):
    from csm_mlx_b200 import CSM, csm_1b

    return CSM(csm_1b(), device=device).load_weights(csm_weights)


@pytest.fixture(scope="session")
def mimi_gpu(mimi_weights, device):
    from csm_mlx_b200 import tokenizers
    from csm_mlx_b200.mimi import Mimi

    m = Mimi(32, device=device).load_pytorch_weights(mimi_weights)
    tokenizers.set_audio_tokenizer(m)
    return m


def snr_db(ref: torch.Tensor, x: torch.Tensor) -> float:
    err = (x.double() - ref.double()).pow(2).mean()
    return float(10 * torch.log10(ref.double().pow(2).mean() / err.clamp_min(1e-300)))
