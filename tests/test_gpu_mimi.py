"""Codec parity on the GPU against golden vectors and the live oracle, for both whole-clip paths: the tensor-core path
(csrc/mimi_tc.cu: tcgen05 GEMMs on bf16 hi+lo planes, three MMAs per K step, ~1e-5 of fp32; the default) and the fp32
CUDA-core path (``CSMB_MIMI_FP32=1``; also what streaming decode runs).

Gates: waveforms match the oracle to SNR ≥ 80 dB (north-star gate: ≥ 40 dB).  RVQ codes are index work: they must be
IDENTICAL to the oracle's, except where the nearest-neighbour search had a float-level near-tie — every disagreeing
frame is refereed in float64 on the oracle's own latent (``oracle.mimi.rvq_disagreement_margins``): at the first
differing codebook the two candidates' distances must agree to ``TIE_TOL`` of |residual|^2, and such frames must stay below
0.5 % of all frames."""
import os

import numpy as np
import pytest
import torch

from oracle import mimi as omimi
from tests.codec_referee import assert_codes_match
from tests.conftest import GOLDEN, snr_db
from tests.workloads import synthetic_audio

pytestmark = pytest.mark.gpu

@pytest.fixture(params=["tc", "fp32"])
def codec_path(request, monkeypatch):
    monkeypatch.setenv("CSMB_MIMI_FP32", "1" if request.param == "fp32" else "0")
    return request.param


def test_decode_vs_golden(mimi_gpu, device, codec_path):
    g = np.load(os.path.join(GOLDEN, "mimi.npz"))
    a = mimi_gpu.decode(torch.from_numpy(g["decode_codes"]).to(device)).cpu()[0, 0]
    assert a.shape == (48000,)
    assert snr_db(torch.from_numpy(g["decode_head"]), a[:9600]) > 80
    assert snr_db(torch.from_numpy(g["decode_tail"]), a[-1920:]) > 80
    assert abs(float(a.double().pow(2).sum()) - float(g["decode_sumsq"])) / float(g["decode_sumsq"]) < 1e-4


def test_decode_batch_long_window_vs_oracle(mimi_gpu, mimi_weights, device, codec_path):
    gen = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 2048, (2, 32, 140), generator=gen)  # 280 latent steps > 250-step attention window
    ref = omimi.decode(codes, mimi_weights)
    got = mimi_gpu.decode(codes.to(device)).cpu()
    assert got.shape == ref.shape == (2, 1, 140 * 1920)
    assert snr_db(ref, got) > 80


def test_decode_clamps_out_of_range_ids(mimi_gpu, mimi_weights, device, codec_path):
    """CSM heads emit ids up to 2050, Mimi has 2048 bins: ids are clamped to the last bin."""
    codes = torch.full((1, 32, 2), 2050)
    ref = omimi.decode(codes.clamp(max=2047), mimi_weights)
    got = mimi_gpu.decode(codes.to(device)).cpu()
    assert snr_db(ref, got) > 80


@pytest.mark.parametrize("use_graph", [False, True])
def test_streaming_equals_offline(mimi_gpu, device, use_graph):
    gen = torch.Generator().manual_seed(4)
    codes = torch.randint(0, 2048, (1, 32, 12), generator=gen).to(device)
    off = mimi_gpu.decode(codes).cpu()
    st = mimi_gpu.new_decode_stream(1, use_graph=use_graph)
    chunks = [st.step(codes[:, :, i:i + 1]).cpu().clone() for i in range(12)]
    assert chunks[0].shape == (1, 1, 1920)
    assert snr_db(off, torch.cat(chunks, -1)) > 80
    if use_graph:
        assert st.graph is not None


def test_decode_step_api_and_reset(mimi_gpu, device):
    gen = torch.Generator().manual_seed(5)
    codes = torch.randint(0, 2048, (1, 32, 3), generator=gen).to(device)
    mimi_gpu.reset_state()
    a = [mimi_gpu.decode_step(codes[:, :, i:i + 1]) for i in range(3)]
    mimi_gpu.reset_state()
    b = [mimi_gpu.decode_step(codes[:, :, i:i + 1]) for i in range(3)]
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert snr_db(mimi_gpu.decode(codes).cpu(), torch.cat(a, -1).cpu()) > 80


@pytest.mark.parametrize("n", [120000, 120000 - 700, 1921])
def test_encode_vs_oracle(mimi_gpu, mimi_weights, device, codec_path, n):
    clip = synthetic_audio(11, 5.0)[:n]
    ref = omimi.encode(clip[None, None], mimi_weights)
    got = mimi_gpu.encode(clip[None, None].to(device)).cpu().long()
    assert got.shape == ref.shape == (1, 32, -(-n // 1920))
    assert_codes_match(got, ref, clip[None, None], mimi_weights)
    if n == 120000:
        g = np.load(os.path.join(GOLDEN, "mimi.npz"))
        assert_codes_match(got, torch.from_numpy(g["encode_codes"])[None], clip[None, None], mimi_weights)


def test_encode_batch_and_roundtrip_consistency(mimi_gpu, mimi_weights, device, codec_path):
    clips = torch.stack([synthetic_audio(100 + i, 2.0) for i in range(3)])[:, None]
    got = mimi_gpu.encode(clips.to(device)).cpu().long()
    ref = omimi.encode(clips, mimi_weights)
    assert_codes_match(got, ref, clips, mimi_weights)
    # decode(encode(x)) on the GPU equals the oracle's decode of the same codes
    assert snr_db(omimi.decode(got, mimi_weights), mimi_gpu.decode(got.to(device)).cpu()) > 80


def test_tensor_core_and_fp32_paths_agree(mimi_gpu, mimi_weights, device, monkeypatch):
    """The two whole-clip implementations against each other on a ragged batch (3 clips x 20 s, a length that is not a
    multiple of the frame): decode SNR > 80 dB, encode codes identical up to refereed near-ties."""
    n = 20 * 24000 - 777
    clips = torch.stack([synthetic_audio(200 + i, 20.0)[:n] for i in range(3)])[:, None]
    out = {}
    for path in ("fp32", "tc"):
        monkeypatch.setenv("CSMB_MIMI_FP32", "1" if path == "fp32" else "0")
        codes = mimi_gpu.encode(clips.to(device))
        out[path] = (codes.cpu(), mimi_gpu.decode(out["fp32"][0].to(device) if path == "tc" else codes).cpu())
    assert out["tc"][0].shape == (3, 32, -(-n // 1920))
    assert_codes_match(out["tc"][0], out["fp32"][0], clips, mimi_weights)
    assert snr_db(out["fp32"][1], out["tc"][1]) > 80


def test_long_batch_codec_roundtrip_shapes(mimi_gpu, device, codec_path):
    """configs[4] in miniature: 2 clips x 60 s (1.44 M samples each: more row tiles than gridDim.y allows) encode and
    decode; self-consistency: decoding the codes of the first 2 s equals decoding the same frames alone (causal codec)."""
    clips = torch.stack([synthetic_audio(100 + i, 60.0) for i in range(2)])[:, None].to(device)
    codes = mimi_gpu.encode(clips)
    assert codes.shape == (2, 32, 750)
    short = mimi_gpu.encode(clips[:, :, : 48000])
    assert torch.equal(codes[:, :, :25], short)  # causal encoder, same kernels and summation order: a prefix encodes to the same codes
    audio = mimi_gpu.decode(codes)
    assert audio.shape == (2, 1, 750 * 1920)
    head = mimi_gpu.decode(codes[:, :, :25].contiguous())
    assert snr_db(head.cpu(), audio[:, :, : 25 * 1920].cpu()) > 80
