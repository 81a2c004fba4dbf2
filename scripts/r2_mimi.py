"""Developer script (gpurun): tensor-core codec path against the fp32 path and the oracle, then timings.
Usage: python scripts/r2_mimi.py [check] [time]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from csm_mlx_b200.mimi import Mimi
from csm_mlx_b200.random_init import random_mimi_weights
from tests.workloads import synthetic_audio

dev = torch.device("cuda", 0)
W = random_mimi_weights()
mimi = Mimi(32, device=dev).load_pytorch_weights(W)
what = sys.argv[1:] or ["check", "time"]


def snr(ref, x):
    return float(10 * torch.log10(ref.double().pow(2).mean() / (x.double() - ref.double()).pow(2).mean().clamp_min(1e-300)))


def both(fn):
    out = {}
    for path in ("fp32", "tc"):
        os.environ["CSMB_MIMI_FP32"] = "1" if path == "fp32" else "0"
        out[path] = fn()
        torch.cuda.synchronize()
    os.environ["CSMB_MIMI_FP32"] = "0"
    return out


if "check" in what:
    from oracle import mimi as omimi
    gen = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 2048, (2, 32, 30), generator=gen)
    o = both(lambda: mimi.decode(codes.to(dev)).cpu())
    ref = omimi.decode(codes, W)
    print(f"decode 2x30 frames: tc vs fp32 {snr(o['fp32'], o['tc']):.1f} dB, tc vs oracle {snr(ref, o['tc']):.1f} dB, "
          f"fp32 vs oracle {snr(ref, o['fp32']):.1f} dB", flush=True)
    # per-stage localisation if something is off: prefix lengths
    for n in (1921, 48000, 120000 - 700):
        clip = synthetic_audio(11, 5.0)[:n][None, None]
        o = both(lambda: mimi.encode(clip.to(dev)).cpu())
        refc = omimi.encode(clip, W)
        print(f"encode n={n}: tc==fp32 {(o['tc'] == o['fp32']).float().mean().item():.4f}  tc==oracle "
              f"{(o['tc'].long() == refc).float().mean().item():.4f}  fp32==oracle {(o['fp32'].long() == refc).float().mean().item():.4f}",
              flush=True)

if "time" in what:
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for nclips, secs in ((4, 20.0), (4, 60.0), (16, 60.0)):
        clips = torch.stack([synthetic_audio(100 + i, secs) for i in range(nclips)])[:, None].to(dev)
        for path in ("tc", "fp32"):
            if path == "fp32" and nclips > 4:
                continue
            os.environ["CSMB_MIMI_FP32"] = "1" if path == "fp32" else "0"
            codes = mimi.encode(clips)
            mimi.decode(codes)
            torch.cuda.synchronize()
            a, b, c = ev(), ev(), ev()
            a.record()
            codes = mimi.encode(clips)
            b.record()
            audio = mimi.decode(codes)
            c.record()
            torch.cuda.synchronize()
            tot = nclips * secs
            te, td = a.elapsed_time(b) / 1e3, b.elapsed_time(c) / 1e3
            print(f"{path} {nclips} x {secs:.0f} s: encode {tot / te:8.0f} audio-s/s ({te * 1e3:.1f} ms, {0.46e9 * 12.5 * tot / te / 1e12:.1f} TFLOP/s)  "
                  f"decode {tot / td:8.0f} audio-s/s ({td * 1e3:.1f} ms, {0.43e9 * 12.5 * tot / td / 1e12:.1f} TFLOP/s)  "
                  f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
            del codes, audio
    os.environ["CSMB_MIMI_FP32"] = "0"
