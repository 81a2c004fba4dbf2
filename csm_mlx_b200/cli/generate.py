"""``python -m csm_mlx_b200.cli.generate TEXT -o out.wav …`` — the ``csm-mlx generate`` command of the reference
(``/root/reference/csm_mlx/cli/generate.py:72-202``) with the same options and defaults, on argparse (typer / rich
are not part of this image).  Weights are a local ``.safetensors`` path (the reference also accepts a Hugging Face
repo id; there is no network here, so a repo id is resolved only through an existing local HF cache)."""

from __future__ import annotations

import argparse
import os
import sys
from typing import List, Optional


def _resolve_weight(value: str) -> str:
    """cli/generate.py:26-45: local file, or ``repo_id`` → ckpt.safetensors from the local Hugging Face cache."""
    if os.path.isfile(value):
        return os.path.abspath(value)
    if os.path.isdir(value) and os.path.isfile(os.path.join(value, "ckpt.safetensors")):
        return os.path.abspath(os.path.join(value, "ckpt.safetensors"))
    try:
        from huggingface_hub import hf_hub_download

        return hf_hub_download(repo_id=value, filename="ckpt.safetensors", local_files_only=True)
    except Exception as e:  # noqa: BLE001
        raise SystemExit(f"Error! weight {value!r} is neither a local file nor a cached Hugging Face repo ({e})")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="csm-mlx generate", description="Generate speech from text using CSM (Conversational Speech Model).")
    p.add_argument("text")
    p.add_argument("--output", "-o", required=True, help="Output audio file path")
    p.add_argument("--model", "-m", default="1b", choices=["1b"], help="Model size")
    p.add_argument("--weight", "-w", default="senstella/csm-1b-mlx", help="Weight file path (HF repo ID or local path)")
    p.add_argument("--adapter", "-a", default=None, help="Path to adapter (adapter_config.json and adapters.safetensors)")
    p.add_argument("--speaker", "-s", type=int, default=0, help="Speaker ID to generate")
    p.add_argument("--max-audio-length", "-l", type=int, default=10000, help="Maximum audio length in miliseconds")
    p.add_argument("--temperature", "--temp", "-t", type=float, default=0.8, help="Sampling temperature")
    p.add_argument("--top-p", "-p", type=float, default=None, help="Top-p sampling parameter")
    p.add_argument("--min-p", type=float, default=None, help="Min-p sampling parameter")
    p.add_argument("--top-k", "-k", type=int, default=50, help="Top-k sampling parameter")
    p.add_argument("--min-tokens-to-keep", "-kt", type=int, default=1, help="Minimum tokens to keep during sampling")
    p.add_argument("--input-speakers", "-is", type=int, nargs="*", default=None, help="List of speaker IDs for context")
    p.add_argument("--input-audios", "-ia", nargs="*", default=None, help="List of audio files for context")
    p.add_argument("--input-texts", "-it", nargs="*", default=None, help="List of text transcripts for context")
    p.add_argument("--seed", type=int, default=None, help="Sampling seed (this implementation's Philox stream)")
    p.add_argument("--synthetic-assets", action="store_true",
                   help="offline smoke runs: seeded random-init CSM / Mimi weights and a stand-in text tokenizer instead "
                        "of the Hugging Face assets (the output is noise)")
    return p


def main(argv: Optional[List[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    input_audios = args.input_audios or []
    input_texts = args.input_texts or []
    input_speakers = args.input_speakers or []
    if len(input_audios) != len(input_texts) or len(input_audios) != len(input_speakers):
        print("Error! All context inputs (input_audios, input_texts, and input_speakers) must have the same length.",
              file=sys.stderr)
        return 1

    from .. import CSM, Segment, csm_1b, generate, load_adapters, make_sampler
    from ..utils import write_audio

    sampler = make_sampler(temp=args.temperature, top_p=args.top_p or 0.0, min_p=args.min_p or 0.0, top_k=args.top_k or -1,
                           min_tokens_to_keep=args.min_tokens_to_keep)           # cli/generate.py:168-174
    csm = CSM(csm_1b())
    if args.synthetic_assets:
        from .. import tokenizers
        from ..mimi import Mimi
        from ..random_init import random_csm_weights, random_mimi_weights

        csm.load_weights(random_csm_weights())
        tokenizers.set_text_tokenizer(tokenizers.SyntheticTextTokenizer())
        tokenizers.set_audio_tokenizer(Mimi(32, device=csm.device).load_pytorch_weights(random_mimi_weights()))
    else:
        csm.load_weights(_resolve_weight(args.weight))
    if args.adapter is not None:
        load_adapters(csm, args.adapter)
    context = [Segment(speaker, text, None, audio)                                # cli/generate.py:186-189
               for audio, text, speaker in zip(input_audios, input_texts, input_speakers)]
    result = generate(csm, args.text, args.speaker, context, args.max_audio_length, sampler=sampler, seed=args.seed)
    write_audio(result, args.output, 24000)
    print(f"Success! Audio saved to: {args.output}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
