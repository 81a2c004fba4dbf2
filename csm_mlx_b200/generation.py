"""Frame generation and the whole-utterance / streaming drivers.  Mirrors
``/root/reference/csm_mlx/generation.py``: ``generate_frame`` (:21-92), ``generate`` (:95-178),
``stream_generate`` (:181-258).  Host logic only — prompt assembly, the frame loop, EOS handling, the API
surface; all arithmetic runs in libcsm_b200.so through ``runtime.LMState`` and ``mimi.Mimi``.

Differences from the reference, all on the host side and deliberate:
* ``sampler=`` (the README/CLI form, README.md:43-52) is accepted next to ``temperature=`` (this fork's form).
* frames are produced with one-frame lookahead: frame t+1 is enqueued before frame t's tokens are inspected
  for EOS on the host, so the GPU never waits for Python; a frame generated past EOS is discarded.
* ``generate_batch`` (not in the reference, which is batch-1 by construction, generation.py:124,156) runs B
  independent utterances in lock-step for the request-sharded throughput configuration.
* conversation caches (``caches.py``): the reference Mimi-encodes and prefills every context segment on every call
  (generation.py:108-121).  Here a model remembers the codes of the segments it has seen and the backbone KV pages of a
  prompt's context rows; the next turn with the same context copies the pages and prefills only the new text rows — same
  tokens (row-invariant prompt pass), a few milliseconds less before the first chunk.  ``set_conversation_cache(model,
  enabled=False)`` or ``CSMB_DISABLE_CONV_CACHE=1`` switches it off.
"""

from __future__ import annotations

import os
import warnings
from typing import Callable, Generator, List, Optional, Sequence, Tuple, Union

import torch

from .caches import ContextCache, KVPrefixCache
from .config import MAX_SEQ_LEN
from .models import CSM
from .runtime import LMState, SamplerSpec
from .sample_utils import DeviceSampler
from .segment import Segment
from .tokenizers import get_audio_tokenizer, tokenize_text_segment

default_stream = None  # the reference exposes an mx.Stream here (generation.py:19); CUDA streams come from torch

LogitsProcessor = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]


def set_conversation_cache(model: CSM, enabled: bool = True, segments: int = 16, prefixes: int = 4,
                           prefix_bytes: int = 512 << 20) -> None:
    """(Re)creates or drops the model's conversation caches: up to ``segments`` tokenised context segments (host memory)
    and up to ``prefixes`` context KV prefixes (device memory, at most ``prefix_bytes``: 64 KiB per context row)."""
    if enabled:
        model.__dict__["_conv_cache"] = (ContextCache(segments, model.n_audio_codebooks), KVPrefixCache(prefixes, prefix_bytes))
    else:
        model.__dict__["_conv_cache"] = None


def _conv_cache(model: CSM) -> Optional[Tuple[ContextCache, KVPrefixCache]]:
    if os.environ.get("CSMB_DISABLE_CONV_CACHE", "0") == "1":
        return None
    if "_conv_cache" not in model.__dict__:
        set_conversation_cache(model)
    return model.__dict__["_conv_cache"]


def make_cache(model: CSM, batch: int = 1, max_len: int = MAX_SEQ_LEN) -> LMState:
    """What ``[KVCache() for _ in model.backbone.layers]`` (generation.py:127) is to the reference."""
    return LMState(model, batch, max_len)


def _resolve_sampler(temperature: float, sampler, seed: Optional[int]):
    """-> (SamplerSpec, host_callable or None)."""
    if sampler is None:
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        return SamplerSpec(temperature=float(temperature), seed=seed), None
    if isinstance(sampler, (DeviceSampler, SamplerSpec)):
        spec = sampler.spec if isinstance(sampler, DeviceSampler) else sampler
        if seed is None and spec.seed is None:   # unseeded sampler: a fresh Philox key per call, like the temperature= form
            seed = int.from_bytes(os.urandom(8), "little")
        if seed is not None:
            spec = SamplerSpec(**{**spec.__dict__, "seed": seed})
        return spec, None
    if callable(sampler):
        return SamplerSpec(temperature=0.0), sampler  # foreign callable: host round trip per codebook
    raise TypeError("sampler must be a csm_mlx.sample_utils.make_sampler(...) object or a callable logits -> ids")


def _frame_after_backbone(state: LMState, frame: torch.Tensor, spec: SamplerSpec, host_sampler,
                          logits_processors: Optional[List[LogitsProcessor]], c0_history: Optional[list]) -> None:
    """generation.py:42-90 given h_last / c0_logits already in `state`."""
    model = state.model
    logits = state.c0_logits
    if logits_processors:
        for proc in logits_processors:
            hist = torch.stack(c0_history, 0) if c0_history else torch.zeros((0,), device=logits.device)
            logits = proc(hist, logits)
    if host_sampler is None:
        state.sample_c0(frame, spec, logits if logits is not state.c0_logits else None)
    else:
        frame[:, 0] = host_sampler(logits).to(torch.int32).reshape(-1)
    if c0_history is not None:
        c0_history.append(frame[:, :1].clone())
    if host_sampler is None:
        state.depth_decode(frame, spec)
    else:
        ncb = model.n_audio_codebooks
        lg = torch.empty((state.batch, ncb, model.n_audio_vocab), device=state.device, dtype=torch.float32)
        for i in range(1, ncb):
            state.depth_decode(frame, spec, logits_out=lg, step_begin=i, step_end=i + 1)
            frame[:, i] = host_sampler(lg[:, i]).to(torch.int32).reshape(-1)


def generate_frame(model: CSM, tokens: torch.Tensor, *, temperature: float = 0.8,
                   token_mask: Optional[torch.Tensor] = None,
                   logits_processors: Optional[List[LogitsProcessor]] = None, cache: Optional[LMState] = None,
                   stream=None, c0_history: Optional[list] = None, sampler=None, seed: Optional[int] = None
                   ) -> torch.Tensor:
    """generation.py:21-92: tokens (B,T,33) [+ mask] -> (B,32) int32 on the model's device.  ``cache`` is an
    ``LMState`` (``make_cache``) carried across calls; ``stream`` is accepted for signature compatibility."""
    B, T, _ = tokens.shape
    mask = token_mask if token_mask is not None else torch.ones_like(tokens)
    state = cache if cache is not None else LMState(model, B, max_len=T)
    spec, host_sampler = _resolve_sampler(temperature, sampler, seed)
    frame = torch.zeros((B, model.n_audio_codebooks), device=model.device, dtype=torch.int32)
    if T == 1 and host_sampler is None and not logits_processors and max(state.pos_host) > 0 and c0_history is None:
        prev = tokens[:, 0, :-1].to(device=model.device, dtype=torch.int32).contiguous()
        if bool((mask[:, 0, :-1] != 0).all()) and not bool((mask[:, 0, -1] != 0).any()):
            state.decode_frame(prev, frame, spec)
            return frame
    state.prefill([tokens[b] for b in range(B)], [mask[b] for b in range(B)])
    _frame_after_backbone(state, frame, spec, host_sampler, logits_processors, c0_history)
    return frame


class _Session:
    """One lock-step generation over B utterances: prefill once, then a frame per ``step()``."""

    def __init__(self, model: CSM, prompts: Sequence[Tuple[torch.Tensor, torch.Tensor]], max_audio_frames: int,
                 spec: SamplerSpec, host_sampler, logits_processors, prefix_rows: Optional[Sequence[int]] = None):
        self.model, self.spec, self.host_sampler, self.procs = model, spec, host_sampler, logits_processors
        B = len(prompts)
        # leading rows of each prompt that are context (shared by the turns of a conversation): KV prefix cache unit
        self.prefix_rows = [0] * B if prefix_rows is None else [int(p) for p in prefix_rows]
        longest = max(int(p[0].shape[0]) for p in prompts)
        self.state = LMState.acquire(model, B, max_len=longest + max_audio_frames + 1)
        self.prompts = prompts
        self.c0_history: Optional[list] = [] if logits_processors else None
        self.prev: Optional[torch.Tensor] = None
        self.B = B
        self.fused = os.environ.get("CSMB_DISABLE_FUSED", "0") != "1"

    def step(self) -> torch.Tensor:
        st, model = self.state, self.model
        frame = torch.zeros((self.B, model.n_audio_codebooks), device=model.device, dtype=torch.int32)
        plain = self.host_sampler is None and not self.procs
        if self.prev is None:
            self._prefill()
            if plain and self.fused and st.fused_supported(self.spec):
                frame = st.first_frame_fused(self.spec)
            else:
                _frame_after_backbone(st, frame, self.spec, self.host_sampler, self.procs, self.c0_history)
        elif plain:
            if self.fused and st.fused_supported(self.spec):
                frame = st.decode_frame_fused(self.prev, self.spec)
            else:
                frame = st.decode_frame_graphed(self.prev, self.spec)
        else:
            st.backbone_step(self.prev)
            _frame_after_backbone(st, frame, self.spec, self.host_sampler, self.procs, self.c0_history)
        self.prev = frame
        return frame

    def _prefill(self) -> None:
        """The prompt pass (generation.py:34-42, T > 1); context rows whose KV pages are in the model's prefix cache are
        copied instead of recomputed, and a context seen for the first time is remembered."""
        st = self.state
        cc = _conv_cache(self.model)
        kvc = cc[1] if cc is not None and st._prefill_fast_ok() else None   # the per-op prompt pass is not row-invariant
        known, keys = [0] * self.B, [None] * self.B
        if kvc is not None:
            ver = int(getattr(self.model, "weights_version", 0))
            for b, ((tok, mask), p) in enumerate(zip(self.prompts, self.prefix_rows)):
                if 0 < p < int(tok.shape[0]) and kvc.capacity > 0:
                    keys[b] = KVPrefixCache.key(tok, mask, p, ver)
                    pages = kvc.get(keys[b])
                    if pages is not None and pages.device == st.kv_pool.device:
                        st.import_kv_prefix(b, pages)
                        st.pos_host[b] = p
                        known[b] = p
        st.prefill([p[0][k:] for p, k in zip(self.prompts, known)], [p[1][k:] for p, k in zip(self.prompts, known)])
        if kvc is not None:
            for b, (k, key) in enumerate(zip(known, keys)):
                if key is not None and k == 0:
                    kvc.put(key, st.export_kv_prefix(b, self.prefix_rows[b]), self.prefix_rows[b])

    def status_word(self) -> torch.Tensor:
        return self.state.status_word()

    @staticmethod
    def raise_if_aborted(host_status: torch.Tensor) -> None:
        """Called on the pinned copy that travels with every frame: a bounded wait that timed out inside a fused kernel
        (its tokens are garbage) surfaces one frame late instead of at the end of the utterance, or never."""
        if int(host_status[0]) != 0:
            raise RuntimeError(f"libcsm_b200: a fused LM kernel aborted (code {int(host_status[0])}); the frame is invalid")

    def close(self) -> None:
        """Raises if the persistent kernel reported an abort; otherwise returns the state to the model's pool."""
        st, self.state = self.state, None
        if st is not None:
            st.check_status()
            st.release()


def _build_prompt(model: CSM, text, speaker: int, context: Sequence[Segment]) -> Tuple[torch.Tensor, torch.Tensor]:
    """generation.py:108-121."""
    tok, mask, _ = _build_prompt_ex(model, text, speaker, context)
    return tok, mask


def _build_prompt_ex(model: CSM, text, speaker: int, context: Sequence[Segment]) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """-> (rows, mask, number of context rows).  Context segments are tokenised through the model's ``ContextCache`` (the rows
    of ``tokenize_segment``, tokenizers.py:88-102; the Mimi encode of a segment's audio runs once per distinct clip)."""
    cc = _conv_cache(model)
    seg_cache = cc[0] if cc is not None else ContextCache(0, model.n_audio_codebooks)
    toks, masks = [], []
    for seg in context:
        t, m = seg_cache.segment_rows(seg)
        toks.append(t)
        masks.append(m)
    n_context = sum(int(t.shape[0]) for t in toks)
    t, m = tokenize_text_segment(text, speaker, n_audio_codebooks=model.n_audio_codebooks)
    toks.append(t)
    masks.append(m)
    return torch.cat(toks, 0).to(torch.int32), torch.cat(masks, 0), n_context


def _check_length(model: CSM, n_rows: int, max_audio_frames: int) -> None:
    """generation.py:131-137."""
    context_window = model.backbone.args.max_position_embeddings or MAX_SEQ_LEN
    max_seq_len = context_window - max_audio_frames
    if n_rows >= max_seq_len:
        raise ValueError(
            f"Inputs too long ({n_rows}), must be below max_seq_len - max_audio_frames: {max_seq_len}")


class _HostMirror:
    """Double-buffered pinned host copies of per-frame results + the event that says they have landed."""

    def __init__(self, shapes_dtypes, device):
        self.bufs = [[torch.empty(s, dtype=d).pin_memory() for s, d in shapes_dtypes] for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.device = device

    def push(self, tensors) -> int:
        slot = self.i & 1
        for dst, src in zip(self.bufs[slot], tensors):
            dst.copy_(src, non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream(self.device))
        self.i += 1
        return slot

    def wait(self, slot: int):
        self.events[slot].synchronize()
        return self.bufs[slot]


class _CodecLane:
    """The codec's streaming step (generation.py:251) on its own CUDA stream: frame t is decoded to audio while
    the LM already computes frame t+1 on the caller's stream.  The frame kernel is launched on ``lm_ctas`` CTAs
    (a few SMs fewer than the GPU has) so the codec's small kernels always have somewhere to run.
    ``CSMB_DISABLE_OVERLAP=1`` keeps everything on the caller's stream."""

    def __init__(self, codec, device: torch.device):
        self.codec, self.device = codec, device
        self.overlap = os.environ.get("CSMB_DISABLE_OVERLAP", "0") != "1"
        self.main = torch.cuda.current_stream(device)
        self.stream = torch.cuda.Stream(device) if self.overlap else self.main
        self._ready = torch.cuda.Event()

    def step(self, frame: torch.Tensor) -> torch.Tensor:
        """frame (B, K) int32 produced on the caller's stream -> (B, 1, 1920) audio valid on ``self.stream``.
        Enter ``torch.cuda.stream(lane.stream)`` for any follow-up work on the result."""
        B, K = frame.shape
        if not self.overlap:
            return self.codec.step(frame.reshape(B, K, 1))
        self._ready.record(self.main)
        self.stream.wait_event(self._ready)
        frame.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            return self.codec.step(frame.reshape(B, K, 1))

    def join(self) -> None:
        """Order the caller's stream after everything enqueued on the lane (before buffers are released)."""
        if self.overlap:
            self.main.wait_stream(self.stream)


def generate_tokens(model: CSM, prompts: Sequence[Tuple[torch.Tensor, torch.Tensor]], max_audio_frames: int, *,
                    temperature: float = 0.8, sampler=None, logits_processors=None, seed: Optional[int] = None,
                    prefix_rows: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
    """The frame loop of generation.py:139-161 for B utterances in lock-step: returns, per utterance, the
    (F_b, 32) int32 CPU tensor of frames before its first all-zero (EOS) frame.  ``prefix_rows[b]``: leading rows of prompt b
    that are conversation context (KV prefix cache unit; 0 = none)."""
    for tok, _ in prompts:
        _check_length(model, int(tok.shape[0]), max_audio_frames)
    spec, host_sampler = _resolve_sampler(temperature, sampler, seed)
    sess = _Session(model, prompts, max_audio_frames, spec, host_sampler, logits_processors, prefix_rows)
    B, ncb = len(prompts), model.n_audio_codebooks
    mirror = _HostMirror([((B, ncb), torch.int32), ((1,), torch.int32)], model.device)
    out: List[List[torch.Tensor]] = [[] for _ in range(B)]
    done = [False] * B
    pending: Optional[int] = None

    def drain(slot: int) -> None:
        host, host_status = mirror.wait(slot)
        sess.raise_if_aborted(host_status)
        for b in range(B):
            if done[b]:
                continue
            if not bool(host[b].any()):
                done[b] = True  # eos (generation.py:151-152)
            else:
                out[b].append(host[b].clone())

    for _ in range(max_audio_frames):
        frame = sess.step()
        slot = mirror.push([frame, sess.status_word()])
        if pending is not None:
            drain(pending)
            if all(done):
                pending = None
                break
        pending = slot
    if pending is not None:
        drain(pending)
    sess.close()
    return [torch.stack(f) if f else torch.zeros((0, ncb), dtype=torch.int32) for f in out]


def generate(model: CSM, text: Union[str, Sequence[int]], speaker: int, context: List[Segment],
             max_audio_length_ms: float = 90_000, *, temperature: float = 0.8,
             logits_processors: Optional[List[LogitsProcessor]] = None, stream=None, sampler=None,
             seed: Optional[int] = None) -> torch.Tensor:
    """generation.py:95-178 -> 1-D float32 audio ``(1920*F,)`` (CPU tensor; ``np.asarray`` works on it)."""
    max_audio_frames = int(max_audio_length_ms / 80)
    tok, mask, n_context = _build_prompt_ex(model, text, speaker, context)
    (frames,) = generate_tokens(model, [(tok, mask)], max_audio_frames, temperature=temperature, sampler=sampler,
                                logits_processors=logits_processors, seed=seed, prefix_rows=[n_context])
    if frames.shape[0] == 0:
        print("[WARN] No samples generated.")
        return torch.zeros((0,), dtype=torch.float32)
    codes = frames.t().unsqueeze(0).to(model.device)  # (1, 32, F)
    audio = get_audio_tokenizer(model.n_audio_codebooks).decode(codes)
    # TODO(reference parity): the reference has an unimplemented watermarking TODO here (generation.py:176)
    return audio.reshape(-1).to("cpu")


def generate_batch(model: CSM, texts: Sequence[Union[str, Sequence[int]]], speakers: Sequence[int],
                   contexts: Optional[Sequence[List[Segment]]] = None, max_audio_length_ms: float = 90_000, *,
                   temperature: float = 0.8, sampler=None, seed: Optional[int] = None,
                   return_tokens: bool = False):
    """B independent utterances in lock-step (request batching; not in the reference).  Returns a list of
    1-D float32 CPU audio tensors (and the per-utterance (F,32) token tensors if ``return_tokens``)."""
    max_audio_frames = int(max_audio_length_ms / 80)
    contexts = contexts if contexts is not None else [[] for _ in texts]
    built = [_build_prompt_ex(model, t, s, c) for t, s, c in zip(texts, speakers, contexts)]
    prompts = [(b[0], b[1]) for b in built]
    frames = generate_tokens(model, prompts, max_audio_frames, temperature=temperature, sampler=sampler, seed=seed,
                             prefix_rows=[b[2] for b in built])
    mimi = get_audio_tokenizer(model.n_audio_codebooks)
    Fmax = max((int(f.shape[0]) for f in frames), default=0)
    audios: List[torch.Tensor] = []
    if Fmax == 0:
        audios = [torch.zeros((0,), dtype=torch.float32) for _ in frames]
    else:
        codes = torch.zeros((len(frames), model.n_audio_codebooks, Fmax), dtype=torch.int32)
        for b, f in enumerate(frames):
            codes[b, :, : f.shape[0]] = f.t()
        audio = mimi.decode(codes.to(model.device)).to("cpu")  # causal codec: padding frames only affect the tail
        audios = [audio[b, 0, : 1920 * int(f.shape[0])].clone() for b, f in enumerate(frames)]
    return (audios, frames) if return_tokens else audios


def stream_generate(model: CSM, text: Union[str, Sequence[int]], speaker: int, context: List[Segment],
                    max_audio_length_ms: float = 90_000, *, temperature: float = 0.8,
                    logits_processors: Optional[List[LogitsProcessor]] = None, stream=None, sampler=None,
                    seed: Optional[int] = None) -> Generator[torch.Tensor, None, None]:
    """generation.py:181-258: yields one ``(1920,)`` float32 CPU chunk per generated frame.  Each generator owns
    its codec streaming state (the reference shares one global Mimi state, tokenizers.py:14-21)."""
    max_audio_frames = int(max_audio_length_ms / 80)
    tok, mask, n_context = _build_prompt_ex(model, text, speaker, context)
    prompt = (tok, mask)
    _check_length(model, int(prompt[0].shape[0]), max_audio_frames)
    spec, host_sampler = _resolve_sampler(temperature, sampler, seed)
    sess = _Session(model, [prompt], max_audio_frames, spec, host_sampler, logits_processors, [n_context])
    ncb = model.n_audio_codebooks
    mimi = get_audio_tokenizer(ncb)
    lane = _CodecLane(mimi.acquire_decode_stream(batch=1), model.device)
    mirror = _HostMirror([((1, ncb), torch.int32), ((1, 1, 1920), torch.float32), ((1,), torch.int32)], model.device)
    pending: Optional[int] = None
    try:
        for _ in range(max_audio_frames):
            frame = sess.step()
            audio = lane.step(frame)
            with torch.cuda.stream(lane.stream):
                slot = mirror.push([frame, audio, sess.status_word()])
            if pending is not None:
                host_frame, host_audio, host_status = mirror.wait(pending)
                sess.raise_if_aborted(host_status)
                if not bool(host_frame.any()):
                    return  # eos: the speculative frame just enqueued is discarded
                yield host_audio.reshape(-1).clone()
            pending = slot
        if pending is not None:
            host_frame, host_audio, host_status = mirror.wait(pending)
            sess.raise_if_aborted(host_status)
            if bool(host_frame.any()):
                yield host_audio.reshape(-1).clone()
    finally:
        lane.join()
        mimi.release_decode_stream(lane.codec)
        sess.close()
