"""CPU oracle for the CSM speech-token generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``csm_mlx_b200/`` (the product) may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker / reported CPU baseline.

What it is: a plain PyTorch-CPU fp32 restatement of the reference path

* ``/root/reference/csm_mlx/generation.py:21-258`` (frame, generate, stream_generate)
* ``/root/reference/csm_mlx/models.py:31-92`` (CSM parameter tree, embed_tokens/embed_audio)
* ``/root/reference/csm_mlx/attention.py:10-253`` (Llama-3 scaled RoPE, GQA attention)
* ``/root/reference/csm_mlx/config.py:3-45`` (hyper-parameters)
* ``/root/reference/csm_mlx/tokenizers.py:43-102`` (frame assembly)

plus restatements of the un-vendored third-party pieces the reference calls:
``mlx_lm.models.llama.LlamaModel`` (pyproject pin mlx-lm>=0.22.0), ``mlx_lm.sample_utils``
and ``moshi_mlx.models.mimi.Mimi`` (pin moshi-mlx>=0.2.3, "mimi_202407" architecture).

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures and its dependencies
(mlx, mlx_lm, moshi_mlx) cannot be installed here, so the reference cannot be run as shipped.  What is done instead:

* reference-owned code — ``attention.py`` (Llama-3 scaled RoPE, attention), ``models.py`` (parameter tree, embeddings),
  ``generation.py`` (``generate_frame``, the ``generate`` driver), ``tokenizers.py`` / ``segment.py`` (frame assembly):
  the reference's OWN modules are imported unmodified from ``/root/reference`` and executed over a small stand-in of
  the mlx API (``oracle/mlx_shim.py``, torch CPU fp32) by ``scripts/make_reference_golden.py``; the outputs are
  committed as ``tests/golden/reference_cfg1.npz`` and ``tests/test_reference_golden.py`` checks the oracle (CPU) and
  the product (GPU) against them: 25 greedy frames of BASELINE configs[0] token-exact, RoPE tables bit-identical,
  prompt rows / context rows identical, hidden states and logits within 2e-5 (oracle) / 1e-4 (product).
* third-party arithmetic the reference calls but does not contain (mlx_lm's Llama block, KV cache and attention
  primitive; moshi_mlx's Mimi): restated from the published algorithms and pinned against the two independent
  implementations importable in this container (HF ``transformers`` ``MimiModel`` and ``CsmForConditionalGeneration``);
  see ``tests/test_oracle_vs_hf.py`` and ``scripts/make_golden.py``.  MLX's own bit patterns (its bf16 kernels, its
  ``mx.random.categorical`` stream) remain unpinned: they cannot be produced without MLX.
"""
