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

PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, golden vectors or fixtures
 and its dependencies (mlx, mlx_lm, moshi_mlx) cannot be installed here,
so the reference itself cannot be run to produce outputs.  The oracle is instead pinned
against two independent implementations that are importable in this container
(HF ``transformers`` ``MimiModel`` and ``CsmForConditionalGeneration``); see
``tests/test_oracle_vs_hf.py`` and ``scripts/make_golden.py``.
"""
