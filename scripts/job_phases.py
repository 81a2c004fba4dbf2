"""Developer timing (run under gpurun): where one configs[3] job (64 utterances x 125 frames through the engine + Mimi decode)
spends its time on one GPU: the admitting step (prompt pass of 64 prompts + first frame-step), the 125 other engine steps, the
batched Mimi decode.  CUDA events on the engine's stream.  Usage: python scripts/job_phases.py [n_requests]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from csm_mlx_b200 import CSM, csm_1b, serving, tokenizers
from csm_mlx_b200.mimi import Mimi
from csm_mlx_b200.random_init import random_csm_weights, random_mimi_weights

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
model = CSM(csm_1b(), device=dev).load_weights(random_csm_weights())
tokenizers.set_audio_tokenizer(Mimi(32, device=dev).load_pytorch_weights(random_mimi_weights()))
frames = 125
prompts = [bench.cfg4_prompt(i) for i in range(N)]
eng = serving.Engine(model, max_batch=N, max_len=32 + frames + 2)


def job():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize(dev)
    ev[0].record()
    rids = [eng.submit_prompt(p[0], p[1], frames) for p in prompts]
    eng.step()
    ev[1].record()
    steps0 = eng.steps
    eng.run()
    ev[2].record()
    eng.audio(rids, to_host=False)
    ev[3].record()
    torch.cuda.synchronize(dev)
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], eng.steps - steps0


for _ in range(2):
    job()
for _ in range(3):
    t, n = job()
    print(f"{N} requests: admitting step {t[0]:.2f} ms | {n} further steps {t[1]:.2f} ms ({t[1] / n:.3f} ms each) | Mimi decode "
          f"{t[2]:.2f} ms | job {sum(t):.2f} ms", flush=True)
