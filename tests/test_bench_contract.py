"""bench.py's output contract on the CPU arm (`--impl reference`: the oracle port timed on the host cores): stdout is
exactly ONE JSON line carrying the keys the driver reads; everything else (library banners, progress) goes to stderr."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_exactly_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-frames", "2"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines[:3]
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "audio_seconds_per_second" and rec["unit"] == "audio-s/s"
    assert rec["higher_is_better"] is True and rec["n_gpus"] == 1 and rec["steps"] == 1 and rec["warmup"] == 0
    assert rec["value"] > 0 and rec["ms_per_step"] > 0 and rec["vs_baseline"] is None
    assert rec["config"]["workload"].startswith("configs[3]") and rec["scaling"] == "strong"
    assert rec["config"]["frames_per_utterance"] == 2
    cb = rec["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == rec["value"] and cb["sample"]
    e2e = rec["e2e"]
    assert e2e["value"] == rec["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_under_torchrun_prints_on_rank0_only():
    """N > 1: the driver launches the reference arm through torchrun like the CUDA arm; rank 0 alone runs and prints, the
    other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--ref-frames", "2"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines[:3]
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["value"] > 0 and rec["cpu_baseline"]["kind"] == "port"
