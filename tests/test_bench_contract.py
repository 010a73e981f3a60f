"""bench.py contract on CPU: the reference arm runs without a GPU and prints one JSON line with the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_json():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env={**os.environ, "FH_BENCH_REF_OPS": "2"})       # 2-operator chunk: seconds, not 19 GB
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "gradients/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def test_bench_without_gpu_fails_loudly():
    """The product arm must not fall back to the CPU: without a CUDA device it exits non-zero."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline",
                          "--no-hbm-regime"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith('{"metric"')]


def test_cpu_closed_form_leg_prints_both_thread_counts():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "cpu-closed-form"], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "cpu-closed-form" and d["1thread"]["threads"] == 1 and d["allcores"]["threads"] == d["cores"]
    assert d["1thread"]["value"] > 0 and d["allcores"]["value"] > 0
    assert abs(d["energy"] - 0.1938641953852845) < 1e-10          # the bench workload's energy (GPU, oracle)
