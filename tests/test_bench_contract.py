"""The bench contract on a machine without a GPU: the reference arm prints one well-formed JSON line, and the
engine arm refuses to run (no CPU fallback) instead of measuring something else."""

from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args: str) -> subprocess.CompletedProcess:
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-updates", "20000", "--no-full-run")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "pauli_term_gate_updates_per_s" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_engine_arm_has_no_cpu_fallback():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the engine arm runs")
    except ImportError:
        pass
    out = _run("--steps", "1", "--warmup", "0")
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not any(ln.startswith("{") for ln in out.stdout.splitlines())
