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


def test_rotation_roofline_picks_the_dominant_rotation_kernel():
    """bench.rotation_roofline: SURVEY 8(d)'s algorithmic figure for whichever rotation kernel took more device time."""
    sys.path.insert(0, ROOT)
    import bench

    prof = {
        "rotate": {"launches": 10, "passes": 10, "ms": 2.0, "term_updates": 1_000_000},
        "segment": {"launches": 8, "passes": 200, "ms": 6.0, "term_updates": 30_000_000},
    }
    r = bench.rotation_roofline(prof, 96, 6531.0)
    assert r["kernel"] == "k_seg_main" and r["bound"] == "hbm" and r["unit"] == "GB/s" and r["traffic"] is None
    assert abs(r["achieved"] - 30_000_000 * 96 / 6.0e-3 / 1e9) < 1e-9
    assert abs(r["frac"] - r["achieved"] / 6531.0) < 1e-12
    assert r["gates_per_launch"] == 100.0 and "not DRAM bytes" in r["note"]
    prof["segment"]["ms"] = 1.0
    r = bench.rotation_roofline(prof, 96, 6531.0)
    assert r["kernel"] == "k_rotate" and "note" not in r
    assert bench.rotation_roofline({}, 96, 6531.0)["achieved"] is None
