"""bench.py's contract that can be checked without a GPU: the reference arm (the reference's
CPU path = OpenCV's BFMatcher + the restated glue) prints exactly one JSON line with the agreed
keys, and the B200 arm refuses to run on the CPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = run("--impl", "reference", "--steps", "2", "--warmup", "1", "--features", "800", "--window", "3")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hamming_comparisons_per_sec" and d["unit"] == "cmp/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["features_per_frame"] == 800 and d["config"]["window"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank_zero_works_under_torchrun():
    p = run("--impl", "reference", "--steps", "1", "--warmup", "1", "--features", "500", "--window", "2",
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
