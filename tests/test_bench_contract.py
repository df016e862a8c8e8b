"""bench.py contract that can be checked without a GPU: the reference arm (CPU port of the
reference's algorithm) prints exactly ONE line on stdout, a JSON object with the keys the driver
reads; under torchrun only rank 0 prints."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                           "rgbnt201", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run({"OMP_NUM_THREADS": "1"})       # what torchrun exports to every rank
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "reid_eval_queries_per_sec" and j["unit"] == "queries/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == j["value"] and cb["cores"] == (os.cpu_count() or 1)
    assert cb["blas_threads"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
