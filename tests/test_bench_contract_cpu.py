"""The reference arm of bench.py on the CPU: exactly one JSON line on stdout with the keys the driver reads (the GPU arm
needs a device; its line is built from the same dictionary layout)."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")   # what torch.distributed.run exports: the arm must not inherit it
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--no-extras"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "KG triples/sec (train step)" and d["unit"] == "triples/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["steps"] == 1 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert "cfg2_transe_ml1m" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"]
    assert cb["cores"] >= 1 and (cb["cores"] > 1 or (os.cpu_count() or 1) == 1), "the arm ran on one thread"
    assert d["e2e"] == {"value": d["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
