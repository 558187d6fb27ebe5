"""bench.py contract checks that need no GPU: the reference arm (the reference's own model on the host cores) prints one
JSON line with the keys the driver reads, and under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    r = _run({}, "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "8")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "meshes/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("train meshes/sec") and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["e2e"]["value"]) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "models.py"))  # the reference's own file, when staged
    assert cb["kind"] == ("reference" if staged else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "batch 8" in cb["sample"] and d["config"]["batch_per_gpu"] == 8  # the config states the batch that ran
    assert "semantichuman_b200" not in r.stderr  # the arm never loads the product package
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
