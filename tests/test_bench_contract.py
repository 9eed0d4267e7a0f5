"""bench.py's reference arm runs without a GPU (it times the oracle port on the host cores): check the JSON contract of
its line here, and that under a multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *args],
                          capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)


def test_reference_arm_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["metric"].startswith("rotations/sec") and line["unit"] == "rotations/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    assert line["scaling"] == "weak" and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert "workload" in line["config"] and "model" not in line["config"]
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["sample"] and cpu["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29561"}, ("--gpus", "2"))
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
