"""bench.py's reference arm runs on the CPU: it must print one JSON line carrying the keys the
driver's contract names (metric/value/unit/config/cpu_baseline/e2e, impl = reference)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT,
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def test_reference_arm_prints_contract_line():
    lines = [l for l in _run("--impl", "reference", "--steps", "100", "--warmup", "3").splitlines()
             if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_s" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["workload"].startswith("C1")
    cb = d["cpu_baseline"]
    # the UNMODIFIED reference classes when a reference tree is present (authoring container:
    # /root/reference; GPU box: the git-ignored install in baseline/_ref), else the oracle port
    assert cb["kind"] in ("port", "reference-classes+pf-port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["port"]["kind"] == "port" and d["port"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert _run("--impl", "reference", "--gpus", "2", "--steps", "10", "--warmup", "3", env=env).strip() == ""
