"""The reference arm of bench.py runs on the CPU (the unmodified reference staged in oracle/_ref, else its dense restatement) and must print ONE JSON
line with the keys the driver reads; rank != 0 under a multi-rank launch prints nothing and exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--ref-cells", "2", "--no-ref-sizes"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "atoms/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("descriptor+energy+force atoms/s")
    assert d["config"]["atoms"] == 32 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    # "reference" = the unmodified package staged in oracle/_ref by __graft_entry__.build(); "port" = the dense restatement
    staged = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "nnmd", "features"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
