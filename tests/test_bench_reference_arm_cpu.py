"""CPU: the contract of `bench.py --impl reference` (it needs no GPU): one JSON line on stdout with the driver's keys, `impl = reference`, a
`cpu_baseline` that says what ran and on how many threads, an `e2e` object repeating the value with zero copies."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c2", "--steps", "2", "--warmup", "1", "--width", "96", "--height", "54"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mray/s" and d["unit"] == "Mray/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C2") and d["config"]["width"] == 96
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "rows" in cb["sample"]
    if cb["kind"] == "reference":                        # the compiled reference integrator ran: the oracle port is reported beside it
        assert "rp_main.rgen" in cb["sample"] and cb["port"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mray/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
