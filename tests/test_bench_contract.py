"""bench.py's reference arm runs without a GPU (it times the CPU oracle port): check, here on CPU, that it
prints exactly ONE JSON line on stdout with the keys the driver's contract names."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_contract_line():
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1",
                           "--warmup", "0", "--gpus", "1", "--no-extra"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pixels/sec" and d["unit"] == "px/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["config"]["workload"].startswith("BASELINE config 2")
    cb = d["cpu_baseline"]
    # "reference": the reference's own functions from baseline/_ref (tools/install_reference.py); "port": the oracle
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == "px/s"
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_other_ranks_of_the_reference_arm_exit_quietly():
    import os

    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1",
                           "--warmup", "0", "--gpus", "2", "--no-extra"], capture_output=True, text=True, timeout=300, cwd=str(ROOT),
                          env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""
