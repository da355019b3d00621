"""bench.py's reference arm runs on the host cores alone (no GPU): its JSON line must keep the contract the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("binaural stream-sec/sec") and d["unit"] == "stream-s/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("C2:") and d["config"]["block"] == 256 and d["config"]["partitions"] == 17
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "streams x" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "stream-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_a_non_zero_rank_exits_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_match_the_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    # bytes per stream per block
    assert bench.algorithmic_bytes(8, 256, 17) == 288768          # C2
    assert bench.algorithmic_bytes(8, 512, 128) == 4214784        # C3
    assert bench.algorithmic_bytes(8, 256, 19, 10) == 322176      # C4, EQ state read + written included
