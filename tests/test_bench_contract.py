"""bench.py contract checks that need no GPU: the reference arm's JSON line (the driver parses it next to ours) and
the algorithmic-byte formulas of SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes(5, 3, 40) == {"solve": 15568, "assemble": 29800, "unfused": 45368}
    assert bench.algorithmic_bytes(5, 3, 7)["solve"] == 3688 and bench.algorithmic_bytes(5, 3, 7)["assemble"] == 5512
    assert bench.algorithmic_bytes(10, 2, 9)["solve"] == 8128 and bench.algorithmic_bytes(10, 2, 9)["assemble"] == 13824


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "agent_qp_solves_per_sec" and line["unit"] == "QP/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["value"] > 0 and line["steps"] == 1 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
