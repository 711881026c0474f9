"""The reference arm of bench.py runs on the CPU (the unmodified reference staged under baseline/_ref by
oracle/install_ref.py, or the oracle port when it is not staged): check the JSON contract the driver reads (one line,
the tier's keys) without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         check=True, capture_output=True, text=True, timeout=600, cwd=ROOT).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "train_audio_seconds_per_second" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["n_gpus"] == 1
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "trainer", "trainer.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_committed_traffic_matches_launch_list(tmp_path):
    """`roofline.traffic` comes from profiles/r02_traffic.json; that file must be what tools/launch_traffic.py derives
    from the committed ncu launch list (no hand-edited numbers)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "traffic.json"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "launch_traffic.py"),
                        os.path.join(root, "profiles", "r02_ncu_launches.csv"), "--json", str(out)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "one step:" in r.stdout
    got, want = json.load(open(out)), json.load(open(os.path.join(root, "profiles", "r02_traffic.json")))
    assert set(got) == set(want)
    for k, v in want.items():
        if k != "source":
            assert got[k] == v, k
