"""bench.py contract pieces that need no GPU: the algorithmic FLOP count (SURVEY.md 8d / BASELINE.md section 3) and the
`--impl reference` arm (CPU restatement of the reference graph): ONE JSON line on stdout with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey():
    assert bench.dense_flops_per_sample(bench.ref_archs()) == 7744400          # fwd 2 862 400 + bwd 4 882 000
    assert bench.dense_flops_per_sample(bench.scaled_archs()) == 124448768     # 2048-wide layers, n_z 64
    conv = bench.conv_archs()
    assert conv[0]["hidden_conv"] and not conv[1]["hidden_conv"] and conv[0]["n_hidden_recog_1"] == 16


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--batch", "128"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "paired samples/sec/train step" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "not TensorFlow" in d["reference_note"]
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1            # the arm honours --steps / --warmup / --gpus
    # `config` is built by ONE function for both arms (the driver compares them)
    import argparse
    a = argparse.Namespace(batch=128, config="ref")
    assert d["config"] == bench.make_config(a, bench.ref_archs(), bench.CONFIGS["ref"][2], 1)


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
