"""bench.py contract checks that need no GPU: the reference arm runs on host cores and prints ONE JSON line with the
keys the driver reads; the native arm's workload table matches BASELINE.json's configurations."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")     # what torchrun exports to its ranks; the arm must override it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config4",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["value"] > 0
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert cb["cores"] == (os.cpu_count() or 1) or "NumPy" in cb["sample"]     # OMP_NUM_THREADS=1 was overridden
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workloads_match_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert bench.WORKLOADS["config2"][:3] == (32, 512, 512) and "batch 32 at 512x512" in cfg[1]
    assert bench.WORKLOADS["config3"][:3] == (16, 1024, 1024) and "batch 16 at 1024x1024" in cfg[2]
    assert bench.WORKLOADS["config4"][:3] == (8, 512, 512) and "batch 8 at 512x512" in cfg[3]
    assert bench.WORKLOADS["config5"][1:3] == (2160, 3840) and "3840x2160" in cfg[4]
    # algorithmic bytes per pixel: 12 in + 4 bytes per output channel
    assert bench.WORKLOADS["config2"][3] == 12 + 4 * 84 and bench.WORKLOADS["config4"][3] == 12 + 4 * 93
    assert bench.WORKLOADS["config3"][3] == 24 and bench.WORKLOADS["config5"][3] == 12 + 4 * 93 + 24


def test_both_arms_print_the_same_config():
    """the driver compares the two arms' `config` dicts (same_config): both come from bench.config_dict"""
    sys.path.insert(0, ROOT)
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_dict(wl, args.gpus)') == 1 and src.count('"config": config_dict(wl, world)') == 1
    c = bench.config_dict("config2", 1)
    assert c["name"] == "config2" and c["batch_per_gpu"] == 32 and c["h"] == 512 and c["w"] == 512
    assert "no flush needed" in c["l2"] and "flushed" in bench.config_dict("config1", 1)["l2"]
    assert bench.WORKLOADS["config1"][:3] == (1, 256, 256)
