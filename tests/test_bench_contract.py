"""bench.py's output contract (the task statement's JSON line) without a GPU: the CPU reference arm runs here for real;
the GPU arm's line is checked on the newest committed capture under profiles/."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "gan_train_frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_line_has_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_bench_cfg2_v*.json")),
                   key=lambda p: int(p.rsplit("_v", 1)[1].split(".")[0]))
    assert files
    lines = [l for l in open(files[-1]) if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS | {"gpu_launches", "roofline", "cpu_baseline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 128 * 100 * (257 + 40) * 4 + 128 * 4 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.02
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_cpu_sample_is_the_whole_minibatch_when_the_host_is_fast_enough():
    """bench.cpu_sample_utterances: the CPU arm runs the config's whole minibatch when one schedule of it fits ~4 s (probed on
    ~3200 frames), else a bounded sample, never fewer utterances than the probe."""
    sys.path.insert(0, ROOT)
    import bench

    class Fake(object):
        def __init__(self, rate):
            self.rate, self.calls = rate, []

        def time_schedule(self, cfg, B, T, steps=1, warmup=0):
            self.calls.append((B, T))
            return self.rate, B * T / self.rate, 16
    cfg = dict(B=128, T=100)
    fast, slow, mid = Fake(7000.0), Fake(500.0), Fake(2000.0)
    assert bench.cpu_sample_utterances(cfg, fast) == 128 and fast.calls == [(32, 100)]
    assert bench.cpu_sample_utterances(cfg, slow) == 32
    assert bench.cpu_sample_utterances(cfg, mid) == 80
    assert bench.cpu_sample_utterances(dict(B=8, T=100), Fake(1.0)) == 8            # already the whole minibatch: no probe
    assert bench.sample_text(128, cfg).startswith("the whole minibatch") and bench.sample_text(32, cfg).startswith("32 of 128")
