"""bench.py contract checks that need no GPU: the reference arm runs the CPU oracle on the bench's own
workload generator (tiny sizes here) and prints ONE JSON line with the keys the driver reads; the
`config` object is built by the same function for both arms."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_line():
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--side", "10", "--nodes", "24",
           "--gn-iters", "2", "--pcg-iters", "5", "--steps", "2", "--warmup", "1", "--ref-budget-s", "60"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "deformation_gn_iters_per_s"
    assert d["unit"] == "GN iters/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["steps_requested"] == 2  # the solves actually timed
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-12
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = d["config"]
    assert cfg["verts"] == 6 * 10 * 10 + 2 and cfg["nodes"] == 24 and cfg["gn_iters"] == 2
    assert "workload" in cfg and "model" not in cfg


def test_config_object_is_shared_by_both_arms():
    sys.path.insert(0, str(ROOT))
    import bench
    src = (ROOT / "bench.py").read_text()
    # both arms call the one builder; neither writes a literal config dict for the headline line
    assert src.count("base_config(args, ") >= 2
    import argparse
    a = argparse.Namespace(side=409, nodes=4096, gn_iters=10, pcg_iters=20, mode="objects")
    c = bench.base_config(a, 1003688, 1058402)
    assert c["verts"] == 1003688 and c["target_points"] == 1058402 and c["mode"] == "objects"
    assert "1003688-vertex" in c["workload"]


def test_clock_sampler_polls_nvml_and_keeps_the_timed_region(monkeypatch):
    """bench.Clocks: NVML polling thread, samples filtered to the timed region, throttle-reason bits
    decoded; without any NVML / nvidia-smi it reports itself unavailable instead of failing."""
    import sys
    import time
    import types
    import bench
    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: ("gpu", i)
    fake.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, k: 1950
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4 | 0x1  # sw_power_cap + gpu_idle
    fake.nvmlDeviceGetPowerUsage = lambda h: 250000
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    c = bench.Clocks(0)
    c.start()
    time.sleep(0.05)
    t0 = time.perf_counter()
    time.sleep(0.06)
    t1 = time.perf_counter()
    time.sleep(0.03)
    c.window(t0, t1)
    r = c.stop()
    assert r["source"] == "nvml" and r["sm_mhz"] == 1950.0 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"] and r["power_w_max"] == 250.0
    assert 3 <= r["in_timed_region"] == r["samples"] <= 20  # ~5 ms period over a 60 ms window
    # no NVML, no nvidia-smi
    broken = types.ModuleType("pynvml")

    def boom():
        raise RuntimeError("no driver")
    broken.nvmlInit = boom
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    monkeypatch.setenv("PATH", "/nonexistent")
    c2 = bench.Clocks(0)
    c2.start()
    r2 = c2.stop()
    assert r2["sm_mhz"] is None and r2["reasons"] == ["nvidia-smi unavailable"]
