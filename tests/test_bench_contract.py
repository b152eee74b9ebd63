"""bench.py's JSON contract on the CPU box: the reference arm (`--impl reference`) runs here (it times the compiled
reference / the oracle on the host cores) and must print one line with the agreed keys and types."""
import json
import os
import subprocess
import sys

import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_line():
    if O.ref() is None:
        pytest.skip("no compiled reference here")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=580)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "output Msamples/s" and line["unit"] == "Msamples/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert isinstance(line["value"], float) and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and isinstance(cb["cores"], int) and cb["cores"] >= 1
    assert cb["value"] == line["value"] and isinstance(cb["sample"], str) and cb["sample"]
    e = line["e2e"]
    assert e == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg2")
    # the config object is exactly the GPU arm's (the driver compares them); remarks about the arm live beside it
    assert set(line["config"]) == {"workload", "channels_per_gpu", "samples_per_channel", "ratio", "taps", "nco_mix", "sharding", "l2"}
    assert "bounded sample" in line["note"]
    assert line["vs_baseline"] is None and line["data"] == "synthetic"


def test_cfg1_cpu_baseline_variants():
    """SURVEY.md 8(d) CPU baselines (A) reference makefile flags and (B) -O2, one thread, on BASELINE configs[0] exactly."""
    if O.ref() is None or O.ref("O0") is None:
        pytest.skip("no compiled reference here")
    sys.path.insert(0, ROOT)
    import bench
    v = bench.cpu_cfg1_variants(bench.WORKLOADS["cfg1"])
    assert set(v) == {"makefile_build", "O2"}
    for k in v.values():
        assert k["cores"] == 1 and k["kind"] == "reference" and k["value"] > 0 and "cfg1 exactly" in k["sample"]
    assert v["O2"]["value"] > v["makefile_build"]["value"]  # the optimised build is the faster one


def test_clock_sampler_samples_densely_even_when_the_power_query_is_slow(monkeypatch):
    """bench.ClockSampler against a stand-in NVML whose power query takes 20 ms (as on the B200 boxes): the clock /
    throttle-reason samples of a 60 ms region must not be held back by it."""
    import time
    import types
    fake = types.ModuleType("pynvml")
    fake.NVML_CLOCK_SM = 1
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByUUID = lambda u: "h"
    fake.nvmlDeviceGetHandleByIndex = lambda i: "h"
    fake.nvmlDeviceGetMaxClockInfo = lambda h, k: 1965
    fake.nvmlDeviceGetClockInfo = lambda h, k: 1900
    fake.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4

    def power(h):
        time.sleep(0.02)
        return 500000
    fake.nvmlDeviceGetPowerUsage = power
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    assert s.nvml is fake
    s.start()
    time.sleep(0.06)
    c = s.stop()
    assert c["samples"] >= 10 and c["sm_mhz"] == 1900.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["power_w"] == 500.0 and c["source"].startswith("nvml")
