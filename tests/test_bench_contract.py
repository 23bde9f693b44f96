"""CPU test of bench.py's output contract on the arm that needs no GPU (--impl reference)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-seconds", "1", "--config", "1"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MSamples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and d["vs_baseline"] is None
    # both thread rows of SURVEY.md 8d, the whole file timed once (config 1 fits), and a label
    # saying what the CPU figure times
    rows = d["cpu_baseline"]["rows"]
    assert len(rows) == 2 and rows[0]["threads"] >= rows[1]["threads"] >= 1
    assert "main.cp:75" in rows[1]["row"]
    assert d["cpu_baseline"]["whole_file"]["value"] > 0
    assert "decode + FIR + peak" in d["cpu_baseline"]["times"]
    # the config object is the one the GPU arm prints for the same arguments, key for key
    sys.path.insert(0, ROOT)
    from bench import CONFIGS, METRIC, config_dict

    assert d["config"] == config_dict(CONFIGS[1], 1, "block") and d["metric"] == METRIC


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_flop_count():
    sys.path.insert(0, ROOT)
    from bench import algorithmic_flop, kernel_order

    # whole file: 2*ch*(taps*frames - H(H+1))  
    taps, frames, ch = 9601, 2_880_000, 2
    H = (taps - 1) // 2
    assert algorithmic_flop(frames, ch, taps, 0, 0) == 2.0 * ch * (taps * frames - H * (H + 1))
    # an interior block with full halo does every tap
    assert algorithmic_flop(1000, 1, taps, H, H) == 2.0 * taps * 1000
    # shorter than the kernel
    assert algorithmic_flop(3, 1, 5, 0, 0) == 2.0 * (3 + 3 + 3)
    assert kernel_order(10.0 / 44100) + 1 == 17641


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_the_full_contract_line():
    """bench.py on the GPU (config 1, a few steps): exactly one JSON line with the base keys,
    the e2e object, the roofline of the FIR kernel and the CPU baseline."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--config", "1",
                        "--cpu-seconds", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "f64"
    assert d["gpu_launches"] == 3 * 3                     # decode + FIR + encode per step, all ours
    sys.path.insert(0, ROOT)
    from bench import CONFIGS, config_dict

    assert d["config"] == config_dict(CONFIGS[1], 1, "block")     # same object as the reference arm's
    p = d["parity"]
    assert p["ok"] is True and p["windows"] >= 4 and p["worst_d3"] <= 1e-12 and p["max_flip_lsb"] <= 1
    assert p["pcm_windows"] >= 2 * p["windows"] and p["global_peak_is_max_of_block_peaks"] is True
    assert d["e2e"]["pipelined"]["matches_serial_arm"] is True and d["e2e"]["copy_ceiling"]["d2h_gbs_per_rank"] > 1
    assert d["roofline"]["traffic_source"] and d["roofline"]["probe_clocks"]["sm_mhz"]
    c = d["cli"]                                           # the shipped C++ host on the same workload's file
    assert c["ok"] is True and c["payload_equals_library"] is True and c["metadata_identical"] is True
    assert 0 < c["wall_s"] < 30 and c["start_up_s"] is not None and c["exit_s"] is not None
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["matches_device_arm"] is True
    assert e["h2d_bytes_per_step"] == 2_880_000 * 6 and e["d2h_bytes_per_step"] == 2_880_000 * 6 + 8
    rf = d["roofline"]
    assert rf["kernel"] == "fir_dmma_kernel" and rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s"
    assert 0.8 < rf["frac"] < 1.05 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                         "sw_thermal_slowdown"}
