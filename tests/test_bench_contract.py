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
