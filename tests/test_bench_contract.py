"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the CPU
implementation the reference calls, timed on the host cores) prints ONE JSON line with the keys the driver reads and with
the same metric / unit / config as the CUDA arm, and the product arm refuses to run without a CUDA device (no fallback)."""
import argparse
import importlib.util
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, _has_gpu


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_config_is_a_pure_function_of_the_command_line():
    b = load_bench()
    a = argparse.Namespace(total_frames=0, frames=2000, unique=0, batch=64, seq_frames=240)
    for world in (1, 2, 8):
        c = b.euroc_config(a, world)
        assert c == b.euroc_config(a, world)
        assert c["frames_per_gpu_per_step"] == 2000 and c["frames_all_gpus_per_step"] == 2000 * world
        assert "configs[1]" in c["workload"] and "752x480" in c["workload"]
    strong = argparse.Namespace(total_frames=20000, frames=2000, unique=0, batch=64, seq_frames=240)
    c = b.euroc_config(strong, 8)
    assert c["frames_all_gpus_per_step"] == 20000 and c["frames_per_gpu_per_step"] == 2500 and "configs[4]" in c["workload"]
    assert b.tracking_config(a, 1) == b.tracking_config(a, 1) and "configs[3]" in b.tracking_config(a, 1)["workload"]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    b = load_bench()
    assert d["impl"] == "reference" and d["metric"] == b.METRIC and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    a = argparse.Namespace(total_frames=0, frames=2000, unique=0, batch=64, seq_frames=240)
    assert d["config"] == b.euroc_config(a, 1), "both arms must describe the same workload"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a CUDA device")
def test_product_arm_fails_loudly_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--frames", "8",
                          "--no-cpu-baseline", "--parity-frames", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0, "bench.py must not produce a number without the CUDA path"
    assert not [l for l in out.stdout.splitlines() if l.startswith("{") and '"value"' in l]
