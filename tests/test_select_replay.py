"""The exact replay of cv::KeyPointsFilter::retainBest (libstdc++ std::nth_element + std::partition) used on the GPU
is checked on the host against the real standard library: serial pieces (select_serial.h: median-of-three,
insertion sort, heap-select fallback) and the lane-parallel formulation of the Hoare pass."""
import os
import subprocess

from conftest import ROOT


def test_select_replay_matches_libstdcxx(tmp_path):
    exe = tmp_path / "select_replay_test"
    src = os.path.join(ROOT, "tests", "cpp", "select_replay_test.cpp")
    inc = os.path.join(ROOT, "orb-slam3-rust_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "MISMATCH" not in out.stdout
