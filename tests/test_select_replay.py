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


def test_device_sincos_formula_matches_libm(tmp_path):
    """BRIEF steering needs (float)cos((double)a), (float)sin((double)a) exactly as the host libm gives them."""
    exe = tmp_path / "sincos_test"
    src = os.path.join(ROOT, "tests", "cpp", "sincos_test.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-o", str(exe), src], check=True)
    out = subprocess.run([str(exe), "16"], capture_output=True, text=True)
    assert out.returncode == 0 and "mismatches 0" in out.stdout, out.stdout
    # the kernel must carry the same constants as the checked formula
    cu = open(os.path.join(ROOT, "orb-slam3-rust_b200", "csrc", "orb_kernels.cuh")).read()
    for const in ("6.36619772367581382433e-01", "1.57079632679489655800e+00", "6.12323399573676603587e-17",
                  "1.58962301576546568060e-10", "-1.13596475577881948265e-11", "4.16666666666666019037e-2"):
        assert const in cu and const in open(src).read()
