"""Oracle vs the committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from cv::ORB
through cv2 + the restated Rust loops) and vs the reference's own known-answer tests for this path."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_struct_equal

FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_golden(oracle, camera_dict, path):
    g = np.load(path)
    n = int(g["n_features"])
    kl, dl = oracle.detect_and_compute(g["left"], n)
    kr, dr = oracle.detect_and_compute(g["right"], n)
    assert_struct_equal(kl, g["kp_left"], "kp_left")
    assert_struct_equal(kr, g["kp_right"], "kp_right")
    assert np.array_equal(dl, g["desc_left"]) and np.array_equal(dr, g["desc_right"])
    m = oracle.match_features(kl, dl, kr, dr, camera_dict)
    assert_struct_equal(m, g["matches"], "matches")
    pts, valid = oracle.triangulate(kl, kr, m, camera_dict)
    assert np.array_equal(valid, g["points_valid"]) and np.array_equal(pts, g["points"])
    assert_struct_equal(oracle.bf_match_crosscheck(dl, dr), g["bf_crosscheck"], "bf_crosscheck")


def test_hamming_kat_from_reference(oracle):
    # /root/reference/src/vocabulary/mod.rs:430-441 and src/loop_closing/corrector.rs:626-633
    a = np.zeros(32, np.uint8)
    c = np.zeros(32, np.uint8)
    assert oracle.descriptor_distance(a, a) == 0
    c[0] = 0xFF
    assert oracle.descriptor_distance(a, c) == 8
    c[1] = 0x0F
    assert oracle.descriptor_distance(a, c) == 12


def test_descriptor_distance_uses_min_cols(oracle):
    # stereo.rs:168: cols = desc1.cols().min(desc2.cols())
    a = np.full(32, 0xFF, np.uint8)
    b = np.zeros(16, np.uint8)
    assert oracle.descriptor_distance(a, b) == 128
