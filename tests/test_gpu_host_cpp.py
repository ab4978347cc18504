"""The compiled C++ host mirror of the reference interface (orb-slam3-rust_b200/host/stereo_processor.hpp) drives the
C ABI end to end; its output checksums must equal the oracle's."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HOST = os.path.join(ROOT, "orb-slam3-rust_b200", "host")


def fnv1a(data: bytes, h=1469598103934665603):
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_cpp_host_mirror_end_to_end(tmp_path, oracle, pairs, camera_dict):
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    l, r = pairs(0)
    (tmp_path / "l.raw").write_bytes(l.tobytes())
    (tmp_path / "r.raw").write_bytes(r.tobytes())
    out = subprocess.run([os.path.join(HOST, "host_demo"), str(tmp_path / "l.raw"), str(tmp_path / "r.raw"), "752", "480"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = dict(re.findall(r"(\w+)=(\w+)", out.stdout))
    kl, dl = oracle.detect_and_compute(l, 1200)
    kr, dr = oracle.detect_and_compute(r, 1200)
    m = oracle.match_features(kl, dl, kr, dr, camera_dict)
    _, valid = oracle.triangulate(kl, kr, m, camera_dict)
    cs, items = oracle.grid_build(kl)
    assert int(got["n_left"]) == len(kl) and int(got["n_right"]) == len(kr) and int(got["n_matches"]) == len(m)
    assert int(got["n_points"]) == int(valid.sum()) and int(got["ts"]) == 42
    assert int(got["kp"], 16) == fnv1a(kr.tobytes(), fnv1a(kl.tobytes()))
    assert int(got["desc"], 16) == fnv1a(dr.tobytes(), fnv1a(dl.tobytes()))
    assert int(got["matches"], 16) == fnv1a(m.tobytes())
    assert int(got["area"]) == len(oracle.grid_query(cs, items, 376.0, 240.0, 15.0))
    # OrbVocabulary mirror: same tiny k = 5, L = 2 tree as host_demo.cpp builds (LCG descriptor bytes)
    lcg, nd = 12345, np.zeros((30, 32), np.uint8)
    for i in range(30):
        for b in range(32):
            lcg = (lcg * 1664525 + 1013904223) & 0xFFFFFFFF
            nd[i, b] = lcg >> 24
    parents = [0] * 5 + [(i - 6) // 5 + 1 for i in range(6, 31)]
    voc = oracle.make_vocabulary_arrays(parents, [0] * 5 + [1] * 25, nd)
    word, leaf, node = oracle.bow_quantize(dl, voc, 1)
    groups = {}
    for i, n in enumerate(node.tolist()):
        groups.setdefault(n, []).append(i)
    h = 1469598103934665603
    for n in sorted(groups):
        for i in groups[n]:
            h = fnv1a(np.array([n, i], np.uint32).tobytes(), h)
    assert int(got["bow_words"]) == len(set(word.tolist())) and int(got["bow_groups"]) == len(groups)
    assert int(got["bow"], 16) == h
    # resident-frame local-map search and search_for_triangulation through the C++ mirror
    pts, valid = oracle.triangulate(kl, kr, m, camera_dict)
    good = np.flatnonzero(valid)
    uv = np.c_[kl["x"][good].astype(np.float64) + 3.0, kl["y"][good].astype(np.float64) - 2.0]
    oi, _ = oracle.search_by_projection(kl, dl, uv, dl[good], 0)
    assert int(got["found"]) == int((oi >= 0).sum()) and int(got["found"]) > 300
    assert int(got["search"], 16) == fnv1a(oi.astype(np.int32).tobytes())
    cam = camera_dict
    Kinv = np.array([[1 / cam["fx"], 0, -cam["cx"] / cam["fx"]], [0, 1 / cam["fy"], -cam["cy"] / cam["fy"]], [0, 0, 1]])
    tx = np.array([[0, 0, 0], [0, 0, -cam["baseline"]], [0, cam["baseline"], 0]])
    tmp = np.zeros((3, 3))
    F = np.zeros((3, 3))
    for i in range(3):  # the same left-to-right accumulation as host_demo.cpp, so F is bit-identical
        for j in range(3):
            for k in range(3):
                tmp[i, j] += tx[i, k] * Kinv[k, j]
    for i in range(3):
        for j in range(3):
            for k in range(3):
                F[i, j] += Kinv[k, i] * tmp[k, j]
    P = oracle.tri_params(F, (1e9, cam["cy"]), 50, cam)
    ref = oracle.search_for_triangulation(kl, dl, np.zeros(len(kl), np.uint8), kr, dr, np.zeros(len(kr), np.uint8), P)
    assert int(got["n_tri"]) == len(ref) and len(ref) > 300
    assert int(got["tri"], 16) == fnv1a(ref.astype(np.uint32).tobytes())
