"""Pins the extraction oracle: the C++ restatement (oracle/orb_oracle.cpp) must equal cv::ORB — the function the
reference calls at src/tracking/frame/stereo.rs:72-73 — bit for bit, whole pipeline and stage by stage."""
import numpy as np
import pytest

from conftest import assert_struct_equal


@pytest.mark.parametrize("seed,w,h,n", [(0, 752, 480, 1200), (1, 752, 480, 1200), (2, 752, 480, 1200),
                                         (3, 1280, 720, 2000), (4, 640, 480, 300)])
def test_detect_and_compute_equals_cv2(oracle, cv2mod, pairs, seed, w, h, n):
    from oracle import cv2_ref

    for img in pairs(seed, w, h):
        k0, d0 = cv2_ref.detect_and_compute(img, n)
        k1, d1 = oracle.detect_and_compute(img, n)
        assert_struct_equal(k1, k0, "keypoints")
        assert np.array_equal(d1, d0)


def test_stage_composition_equals_single_call(oracle, pairs):
    img = pairs(5)[0]
    st = oracle.detect_stages(img, 1200)
    k, d = oracle.detect_and_compute(img, 1200)
    assert_struct_equal(st["keypoints"], k, "keypoints")
    assert np.array_equal(st["descriptors"], d)


def test_level_geometry(oracle):
    lv = oracle.level_geometry(752, 480, 1200)
    assert [(int(l["w"]), int(l["h"])) for l in lv] == [(752, 480), (627, 400), (522, 333), (435, 278), (363, 231),
                                                        (302, 193), (252, 161), (210, 134)]
    assert [int(l["nfeatures"]) for l in lv] == [261, 217, 181, 151, 126, 105, 87, 72]
    lv = oracle.level_geometry(1280, 720, 2000)
    assert [(int(l["w"]), int(l["h"])) for l in lv] == [(1280, 720), (1067, 600), (889, 500), (741, 417), (617, 347),
                                                        (514, 289), (429, 241), (357, 201)]
    assert [int(l["nfeatures"]) for l in lv] == [434, 362, 302, 251, 209, 175, 145, 122]


def test_resize_equals_cv2(oracle, cv2mod, pairs):
    img = pairs(0)[0]
    lv = oracle.level_geometry(752, 480, 1200)
    prev = img
    for l in range(1, 8):
        w, h = int(lv[l]["w"]), int(lv[l]["h"])
        ref = cv2mod.resize(prev, (w, h), interpolation=cv2mod.INTER_LINEAR_EXACT)
        got = oracle.resize_linear_exact(prev, w, h)
        assert np.array_equal(got, ref), f"level {l}"
        prev = ref


def test_fast_equals_cv2(oracle, cv2mod, pairs):
    img = pairs(1)[1]
    fd = cv2mod.FastFeatureDetector_create(20, True)
    ref = fd.detect(img, None)
    got = oracle.fast_nms(img, 20)
    assert len(got) == len(ref)
    assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in ref] == [
        (int(c["x"]), int(c["y"]), int(c["resp"])) for c in got]


def test_blur_equals_sepfilter(oracle, cv2mod, pairs):
    img = pairs(2)[0]
    k = cv2mod.getGaussianKernel(7, 2.0, cv2mod.CV_32F)
    ref = cv2mod.sepFilter2D(img, -1, k, k, borderType=cv2mod.BORDER_REFLECT_101)
    assert np.array_equal(oracle.blur7(img), ref)


def test_bf_crosscheck_equals_cv2(oracle, cv2mod, pairs):
    l, r = pairs(0)
    _, dl = oracle.detect_and_compute(l, 1200)
    _, dr = oracle.detect_and_compute(r, 1200)
    ref = cv2mod.BFMatcher(cv2mod.NORM_HAMMING, True).match(dl, dr)
    got = oracle.bf_match_crosscheck(dl, dr)
    assert [(m.queryIdx, m.trainIdx, int(m.distance)) for m in ref] == [
        (int(m["query_idx"]), int(m["train_idx"]), int(m["distance"])) for m in got]


def test_response_ties_keep_more_than_nfeatures_like_cv2(oracle, cv2mod):
    """retainBest keeps every keypoint tied with the n-th response: the oracle must return the same surplus, in the
    same (libstdc++ nth_element/partition) order, as cv::ORB."""
    from oracle import cv2_ref

    yy, xx = np.mgrid[0:480, 0:752]
    dx, dy = np.abs((xx % 12) - 6), np.abs((yy % 12) - 6)
    img = (20 + 200 * np.clip(1 - (dx + dy) / 4.0, 0, 1)).astype(np.uint8)
    k0, d0 = cv2_ref.detect_and_compute(img, 100)
    k1, d1 = oracle.detect_and_compute(img, 100)
    assert len(k0) > 100
    assert_struct_equal(k1, k0, "keypoints")
    assert np.array_equal(d1, d0)
