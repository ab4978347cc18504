"""The reference's actual extraction engine, reached through Python cv2 — TEST INFRASTRUCTURE ONLY.

/root/reference/src/tracking/frame/stereo.rs:38-48 builds
    features2d::ORB::create(n_features, 1.2, 8, 31, 0, 2, HARRIS_SCORE, 31, 20)
and :72-73 calls detect_and_compute(image, empty_mask, ..., useProvidedKeypoints=false).  cv2.ORB_create(...) /
detectAndCompute is the same C++ cv::ORB_Impl, so its output IS the reference output for row a1 of SURVEY.md §8.
`stereo_process` appends the restated Rust loops (oracle/stereo_oracle.cpp) to form the whole
StereoProcessor::process (stereo.rs:52-66).
"""
from __future__ import annotations

import numpy as np

from . import orb_oracle as oo


def cv2_module():
    import cv2  # noqa: PLC0415  (optional dependency of the test infrastructure)

    return cv2


def make_orb(nfeatures=1200):
    cv2 = cv2_module()
    return cv2.ORB_create(int(nfeatures), 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)


def keypoints_to_array(kps):
    out = np.zeros(len(kps), dtype=oo.KP_DTYPE)
    for i, k in enumerate(kps):
        out[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return out


def detect_and_compute(img, nfeatures=1200, orb=None):
    """-> (keypoints[KP_DTYPE], descriptors[n,32] u8) exactly as the reference's detect_features returns them."""
    orb = orb or make_orb(nfeatures)
    kps, desc = orb.detectAndCompute(np.ascontiguousarray(img, dtype=np.uint8), None)
    if desc is None:
        desc = np.zeros((0, 32), dtype=np.uint8)
    return keypoints_to_array(kps), desc


def stereo_process(left, right, camera, nfeatures=1200, orb=None):
    """StereoProcessor::process (stereo.rs:52-66): cv::ORB on L then R, match_features, triangulate."""
    orb = orb or make_orb(nfeatures)
    kl, dl = detect_and_compute(left, nfeatures, orb)
    kr, dr = detect_and_compute(right, nfeatures, orb)
    m = oo.match_features(kl, dl, kr, dr, camera)
    pts, valid = oo.triangulate(kl, kr, m, camera)
    return dict(kp_left=kl, desc_left=dl, kp_right=kr, desc_right=dr, matches=m, points=pts, points_valid=valid)
