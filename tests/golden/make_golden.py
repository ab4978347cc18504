"""Generates tests/golden/*.npz by running the reference's actual extraction engine (cv::ORB through cv2, the same
C++ function /root/reference/src/tracking/frame/stereo.rs:72-73 calls) on fixed inputs, followed by the restated Rust
loops.  Run in the build container:  python tests/golden/make_golden.py
The inputs are stored with the outputs so the fixtures do not depend on the synthetic generator's float paths."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "orb-slam3-rust_b200"))
import cv2  # noqa: E402
from oracle import cv2_ref, orb_oracle as oo  # noqa: E402
from vslam_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def one(name, seed, w, h, n):
    l, r = synth.make_stereo_pair(seed, w, h)
    l, r = l.numpy(), r.numpy()
    res = cv2_ref.stereo_process(l, r, synth.EUROC_CAMERA, n)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(res["desc_left"], res["desc_right"])
    bfm = np.array([(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in bf], dtype=oo.DMATCH_DTYPE)
    np.savez_compressed(
        os.path.join(HERE, name), left=l, right=r, n_features=n, cv2_version=cv2.__version__,
        kp_left=res["kp_left"], desc_left=res["desc_left"], kp_right=res["kp_right"], desc_right=res["desc_right"],
        matches=res["matches"], points=res["points"], points_valid=res["points_valid"], bf_crosscheck=bfm,
    )
    print(name, len(res["kp_left"]), len(res["kp_right"]), len(res["matches"]), len(bfm))


if __name__ == "__main__":
    one("stereo_480x320_n500_seed7.npz", 7, 480, 320, 500)
    one("stereo_752x480_n1200_seed0.npz", 0, 752, 480, 1200)
