"""`classic` mode: per-cell FAST with threshold retry, quadtree distribution, row-band stereo with
SAD refinement and parabolic sub-pixel disparity.

NOT REFERENCE PARITY.  jurmy24/orb-slam3-rust contains none of these stages; the specification is this repository's
own CPU model (oracle/classic_model.cpp, written from the published ORB-SLAM3 algorithm).  The CPU tests below check
that model's invariants; the GPU tests compare the CUDA kernels with the model bit for bit."""
import numpy as np
import pytest

from conftest import assert_struct_equal


def low_contrast(img):
    """Squeezes the image into a narrow grey band so that most 35-px cells have no corner at iniThFAST = 20 and fall back
    to minThFAST = 7."""
    return (96.0 + (img.astype(np.float32) - 128.0) * 0.22).round().clip(0, 255).astype(np.uint8)


def test_model_extraction_invariants(oracle, pairs):
    left, _ = pairs(0)
    kp, desc = oracle.classic_detect_and_compute(left, 1200)
    lv = oracle.level_geometry(752, 480, 1200)
    per_level = np.bincount(kp["octave"], minlength=8)
    # the quadtree stops at >= N_l nodes and a split adds at most 3: N_l .. N_l + 2 keypoints on a feature-rich level
    assert all(lv[l]["nfeatures"] <= per_level[l] <= lv[l]["nfeatures"] + 2 for l in range(8))
    assert desc.shape == (len(kp), 32) and (kp["class_id"] == -1).all()
    # keypoints are >= 19 px (EDGE_THRESHOLD) inside their level and at most one per pixel
    for l in range(8):
        k = kp[kp["octave"] == l]
        x = np.rint(k["x"] / lv[l]["scale"]).astype(int)
        y = np.rint(k["y"] / lv[l]["scale"]).astype(int)
        assert x.min() >= 19 and x.max() < lv[l]["w"] - 19 and y.min() >= 19 and y.max() < lv[l]["h"] - 19
        assert len(set(zip(x, y))) == len(k)
    # the distribution is spatially even: every 94 x 96 px block of level 0 holds keypoints
    k0 = kp[kp["octave"] == 0]
    hist, _, _ = np.histogram2d(k0["y"], k0["x"], bins=[5, 8], range=[[0, 480], [0, 752]])
    assert hist.min() >= 2


def test_model_threshold_retry(oracle, pairs):
    left, _ = pairs(1)
    pyr0 = left
    c_hi, keep_hi = oracle.classic_level_candidates(pyr0, 261)
    assert (c_hi[:, 2] + 1 > 20).all()  # textured scene: every cell has corners at iniThFAST, nothing below it
    dim = low_contrast(left)
    c_lo, keep_lo = oracle.classic_level_candidates(dim, 261)
    below = (c_lo[:, 2] + 1 <= 20).sum()
    assert below > 200 and (c_lo[:, 2] + 1 > 7).all()  # retry cells contribute corners with 7 < S <= 20
    # a cell is either an iniTh cell or a retry cell, never mixed
    w_cell = int(np.ceil((752 - 32) / ((752 - 32) // 35)))
    h_cell = int(np.ceil((480 - 32) / ((480 - 32) // 35)))
    cell = (c_lo[:, 1] - 19) // h_cell * 100 + (c_lo[:, 0] - 19) // w_cell
    for cid in np.unique(cell)[:50]:
        s = c_lo[cell == cid, 2] + 1
        assert (s > 20).all() or (s <= 20).all()
    assert 261 <= len(keep_lo) <= 263 or len(keep_lo) == len(c_lo)


def test_model_stereo_subpixel(oracle, pairs, camera_dict):
    left, right = pairs(2)
    kl, dl = oracle.classic_detect_and_compute(left, 1200)
    kr, dr = oracle.classic_detect_and_compute(right, 1200)
    mi, od, ur, dp = oracle.classic_stereo(kl, dl, kr, dr, left, right, camera_dict)
    ok = mi >= 0
    assert ok.sum() > 400 and (od[ok] < 75).all()
    disp = kl["x"][ok] - ur[ok]
    # the synthetic scene has three disparity bands (6 / 14 / 30 px): the refined disparities sit within 0.5 px of them
    band = np.array([6.0, 14.0, 30.0])[np.clip((kl["y"][ok] // 160).astype(int), 0, 2)]
    assert np.mean(np.abs(disp - band) < 0.5) > 0.9
    fxb = np.float32(camera_dict["fx"] * camera_dict["baseline"])
    assert np.allclose(dp[ok], fxb / disp, rtol=1e-6)
    # +-1 octave consistency
    assert (np.abs(kl["octave"][ok] - kr["octave"][mi[ok]]) <= 1).all()


# ----------------------------------------------------------------------------------------------- GPU vs the model
def _check_classic(sp, oracle, left, right, cam, nfeat):
    fr = sp.process(left, right)
    ur, dp = sp.classic_stereo_extra(0)
    kl, dl = oracle.classic_detect_and_compute(left, nfeat)
    kr, dr = oracle.classic_detect_and_compute(right, nfeat)
    assert_struct_equal(fr.left_features.keypoints, kl, "classic left keypoints")
    assert_struct_equal(fr.right_features.keypoints, kr, "classic right keypoints")
    assert np.array_equal(fr.left_features.descriptors, dl) and np.array_equal(fr.right_features.descriptors, dr)
    mi, od, mur, mdp = oracle.classic_stereo(kl, dl, kr, dr, left, right, cam, nfeat)
    ok = np.flatnonzero(mi >= 0)
    assert list(fr.matches_lr["query_idx"]) == list(ok)
    assert list(fr.matches_lr["train_idx"]) == list(mi[ok])
    assert list(fr.matches_lr["distance"]) == [float(v) for v in od[ok]]
    n = len(kl)
    assert np.array_equal(ur[:n].view(np.uint32), mur.view(np.uint32)), "u_right differs from the model"
    assert np.array_equal(dp[:n].view(np.uint32), mdp.view(np.uint32)), "depth differs from the model"
    assert np.max(np.abs(ur[:n] - mur), initial=0.0) <= 1e-4  # north_star tolerance (in fact exact)
    assert np.array_equal(fr.points_valid[:n], (mi >= 0).astype(np.uint8))
    z = mdp[ok].astype(np.float64)
    x = (kl["x"][ok].astype(np.float64) - cam["cx"]) * z / cam["fx"]
    y = (kl["y"][ok].astype(np.float64) - cam["cy"]) * z / cam["fy"]
    assert np.array_equal(fr.points_cam[ok], np.c_[x, y, z])
    return fr, len(ok)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 3])
def test_classic_matches_model_752x480(oracle, pairs, camera_dict, seed):
    from vslam_b200 import CameraModel, StereoProcessor

    left, right = pairs(seed)
    with StereoProcessor(CameraModel(**camera_dict), 1200, mode="classic", max_batch=2) as sp:
        fr, nm = _check_classic(sp, oracle, left, right, camera_dict, 1200)
        assert nm > 400
        lv = oracle.level_geometry(752, 480, 1200)
        pyr = oracle.build_pyramid(left, lv)
        for l in (0, 3, 7):
            ref, _ = oracle.classic_level_candidates(pyr[l], int(lv[l]["nfeatures"]))
            assert np.array_equal(sp.debug_classic_candidates(0, l), ref), f"per-cell candidates of level {l}"


@pytest.mark.gpu
def test_classic_threshold_retry_and_sparse_scenes(oracle, pairs, camera_dict):
    from vslam_b200 import CameraModel, StereoProcessor

    left, right = pairs(5)
    with StereoProcessor(CameraModel(**camera_dict), 1200, mode="classic", max_batch=2) as sp:
        _check_classic(sp, oracle, low_contrast(left), low_contrast(right), camera_dict, 1200)
        # almost empty scene: a few rectangles on a flat background (fewer candidates than N: the quadtree keeps all)
        img = np.full((480, 752), 90, np.uint8)
        rng = np.random.default_rng(3)
        for _ in range(12):
            x, y = int(rng.integers(40, 650)), int(rng.integers(40, 400))
            img[y:y + 30, x:x + 50] = int(rng.integers(150, 255))
        fr, _ = _check_classic(sp, oracle, img, np.roll(img, -9, axis=1), camera_dict, 1200)
        assert 0 < len(fr.left_features.keypoints) < 400
        blank = np.full((480, 752), 7, np.uint8)
        fr = sp.process(blank, blank)
        assert len(fr.left_features.keypoints) == 0 and len(fr.matches_lr) == 0


@pytest.mark.gpu
def test_classic_other_sizes_and_batch(oracle, pairs, camera_dict):
    from vslam_b200 import CameraModel, StereoProcessor

    left, right = pairs(7, 1280, 720)
    with StereoProcessor(CameraModel(**camera_dict), 2000, 1280, 720, mode="classic", max_batch=2) as sp:
        _check_classic(sp, oracle, left, right, camera_dict, 2000)
    # a batch through the pipelined entry point equals frame-by-frame processing
    with StereoProcessor(CameraModel(**camera_dict), 1200, mode="classic", max_batch=2) as sp:
        ls = np.stack([pairs(s)[0] for s in (10, 11, 12)])
        rs = np.stack([pairs(s)[1] for s in (10, 11, 12)])
        frs = sp.process_batch(ls, rs)
        for i, fr in enumerate(frs):
            kl, dl = oracle.classic_detect_and_compute(ls[i], 1200)
            assert_struct_equal(fr.left_features.keypoints, kl, f"batch frame {i}")
            assert np.array_equal(fr.left_features.descriptors, dl)
