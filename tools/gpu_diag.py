"""Stage-by-stage GPU-vs-oracle diff (development aid; prints everything, never stops at the first mismatch)."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "orb-slam3-rust_b200"))
import numpy as np
from vslam_b200 import CameraModel, StereoProcessor, FeatureSet, synth, STAGE_PYRAMID, STAGE_BLURRED
from oracle import orb_oracle as oo

def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        print(f"  [{name}] SHAPE {a.shape} vs {b.shape}"); return False
    if a.dtype.names:
        ok = a.tobytes() == b.tobytes()
        if not ok:
            for f in a.dtype.names:
                ne = np.flatnonzero(a[f].view(np.uint32) != b[f].view(np.uint32))
                if len(ne): print(f"  [{name}] field {f}: {len(ne)} differ, first idx {ne[0]}: gpu {a[f][ne[0]]} ref {b[f][ne[0]]}")
        else: print(f"  [{name}] ok ({len(a)})")
        return ok
    ne = np.flatnonzero(a.ravel() != b.ravel())
    if len(ne): print(f"  [{name}] {len(ne)}/{a.size} differ, first flat idx {ne[0]}: gpu {a.ravel()[ne[0]]} ref {b.ravel()[ne[0]]}")
    else: print(f"  [{name}] ok ({a.size})")
    return len(ne) == 0

def run(W, H, N, seed):
    print(f"=== {W}x{H} N={N} seed={seed}")
    cam = CameraModel(**synth.EUROC_CAMERA)
    L, R = synth.make_stereo_pair(seed, W, H); L = L.numpy(); R = R.numpy()
    sp = StereoProcessor(cam, N, W, H, max_batch=2)
    t = time.time(); fr = sp.process(L, R, 7); print("process %.1f ms, launches %d" % ((time.time()-t)*1e3, sp.kernel_launches()))
    for eye, img in enumerate((L, R)):
        st = oo.detect_stages(img, N)
        print(f" eye {eye}: oracle kps {len(st['keypoints'])}")
        for l in range(8):
            cmp(f"pyr{l}", sp.debug_level_image(STAGE_PYRAMID, eye, l), st["pyramid"][l])
        for l in range(8):
            c = sp.debug_fast_corners(eye, l); o = st["fast"][l]
            ref = np.stack([o["x"], o["y"], o["resp"].astype(np.int32)], 1) if len(o) else np.zeros((0,3),np.int32)
            cmp(f"fast{l}", c, ref)
        for l in range(8):
            cmp(f"blur{l}", sp.debug_level_image(STAGE_BLURRED, eye, l), st["blurred"][l])
        fs = fr.left_features if eye == 0 else fr.right_features
        cmp("keypoints", fs.keypoints, st["keypoints"])
        if len(fs.keypoints) == len(st["keypoints"]):
            cmp("descriptors", fs.descriptors, st["descriptors"])
            if (fs.keypoints["x"] == st["keypoints"]["x"]).all():
                db = np.unpackbits(fs.descriptors ^ st["descriptors"], axis=1).sum(1)
                print("   descriptor rows differing:", int((db > 0).sum()), "bits", int(db.sum()))
    kl, dl = oo.detect_and_compute(L, N); kr, dr = oo.detect_and_compute(R, N)
    m = oo.match_features(kl, dl, kr, dr, synth.EUROC_CAMERA); pts, valid = oo.triangulate(kl, kr, m, synth.EUROC_CAMERA)
    cmp("matches(pipeline)", fr.matches_lr, m); cmp("pts", fr.points_cam, pts); cmp("valid", fr.points_valid, valid)
    # standalone pieces on oracle inputs
    try:
        cmp("match_features(standalone)", sp.match_features(FeatureSet(kl, dl), FeatureSet(kr, dr)), m)
        p2, v2 = sp.triangulate(FeatureSet(kl, dl), FeatureSet(kr, dr), m); cmp("triangulate", p2, pts); cmp("tri valid", v2, valid)
        n = min(len(dl), len(dr)); hd = sp.descriptor_distances(dl[:n], dr[:n])
        cmp("hamming", hd, np.array([oo.descriptor_distance(dl[i], dr[i]) for i in range(n)], dtype=np.uint32))
        cs, items = sp.grid_build(kl); ocs, oit = oo.grid_build(kl); cmp("grid cell_start", cs, ocs); cmp("grid items", items, oit)
        rng = np.random.default_rng(5)
        q = np.c_[rng.uniform(-60, 820, 300), rng.uniform(-60, 540, 300), np.full(300, 15.0)]
        got = sp.grid_query(kl, q); bad = 0
        for i in range(len(q)):
            ref = oo.grid_query(ocs, oit, *q[i])
            if not np.array_equal(got[i], ref): bad += 1
        print(f"  [grid_query] {bad}/300 differ")
        M = 1500
        src = rng.integers(0, len(kl), M)
        uv = np.c_[kl["x"][src] + rng.normal(0, 6, M), kl["y"][src] + rng.normal(0, 6, M)].astype(np.float64)
        uv[:50] += rng.uniform(-900, 900, (50, 2))
        md = dl[src].copy(); flip = rng.integers(0, 256, (M, 12)); 
        for i in range(M):
            for b in flip[i][: rng.integers(0, 12)]: md[i, b // 8] ^= 1 << (b % 8)
        for mode in (0, 1):
            gi, gd = sp.search_by_projection(FeatureSet(kl, dl), uv, md, mode=mode)
            oi, od = oo.search_by_projection(kl, dl, uv, md, mode, bound_w=2*cam.cx, bound_h=2*cam.cy)
            cmp(f"search mode{mode} idx", gi, oi); cmp(f"search mode{mode} dist", gd, od); print("   matched", int((oi >= 0).sum()))
        cmp("bf crosscheck", sp.bf_match_crosscheck(dl, dr), oo.bf_match_crosscheck(dl, dr))
        fs = sp.detect_features(L); cmp("mono keypoints", fs.keypoints, kl); cmp("mono desc", fs.descriptors, dl)
    except Exception:
        traceback.print_exc()
    sp.close()

if __name__ == "__main__":
    run(752, 480, 1200, 0)
    run(1280, 720, 2000, 3)
    # multi-chunk batch
    print("=== batch 5 frames, max_batch 2")
    cam = CameraModel(**synth.EUROC_CAMERA)
    Ls, Rs = synth.make_stereo_batch(range(10, 15)); Ls = Ls.numpy(); Rs = Rs.numpy()
    sp = StereoProcessor(cam, 1200, max_batch=2)
    frs = sp.process_batch(Ls, Rs, timestamps_ns=np.arange(5) + 100)
    for i, fr in enumerate(frs):
        kl, dl = oo.detect_and_compute(Ls[i], 1200); kr, dr = oo.detect_and_compute(Rs[i], 1200)
        m = oo.match_features(kl, dl, kr, dr, synth.EUROC_CAMERA)
        ok = fr.left_features.keypoints.tobytes() == kl.tobytes() and fr.right_features.keypoints.tobytes() == kr.tobytes() and np.array_equal(fr.left_features.descriptors, dl) and np.array_equal(fr.right_features.descriptors, dr) and fr.matches_lr.tobytes() == m.tobytes()
        print(f"  frame {i}: {'ok' if ok else 'MISMATCH'} ts={fr.timestamp_ns} n={len(kl)},{len(kr)},{len(m)}")
    # throughput probe
    import torch
    n = 256
    Lb, Rb = synth.make_stereo_batch(range(64), device="cuda")
    Lb = Lb.repeat(n // 64, 1, 1).contiguous(); Rb = Rb.repeat(n // 64, 1, 1).contiguous()
    sp2 = StereoProcessor(cam, 1200, max_batch=64)
    for it in range(3):
        torch.cuda.synchronize(); t = time.time()
        sp2.process_batch_device(n, Lb.data_ptr(), Rb.data_ptr(), 752 * 480, 752, None)
        torch.cuda.synchronize(); dt = time.time() - t
        print(f"  device-resident {n} frames: {dt*1e3:.1f} ms -> {n/dt:.0f} frames/s")
    sp2.set_profiling(True)
    sp2.process_batch_device(n, Lb.data_ptr(), Rb.data_ptr(), 752 * 480, 752, None)
    ms = sp2.last_stage_ms(); print("  stage ms [pyr,fast,select,blur,brief,match,total] per %d frames:" % n, np.round(ms, 3))
    sp2.set_profiling(False)
    Lh = Lb.cpu().pin_memory(); Rh = Rb.cpu().pin_memory()
    ob = sp2.new_output_buffers(n)
    for it in range(2):
        t = time.time(); sp2.process_batch(Lh.numpy(), Rh.numpy(), out=ob, copy=False); dt = time.time() - t
        print(f"  host e2e {n} frames: {dt*1e3:.1f} ms -> {n/dt:.0f} frames/s")
