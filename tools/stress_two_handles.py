"""Stress of the threading contract (one handle per host thread): two handles of different geometry created and driven
from two threads at once, repeated; every result is compared with the first iteration's (and that one with the oracle).
usage: python tools/stress_two_handles.py [iterations] [threads_mode: 2 | 1a | 1b]"""
import sys, os, threading, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "orb-slam3-rust_b200")]
from vslam_b200 import CameraModel, StereoProcessor, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = sys.argv[2] if len(sys.argv) > 2 else "2"
cam = CameraModel(**synth.EUROC_CAMERA)

def pair(seed, w=752, h=480):
    l, r = synth.make_stereo_pair(seed, w, h)
    return l.numpy(), r.numpy()

jobs = [(752, 480, 1200, [pair(s) for s in (30, 31, 32, 33)]), (640, 427, 600, [pair(s, 640, 427) for s in (34, 35, 36, 37)])]
if mode == "1a": jobs = jobs[:1]
if mode == "1b": jobs = jobs[1:]

def digest(fr):
    parts = [fr.left_features.keypoints, fr.right_features.keypoints, fr.left_features.descriptors, fr.right_features.descriptors,
             fr.matches_lr, fr.points_cam, fr.points_valid]
    return [hashlib.sha1(np.ascontiguousarray(p).tobytes()).hexdigest()[:12] for p in parts] + [len(fr.left_features.keypoints), len(fr.right_features.keypoints)]

golden = {}
bad = 0
for it in range(iters):
    results, errors = {}, []
    def work(k, w, h, nfeat, frames):
        try:
            with StereoProcessor(cam, nfeat, w, h, max_batch=2) as sp:
                outs = []
                for rep in range(3):
                    o = [sp.process(l, r) for l, r in frames[:2]]
                    o += sp.process_batch(np.stack([f[0] for f in frames[2:]]), np.stack([f[1] for f in frames[2:]]))
                    outs.append([digest(x) for x in o])
                results[k] = outs
        except Exception as e:
            errors.append(repr(e))
    th = [threading.Thread(target=work, args=(k, *j)) for k, j in enumerate(jobs)]
    for t in th: t.start()
    for t in th: t.join()
    if errors:
        print("iter", it, "ERRORS", errors, flush=True); bad += 1; continue
    for k in results:
        for rep, outs in enumerate(results[k]):
            for fi, d in enumerate(outs):
                g = golden.setdefault((k, fi), d)
                if g != d:
                    bad += 1
                    names = ["kpL", "kpR", "descL", "descR", "matches", "pts", "valid", "nL", "nR"]
                    diff = [n for n, a, b in zip(names, g, d) if a != b]
                    print(f"iter {it} handle {k} rep {rep} frame {fi} ({'single' if fi < 2 else 'batch'}): differs in {diff} n={d[-2:]} golden n={g[-2:]}", flush=True)
print("iterations", iters, "mismatching results", bad)
